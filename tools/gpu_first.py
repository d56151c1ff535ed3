"""Dev script: first contact with the GPU — parity vs the compiled reference + rough timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import xs3d
from xs3d import synthetic as syn
from refapi import Ref, Oracle, have_ref

print("torch", torch.__version__, "cuda", torch.cuda.is_available(), torch.cuda.get_device_name(0))
ref = Ref() if have_ref() else Oracle()
N = int(os.environ.get("XS_N", "512"))
NQ = int(os.environ.get("XS_NQ", "10000"))
t = time.time(); vol, verts, tans = syn.make_neurite(N); print("neurite", N, vol.sum(), len(verts), "%.1fs" % (time.time() - t))
pos, nrm = syn.draw_queries(verts, tans, NQ)
w = [32, 32, 40]
t = time.time(); V = xs3d.Volume(vol); print("volume create+load %.3fs" % (time.time() - t))
t = time.time(); V.load(vol); print("reload %.4fs" % (time.time() - t))
for rep in range(3):
  t = time.time(); a, c = V.cross_sectional_area_batch(pos, nrm, w, return_contact=True); dt = time.time() - t
  print("batch host->host %.4fs  %.0f xs/s" % (dt, NQ / dt), V.stats())
dpos = torch.from_numpy(pos).cuda(); dnrm = torch.from_numpy(nrm).cuda()
for rep in range(5):
  torch.cuda.synchronize(); t = time.time(); da, dc = V.cross_sectional_area_batch(dpos, dnrm, w, return_contact=True); torch.cuda.synchronize(); dt = time.time() - t
  print("batch device %.5fs  %.0f xs/s" % (dt, NQ / dt))
nchk = min(NQ, int(os.environ.get("XS_CHK", "3000")))
t = time.time(); ra, rc = ref.area_batch(vol, pos[:nchk], nrm[:nchk], w); dt = time.time() - t
print("reference 1 thread: %.0f xs/s" % (nchk / dt))
eq = (a[:nchk].view(np.uint32) == ra.view(np.uint32))
print("AREA bit-equal %d / %d   contact equal %d / %d" % (eq.sum(), nchk, (c[:nchk] == rc).sum(), nchk))
print("device path equal to host path:", np.array_equal(da.cpu().numpy().view(np.uint32), a.view(np.uint32)), np.array_equal(dc.cpu().numpy(), c))
bad = np.nonzero(~eq)[0][:10]
for b in bad: print("  mismatch q", b, pos[b], nrm[b], a[b], ra[b], c[b], rc[b])
# small random cases through the single-call drop-in
rng = np.random.default_rng(3)
nb = 0; tot = 0
for it in range(300):
  shape = rng.integers(1, 20, size=3)
  v = np.asfortranarray(rng.random(shape) < rng.uniform(0.3, 0.95))
  fg = np.argwhere(v)
  if len(fg) == 0: continue
  p = fg[rng.integers(len(fg))].astype(np.float32)
  k = rng.integers(4)
  n = rng.normal(size=3).astype(np.float32)
  if k == 0: n = (np.eye(3)[rng.integers(3)] * rng.choice([-1, 1])).astype(np.float32)
  if k == 1: n[rng.integers(3)] = 0
  if not np.any(n): n = np.array([0, 0, 1], np.float32)
  ww = [[1, 1, 1], [32, 32, 40], [0.5, 2, 1.25]][rng.integers(3)]
  g = xs3d.cross_sectional_area(v, p, n, ww, return_contact=True)
  r = ref.area(v, p, n, ww)
  tot += 1
  if np.float32(g[0]).view(np.uint32) != r[0].view(np.uint32) or g[1] != r[1]:
    nb += 1
    if nb < 8: print("  small mismatch", shape, p, n, ww, g, r)
  if it % 5 == 0:
    s, sc = xs3d.cross_section(v, p, n, ww, return_contact=True)
    rs, rsc = ref.section(v, p, n, ww, method=0)
    if not np.array_equal(s.view(np.uint32), rs.view(np.uint32)) or sc != rsc:
      nb += 1
      if nb < 8: print("  section mismatch", shape, p, n, ww, sc, rsc, np.abs(s - rs).max())
print("small cases: %d, mismatches %d" % (tot, nb))
# big single queries: sphere config
sph = syn.make_sphere()
t = time.time(); g = xs3d.cross_sectional_area(sph, [250, 250, 250], [0.01, 0.033, 0.9], [32, 32, 40], return_contact=True); dt = time.time() - t
print("sphere-500 single call:", g, "%.4fs" % dt, "(reference 128656112.0, 0)")
xs3d.set_shape(sph)
for rep in range(2):
  t = time.time(); g = xs3d.cross_sectional_area(sph, [250, 250, 250], [0.01, 0.033, 0.9], [32, 32, 40], return_contact=True, use_persistent_data=True); dt = time.time() - t
  print("sphere-500 resident:", g, "%.4fs" % dt)
r = ref.area(sph, [250, 250, 250], [0.01, 0.033, 0.9], [32, 32, 40])
print("ref:", r, "equal:", np.float32(g[0]).view(np.uint32) == r[0].view(np.uint32))
t = time.time(); s, sc = xs3d.cross_section(sph, [250, 250, 250], [0.01, 0.033, 0.9], [32, 32, 40], return_contact=True); dt = time.time() - t
print("sphere cross_section: nnz", np.count_nonzero(s), "sum64", s.sum(dtype=np.float64), "%.3fs" % dt, "(ref nnz 131559, sum 128692269.153)")
xs3d.clear_shape()
