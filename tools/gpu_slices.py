"""Dev script: batched slices at scale (BASELINE configs 4/5 on one GPU): 1024^3 uint64 label volume,
10k cropped planes and a batch of un-cropped planes, plan and fill timed separately (CUDA events, device-resident
output) and end to end (host output); XS_N=512 for a quicker run."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import xs3d
from xs3d import synthetic as syn
from refapi import Ref, Oracle, have_ref
ref = Oracle()      # (the compiled reference reads out of bounds on planes that graze a volume face — its documented UB — and may crash)
N = int(os.environ.get("XS_N", "1024"))
w = [32, 32, 40]
vol, verts, tans = syn.make_neurite(N, seed=0)
lab = syn.make_labels(vol, np.uint64, hashed=True); del vol
qp, qn = syn.draw_queries(verts, tans, 10000, seed=2)
t = time.time(); V = xs3d.Volume(labels=lab); print("label upload (%.1f GiB): %.2fs" % (lab.nbytes / 2**30, time.time() - t), flush=True)
V.set_stream(torch.cuda.current_stream().cuda_stream)
dp, dn = torch.from_numpy(qp).cuda(), torch.from_numpy(qn).cuda()
for crop, n in ((100.0, 10000), (1000.0, 10000), (float("inf"), 200)):
  for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    shapes, offsets = V.slice_plan(dp[:n], dn[:n], w, True, crop)
    torch.cuda.synchronize(); t1 = time.time()
    out = torch.empty(int(offsets[-1]), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize(); t2 = time.time()
    V.slice_fill(0, n, out)
    torch.cuda.synchronize(); t3 = time.time()
  tot = (t1 - t0) + (t3 - t2)
  print("crop=%s: %d planes, %.1f M output elements (%.1f MB): plan %.3f ms, fill %.3f ms (%.0f GB/s of output) -> %.0f planes/s resident; shape[0]=%s" % (
    crop, n, offsets[-1] / 1e6, offsets[-1] * 8 / 1e6, (t1 - t0) * 1e3, (t3 - t2) * 1e3, offsets[-1] * 8 / (t3 - t2) / 1e9, n / tot, tuple(shapes[0])), flush=True)
  t = time.time(); r = V.slice_batch(qp[:n], qn[:n], w, crop=crop); dt = time.time() - t
  print("   end to end (host queries in, ragged host arrays out): %.4fs -> %.0f planes/s" % (dt, n / dt), flush=True)
  k = 40 if crop != 100.0 else 200
  t = time.time()
  for i in range(k):
    o = ref.projection(lab, qp[i], qn[i], w, True, crop)
    assert np.array_equal(o, np.asfortranarray(r[i])), i
  print("   reference, 1 core: %.0f planes/s; first %d planes bit-identical" % (k / (time.time() - t), k), flush=True)
t = time.time(); g = xs3d.slice(lab, qp[0], qn[0], w); print("drop-in xs3d.slice(ndarray) un-cropped incl. label upload: %.3fs" % (time.time() - t))
t = time.time(); g = xs3d.slice(V, qp[0], qn[0], w); print("xs3d.slice(Volume) un-cropped, labels resident: %.4fs" % (time.time() - t))
