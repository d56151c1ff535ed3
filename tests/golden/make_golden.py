"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference (dev container only).

  python tests/golden/make_golden.py

Imports the reference Python package from /root/reference with its own extension compiled as-is into
oracle/_ref (make -C oracle ref).  Inputs are regenerated from seeds by tests/cases.py, so only the
reference OUTPUTS (and the small packed volumes, for robustness against numpy RNG changes) are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200", "xs3d"))

import xs3d  # the reference package
import cases
import synthetic
from refapi import Oracle

assert xs3d.__file__.startswith("/root/reference"), xs3d.__file__
orc = Oracle()
out = {}

# ---- small randomised volumes: area (fast + slow), contact, cross_section methods 0/1/2 ---------------
rng = np.random.default_rng(2024)
NC = 160
shapes, packed, offs = [], [], [0]
pos, nrm, ani = [], [], []
area, contact, sarea, scontact = [], [], [], []
sec = {m: {"idx": [], "val": [], "off": [0], "contact": []} for m in (0, 1, 2)}
for i in range(NC):
  vol, p, n, w = cases.rand_case(rng, max_side=12)
  shapes.append(vol.shape)
  bits_ = np.packbits(vol.ravel(order="F"))
  packed.append(bits_)
  offs.append(offs[-1] + len(bits_))
  pos.append(p); nrm.append(n); ani.append(w)
  a, c = xs3d.cross_sectional_area(vol, p, n, w, return_contact=True)
  area.append(np.float32(a)); contact.append(c)
  a, c = xs3d.cross_sectional_area(vol, p, n, w, return_contact=True, slow_method=True)
  sarea.append(np.float32(a)); scontact.append(c)
  for m in (0, 1, 2):
    img, c = xs3d.cross_section(vol, p, n, w, return_contact=True, method=m)
    flat = img.ravel(order="F")
    nz = np.flatnonzero(flat.view(np.uint32))
    sec[m]["idx"].append(nz.astype(np.uint32)); sec[m]["val"].append(flat[nz])
    sec[m]["off"].append(sec[m]["off"][-1] + len(nz)); sec[m]["contact"].append(c)
out.update(small_shapes=np.array(shapes, np.int32), small_packed=np.concatenate(packed), small_off=np.array(offs, np.int64),
           small_pos=np.array(pos, np.float32), small_nrm=np.array(nrm, np.float32), small_ani=np.array(ani, np.float32),
           small_area=np.array(area, np.float32), small_contact=np.array(contact, np.uint8),
           small_slow_area=np.array(sarea, np.float32), small_slow_contact=np.array(scontact, np.uint8))
for m in (0, 1, 2):
  out[f"sec{m}_idx"] = np.concatenate(sec[m]["idx"]); out[f"sec{m}_val"] = np.concatenate(sec[m]["val"])
  out[f"sec{m}_off"] = np.array(sec[m]["off"], np.int64); out[f"sec{m}_contact"] = np.array(sec[m]["contact"], np.uint8)

# ---- slices: shapes + raw bytes; `ub` marks cases where the reference itself hits its -0.5 UB cast --------
rng = np.random.default_rng(77)
NS = 200
sl_shapes, sl_bytes, sl_off, sl_ub = [], [], [0], []
for i in range(NS):
  lab, p, n, w, sb, crop = cases.rand_labels_case(rng, max_side=16)
  r = xs3d.slice(lab, p, n, w, sb, crop)
  orc.lib.oracle_projection_clamped()
  orc.projection(lab, p, n, w, sb, crop)
  sl_ub.append(orc.lib.oracle_projection_clamped() > 0)
  b = np.frombuffer(r.tobytes(order="F"), np.uint8)
  sl_shapes.append(r.shape); sl_bytes.append(b); sl_off.append(sl_off[-1] + len(b))
out.update(slice_shapes=np.array(sl_shapes, np.int64), slice_bytes=np.concatenate(sl_bytes), slice_off=np.array(sl_off, np.int64),
           slice_ub=np.array(sl_ub, bool))

# ---- synthetic neurite 256^3: 1000 skeleton queries ---------------------------------------------------------
vol, verts, tans = synthetic.make_neurite(256, seed=0)
qp, qn = synthetic.draw_queries(verts, tans, 1000, seed=1)
na = np.zeros(1000, np.float32); nc = np.zeros(1000, np.uint8)
for i in range(1000):
  a, c = xs3d.cross_sectional_area(vol, qp[i], qn[i], [32, 32, 40], return_contact=True)
  na[i] = a; nc[i] = c
out.update(neurite256_fg=np.int64(vol.sum()), neurite256_nverts=np.int64(len(verts)), neurite256_area=na, neurite256_contact=nc)

# ---- BASELINE configs 0 and 1 -----------------------------------------------------------------------------------
sph = synthetic.make_sphere()
a, c = xs3d.cross_sectional_area(sph, [250, 250, 250], [0.01, 0.033, 0.9], [32, 32, 40], return_contact=True)
img = xs3d.cross_section(sph, [250, 250, 250], [0.01, 0.033, 0.9], [32, 32, 40])
out.update(sphere_area=np.float32(a), sphere_contact=np.uint8(c), sphere_section_nnz=np.int64(np.count_nonzero(img)),
           sphere_section_sum=np.float64(img.sum(dtype=np.float64)))
del img, sph
cyl, axis = synthetic.make_cylinder()
cv = [(256, 256, 256), (496, 352, 316), (12, 158, 195)]
ca, cc = [], []
for v in cv:
  a, c = xs3d.cross_sectional_area(cyl, v, axis, [32, 32, 40], return_contact=True)
  ca.append(np.float32(a)); cc.append(c)
img, c = xs3d.cross_section(cyl, cv[0], axis, [32, 32, 40], return_contact=True)
out.update(cyl_verts=np.array(cv, np.float32), cyl_axis=axis, cyl_area=np.array(ca, np.float32), cyl_contact=np.array(cc, np.uint8),
           cyl_section_nnz=np.int64(np.count_nonzero(img)), cyl_section_sum=np.float64(img.sum(dtype=np.float64)))

path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
for k in ("sphere_area", "sphere_contact", "sphere_section_nnz", "sphere_section_sum", "cyl_area", "cyl_contact", "cyl_section_nnz", "cyl_section_sum", "neurite256_fg", "neurite256_nverts"):
  print(k, out[k])
print("slice ub cases", int(np.sum(sl_ub)), "of", NS)
