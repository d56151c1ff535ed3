"""Dev script: latency of single large sections (BASELINE configs 0 and 1) with a resident volume."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import xs3d
from xs3d import synthetic as syn
from refapi import Ref, Oracle, have_ref
ref = Ref() if have_ref() else Oracle()
w = [32, 32, 40]
sph = syn.make_sphere()
V = xs3d.Volume(sph)
p = np.array([250, 250, 250], np.float32); n = np.array([0.01, 0.033, 0.9], np.float32)
V.cross_sectional_area(p, n, w)
for _ in range(3):
  t = time.time(); r = V.cross_sectional_area(p, n, w, return_contact=True); dt = time.time() - t
  print("config 0 sphere-500 single query (resident): %.2f ms" % (dt * 1e3), r, V.timing(), V.stats()["samples"])
pr = V.profile()["cta_tier"]; tot = sum(list(pr.values())[:3]); print({k: "%.1f%%" % (100 * v / tot) for k, v in pr.items()})
t = time.time(); rr = ref.area(sph, p, n, w); print("reference: %.2f ms" % ((time.time() - t) * 1e3), rr)
for _ in range(4):
  t = time.time(); idx, val, ct = V.cross_section_sparse(p, n, w); print("cross_section sparse (resident): %.2f ms nnz %d" % ((time.time() - t) * 1e3, len(idx)))
t = time.time(); img = xs3d.cross_section(sph, p, n, w); print("cross_section drop-in (upload + dense 500 MB host image): %.1f ms" % ((time.time() - t) * 1e3))
t = time.time(); rimg, _ = ref.section(sph, p, n, w); print("reference cross_section: %.1f ms" % ((time.time() - t) * 1e3), np.array_equal(img.view(np.uint32), rimg.view(np.uint32)))
V.close()
cyl, axis = syn.make_cylinder()
V = xs3d.Volume(cyl)
for v in [(256, 256, 256), (496, 352, 316)]:
  V.cross_sectional_area(v, axis, w)
  t = time.time(); r = V.cross_sectional_area(v, axis, w, return_contact=True); dt = time.time() - t
  t = time.time(); rr = ref.area(cyl, v, axis, w); dr = time.time() - t
  print("config 1 cylinder", v, "%.2f ms" % (dt * 1e3), r, "reference %.2f ms" % (dr * 1e3), rr)
  V.cross_section_sparse(v, axis, w)
  t = time.time(); idx, val, ct = V.cross_section_sparse(v, axis, w); dt = time.time() - t
  t = time.time(); rimg, rct = ref.section(cyl, v, axis, w); dr = time.time() - t
  print("   cross_section sparse (resident) %.2f ms nnz %d contact %d; reference (dense 512 MB image) %.1f ms" % (dt * 1e3, len(idx), ct, dr * 1e3))
