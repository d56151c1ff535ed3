"""The reference's one published benchmark (benchmarks/perf.cpp:5-26, version_timings.log): 500 x cross_sectional_area on a
solid 500^3 volume, pos=(i,250,250), normal (1,1,1), anisotropy 1.  Published: 15.74 s on a MacBook Pro M3 (v1.11.0)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import xs3d
from refapi import Ref, Oracle, have_ref
solid = np.ones((500, 500, 500), bool, order="F")
pos = np.array([[i, 250, 250] for i in range(500)], np.float32)
nrm = np.tile(np.array([1, 1, 1], np.float32), (500, 1))
t = time.time(); V = xs3d.Volume(solid); t_up = time.time() - t
V.cross_sectional_area_batch(pos[:150], nrm[:150], [1, 1, 1])
t = time.time(); a, c = V.cross_sectional_area_batch(pos, nrm, [1, 1, 1], return_contact=True); dt = time.time() - t
print("B200: 500 planes in %.3f s (%.2f ms/plane) + %.3f s upload/ingest; stats %s" % (dt, dt / 500 * 1e3, t_up, V.stats()))
ref = Ref() if have_ref() else Oracle()
idx = list(range(0, 500, 50))
t = time.time(); ra, rc = ref.area_batch(solid, pos[idx], nrm[idx], [1, 1, 1]); dr = time.time() - t
print("reference here, 1 core: %.1f ms/plane on %d sampled planes -> %.1f s extrapolated for 500 (published: 15.74 s on an M3)" % (dr / len(idx) * 1e3, len(idx), dr / len(idx) * 500))
print("bit-identical on the sampled planes:", np.array_equal(ra.view(np.uint32), a[idx].view(np.uint32)), np.array_equal(rc, c[idx]))
