"""Dev script: a larger randomised parity run than the test-suite's — areas + contacts of the CUDA path (through the
Python host layer and the C ABI) against the compiled reference (oracle/_ref) or the C restatement, bit for bit.
  python tools/gpu_stress.py [n_small] [n_neurite_volumes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import xs3d
from xs3d import synthetic as syn
import cases
from refapi import Ref, Oracle, have_ref
chk = Ref() if have_ref() else Oracle()
print("checker:", type(chk).__name__)
n_small = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n_vol = int(sys.argv[2]) if len(sys.argv) > 2 else 12
bits = lambda a: np.asarray(a, np.float32).view(np.uint32)
rng = np.random.default_rng(2026)
t0 = time.time(); bad = 0
for it in range(n_small):
  vol, p, n, w = cases.rand_case(rng, max_side=int(rng.integers(4, 40)))
  if it % 3 == 0: vol = np.ascontiguousarray(vol)                       # C-order ingest
  a, c = chk.area(np.asfortranarray(vol), p, n, w)
  a2, c2 = xs3d.cross_sectional_area(vol, p, n, w, return_contact=True)
  if bits(a) != bits(a2) or c != c2:
    bad += 1; print("MISMATCH small", it, vol.shape, p, n, w, a, a2, c, c2)
print("small random volumes: %d cases, %d mismatches (%.1fs)" % (n_small, bad, time.time() - t0), flush=True)
t0 = time.time(); total = 0
for k in range(n_vol):
  N = int(rng.choice([96, 128, 160, 200, 256, 320]))
  vol, verts, tans = syn.make_neurite(N, seed=100 + k, n_branches=int(rng.integers(3, 9)))
  if len(verts) == 0: continue
  nq = 1500
  qp, qn = syn.draw_queries(verts, tans, nq, seed=k)
  qp = qp + rng.uniform(-0.45, 0.45, size=qp.shape).astype(np.float32) * (rng.random((nq, 1)) < 0.3)   # fractional vertices
  qn = qn + rng.normal(scale=0.3, size=qn.shape).astype(np.float32) * (rng.random((nq, 1)) < 0.5)     # tilted planes
  qn[np.all(qn == 0, axis=1)] = (0, 0, 1)
  w = cases.ANISOTROPIES[int(rng.integers(len(cases.ANISOTROPIES)))]
  oa, oc = chk.area_batch(vol, qp, qn, w)
  with xs3d.Volume(vol) as V:
    a, c = V.cross_sectional_area_batch(qp, qn, w, return_contact=True)
    st = V.stats()
  m = int((bits(a) != bits(oa)).sum() + (c != oc).sum())
  bad += m; total += nq
  print("neurite %d^3 seed %d: %d queries, %d mismatches, to_mid %d to_wide %d" % (N, 100 + k, nq, m, st["escalated_mid"], st["escalated_big"]), flush=True)
print("neurite volumes: %d queries in %.1fs; TOTAL mismatches: %d" % (total, time.time() - t0, bad))
sys.exit(1 if bad else 0)
