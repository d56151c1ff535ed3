"""Dev script: throughput on BASELINE configs 3 and 4 (1024^3 neurite sweep; 1024^3 uint64 slices, crop = 100 nm)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import xs3d
from xs3d import synthetic as syn
from refapi import Ref, Oracle, have_ref
ref = Ref() if have_ref() else Oracle()
w = [32, 32, 40]
t = time.time(); vol, verts, tans = syn.make_neurite(1024, seed=0); print("neurite 1024^3: %d fg voxels, %d skeleton vertices (%.1fs)" % (int(vol.sum()), len(verts), time.time() - t), flush=True)
qp, qn = syn.draw_queries(verts, tans, 10000, seed=1)
t = time.time(); V = xs3d.Volume(vol); print("upload + ingest 1 GiB: %.3fs, ingest kernel %.3f ms" % (time.time() - t, V.timing()["ingest_ms"]))
V.set_stream(torch.cuda.current_stream().cuda_stream)
dp = torch.from_numpy(qp).cuda(); dn = torch.from_numpy(qn).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for rep in range(8):
  flush.fill_(rep)
  s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
  s.record(); a, c = V.cross_sectional_area_batch(dp, dn, w, return_contact=True); e.record(); torch.cuda.synchronize()
  ts.append(s.elapsed_time(e))
ms = float(np.median(ts[3:]))
print("config 3: 10k-query sweep of the 1024^3 neurite, resident: %.3f ms -> %.2f M xs/s" % (ms, 10000 / ms / 1e3), V.timing(), V.stats(), flush=True)
t = time.time(); ra, rc = ref.area_batch(vol, qp[:400], qn[:400], w); dt = time.time() - t
print("reference, 1 core, 400 queries: %.0f xs/s; bit-identical: %s" % (400 / dt, np.array_equal(ra.view(np.uint32), a[:400].cpu().numpy().view(np.uint32)) and np.array_equal(rc, c[:400].cpu().numpy())), flush=True)
V.close()
import psutil
if psutil.virtual_memory().available > 40 * 2 ** 30:
  t = time.time(); lab = syn.make_labels(vol, np.uint64, hashed=True); del vol; print("labels 8 GiB built in %.1fs" % (time.time() - t), flush=True)
  qp2, qn2 = syn.draw_queries(verts, tans, 10000, seed=2)
  t = time.time(); V = xs3d.Volume(labels=lab); print("label upload: %.2fs" % (time.time() - t), flush=True)
  for crop in (100, 1000):
    n = 10000 if crop == 100 else 1000
    V.slice_batch(qp2[:n], qn2[:n], w, crop=crop)
    t = time.time(); r = V.slice_batch(qp2[:n], qn2[:n], w, crop=crop); dt = time.time() - t
    print("config 4: slice_batch crop=%d: %d planes in %.4fs -> %.0f planes/s, shape[0]=%s" % (crop, n, dt, n / dt, r[0].shape), flush=True)
  # the same 8 GiB label volume as the source of per-object masks: mask of one label built on the device
  for L in (int(lab[tuple(np.round(qp2[0]).astype(int))]), 12345):
    V.select_label(L)
    best = 1e9
    for i in range(3):
      flush.fill_(i); torch.cuda.synchronize()
      V.select_label(L); best = min(best, V.timing()["ingest_ms"])
    print("select_label on the 1024^3 uint64 volume: %.3f ms (%.0f GB/s)" % (best, (lab.nbytes + lab.size / 8) / best / 1e6), flush=True)
  t = time.time(); k = 200
  for i in range(k): ref.projection(lab, qp2[i], qn2[i], w, True, 100.0)
  print("reference slice crop=100: %.0f planes/s (1 core)" % (k / (time.time() - t)))
  V.close()
else:
  print("config 4 skipped: not enough host memory")
