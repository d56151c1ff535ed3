"""Dev script: xs3d_volume_select_label (label volume -> occupancy bytes of one label) timings, 512^3, every item size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
import numpy as np, torch
import xs3d
from xs3d import synthetic as syn
vol, verts, tans = syn.make_neurite(512)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dt in (np.uint8, np.uint16, np.uint32, np.uint64):
  lab = syn.make_labels(vol, dt)
  val = int(lab.max())
  for name, arr in (("F", lab), ("C", np.ascontiguousarray(lab))):
    with xs3d.Volume(labels=arr) as V:
      best = 1e9
      for i in range(3):
        flush.fill_(i); torch.cuda.synchronize()
        V.select_label(val)
        best = min(best, V.timing()["ingest_ms"])
      nbytes = lab.nbytes + lab.size // 8
      print("select_label %-6s %s order: %.4f ms  (%.0f GB/s of %d MB)" % (np.dtype(dt).name, name, best, nbytes / best / 1e6, nbytes >> 20))
