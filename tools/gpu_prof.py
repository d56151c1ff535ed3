"""Dev script: batch timing + per-phase cycle profile."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
import numpy as np, torch
import xs3d
from xs3d import synthetic as syn
N = int(os.environ.get("XS_N", "512")); NQ = int(os.environ.get("XS_NQ", "10000"))
vol, verts, tans = syn.make_neurite(N)
pos, nrm = syn.draw_queries(verts, tans, NQ)
w = [32, 32, 40]
V = xs3d.Volume(vol)
dpos = torch.from_numpy(pos).cuda(); dnrm = torch.from_numpy(nrm).cuda()
for rep in range(4):
  torch.cuda.synchronize(); t = time.time(); da, dc = V.cross_sectional_area_batch(dpos, dnrm, w, return_contact=True); torch.cuda.synchronize(); dt = time.time() - t
  print("batch device %.5fs  %.0f xs/s" % (dt, NQ / dt))
print(V.stats())
p = V.profile()
for tier in p:
  tot = sum(p[tier][k] for k in ("flood", "claims", "emit"))
  print(tier, {k: "%.1f%%" % (100.0 * v / max(tot, 1)) for k, v in p[tier].items()}, "total Mcycles %.1f" % (tot / 1e6))
st = V.stats()
nq0 = NQ - st["escalated_mid"]
print("warp tier: mean cycles/query %.0f ; cta tier: mean cycles/query %.0f" % (sum(list(p["warp_tier"].values())[:3]) / max(nq0, 1), sum(list(p["cta_tier"].values())[:3]) / max(st["escalated_mid"], 1)))
print(V.timing())
if os.environ.get("XS_MID_ALONE"):
  # the CTA tiers alone: only the queries the pre-pass sends to the mid tier (sizes from a first full batch)
  a_all = da.cpu().numpy()
  big = np.argsort(-a_all)[:st["escalated_mid"]]
  bp = torch.from_numpy(pos[big]).cuda(); bn = torch.from_numpy(nrm[big]).cuda()
  for rep in range(3):
    torch.cuda.synchronize(); t = time.time(); V.cross_sectional_area_batch(bp, bn, w); torch.cuda.synchronize(); print("largest %d queries alone: %.5fs" % (len(big), time.time() - t))
  print(V.stats()); p2 = V.profile(); print(p2); print(V.timing())
t = time.time(); V.load(vol); print("upload+ingest %.4fs" % (time.time() - t))
dv = torch.from_numpy(vol.view(np.uint8).ravel(order="K")).cuda()
torch.cuda.synchronize()
for rep in range(3):
  t = time.time(); V._lib.xs3d_volume_load(V._h, dv.data_ptr(), 1, 0); torch.cuda.synchronize(); print("ingest from device %.6fs  %.0f GB/s" % (time.time() - t, vol.size * 1.125 / (time.time() - t) / 1e9))
