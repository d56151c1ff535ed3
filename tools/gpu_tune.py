"""Dev script: sweep of the scheduling knobs (mid-tier CTA count, classification threshold) on the bench workload."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
  sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
  import numpy as np, torch
  import xs3d
  from xs3d import synthetic as syn
  vol, verts, tans = syn.make_neurite(512)
  pos, nrm = syn.draw_queries(verts, tans, 10000)
  torch.cuda.set_stream(torch.cuda.Stream())
  V = xs3d.Volume(vol)
  V.set_stream(torch.cuda.current_stream().cuda_stream)
  dpos = torch.from_numpy(pos).cuda(); dnrm = torch.from_numpy(nrm).cuda()
  flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
  ts = []
  for rep in range(8):
    flush.fill_(rep)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); V.cross_sectional_area_batch(dpos, dnrm, [32, 32, 40], return_contact=True); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
  t = V.timing(); st = V.stats()
  print("%-28s step %.3f ms (min %.3f) | warp %.3f mid %.3f exposed %.3f pre %.3f poly %.3f comb %.3f | to_mid %d mid->big %d" % (
      os.environ.get("TAG", ""), float(np.median(ts[3:])), min(ts), t["warp_tier_ms"], t["mid_tier_ms"], t["mid_exposed_ms"], t["prepass_ms"],
      t["polygon_ms"], t["combine_ms"], st["escalated_mid"], st["escalated_big"]), flush=True)
  sys.exit(0)
for key in [int(k) for k in os.environ.get("XS_KEYS", "18,22,26,30,34,40,48").split(",")]:
  env = dict(os.environ, XS3D_MID_KEY=str(key), TAG="key=%d" % key)
  subprocess.run([sys.executable, __file__, "child"], env=env)
