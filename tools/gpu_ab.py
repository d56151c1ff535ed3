"""Dev script: A/B of library builds on the bench sweep.  usage: gpu_ab.py [reps] lib_a.so lib_b.so ...
Each build runs in its own process (XS3D_LIB); prints the step time, the per-kernel times and a hash of the
results (equal hashes = bit-identical areas and contacts)."""
import hashlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
  sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
  import numpy as np, torch
  import xs3d
  from xs3d import synthetic as syn
  N = int(os.environ.get("XS_N", "512"))
  vol, verts, tans = syn.make_neurite(N)
  pos, nrm = syn.draw_queries(verts, tans, 10000)
  torch.cuda.set_stream(torch.cuda.Stream())
  V = xs3d.Volume(vol)
  V.set_stream(torch.cuda.current_stream().cuda_stream)
  dpos = torch.from_numpy(pos).cuda(); dnrm = torch.from_numpy(nrm).cuda()
  flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
  ts = []
  for rep in range(int(os.environ.get("XS_REPS", "24"))):
    flush.fill_(rep)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); a, c = V.cross_sectional_area_batch(dpos, dnrm, [32, 32, 40], return_contact=True); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
  t = V.timing()
  h = hashlib.sha1(a.cpu().numpy().tobytes() + c.cpu().numpy().tobytes()).hexdigest()[:12]
  print("%-34s step %.3f ms (min %.3f) | warp %.3f mid %.3f exposed %.3f pre %.3f poly %.3f comb %.3f p2 %.3f | %s" % (
      os.environ.get("TAG", ""), float(np.median(ts[4:])), min(ts), t["warp_tier_ms"], t["mid_tier_ms"], t["mid_exposed_ms"], t["prepass_ms"],
      t["polygon_ms"], t["combine_ms"], t.get("p2_ms", 0.0), h), flush=True)
  sys.exit(0)
args = sys.argv[1:]
rounds = int(args.pop(0)) if args and args[0].isdigit() else 2
for r in range(rounds):
  for lib in args:
    env = dict(os.environ, XS3D_LIB=os.path.abspath(lib), TAG=os.path.basename(lib))
    subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
