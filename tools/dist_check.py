"""Run under torchrun on N GPUs: NCCL broadcast of the occupancy bytes + query sharding + all-gather, checked
against the single-GPU result and the oracle.  usage: torchrun --nproc-per-node N tools/dist_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.distributed as dist
import xs3d
from xs3d import synthetic as syn, dist as xdist
from refapi import Oracle
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
vol, verts, tans = syn.make_neurite(256, seed=0)
pos, nrm = syn.draw_queries(verts, tans, 2001, seed=5)
w = [32, 32, 40]
V = xs3d.Volume(shape=vol.shape, device=local)
xdist.broadcast_volume(V, vol if rank == 0 else None, src=0)        # only rank 0 passes the image
compute = lambda p, n: V.cross_sectional_area_batch(p, n, w, return_contact=True)
a, c, (lo, hi) = xdist.sharded_area_batch(compute, pos, nrm)
oa, oc = Oracle().area_batch(vol, pos, nrm, w)
ok = np.array_equal(a.view(np.uint32), oa.view(np.uint32)) and np.array_equal(c, oc)
print(f"rank {rank}/{world}: shard [{lo},{hi}) gathered {len(a)} results, bit-identical to the oracle: {ok}", flush=True)
# slices (BASELINE config 5): rank 0 uploads the label volume, it travels by NCCL broadcast, the planes are sharded;
# every rank checks every cut-out of its share against the oracle
lab = syn.make_labels(vol, np.uint64, hashed=True)
xdist.broadcast_labels(V, lab if rank == 0 else None, src=0)
spos, snrm = syn.draw_queries(verts, tans, 203, seed=6)
mine, (slo, shi) = xdist.sharded_slice_batch(lambda p, n: V.slice_batch(p, n, w, crop=150.0), spos, snrm)
orc = Oracle()
ok_s = all(np.array_equal(np.asfortranarray(mine[i]), orc.projection(lab, spos[slo + i], snrm[slo + i], w, True, 150.0)) for i in range(shi - slo))
big = V.slice_batch(spos[slo:slo + 2], snrm[slo:slo + 2], w)            # two un-cropped planes per rank
ok_s = ok_s and all(np.array_equal(np.asfortranarray(big[i]), orc.projection(lab, spos[slo + i], snrm[slo + i], w, True, float("inf"))) for i in range(2))
shapes = xdist.gather_slice_shapes(mine.shapes, len(spos))
ok_s = ok_s and shapes.shape == (203, 2) and np.array_equal(shapes[slo:shi], mine.shapes)
print(f"rank {rank}/{world}: slice shard [{slo},{shi}) of 203 planes + 2 un-cropped, bit-identical to the oracle: {ok_s}", flush=True)
ok = ok and ok_s
t = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
