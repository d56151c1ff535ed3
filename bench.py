#!/usr/bin/env python
"""bench.py — cross-sections/sec for a batched skeleton sweep on a 512^3 synthetic neurite (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm (default).  One step = one pass of the hot path over one batch: QUERIES (10 000) (vertex, normal)
queries of the skeleton sweep against the 512^3 neurite volume.
  value : whole-job throughput with inputs resident in HBM (volume ingested, queries and results on the
          device), timed with CUDA events on the launching stream, L2 flushed between steps, max over ranks.
  e2e   : the same sweep through the C ABI with HOST buffers: every step takes the 128 MiB boolean volume and the
          queries from pinned host memory (xs3d_pack_bits packs the volume on the host cores, 16 MiB cross PCIe;
          the packing + upload of step k+1 run beside the sweep of step k), ingests, sweeps, and reads areas +
          contact bits back.  N > 1: ONE host volume — rank 0 packs / uploads / ingests, the occupancy bytes are
          NCCL-broadcast every step, every rank sweeps its own queries.  Timed over >= 100 steps (pipeline fill
          amortised); every sweep is preceded, on its stream and inside the timed region, by an L2-evicting write
          (enqueued behind the previous sweep through xs3d_volume_set_after_enqueue).
  clocks : SM clock / throttle reasons sampled by an in-process NVML thread while the timed regions run.
  roofline     : the dominant kernel (warp-tier query kernel), algorithmic bytes / CUDA-event duration vs the
                 measured HBM peak (MEASURED_PEAKS.json).  The mid / big tiers run beside it on a second stream.
  cpu_baseline : the compiled reference (oracle/_ref, else the C port) fanned out over all host cores.
N > 1: one process per GPU (torchrun), queries partitioned with no data-path collective (weak scaling:
each rank sweeps its own 10 000-query draw); the volume's occupancy bytes are NCCL-broadcast once at setup.

Reference arm (--impl reference): the reference's own CPU implementation of the path (oracle/_ref when it
was compiled, else the port) on all host cores, same workload/metric; rank 0 only.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))

QUERIES = 10000
VOLUME_N = 512
ANISOTROPY = (32.0, 32.0, 40.0)
METRIC = "cross-sections/sec (batched skeleton vertices, 512^3 vol)"
WORKLOAD = "neurite512_sweep10k"
SUSTAINED_SWEEPS = 200
E2E_MIN_STEPS = 100          # steps of the pipelined end-to-end loop per timed region (>= --steps)
# the same `config` object in both arms (the driver compares them key by key)
CONFIG = {"workload": WORKLOAD, "volume": [VOLUME_N] * 3, "queries_per_step": QUERIES, "anisotropy": list(ANISOTROPY)}


def make_workload(rank=0):
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(VOLUME_N, seed=0)
  pos, nrm = synthetic.draw_queries(verts, tans, QUERIES, seed=1 + rank)
  return vol, pos, nrm


# ------------------------------------------------------------------------------------------------ CPU arm
_G = {}


def _cpu_chunk(args):
  lo, hi = args
  return _G["chk"].area_batch(_G["vol"], _G["pos"][lo:hi], _G["nrm"][lo:hi], ANISOTROPY, nthreads=1)


def cpu_sweep_rate(vol, pos, nrm, passes, warm=1):
  """The single-query reference fanned out over all host cores with fork()ed workers (the volume is
  inherited copy-on-write), one contiguous chunk per worker.  Returns
  (xs/s, cores, kind, seconds per pass list)."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import refapi
  if refapi.have_ref():
    chk, kind = refapi.Ref(), "reference"
  else:
    if not refapi.have_oracle():
      subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    chk, kind = refapi.Oracle(), "port"
  cores = os.cpu_count() or 1
  _G.update(chk=chk, vol=vol, pos=pos, nrm=nrm)
  n = len(pos)
  chunks = [(i * n // cores, (i + 1) * n // cores) for i in range(cores)]
  times = []
  with mp.get_context("fork").Pool(cores) as pool:
    for _ in range(warm):
      pool.map(_cpu_chunk, chunks)
    for _ in range(passes):
      t = time.perf_counter()
      pool.map(_cpu_chunk, chunks)
      times.append(time.perf_counter() - t)
  # one core, for scale: the first 1000 queries in this process
  t = time.perf_counter()
  chk.area_batch(vol, pos[:1000], nrm[:1000], ANISOTROPY, nthreads=1)
  _G["one_core_rate"] = min(1000, n) / (time.perf_counter() - t)
  return n / (sum(times) / len(times)), cores, kind, times


def run_reference_arm(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  vol, pos, nrm = make_workload(0)
  rate, cores, kind, times = cpu_sweep_rate(vol, pos, nrm, passes=max(1, args.steps), warm=max(1, args.warmup))
  ms = 1000.0 * sum(times) / len(times)
  line = {
    "impl": "reference", "metric": METRIC, "value": rate, "unit": "cross-sections/s", "n_gpus": args.gpus,
    "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": dict(CONFIG),
    "cpu_baseline": {"value": rate, "unit": "cross-sections/s", "cores": cores, "kind": kind, "one_core": _G.get("one_core_rate"),
                     "sample": f"{len(times)} passes of the full {QUERIES}-query sweep, fork()ed workers, one chunk per core"},
    "e2e": {"value": rate, "unit": "cross-sections/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "gpu_launches": 0,
  }
  print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
  """SM clock and throttle reasons sampled while the timed regions run.  In-process NVML (one nvmlInit, then three cheap
  queries every 50 ms): a looping `nvidia-smi -lms 100` process took seconds per sample on the bench boxes and, hammering
  the driver the whole time, slowed the host-bound end-to-end loop by 25 % (1.08 against 0.86 ms per step; it is the
  fallback when the NVML binding is missing)."""
  Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
  REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

  def __init__(self, index):
    self.index, self.rows, self.proc = index, [], None
    self.sm, self.mx, self.reasons, self.how = [], 0, set(), None
    self._stop = threading.Event()
    self._thread = None

  def start(self):
    try:
      import pynvml
      pynvml.nvmlInit()
      h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
      self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
      get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

      def poll():
        while not self._stop.is_set():
          try:
            self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            mask = int(get_reasons(h))
            for name, bit in self.REASONS:
              if mask & bit:
                self.reasons.add(name)
          except Exception:
            pass
          self._stop.wait(0.05)
      self.how = "nvml"
      self._thread = threading.Thread(target=poll, daemon=True)
      self._thread.start()
      return
    except Exception:
      self.how = None
    try:
      self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.how = "nvidia-smi"
      threading.Thread(target=self._read, daemon=True).start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(",")])

  def stop(self):
    self._stop.set()
    if self._thread is not None:
      self._thread.join(1.0)
    if self.proc:
      self.proc.terminate()
    sm, mx, reasons = list(self.sm), self.mx, set(self.reasons)
    for r in self.rows:
      try:
        sm.append(float(r[1])); mx = max(mx, float(r[2]))
        for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
          if val.lower().startswith("active"):
            reasons.add(name)
      except Exception:
        pass
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm), "source": self.how}


def slice_record(args, V0, world, rank, local, dev, barrier):
  """BASELINE config 5 beside the headline: xs3d.slice of a 1024^3 uint64 label volume, crop = 100 nm, 10 000 planes
  sharded over the ranks.  The 8 GiB volume is uploaded by rank 0 once and NCCL-broadcast (outside the timed region: it is
  the resident data set, like the ingested volume of `value`); every step each rank plans + gathers its planes —
  device-resident queries and output for `value`, host queries in / ragged host arrays out for `e2e`."""
  import torch
  import torch.distributed as dist
  import psutil
  import xs3d
  from xs3d import synthetic, dist as xdist
  ok_mem = psutil.virtual_memory().available > 40 * 2 ** 30 and torch.cuda.mem_get_info(dev)[0] > 24 * 2 ** 30
  if world > 1:
    t = torch.tensor([int(ok_mem)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_mem = bool(int(t.item()))
  if not ok_mem:
    return {"skipped": "needs ~20 GiB of free host memory on rank 0 and 24 GiB of HBM"}
  N, NP, CROP = 1024, 10000, 100.0
  if rank == 0:
    vol, verts, tans = synthetic.make_neurite(N, seed=0)
    lab = synthetic.make_labels(vol, np.uint64, hashed=True)
    del vol
    qp, qn = synthetic.draw_queries(verts, tans, NP, seed=2)
  else:
    lab, qp, qn = None, np.zeros((NP, 3), np.float32), np.zeros((NP, 3), np.float32)
  V = xs3d.Volume(shape=(N, N, N), device=local)
  V.set_stream(torch.cuda.current_stream().cuda_stream)
  t0 = time.perf_counter()
  if world > 1:
    xdist.broadcast_labels(V, lab, src=0)
    tq = torch.from_numpy(np.stack([qp, qn])).to(dev)
    dist.broadcast(tq, src=0)
    qp, qn = tq[0].cpu().numpy(), tq[1].cpu().numpy()
  else:
    V.load_labels(lab)
  t_setup = time.perf_counter() - t0
  lo, hi = xdist.shard_bounds(NP, world, rank)
  dp, dn = torch.from_numpy(qp[lo:hi]).to(dev), torch.from_numpy(qn[lo:hi]).to(dev)
  shapes, offsets = V.slice_plan(dp, dn, ANISOTROPY, True, CROP)
  out = torch.empty(int(offsets[-1]), dtype=torch.int64, device=dev)

  def step_dev():
    V.slice_plan(dp, dn, ANISOTROPY, True, CROP)
    V.slice_fill(0, hi - lo, out)

  def step_host():
    return V.slice_batch(qp[lo:hi], qn[lo:hi], ANISOTROPY, crop=CROP)

  res = {}
  for name, fn in (("value", step_dev), ("e2e", step_host)):
    for _ in range(3):
      fn()
    barrier()
    tot = 0.0
    for i in range(args.steps):
      s_ = torch.cuda.Event(enable_timing=True); e_ = torch.cuda.Event(enable_timing=True)
      s_.record(); r = fn(); e_.record()
      torch.cuda.synchronize()
      tot += s_.elapsed_time(e_)
    barrier()
    if world > 1:
      tt = torch.tensor([tot], dtype=torch.float64, device=dev)
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      tot = float(tt.item())
    res[name] = {"value": NP * args.steps / (tot / 1000.0), "unit": "planes/s", "ms_per_step": tot / args.steps}
  # every rank: a sample of its own cut-outs against the CPU checker (rank 0 has the labels; the others check theirs
  # against rank 0's re-computation of the same planes, gathered as hashes)
  r = step_host()
  sel = np.linspace(0, hi - lo - 1, 25).astype(np.int64)
  digest = np.array([[int(np.asarray(r[int(i)]).view(np.uint64).sum() & np.uint64(0x7fffffffffffffff)), r[int(i)].shape[0], r[int(i)].shape[1]] for i in sel], dtype=np.int64)
  ok = True
  if world > 1:
    allsel = [torch.zeros((25, 4), dtype=torch.int64, device=dev) for _ in range(world)]
    mine = torch.from_numpy(np.concatenate([(sel + lo)[:, None], digest], axis=1)).to(dev)
    dist.all_gather(allsel, mine)
    gathered = torch.cat(allsel).cpu().numpy()
  else:
    gathered = np.concatenate([(sel + lo)[:, None], digest], axis=1)
  if rank == 0:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refapi
    chk = refapi.Oracle()
    for row in gathered:
      o = chk.projection(lab, qp[row[0]], qn[row[0]], ANISOTROPY, True, CROP)
      ok = ok and tuple(row[2:]) == o.shape and int(o.view(np.uint64).sum() & np.uint64(0x7fffffffffffffff)) == int(row[1])
  V.close()
  res.update({"workload": "xs3d.slice, 1024^3 uint64 labels, crop = 100 nm, 10 000 planes sharded over the ranks (BASELINE config 5)",
              "planes_per_rank": hi - lo, "output_elements_per_step_rank0": int(offsets[-1]), "label_upload_and_broadcast_s": t_setup,
              "checked_planes": int(len(gathered)), "bit_exact_vs_cpu_checker": bool(ok),
              "collective": "NCCL broadcast of the 8 GiB label volume once, at setup; none in the timed region (planes are consumed where they are made)"})
  return res


def run_own_arm(args):
  import torch
  import torch.distributed as dist
  import xs3d
  from xs3d import dist as xdist

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
  if rank == 0 and not os.environ.get("XS_BENCH_TWO_HANDLES"):
    # the e2e loop packs the (single) host volume on rank 0: all host cores but one per rank (each rank's main thread
    # spins on its GPU); the packing thread itself is one of the pool's workers
    # (measured: 16-core box, N = 1: 14-15 threads best; 32-core box, N = 8: 16 threads 1.15-1.22 ms per step, 24 threads 1.37 ms —
    # the other ranks' main threads and the broadcasters need cores too)
    os.environ.setdefault("XS3D_HOST_THREADS", str(max(4, min(16, (os.cpu_count() or 8) - world - 1))))
  if world > 1 and rank != 0 and os.environ.get("XS_BENCH_POLL_RANKS"):
    # dev knob: the ranks that do not pack poll their GPU with short sleeps instead of spinning (their cores go to rank 0's pool)
    os.environ.setdefault("XS3D_POLL_SLEEP_US", os.environ["XS_BENCH_POLL_RANKS"])
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  # everything below runs on a stream of its own, made torch's current stream: the library launches there
  # (xs3d_volume_set_stream), torch's CUDA events bracket its work, and — unlike the legacy default stream — the stream can
  # be captured, so the library replays each batch's launch sequence as a CUDA graph
  bench_stream = torch.cuda.Stream(device=dev)
  torch.cuda.set_stream(bench_stream)

  vol, pos, nrm = make_workload(rank)
  w = np.asarray(ANISOTROPY, np.float32)
  V = xs3d.Volume(shape=vol.shape, device=local)
  V.set_stream(torch.cuda.current_stream().cuda_stream)
  if world > 1:
    xdist.broadcast_volume(V, vol if rank == 0 else None, src=0)   # one NCCL broadcast of the 16 MiB occupancy bytes
  else:
    V.load(vol)
  dpos = torch.from_numpy(pos).to(dev)
  dnrm = torch.from_numpy(nrm).to(dev)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  # ---- resident-in-HBM timing (`value`): the C ABI with DEVICE pointers, outputs preallocated (no per-step Python /
  #      allocator work inside the timed region)
  import ctypes as C
  from xs3d import _abi
  L = _abi.lib()
  darea = torch.empty(QUERIES, dtype=torch.float32, device=dev)
  dct = torch.empty(QUERIES, dtype=torch.uint8, device=dev)
  wp = _abi.fptr(w)

  def step_resident():
    _abi.check(L.xs3d_area_batch(V._h, QUERIES, dpos.data_ptr(), dnrm.data_ptr(), wp, _abi.DEVICE, darea.data_ptr(), dct.data_ptr()))
    return darea, dct

  for _ in range(max(3, args.warmup)):
    step_resident()
  sampler = ClockSampler(local)
  if rank == 0 and not os.environ.get("XS_BENCH_NO_SMI"):
    sampler.start()
  barrier()
  starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
  ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
  kern = {"prepass_ms": 0.0, "warp_tier_ms": 0.0, "polygon_ms": 0.0, "combine_ms": 0.0, "mid_exposed_ms": 0.0,
          "mid_tier_ms": 0.0, "mid_second_launch_ms": 0.0, "p2_ms": 0.0, "second_stream_poly_combine_ms": 0.0, "second_round_ms": 0.0, "worst_case_ms": 0.0}
  launches = 0
  stats = None
  for i in range(args.steps):
    flush.fill_(i & 255)                 # evict L2 between timed iterations (untimed)
    starts[i].record()
    step_resident()
    ends[i].record()
    torch.cuda.synchronize()
    t = V.timing()
    for k in kern:
      kern[k] += t[k]
    stats = V.stats()
    launches += stats["launches"]
  barrier()
  ms_total = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
  if world > 1:
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
  value = world * QUERIES * args.steps / (ms_total / 1000.0)

  # ---- the same resident step over a longer region (>= ~150 ms of sweeps, so that the clock sampler sees the load):
  #      SUSTAINED_SWEEPS back-to-back flush + sweep pairs, only the sweeps between event pairs are summed
  n_sus = max(args.steps, SUSTAINED_SWEEPS)
  sus_s = [torch.cuda.Event(enable_timing=True) for _ in range(n_sus)]
  sus_e = [torch.cuda.Event(enable_timing=True) for _ in range(n_sus)]
  barrier()
  for i in range(n_sus):
    flush.fill_(i & 255)
    sus_s[i].record()
    step_resident()
    sus_e[i].record()
  barrier()
  sus_ms = sorted(s.elapsed_time(e) for s, e in zip(sus_s, sus_e))
  sus_total = sum(sus_ms)
  if world > 1:
    tt = torch.tensor([sus_total], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    sus_total = float(tt.item())
  sustained = {"sweeps": n_sus, "value": world * QUERIES * n_sus / (sus_total / 1000.0), "ms_per_step": sus_total / n_sus,
               "ms_median_rank0": sus_ms[len(sus_ms) // 2], "ms_min_rank0": sus_ms[0], "ms_max_rank0": sus_ms[-1],
               "what": "the `value` step repeated over a longer timed region (L2 flushed before every sweep, CUDA events around each sweep, max over ranks of the per-rank totals)"}

  # ---- strong scaling (BASELINE config 3/4 wording: ONE 10k-query list sharded over the ranks): rank r sweeps its
  #      contiguous slice of rank 0's draw with device-resident queries, then areas + contacts are all-gathered (NCCL)
  #      inside the timed region, so every rank ends up with all 10 000 answers
  pos0, nrm0 = (pos, nrm) if rank == 0 else make_workload(0)[1:]
  lo, hi = xdist.shard_bounds(QUERIES, world, rank)
  kmax = xdist.shard_bounds(QUERIES, world, 0)[1]
  spos = torch.from_numpy(np.ascontiguousarray(pos0[lo:hi])).to(dev)
  snrm = torch.from_numpy(np.ascontiguousarray(nrm0[lo:hi])).to(dev)
  sarea = torch.zeros(kmax, dtype=torch.float32, device=dev)
  sct = torch.zeros(kmax, dtype=torch.uint8, device=dev)
  garea = torch.empty(world * kmax, dtype=torch.float32, device=dev)
  gct = torch.empty(world * kmax, dtype=torch.uint8, device=dev)

  def step_strong():
    _abi.check(L.xs3d_area_batch(V._h, hi - lo, spos.data_ptr(), snrm.data_ptr(), wp, _abi.DEVICE, sarea.data_ptr(), sct.data_ptr()))
    if world > 1:
      dist.all_gather_into_tensor(garea, sarea)
      dist.all_gather_into_tensor(gct, sct)

  for _ in range(max(3, args.warmup)):
    step_strong()
  barrier()
  st_tot = 0.0
  for i in range(args.steps):
    flush.fill_(i & 255)
    s_ = torch.cuda.Event(enable_timing=True); e_ = torch.cuda.Event(enable_timing=True)
    s_.record(); step_strong(); e_.record()
    torch.cuda.synchronize()
    st_tot += s_.elapsed_time(e_)
  barrier()
  strong_kernel_ms = V.timing()
  if world > 1:
    tt = torch.tensor([st_tot], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    st_tot = float(tt.item())
  strong = {"queries_total": QUERIES, "queries_per_rank": hi - lo, "value": QUERIES * args.steps / (st_tot / 1000.0), "ms_per_step": st_tot / args.steps,
            "collective": "all-gather of areas (4 B) + contacts (1 B) per query inside the timed region" if world > 1 else "none (one rank)",
            "rank0_kernels_ms_last_step": {k: strong_kernel_ms[k] for k in ("prepass_ms", "warp_tier_ms", "polygon_ms", "combine_ms", "mid_tier_ms", "mid_exposed_ms")},
            "limiter": "the longest single section: the mid tier walks it on one SM (latency chain) while the other SMs run out of queries; plus the fixed pre-pass / polygon / combine launches"}
  # every rank: a 200-query sample of what ITS GPU computed (weak draw and strong shard) against the CPU checker,
  # outside the timed regions; the gathered strong result against rank 0's own full sweep
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import refapi
  chk = refapi.Ref() if refapi.have_ref() else refapi.Oracle()
  da_, dc_ = step_resident()
  torch.cuda.synchronize()
  ga, gc = da_.cpu().numpy(), dc_.cpu().numpy()
  sel = np.linspace(0, QUERIES - 1, 200).astype(np.int64)
  oa, oc = chk.area_batch(vol, pos[sel], nrm[sel], ANISOTROPY)
  ok_weak = bool(np.array_equal(ga[sel].view(np.uint32), oa.view(np.uint32)) and np.array_equal(gc[sel], oc))
  step_strong()
  torch.cuda.synchronize()
  ssel = np.linspace(0, hi - lo - 1, min(200, hi - lo)).astype(np.int64)
  oa, oc = chk.area_batch(vol, pos0[lo:hi][ssel], nrm0[lo:hi][ssel], ANISOTROPY)
  ok_strong = bool(np.array_equal(sarea.cpu().numpy()[ssel].view(np.uint32), oa.view(np.uint32)) and np.array_equal(sct.cpu().numpy()[ssel], oc))
  if world > 1:
    full_a = torch.cat([garea[r * kmax: r * kmax + (xdist.shard_bounds(QUERIES, world, r)[1] - xdist.shard_bounds(QUERIES, world, r)[0])] for r in range(world)])
    full_c = torch.cat([gct[r * kmax: r * kmax + (xdist.shard_bounds(QUERIES, world, r)[1] - xdist.shard_bounds(QUERIES, world, r)[0])] for r in range(world)])
    ref_a = da_ if rank == 0 else torch.empty_like(da_)
    ref_c = dc_ if rank == 0 else torch.empty_like(dc_)
    dist.broadcast(ref_a, src=0); dist.broadcast(ref_c, src=0)
    ok_strong = ok_strong and bool(torch.equal(full_a.view(torch.int32), ref_a.view(torch.int32)) and torch.equal(full_c, ref_c))
  oks = torch.tensor([int(ok_weak), int(ok_strong)], dtype=torch.int32, device=dev)
  if world > 1:
    dist.all_reduce(oks, op=dist.ReduceOp.MIN)
  rank_checks = {"checker": "reference" if refapi.have_ref() else "port", "queries_per_rank_checked": int(len(sel) + len(ssel)),
                 "all_ranks_bit_exact": bool(int(oks[0]) == 1 and int(oks[1]) == 1),
                 "what": "every rank compares 200 queries of its own weak-scaling sweep and 200 of its strong-scaling shard with the CPU checker (float32 bit patterns + contact bytes); the all-gathered strong result equals rank 0's single-GPU sweep of the same list"}
  assert rank_checks["all_ranks_bit_exact"], "a rank's results differ from the CPU checker"

  # ---- end to end through the C ABI with host buffers (`e2e`): volume + queries up, results down, every step
  hvol = torch.from_numpy(np.ascontiguousarray(vol.view(np.uint8).ravel(order="F"))).pin_memory()
  hpos = torch.from_numpy(pos.copy()).pin_memory()
  hnrm = torch.from_numpy(nrm.copy()).pin_memory()
  harea = torch.empty(QUERIES, dtype=torch.float32).pin_memory()
  hct = torch.empty(QUERIES, dtype=torch.uint8).pin_memory()

  def step_e2e():
    _abi.check(L.xs3d_area_sweep_host(V._h, hvol.data_ptr(), 0, QUERIES, hpos.data_ptr(), hnrm.data_ptr(), _abi.fptr(w), harea.data_ptr(), hct.data_ptr()))

  def step_e2e_resident():
    _abi.check(L.xs3d_area_batch(V._h, QUERIES, hpos.data_ptr(), hnrm.data_ptr(), _abi.fptr(w), _abi.HOST, harea.data_ptr(), hct.data_ptr()))

  def time_steps(fn):
    for _ in range(max(3, args.warmup)):
      fn()
    barrier()
    tot = 0.0
    for i in range(args.steps):
      flush.fill_(i & 255)
      s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
      s.record(); fn(); e.record()
      torch.cuda.synchronize()
      tot += s.elapsed_time(e)
    barrier()
    if world > 1:
      tt = torch.tensor([tot], dtype=torch.float64, device=dev)
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      tot = float(tt.item())
    return tot

  ms_e2e = time_steps(step_e2e)
  ingest_e2e_ms = V.timing()["ingest_ms"]            # the ingest kernel of the e2e path (bits -> occupancy bytes when the upload is packed)
  dvol = hvol.to(dev)
  ingest_ms = 1e9
  for i in range(5):                                 # the plain uint8 -> occupancy-byte kernel on a device-resident image (HBM-bound)
    flush.fill_(i)
    _abi.check(L.xs3d_volume_load(V._h, dvol.data_ptr(), _abi.DEVICE, 0))
    ingest_ms = min(ingest_ms, V.timing()["ingest_ms"])
  del dvol
  ms_e2e_res = time_steps(step_e2e_resident)

  import threading as _th
  # the pipelined loop is timed over at least E2E_MIN_STEPS steps: its first step has no sweep to hide the packing of
  # image 0 behind (one extra packing time per timed region: +5 % over 20 steps, +1 % over 100)
  psteps = max(E2E_MIN_STEPS, args.steps)
  staged_mode = world > 1 or not os.environ.get("XS_BENCH_TWO_HANDLES")
  if not staged_mode:
    # ---- the same end-to-end step, pipelined across sweeps: two resident-volume handles on their own streams,
    #      driven by two host threads, so the upload of sweep k+1 overlaps the kernels of sweep k.  Every step
    #      still takes its own volume + queries from pinned host memory and reads its results back.
    V2 = [xs3d.Volume(shape=vol.shape, device=local) for _ in range(2)]
    outs = [(torch.empty(QUERIES, dtype=torch.float32).pin_memory(), torch.empty(QUERIES, dtype=torch.uint8).pin_memory()) for _ in range(2)]
    pstreams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    for k in range(2):
      V2[k].set_stream(pstreams[k].cuda_stream)

    def sweep(slot):
      torch.cuda.set_device(local)
      with torch.cuda.stream(pstreams[slot]):
        flush.fill_(slot)                    # L2 eviction before every sweep, on the sweep's own stream (inside the timed region)
      _abi.check(L.xs3d_area_sweep_host(V2[slot]._h, hvol.data_ptr(), 0, QUERIES, hpos.data_ptr(), hnrm.data_ptr(), _abi.fptr(w),
                                        outs[slot][0].data_ptr(), outs[slot][1].data_ptr()))

    def run_e2e(nsteps):
      def worker(slot):
        for _ in range(slot, nsteps, 2):
          sweep(slot)
      ts = [_th.Thread(target=worker, args=(k,)) for k in range(2)]
      for t in ts:
        t.start()
      for t in ts:
        t.join()
    e2e_mode = "two sweeps in flight (two Volume handles / streams / host threads): upload of sweep k+1 overlaps the kernels of sweep k"
  else:
    # ---- N ranks, ONE volume in host memory: every step, rank 0 bit-packs it on its host cores and starts
    #      its upload (a helper thread does both for the image of step k+1, into the other pinned / staging buffer, while
    #      step k runs) and ingests it; for N > 1 the occupancy bytes are NCCL-broadcast to every rank; every rank sweeps
    #      its own 10 000 queries from pinned host buffers and reads its results back.  The other ranks' host cores stay
    #      idle: the box has one memory system, which the packing saturates.
    feed = xdist.VolumeFeed(V)               # two pinned bit buffers + the volume's staging buffers, reused by every call
    outs = None

    def run_e2e(nsteps):
      # L2 eviction before every sweep, on the sweep's stream, inside the timed region: the write for step 0 is enqueued by
      # before_step, the write for step k+1 by the library's after-enqueue hook during step k's call — behind sweep k's
      # kernels and result copies — so that the GPU runs it while the host is on its way back with the results of step k
      # (the host waits for an event behind the result copies, not for the stream).  One write per sweep either way.
      done = [0]

      def evict_for_next_step():
        done[0] += 1
        if done[0] < nsteps:
          flush.fill_(done[0] & 255)
      hook = not os.environ.get("XS_BENCH_NO_HOOK")
      if hook:
        V.set_after_enqueue(evict_for_next_step)
      try:
        xdist.stream_volumes(V, [vol] * nsteps if rank == 0 else None, nsteps, lambda k: step_e2e_resident(),
                             before_step=(lambda k: flush.fill_(0) if k == 0 else None) if hook else (lambda k: flush.fill_(k & 255)),
                             feed=feed)
      finally:
        V.set_after_enqueue(None)
    e2e_mode = ("one volume in host memory, taken afresh every step: rank 0 bit-packs and uploads it (double-buffered: image k+1 is packed and copied beside the sweep of image k) and ingests it, "
                + ("the 16 MiB occupancy bytes are NCCL-broadcast, " if world > 1 else "") + "every rank sweeps its own 10 000 host-resident queries and reads its results back")

  run_e2e(max(4, args.warmup))
  barrier()
  s_ev = torch.cuda.Event(enable_timing=True); e_ev = torch.cuda.Event(enable_timing=True)
  s_ev.record()
  run_e2e(psteps)
  torch.cuda.synchronize()
  e_ev.record()
  torch.cuda.synchronize()
  ms_pipe = s_ev.elapsed_time(e_ev)
  barrier()
  if world > 1:
    tt = torch.tensor([ms_pipe], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_pipe = float(tt.item())
  if not staged_mode:
    assert np.array_equal(outs[0][0].numpy().view(np.uint32), harea.numpy().view(np.uint32)) and np.array_equal(outs[1][1].numpy(), hct.numpy())
    for vv in V2:
      vv.close()
  L.xs_pack_threads.restype = C.c_int
  pack_threads = int(L.xs_pack_threads())
  packed = pack_threads >= 6 and not os.environ.get("XS3D_NO_PACK")     # use_packed_upload(), xs_device.cu
  h2d_volume = (vol.size + 7) // 8 if (packed or staged_mode) else vol.size
  e2e_serial_value = world * QUERIES * args.steps / (ms_e2e / 1000.0)
  e2e_value = world * QUERIES * psteps / (ms_pipe / 1000.0)
  e2e_res_value = world * QUERIES * args.steps / (ms_e2e_res / 1000.0)
  clocks = sampler.stop() if rank == 0 else None

  # parity spot check inside the bench (results of the e2e path vs the resident path)
  da, dc = step_resident()
  assert np.array_equal(da.cpu().numpy().view(np.uint32), harea.numpy().view(np.uint32)) and np.array_equal(dc.cpu().numpy(), hct.numpy())

  slices = slice_record(args, V, world, rank, local, dev, barrier) if not os.environ.get("XS_BENCH_NO_SLICES") else {"skipped": "XS_BENCH_NO_SLICES"}

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  # ---- roofline of the dominant kernel
  peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(peaks_path):
    peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
  else:
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
  # dominant kernel of the launching stream's critical path (the mid / big tiers run beside it on a second stream;
  # what they add to the step is kernels_ms_per_step.mid_exposed_ms)
  dom = max(("warp_tier_ms", "polygon_ms"), key=lambda k: kern[k])
  dom_ms = kern[dom] / args.steps
  # algorithmic bytes (SURVEY.md §8d formula for the reference's uint8 volume): one byte per lattice
  # sample, 8 bytes per contributing 2x2x2 block, 24 B query in, 5 B result out — for the queries THIS kernel handled
  # (the CTA tiers report their share of the samples; blocks are apportioned by the same ratio).
  cta_frac = stats["samples_cta_tiers"] / max(stats["samples"], 1)
  q_cta = stats["escalated_mid"]
  if dom == "warp_tier_ms":
    frac, nq_k = 1.0 - cta_frac, QUERIES - q_cta
  elif dom in ("mid_tier_ms", "big_tier_ms"):
    frac, nq_k = cta_frac, q_cta
  else:
    frac, nq_k = 1.0, QUERIES
  if dom == "polygon_ms":
    alg_bytes = stats["polygons"] * (12 + 4)            # one 12-byte record in, one float out per polygon
  else:
    alg_bytes = int(frac * (stats["samples"] * 1 + 8 * stats["blocks"])) + 29 * nq_k
  achieved = alg_bytes / (dom_ms / 1000.0) / 1e9
  kname = {"warp_tier_ms": "area_warpn_kernel", "mid_tier_ms": "area_mid_kernel", "big_tier_ms": "area_block_kernel<BigTier>", "polygon_ms": "poly_eval_kernel"}[dom]
  # ncu figures (dram__bytes_read.sum + dram__bytes_write.sum per launch, IPC, lanes, warp instructions) come from
  # profiles/ncu_current.json, which tools/ncu_current.py writes from the latest `ncu --set full` capture of this
  # workload; absent file -> null (never a literal typed in here)
  ncu = {}
  ncu_path = os.path.join(ROOT, "profiles", "ncu_current.json")
  if os.path.exists(ncu_path):
    ncu = json.load(open(ncu_path))
  nk = ncu.get("kernels", {})
  def ncu_of(name, key):
    return nk.get(name, {}).get(key)
  roofline = {"bound": "hbm", "kernel": kname,
              "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_of(kname, "dram_bytes"),
              "traffic_source": ncu.get("capture"),
              "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms, "queries_in_kernel": nq_k,
              "note": "latency/issue-bound graph walk over the L2-resident 16 MiB occupancy volume, not an HBM stream; the HBM-bound kernel of the path is the ingest, reported beside it",
              "issue_rate": {"what": "the bound that actually applies to the dominant kernel, from the ncu capture named in traffic_source (not measured live)",
                             "ipc_per_sm": ncu_of(kname, "ipc"), "peak_ipc_per_sm": 4.0, "active_lanes": ncu_of(kname, "active_lanes"), "l2_hit_pct": ncu_of(kname, "l2_hit_pct"),
                             "warp_instructions_per_query": (ncu_of(kname, "warp_inst") / ncu["queries"]) if ncu_of(kname, "warp_inst") and ncu.get("queries") else None},
              "ingest_kernel": {"kernel": "ingest_f16_kernel (uint8 image in HBM -> occupancy bytes)", "ms": ingest_ms,
                                "achieved": (vol.size * 1.125) / (ingest_ms / 1000.0) / 1e9, "unit": "GB/s",
                                "frac": (vol.size * 1.125) / (ingest_ms / 1000.0) / 1e9 / peak, "algorithmic_bytes": int(vol.size * 1.125),
                                "traffic": ncu_of("ingest_f16_kernel", "dram_bytes")},
              "ingest_e2e_kernel": {"kernel": "ingest_bits_f32_kernel (bit-packed image -> occupancy bytes)" if packed else "ingest_f16_kernel",
                                    "ms": ingest_e2e_ms, "algorithmic_bytes": int(vol.size * (0.25 if packed else 1.125)),
                                    "traffic": ncu_of("ingest_bits_f32_kernel" if packed else "ingest_f16_kernel", "dram_bytes"),
                                    "achieved": vol.size * (0.25 if packed else 1.125) / (ingest_e2e_ms / 1000.0) / 1e9 if ingest_e2e_ms > 0 else None,
                                    "unit": "GB/s",
                                    "frac": vol.size * (0.25 if packed else 1.125) / (ingest_e2e_ms / 1000.0) / 1e9 / peak if ingest_e2e_ms > 0 else None}}

  # ---- CPU baseline on this box's host cores (bounded: a few passes of the same sweep)
  rate, cores, kind, times = cpu_sweep_rate(vol, pos, nrm, passes=10, warm=2)    # ~15-25 core-seconds; the reference arm runs --steps passes
  line = {
    "metric": METRIC, "value": value, "unit": "cross-sections/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
    "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
    "dtype": "f32", "data": "synthetic",
    "config": dict(CONFIG),
    "run": {"queries_per_rank": QUERIES, "parallelism": f"query-sharded x{world}",
            "l2": "flushed between timed iterations (256 MiB write); in the pipelined e2e loop the write runs on the sweep's stream before every sweep, inside the timed region"},
    "sustained": sustained,
    "strong_scaling": strong,
    "rank_checks": rank_checks,
    "slices": slices,
    "e2e": {"value": e2e_value, "unit": "cross-sections/s", "h2d_bytes_per_step": int(h2d_volume + world * 2 * pos.nbytes),
            "d2h_bytes_per_step": int(world * QUERIES * 5), "ms_per_step": ms_pipe / psteps, "steps": psteps,
            "host_input_bytes_per_step": int(vol.size + world * 2 * pos.nbytes),
            "upload": (f"bit-packed on {pack_threads} host threads{' of rank 0' if world > 1 else ''} (xs_hostpack.cpp): the 128 MiB boolean image is read once by the host cores, 16 MiB cross PCIe"
                       if (packed or staged_mode) else "plain bytes (too few host threads for the bit-packed upload)"),
            "includes": "every step: 128 MiB volume + queries taken from pinned host memory, upload, ingest, sweep, areas + contacts read back",
            "mode": e2e_mode,
            ("one_sweep_at_a_time" if world == 1 else "every_rank_uploads_its_own_copy"): {"value": e2e_serial_value, "ms_per_step": ms_e2e / args.steps},
            "volume_resident": {"value": e2e_res_value, "ms_per_step": ms_e2e_res / args.steps, "h2d_bytes_per_step": int(2 * pos.nbytes)}},
    "gpu_launches": int(launches) * world,
    "kernels_ms_per_step": {k: v / args.steps for k, v in kern.items()},
    "kernels_note": "launching stream: prepass -> warp_tier -> polygon -> combine -> mid_exposed (wait for the second stream); second stream, concurrent: mid_tier (floods) -> mid_second_launch (late hand-overs, after the warp tier) -> p2 (claims / ownership / records of the parked sections) -> second_stream_poly_combine; second_round: wide / big tiers after a host round trip, only when a section left the mid tier's window",
    "work_per_step": {k: int(stats[k]) for k in ("samples", "blocks", "polygons", "cells_tested", "escalated_mid", "escalated_big", "escalated_worst_case")},
    "roofline": roofline,
    "cpu_baseline": {"value": rate, "unit": "cross-sections/s", "cores": cores, "kind": kind, "one_core": _G.get("one_core_rate"),
                     "sample": f"{len(times)} passes of the full {QUERIES}-query sweep, fork()ed workers, one chunk per core"},
    "clocks": clocks,
  }
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="own", choices=["own", "reference"])
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference_arm(args)
  else:
    run_own_arm(args)


if __name__ == "__main__":
  main()
