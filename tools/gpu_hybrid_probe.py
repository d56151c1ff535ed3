"""Dev probe: can part of a 128 MiB boolean image go up raw (DMA) while the host pool packs the rest?  Prints the time of
packing a fraction of the image alone, of the raw copy of the remainder alone, and of both at once."""
import os, sys, time, threading, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
import numpy as np, torch
from xs3d import _abi
lib = C.CDLL(_abi.LIB_PATH)
lib.xs_pack_bits.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
N = 512 ** 3
hvol = torch.randint(0, 2, (N,), dtype=torch.uint8).pin_memory()
hb = torch.empty(N // 8 + 64, dtype=torch.uint8).pin_memory()
d = torch.empty(N, dtype=torch.uint8, device="cuda")
db = torch.empty(N // 8 + 64, dtype=torch.uint8, device="cuda")
T = time.perf_counter
def pack(lo, hi):
  lib.xs_pack_bits(hvol.data_ptr(), lo, hi, hb.data_ptr())
for frac in (0.0, 0.1, 0.15, 0.2, 0.25, 0.3):
  cut = (int(N * (1 - frac)) // (1 << 20)) << 20
  res = []
  for rep in range(6):
    torch.cuda.synchronize()
    t0 = T(); pack(0, cut); tp = T() - t0
    t0 = T()
    if frac: d[cut:].copy_(hvol[cut:], non_blocking=True)
    torch.cuda.synchronize(); tc = T() - t0
    t0 = T()
    if frac: d[cut:].copy_(hvol[cut:], non_blocking=True)
    pack(0, cut)
    db[: cut // 8].copy_(hb[: cut // 8], non_blocking=True)
    torch.cuda.synchronize(); tb = T() - t0
    res.append((tp, tc, tb))
  tp, tc, tb = np.median(np.array(res[2:]), axis=0)
  print("raw fraction %.2f: pack alone %.3f ms, raw copy alone %.3f ms, both + bits copy %.3f ms" % (frac, tp * 1e3, tc * 1e3, tb * 1e3))
