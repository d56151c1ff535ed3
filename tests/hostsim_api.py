"""Loader for tests/hostsim: the kernel headers compiled as plain C++ (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
OUT = os.path.join(ROOT, "tests", "hostsim", "_build", "libxs3d_hostsim.so")
CSRC = os.path.join(ROOT, "cross-section_b200", "csrc")

fp = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
u64 = C.c_uint64


def _f3(a):
  return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(3))


class HostSim:
  def __init__(self, lib):
    self.lib = lib
    lib.hostsim_area_batch.argtypes = [u8p, u64, u64, u64, u64, fp, fp, fp, fp, u8p, C.c_int, C.POINTER(C.c_int64)]
    lib.hostsim_projection.argtypes = [C.c_void_p, u64, u64, u64, C.c_int, C.c_int, fp, fp, fp, C.c_int, C.c_float,
                                       C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    lib.hostsim_projection.restype = C.c_int
    lib.hostsim_plane_size.argtypes = [u64, u64, u64, fp, C.c_float]
    lib.hostsim_plane_size.restype = C.c_longlong
    lib.hostsim_round.argtypes = [C.c_float]
    lib.hostsim_round.restype = C.c_float
    lib.hostsim_is_26_connected.argtypes = [C.c_uint8, C.c_uint8, C.c_int, C.c_int, C.c_int]
    lib.hostsim_voxel_area.argtypes = [u64, u64, u64, fp, fp, fp]
    lib.hostsim_voxel_area.restype = C.c_float
    lib.hostsim_block_area.argtypes = [C.c_uint8, fp, fp, fp, fp]
    lib.hostsim_block_area.restype = C.c_float
    lib.hostsim_miss_check.argtypes = [u64, u64, C.c_float, C.POINTER(C.c_int64)]
    lib.hostsim_miss_check.restype = None

  def miss_check(self, n_cases, seed, span):
    out = (C.c_int64 * 5)()
    self.lib.hostsim_miss_check(n_cases, seed, C.c_float(span), out)
    return dict(zip(("cases", "violations", "miss", "far", "polygons"), list(out)))

  def area_batch(self, vol, pos, nrm, w, mode=0):
    a = np.asfortranarray(vol)
    a = a.view(np.uint8) if a.dtype == bool else a
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1, 3)
    w = _f3(w)
    nq = len(pos)
    area = np.zeros(nq, np.float32)
    ct = np.zeros(nq, np.uint8)
    st = (C.c_int64 * 3)()
    self.lib.hostsim_area_batch(a.ctypes.data_as(u8p), *a.shape, nq, pos.ctypes.data_as(fp), nrm.ctypes.data_as(fp),
                                w.ctypes.data_as(fp), area.ctypes.data_as(fp), ct.ctypes.data_as(u8p), mode, st)
    return area, ct, list(st)

  def projection(self, labels, pos, normal, w, sb=True, crop=float("inf"), mode=0):
    if not labels.flags.f_contiguous:
      labels = np.ascontiguousarray(labels)
    c_order = bool(labels.flags.c_contiguous)
    if crop == 0:
      return np.zeros((0, 0), labels.dtype, order="F")
    ww = _f3(w)
    psx = self.lib.hostsim_plane_size(*labels.shape, ww.ctypes.data_as(fp), C.c_float(crop))
    udt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[labels.dtype.itemsize]
    out = np.zeros((psx, psx), udt, order="F")
    bbox = (C.c_int64 * 4)()
    p, n = _f3(pos), _f3(normal)
    self.lib.hostsim_projection(labels.ctypes.data, *labels.shape, labels.dtype.itemsize, int(c_order),
                                p.ctypes.data_as(fp), n.ctypes.data_as(fp), ww.ctypes.data_as(fp), int(sb),
                                C.c_float(crop), out.ctypes.data, bbox, mode)
    return np.asfortranarray(out[bbox[0]:bbox[1] + 1, bbox[2]:bbox[3] + 1]).view(labels.dtype)


def load():
  deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
  if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-I", CSRC, SRC, "-o", OUT])
  return HostSim(C.CDLL(OUT))
