"""Build libxs3d_b200.so for sm_100a with nvcc (in-tree, so the .so travels with the repo snapshot).

  python cross-section_b200/build.py [--force]

Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo; -fmad=false is belt and braces (all float
arithmetic already goes through __f*_rn intrinsics, xs_math.cuh).  Each .cu is compiled to its own
object (in parallel) and linked with the shared CUDA runtime.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
# dev aid for A/B runs: XS3D_BUILD_TAG=x XS3D_EXTRA_FLAGS="-DFOO=1" builds lib/libxs3d_b200_x.so beside the product library
TAG = os.environ.get("XS3D_BUILD_TAG", "")
OUT = os.path.join(HERE, "lib", "libxs3d_b200%s.so" % ("_" + TAG if TAG else ""))
OBJ = os.path.join(HERE, "lib", "obj" + ("_" + TAG if TAG else ""))
UNITS = ["xs_device.cu", "xs_slow.cu", "xs_slice.cu", "xs_section.cu", "xs_skeleton.cu", "xs_twod.cu", "xs_hostpack.cpp"]   # .cpp: host-only code, compiled by g++
CXX = os.environ.get("CXX", "g++")
NVCC = os.environ.get("NVCC", "nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("XS3D_EXTRA_FLAGS", "").split()


def _deps():
  d = [os.path.join(SRC, f) for f in os.listdir(SRC)]
  d.append(os.path.join(os.path.dirname(HERE), "include", "xs3d_b200.h"))
  return d


def _compile(unit):
  obj = os.path.join(OBJ, os.path.splitext(unit)[0] + ".o")
  src = os.path.join(SRC, unit)
  newest = max(os.path.getmtime(p) for p in _deps())
  if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
    return obj, ""
  if unit.endswith(".cpp"):
    r = subprocess.run([CXX, "-std=c++17", "-O3", "-fPIC", "-pthread", "-c", src, "-o", obj], capture_output=True, text=True)
  else:
    r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError(f"compiler failed on {unit}:\n{r.stdout}\n{r.stderr}")
  return obj, r.stderr


def build(force=False, verbose=False):
  os.makedirs(OBJ, exist_ok=True)
  if force:
    for f in os.listdir(OBJ):
      os.remove(os.path.join(OBJ, f))
  with cf.ThreadPoolExecutor(len(UNITS)) as ex:
    results = list(ex.map(_compile, UNITS))
  objs = [o for o, _ in results]
  log = "\n".join(l for _, l in results if l)
  if log:
    with open(os.path.join(HERE, "lib", "ptxas%s.log" % ("_" + TAG if TAG else "")), "w") as f:
      f.write(log)
    if verbose:
      print(log)
  if force or not os.path.exists(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs):
    r = subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-lcudart", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
  return OUT


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose=True))
