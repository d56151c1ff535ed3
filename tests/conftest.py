import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
  try:
    import torch
    return torch.cuda.is_available()
  except Exception:
    return False


def pytest_collection_modifyitems(config, items):
  if _has_gpu():
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
  """The plain-C restatement, (re)built on demand (gcc is on every box)."""
  import refapi
  src = os.path.join(ROOT, "oracle", "xs3d_oracle.c")
  if not refapi.have_oracle() or os.path.getmtime(refapi.ORACLE_SO) < os.path.getmtime(src):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
  return refapi.Oracle()


@pytest.fixture(scope="session")
def ref():
  """The compiled, unmodified reference (oracle/_ref); prebuilt in the dev container, travels to the GPU box."""
  import refapi
  if not refapi.have_ref():
    pytest.skip("oracle/_ref not built (needs /root/reference)")
  return refapi.Ref()


@pytest.fixture(scope="session")
def hostsim():
  import hostsim_api
  return hostsim_api.load()


@pytest.fixture(scope="session")
def xs3d_gpu():
  import xs3d
  return xs3d


@pytest.fixture(scope="session")
def golden():
  return np.load(os.path.join(GOLDEN, "golden_v1.npz"))


def bits(a):
  return np.asarray(a, dtype=np.float32).view(np.uint32)
