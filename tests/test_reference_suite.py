"""GPU (-m gpu): the reference's OWN test file (/root/reference/automated_test.py:7-465, what its CI runs:
.github/workflows/tests.yml:31-33) executed unchanged on the CUDA path, two ways:

  shim    : the reference's own Python package (xs3d/__init__.py) with tests/shim/fastxs3d.py in place of its pybind11
            extension — i.e. through the five C-ABI entry points a maintainer would bind;
  package : this repo's host layer (cross-section_b200/xs3d) as the drop-in `import xs3d`.

The test file and the reference package are copied, as they are, into oracle/_ref/pytests by oracle/Makefile in the dev
container (git-ignored, travels to the GPU box like the compiled reference); tests/shim/xs_refsuite_plugin.py sub-samples
test_sphere's 500 k-normal loop and deselects test_2d for the shim run only (there the reference's own pure-Python twod.py
would run, on the CPU, and needs cc3d); through the package test_2d runs on the device (xs_twod.cu)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYTESTS = os.path.join(ROOT, "oracle", "_ref", "pytests")
SHIM = os.path.join(ROOT, "tests", "shim")


def _run(expect, path):
  test_file = os.path.join(PYTESTS, "automated_test.py")
  if not os.path.exists(test_file):
    pytest.skip("oracle/_ref/pytests not built (needs /root/reference: make -C oracle ref)")
  env = dict(os.environ)
  env["PYTHONPATH"] = os.pathsep.join(path + [SHIM])
  env["XS_REFSUITE_EXPECT"] = expect
  r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "xs_refsuite_plugin", "-p", "no:cacheprovider",
                      "--rootdir", PYTESTS, "-c", os.devnull, test_file],
                     env=env, cwd=PYTESTS, capture_output=True, text=True, timeout=1500)
  tail = (r.stdout + r.stderr)[-3000:]
  assert r.returncode == 0, tail
  assert " passed" in r.stdout and "failed" not in r.stdout, tail
  return r.stdout


def test_reference_tests_through_the_fastxs3d_shim():
  out = _run("shim", [SHIM, os.path.join(PYTESTS, "refpkg")])
  print(out[-400:])


def test_reference_tests_through_the_package():
  out = _run("package", [os.path.join(ROOT, "cross-section_b200")])
  print(out[-400:])
