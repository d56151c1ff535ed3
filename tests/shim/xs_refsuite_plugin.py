"""pytest plugin for running the reference's automated_test.py unchanged on the CUDA path (tests/test_reference_suite.py):
  * test_2d runs through this repo's package (its 2-D path is xs_twod.cu); it is deselected for the shim run, where the
    reference's own pure-Python xs3d/twod.py would run on the CPU — and needs the third-party cc3d, absent here;
  * test_sphere's 1000 x 500 normal loop (automated_test.py:229-230) is sub-sampled to 40 x 20 by giving the test
    module its own `range` (loops of < 500 iterations run in full);
  * XS_REFSUITE_EXPECT says which `fastxs3d` / `xs3d` must have been imported, so a run can never silently test the
    reference's own extension."""
import builtins
import os
import sys


def _sub_range(*a):
  r = builtins.range(*a)
  if len(r) >= 500:
    return r[:: len(r) // 20 if len(r) < 1000 else len(r) // 40]
  return r


def pytest_collection_modifyitems(session, config, items):
  keep, drop = [], []
  drop_2d = os.environ.get("XS_REFSUITE_EXPECT", "") != "package"
  for it in items:
    (drop if (drop_2d and it.name.startswith("test_2d")) else keep).append(it)
    it.module.range = _sub_range
  if drop:
    config.hook.pytest_deselected(items=drop)
    items[:] = keep
  expect = os.environ.get("XS_REFSUITE_EXPECT", "")
  import xs3d
  if expect == "shim":
    import fastxs3d
    assert fastxs3d.__file__.endswith(os.path.join("tests", "shim", "fastxs3d.py")), fastxs3d.__file__
    assert os.path.join("refpkg", "xs3d") in xs3d.__file__, xs3d.__file__
  elif expect == "package":
    assert os.path.join("cross-section_b200", "xs3d") in xs3d.__file__, xs3d.__file__
    assert "fastxs3d" not in sys.modules
