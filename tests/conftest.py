import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason = "no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope = "session", autouse = True)
def _built_library():
  """The shared library is built in-tree (git-ignored, shipped with the snapshot); build it if a fresh clone lacks it."""
  lib_path = os.path.join(ROOT, "fasterrcnn_b200", "libfrcnn_sm100.so")
  if not os.path.exists(lib_path):
    import __graft_entry__ as g
    g.build()
  yield


@pytest.fixture(scope = "session")
def golden_dir():
  return os.path.join(ROOT, "tests", "golden")
