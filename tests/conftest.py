import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
  try:
    import torch
    has_gpu = torch.cuda.is_available()
  except Exception:
    has_gpu = False
  if has_gpu:
    return
  skip = pytest.mark.skip(reason="no CUDA device in this container")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
  import numpy as np
  return np.load(os.path.join(GOLDEN, "reference_golden.npz"))


@pytest.fixture(scope="session")
def knn_golden():
  """Ids returned by the reference's own show_knn.calc_nn (tests/golden/make_knn_golden.py) + the seeded row generator."""
  import importlib.util
  import numpy as np
  spec = importlib.util.spec_from_file_location("_make_knn_golden", os.path.join(GOLDEN, "make_knn_golden.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return {"ids": np.load(os.path.join(GOLDEN, "knn_ids_golden.npz")), "cases": mod.CASES, "rows": mod.rows}
