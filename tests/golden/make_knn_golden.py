"""Golden exact-KNN ids produced by RUNNING the reference's own brute-force search (build container only).

`show_knn.calc_nn` (show_knn.py:63-79) is the one exact nearest-neighbour statement the reference contains that is pure
numpy: per query `all_embedding.dot(all_embedding[index])`, `np.argsort(-dist)[:nearestN]`.  It imports cv2 / matplotlib /
requests-based helpers at module top, which are replaced by empty import shims (nothing of them is touched by calc_nn);
`nearestN` is the module's own global.  The ids it returns are what `faiss.IndexFlatIP.search` must return for the same
rows (and, the rows being unit vectors, what IndexFlatL2 returns).  The rows are regenerated from their seed by the tests,
only the ids are committed.

  python tests/golden/make_knn_golden.py          # rewrites tests/golden/knn_ids_golden.npz
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {          # name: (seed, N, d, number of queries, k, clustered)
    "gauss": (4, 3000, 256, 96, 20, False),
    "clustered": (5, 2500, 256, 64, 30, True),
    "raw_width": (6, 1200, 1628, 48, 26, False),
}


def rows(seed, N, d, clustered):
  """Unit-norm fp32 rows: the C4 generator of SURVEY 8(d), or 50 Gaussian centres with within-cluster cosine ~0.8 (dense
  neighbourhoods: many candidates within a few 1e-4 of the k-th inner product)."""
  rng = np.random.RandomState(seed)
  if clustered:
    centres = rng.standard_normal((50, d))
    x = centres[rng.randint(0, 50, N)] + 0.5 * rng.standard_normal((N, d))
  else:
    x = rng.standard_normal((N, d))
  x = x.astype(np.float32)
  x /= np.linalg.norm(x, axis=1, keepdims=True)            # faiss_knn.py:100-104 (numpy row normalisation)
  return x


def main():
  for name in ("cv2", "matplotlib", "matplotlib.pyplot", "requests", "xml", "xml.dom", "xml.dom.minidom"):
    if name not in sys.modules:
      try:
        __import__(name)
      except Exception:
        sys.modules[name] = types.ModuleType(name)
  sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
  sys.path.insert(0, REF)
  import show_knn
  gold = {}
  for name, (seed, N, d, nq, k, clustered) in CASES.items():
    X = rows(seed, N, d, clustered)
    queries = np.random.RandomState(seed + 100).choice(N, nq, replace=False)
    show_knn.nearestN = k
    ident = {i: str(i) for i in range(N)}                  # decode_map: index -> "guid" string; str(index) keeps the ids
    with contextlib.redirect_stdout(io.StringIO()):
      flat = show_knn.calc_nn(queries, X, ident)
    gold[name + "_queries"] = queries.astype(np.int64)
    gold[name + "_ids"] = np.asarray([int(g) for g in flat], np.int64).reshape(nq, k)
    # gap between the k-th and (k+1)-th inner product: rows where it is below fp32 noise are the "exact tie" exemption
    ip = X[queries] @ X.T
    srt = -np.sort(-ip, axis=1)
    gold[name + "_gap"] = (srt[:, :k] - srt[:, 1:k + 1]).min(axis=1).astype(np.float32)
  np.savez(os.path.join(OUT, "knn_ids_golden.npz"), **gold)
  print({k: v.shape for k, v in gold.items()})


if __name__ == "__main__":
  main()
