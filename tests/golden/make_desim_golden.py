"""Golden vectors for the de-similarity post-filter, produced by RUNNING the reference's own faiss_knn.iter_desim_mp /
add_invalid_row / desim (faiss_knn.py:134-244) in the build container (numpy + multiprocessing only; `tensorflow` and
`faiss` are import shims as in make_golden.py; `np.int`, removed from numpy >= 1.24, is aliased to `int` for
faiss_knn.py:164).  Output: tests/golden/desim_golden.npz.

  python tests/golden/make_desim_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, install_shims  # noqa: E402


def knn_like(rng, n, k, n_ids, self_first=True, holes=0.0):
  """Rows of distinct ids (as a KNN returns them), optionally with the row's own id first and -1 padding at the end."""
  I = np.empty((n, k), np.int64)
  for r in range(n):
    ids = rng.choice(n_ids, size=k, replace=False)
    if self_first and r < n_ids:
      ids = np.concatenate(([r], ids[ids != r][:k - 1]))
    I[r] = ids
    if holes and rng.rand() < holes:
      I[r, rng.randint(1, k):] = -1
  return I


def main():
  install_shims()
  sys.path.insert(0, REF)
  os.makedirs("./logs", exist_ok=True)
  if not hasattr(np, "int"):
    np.int = int
  import faiss_knn
  gold = {}
  cases = [  # name, n, ke, kf, id range, fI_end, process_num, holes
      ("a", 60, 9, 6, 60, 31, 1, 0.0),
      ("b", 211, 13, 8, 211, 5, 3, 0.2),       # f_end cuts the feature lists; some rows end in -1 padding
      ("c", 40, 81, 26, 40, 31, 7, 0.0),       # more columns than distinct ids cannot happen -> ids drawn from 0..n-1 w/o repl.
      ("d", 97, 20, 31, 97, 31, 200, 0.1),     # more processes than rows: every patch but the last is empty
  ]
  for name, n, ke, kf, nid, f_end, procs, holes in cases:
    rng = np.random.RandomState(len(name) * 7 + n)
    ke_eff = min(ke, nid)
    eI = knn_like(rng, n, ke_eff, nid, True, holes)
    fI = knn_like(rng, n, min(kf, nid), nid, True, holes)
    # make the two neighbourhoods overlap the way real ones do: copy a few embedding neighbours into the feature lists
    for r in range(n):
      take = rng.randint(0, 4)
      if take:
        fI[r, 1:1 + take] = eI[r, 1:1 + take]
    fD = np.sort(rng.rand(n, fI.shape[1]).astype(np.float32) * 2.0, axis=1)
    fD[:, 0] = 0.0
    gold[name + "_eI"], gold[name + "_fI"], gold[name + "_fD"] = eI.copy(), fI.copy(), fD.copy()
    gold[name + "_args"] = np.array([f_end, procs], np.int64)
    out = faiss_knn.iter_desim_mp(eI.copy(), fI.copy(), fD.copy(), fD_threshold=1.4, fI_end=f_end, process_num=procs)
    gold[name + "_out"] = np.asarray(out)
    gold[name + "_simple"] = faiss_knn.desim(eI.copy(), fI.copy())
  np.savez_compressed(os.path.join(OUT, "desim_golden.npz"), **gold)
  print("desim golden:", {k: v.shape for k, v in gold.items() if k.endswith("_out")})


if __name__ == "__main__":
  main()
