"""Generate golden vectors by RUNNING the reference's own Python (build container only).

/root/reference is read-only and does not travel to the GPU box, so this script is
run once here and its small outputs are committed under tests/golden/.  TensorFlow and
faiss are not installable, so only the reference functions that are pure numpy/Python
are executed; ``tensorflow`` and ``faiss`` are replaced by import shims that expose
nothing but ``logging`` and ``flags`` (what those modules touch at import time).

  python tests/golden/make_golden.py          # rewrites tests/golden/*.npz|*.txt|*.json
"""
import io
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_shims():
  import logging as pylogging
  from absl import flags as absl_flags

  tf = types.ModuleType("tensorflow")
  lg = types.ModuleType("tensorflow.logging")
  for name in ("debug", "info", "warning", "error", "warn"):
    setattr(lg, name, getattr(pylogging, name if name != "warn" else "warning"))
  lg.DEBUG = pylogging.DEBUG
  lg.set_verbosity = lambda *_: None
  tf.logging = lg
  tf.flags = absl_flags
  tf.app = types.SimpleNamespace(run=lambda *a, **k: None)
  tf.gfile = types.SimpleNamespace(Glob=lambda p: sorted(__import__("glob").glob(p)))
  sys.modules["tensorflow"] = tf
  sys.modules["tensorflow.logging"] = lg
  sys.modules["tensorflow.flags"] = absl_flags
  sys.modules["faiss"] = types.ModuleType("faiss")
  return absl_flags


def main():
  flags = install_shims()
  sys.path.insert(0, REF)
  os.makedirs("./logs", exist_ok=True)
  import imitation_data
  import parse_data
  import online_data
  import evaluate
  import faiss_knn
  flags.FLAGS(["make_golden", "--feature_size=12"])

  gold = {}

  # -- synthetic features: imitation_data.gen_features with the global RNG seeded --
  np.random.seed(1234)
  feats = imitation_data.gen_features(64, 12)               # float64, rounded to 8 decimals
  gold["gen_features_seed1234"] = feats
  np.random.seed(7)
  gold["gen_triplets_seed7"] = imitation_data.gen_triplets(5, 4)

  # -- gather: inputs.py:158 is FEATURES[np.asarray(guid_triplets)] (numpy semantics) --
  features32 = feats.astype(np.float32)
  np.random.seed(99)
  neg_iter = parse_data.yield_negative_index(len(features32), putback=True)
  pairs = np.array([[3, 9], [10, 11], [0, 63], [5, 5], [62, 1], [17, 40], [40, 17], [8, 2]])
  trip = []
  for a, p in pairs:                                         # inputs.py:123-129 verbatim loop
    t = [int(a), int(p)]
    neg = neg_iter.__next__()
    while neg in t:
      neg = neg_iter.__next__()
    t.append(neg)
    trip.append(t)
  gold["sampler_seed99_pairs"] = pairs
  gold["sampler_seed99_triplets"] = np.asarray(trip, np.int64)
  gold["gather_features"] = features32
  gold["gather_out"] = features32[np.asarray(trip)]          # [8,3,12]

  # -- evaluate.Evaluation: _rencode + mean_dist (evaluate.py:34-73) --
  rng = np.random.RandomState(5)
  vec = evaluate.l2_normalize(rng.standard_normal((64, 16))).astype(np.float32)
  cow = [[3, 9], [10, 11], [9, 40], [62, 3]]
  ev = evaluate.Evaluation(features32, cow)
  gold["eval_vectors"] = vec
  gold["eval_cowatches"] = np.asarray(cow)
  gold["eval_rencoded_features"] = ev.features
  gold["eval_rencoded_cowatches"] = np.asarray(ev.cowatches)
  gold["eval_mean_dist"] = np.float64(ev.mean_dist(vec, np.asarray(cow)))

  # -- knn_result writer: faiss_knn.write_process (faiss_knn.py:267-283) --
  D = np.array([[0.0, 0.25, 0.5, 1.39999, 1.4, 0.7],
                [0.0, 0.1, 1.5, 0.3, 0.0, 0.9],
                [1e-7, 0.33333334, 0.2, 0.6, 0.8, 1.2]], np.float32)
  I = np.array([[0, 4, 2, 1, 3, 0],
                [1, 0, 2, 3, 4, 2],
                [2, 1, -1, 4, 0, 3]], np.int64)
  decode = {i: "guid%02d" % i for i in range(5)}
  faiss_knn.DECODE_MAP = decode
  tmp = os.path.join(OUT, "_tmp_knn")
  os.makedirs(tmp, exist_ok=True)
  faiss_knn.write_process(tmp, 0, 0, D, I, "knn_split")
  with open(os.path.join(tmp, "knn_split0")) as f:
    knn_text = f.read()
  os.remove(os.path.join(tmp, "knn_split0"))
  os.rmdir(tmp)
  gold["knn_D"], gold["knn_I"] = D, I
  with open(os.path.join(OUT, "knn_split0.txt"), "w") as f:
    f.write(knn_text)
  with open(os.path.join(OUT, "knn_decode_map.json"), "w") as f:
    json.dump({str(k): v for k, v in decode.items()}, f)

  # -- feature text reader + cowatch loader (online_data.py:48-84, 125-142) --
  txt = os.path.join(OUT, "features_small.txt")
  with open(txt, "w") as f:
    for i in range(6):
      f.write("g%d#" % i + ",".join("%.6f" % v for v in feats[i]) + "\n")
    f.write("bad#1,2,3\n")                                   # wrong width -> dropped
    f.write("g7#" + ",".join("%.6f" % v for v in feats[7]) + "\n")
  fe, enc, dec = online_data.read_features_txt(txt)
  gold["read_features_txt"] = fe
  with open(os.path.join(OUT, "features_small_maps.json"), "w") as f:
    json.dump({"encode": enc, "decode": {str(k): v for k, v in dec.items()}}, f)
  cw = os.path.join(OUT, "cowatches_small.eval")
  with open(cw, "w") as f:
    f.write("3,9\n10,11\nxx,1\n9,40\n")
  gold["load_cowatches"] = np.asarray(online_data.load_cowatches(cw))

  # -- de-similarity filter pieces that are pure numpy (faiss_knn.py:134-154) --
  eI = np.array([[0, 5, 7, 9], [1, 2, 3, 4], [2, 8, 6, 0]], np.int64)
  fI = np.array([[0, 7, 3], [1, 9, 4], [2, 0, 5]], np.int64)
  fD = np.array([[0.0, 0.5, 1.5], [0.0, 1.45, 0.2], [0.0, 0.3, 0.1]], np.float32)
  gold["desim_eI_in"], gold["desim_fI_in"], gold["desim_fD_in"] = eI.copy(), fI.copy(), fD.copy()
  gold["fliter_fI_out"] = faiss_knn.fliter_fI(fI.copy(), fD, 1.4)
  gold["desim_out"] = faiss_knn.desim(eI.copy(), fI.copy())

  np.savez(os.path.join(OUT, "reference_golden.npz"), **gold)

  # -- known answers derived from the reference's loss fixture (tests/test_losses.py:13-18) --
  fixture = np.array([[[1, 1], [2, 2], [5, 5]], [[1, 1], [5, 5], [2, 2]], [[1, 1], [2, 2], [5, 5]],
                      [[1, 1], [5, 5], [2, 2]], [[1, 1], [5, 5], [2, 2]]], np.float32)
  with open(os.path.join(OUT, "loss_fixture.json"), "w") as f:
    json.dump({"source": "tests/test_losses.py:13-18; values derived by hand from losses.py:33-38",
               "triplets": fixture.tolist(),
               "pos_dist": [2, 32, 2, 32, 32], "neg_dist": [32, 2, 32, 2, 2],
               "hinge_dist@0.1": [0, 30.1, 0, 30.1, 30.1], "hinge_loss@0.1": 18.06,
               "hinge_loss@0.8": 18.48}, f, indent=1)
  print("golden written to", OUT, "keys:", sorted(gold))


if __name__ == "__main__":
  main()
