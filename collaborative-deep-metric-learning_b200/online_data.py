"""File-format readers of the reference's online_data.py that sit either side of the hot path:
read_features_txt (:48-84), read_features_npy (:87-93), load_cowatches (:125-142) and the flags it defines
(:30-45; `feature_size` reaches train.py through this module -- SURVEY.md Q12).  The txt->npy ETL is out of scope."""
import logging

import numpy as np
from absl import flags

FLAGS = flags.FLAGS


def _define(kind, name, default, doc):
  if name not in FLAGS:
    getattr(flags, "DEFINE_" + kind)(name, default, doc)


_define("string", "base_save_dir", "", "root of the generated training set")
_define("string", "training_click_records", "", "watched-guid text file")
_define("string", "training_dense_feature", "", "feature text file")
_define("integer", "feature_size", 1628, "width of a feature row: 1500 visual (+128 doc)")
_define("integer", "threshold", 1, "cowatch count threshold")
_define("integer", "split_num", 10, "number of *.train shards")
_define("boolean", "unique", False, "keep each cowatch once")


def feature_size():
  try:
    return FLAGS.feature_size
  except flags.UnparsedFlagAccessError:
    return FLAGS["feature_size"].default


def read_features_txt(filename, width=None):
  """Lines 'guid#f1,f2,...' -> (float32 [n,width], {guid:i}, {i:guid}); rows of the wrong width are dropped."""
  width = feature_size() if width is None else width
  rows, encode_map, decode_map = [], {}, {}
  with open(filename, "r") as f:
    for line in f:
      line = line.rstrip("\n")
      parts = line.split("#")
      if len(parts) != 2:
        logging.warning("read_features_txt: malformed line dropped")
        continue
      try:
        vals = np.array(parts[1].split(","), dtype=np.float64)
      except ValueError as e:
        logging.warning("read_features_txt: drop feature. %s", e)
        continue
      if vals.shape[0] != width:
        continue
      encode_map[parts[0]] = len(rows)
      decode_map[len(rows)] = parts[0]
      rows.append(vals.astype(np.float32))
  feats = np.stack(rows).astype(np.float32) if rows else np.zeros((0, width), np.float32)
  return feats, encode_map, decode_map


def read_features_npy(filename):
  """features.npy: row i is the feature of guid index i."""
  return np.load(filename)


def load_cowatches(filename):
  """Lines 'a,b' -> [[a,b],...]; unparsable lines are skipped with a warning."""
  cowatches = []
  with open(filename, "r") as f:
    for line in f:
      ids = line.strip().split(",")
      try:
        cowatches.append([int(ids[0]), int(ids[1])])
      except (ValueError, IndexError) as e:
        logging.warning(str(e))
  logging.info("online_data load_cowatches num:%d", len(cowatches))
  return cowatches
