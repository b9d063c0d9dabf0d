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


def read_features_txt(filename, width=None, num_threads=0):
  """Lines 'guid#f1,f2,...' -> (float32 [n,width], {guid:i}, {i:guid}); lines without exactly one '#', with a field that
  float() would reject or with another field count are dropped (online_data.py:48-84).  Parsed by libcdml's host code
  (cdml_parse_features_txt: the file image is split into lines and parsed on all cores)."""
  from . import _lib
  width = feature_size() if width is None else width
  data = np.fromfile(filename, dtype=np.uint8)
  if data.size == 0:
    return np.zeros((0, width), np.float32), {}, {}
  max_rows = int(np.count_nonzero(data == 10)) + 1
  feats = np.empty((max_rows, width), np.float32)
  gb, gl = np.empty(max_rows, np.int64), np.empty(max_rows, np.int32)
  n = _lib.load().cdml_parse_features_txt(data.ctypes.data, data.size, int(width), feats.ctypes.data, max_rows,
                                          gb.ctypes.data, gl.ctypes.data, int(num_threads))
  if n < 0:
    _lib.check(-1)
  encode_map, decode_map = {}, {}
  view = memoryview(data)
  for i in range(n):
    guid = bytes(view[gb[i]:gb[i] + gl[i]]).decode("utf-8")
    encode_map[guid] = i
    decode_map[i] = guid
  logging.info("read_features_txt drop features count:%d", max_rows - n)
  return feats[:n].copy() if n < max_rows else feats, encode_map, decode_map


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
