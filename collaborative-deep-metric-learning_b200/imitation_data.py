"""Synthetic data -- mirror of the reference's imitation_data.py (constants :8-11, gen_features :41-53,
gen_watched_guids :56-85, gen_triplets :88-93, arrays_to_dict :96-106).  Uses numpy's global RNG exactly like the
reference, so `np.random.seed(s)` before a call reproduces the reference's output (tests/golden)."""
import random

import numpy as np

num_uid = 30000
num_guid = 10000
feature_size = 1500


def gen_unique_id_array(low, high, size, dtype=None):
  """`size` distinct integers from [low, high] in random order."""
  if low > high:
    raise ValueError("low > high")
  if high - low + 1 < size:
    raise ValueError("size is larger than high-low+1")
  if size < 0:
    raise ValueError("size is negative")
  ids = np.array(random.sample(range(low, high + 1), size))
  return ids.astype(dtype) if dtype else ids


def gen_features(num_feature, feature_size, decimals=8):
  """[num_feature, feature_size] uniform [0,1) rounded to `decimals` (float64)."""
  return np.around(np.random.random((num_feature, feature_size)), decimals)


def gen_watched_guids(guids, low, high):
  return np.random.choice(guids, random.randint(low, high)).tolist()


def gen_all_watched_guids(guids, num_cowatch, low=2, high=30):
  return [gen_watched_guids(guids, low, high) for _ in range(num_cowatch)]


def gen_triplets(batch_size, feature_size):
  return np.reshape(gen_features(batch_size * 3, feature_size), [batch_size, 3, feature_size])


def arrays_to_dict(array_1d, array_2d):
  return dict(zip(array_1d, array_2d))
