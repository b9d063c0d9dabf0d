"""Evaluation -- mirror of the reference's evaluate.py (l2_normalize :14-17, Evaluation :20-90).  `mean_dist` runs on
the GPU (cdml_mean_pair_dist) when given a CUDA tensor, i.e. inside the training loop; numpy inputs take the same
arithmetic through a device round trip."""
import logging

import numpy as np
import torch

from . import ops
from .parse_data import get_unique_watched_guids


def l2_normalize(a, axis=-1, order=2):
  l2 = np.atleast_1d(np.linalg.norm(a, order, axis))
  l2[l2 == 0] = 1
  return a / np.expand_dims(l2, axis)


class Evaluation():
  def __init__(self, features, cowatches):
    """features: [G,F] original features; cowatches: index pairs into them (evaluate.py:21-32)."""
    try:
      self.features, self.cowatches = self._rencode(features, cowatches)
    except Exception as e:  # the reference swallows and disables evaluation
      logging.warning("Evaluation.__init__ features or cowatches %s", e)
      self.features, self.cowatches = None, None

  def _rencode(self, features, cowatches):
    """Keep only the guids the pairs touch, renumbered 0..U-1 in ascending order (evaluate.py:34-55)."""
    sorted_indexes = np.sort(np.asarray(get_unique_watched_guids(cowatches)))
    eval_features = features[sorted_indexes]
    index_map = {int(old): new for new, old in enumerate(sorted_indexes)}
    return eval_features, [[index_map[int(i)] for i in pair] for pair in cowatches]

  def mean_dist(self, vectors, cowatches):
    """mean over pairs of sum_d (v_a - v_b)^2 (evaluate.py:57-73)."""
    pairs = torch.as_tensor(np.asarray(cowatches, np.int64))
    if torch.is_tensor(vectors) and vectors.is_cuda:
      return float(ops.mean_pair_dist(vectors.float().contiguous(), pairs.to(vectors.device)).item())
    v = torch.as_tensor(np.asarray(vectors, np.float32)).cuda()
    return float(ops.mean_pair_dist(v, pairs.cuda()).item())

  def mean_cos_dist(self, vectors, cowatches):
    """mean over pairs of v_a . v_b (evaluate.py:75-90); cheap, host numpy."""
    co = np.asarray(vectors)[np.asarray(cowatches)]
    return float(np.mean(np.sum(co[:, 0, :] * co[:, 1, :], axis=-1)))
