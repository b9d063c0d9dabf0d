"""The two parse_data.py helpers the hot path touches: the negative-index stream (parse_data.py:292-303) and the
unique-guid scan used by the evaluator (parse_data.py:17-26).  The cowatch-mining ETL of that file runs once per
day on the host and is out of scope."""
from itertools import cycle

import numpy as np


def get_unique_watched_guids(all_watched_guids):
  """Distinct elements of a list of lists, as a list."""
  seen = set()
  for watched in all_watched_guids:
    seen.update(watched)
  return list(seen)


def yield_negative_index(size, putback=False):
  """Endless stream of integers in [0,size): with replacement (`np.random.randint` per draw, global RNG -- the
  reference's reader uses this form, inputs.py:105) or a shuffled cycle."""
  if putback:
    while True:
      yield np.random.randint(0, size)
  else:
    indexes = list(range(size))
    np.random.shuffle(indexes)
    for neg_index in cycle(indexes):
      yield neg_index
