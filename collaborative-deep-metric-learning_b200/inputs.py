"""Triplet reader -- host-side mirror of the reference's inputs.py (MPTripletPipe, inputs.py:62-172).

Same surface: `MPTripletPipe(cowatch_file_patten, feature_file, wait_times)`, `.cowatch_num`, `.create_pipe(num_epochs,
batch_size, queue_length)`, `.get_batch() -> np.float32 [B,3,F] | None`, module global `FEATURES`.

What changed underneath:
  * the feature table is uploaded once and stays resident in HBM; `get_batch` gathers rows with libcdml's
    cdml_gather_rows (bit-exact with `FEATURES[np.asarray(guid_triplets)]`, inputs.py:158) and copies the batch back,
    while `get_batch_indices()` hands the [B,3] int64 index triplets to the fused training step so that features
    never cross PCIe;
  * the reader keeps the per-file semantics of `subprocess` (inputs.py:102-142) -- every batch comes from ONE file,
    a file is re-read `num_epochs` times back to back, the trailing partial batch is dropped, each negative is
    `randint(0,G)` re-drawn while it equals the anchor or the positive -- but runs in-process, files served
    round-robin, from a seeded RandomState (the reference forks unseeded workers that all replay one stream, Q10);
  * rank/world sharding for data-parallel training: rank r serves batches r, r+world, ... of that round-robin order, in
    complete rounds of `world` batches (a trailing incomplete round is dropped so every rank runs the same step count);
  * `device_reader=True`: the pairs of every file stay resident in HBM and `get_batch_indices_device()`
    produces the [B,3] triplets on the device (cdml_sample_triplets: same file order, wrap-around and negative rule, a
    counter-based Philox generator per stream position instead of numpy's Mersenne Twister) -- no host work per step.
"""
import glob
import logging

import numpy as np
import torch

from . import ops
from .online_data import read_features_npy

FEATURES = {}
_DEVICE_FEATURES = None


class BasePipe(object):
  """Inherit from this class when implementing new readers (inputs.py:24-29)."""

  def create_pipe(self, unused_data, **unused_params):
    raise NotImplementedError()


def _load_pairs(path):
  """'a,p' per line -> int64 [n,2] (lines that do not parse are dropped, as the reader's except-branch does)."""
  try:
    arr = np.loadtxt(path, delimiter=",", dtype=np.int64, ndmin=2)
    if arr.shape[1] >= 2:
      return np.ascontiguousarray(arr[:, :2])
  except Exception:
    pass
  rows = []
  with open(path, "r") as f:
    for line in f:
      parts = line.strip().split(",")
      try:
        rows.append((int(parts[0]), int(parts[1])))
      except (ValueError, IndexError):
        continue
  return np.asarray(rows, np.int64).reshape(-1, 2)


class MPTripletPipe(object):
  def __init__(self, cowatch_file_patten, feature_file=None, wait_times=30, seed=1, rank=0, world=1, device=None,
               device_reader=False):
    global FEATURES, _DEVICE_FEATURES
    self.cowatch_files = sorted(glob.glob(cowatch_file_patten))
    self._pairs = [_load_pairs(f) for f in self.cowatch_files]
    self.cowatch_num = int(sum(len(p) for p in self._pairs))   # `wc -l` over the files (inputs.py:79-86)
    if feature_file is not None:
      FEATURES = read_features_npy(feature_file)
    self.wait_times = wait_times
    self.seed, self.rank, self.world = seed, rank, world
    self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device()) \
        if torch.cuda.is_available() else None
    _DEVICE_FEATURES = None
    self._iter = None
    self.device_reader = bool(device_reader)
    self._dev_pairs, self._dev_iter = {}, None
    logging.info("MPTripletPipe __init__ cowatch_files: %s", self.cowatch_files)

  # the device copy of FEATURES is created lazily so CPU-only tests of the reader logic need no GPU
  def device_features(self):
    global _DEVICE_FEATURES
    if _DEVICE_FEATURES is None:
      if self.device is None:
        raise RuntimeError("MPTripletPipe.get_batch gathers on the GPU; no CUDA device is visible")
      _DEVICE_FEATURES = torch.as_tensor(np.ascontiguousarray(FEATURES, dtype=np.float32)).to(self.device)
    return _DEVICE_FEATURES

  def create_pipe(self, num_epochs, batch_size, queue_length=2 ** 14):
    self.batch_size, self.num_epochs = batch_size, num_epochs
    self._iter = self._index_batches()
    self._dev_iter = self._position_batches()

  def _file_batches(self, pairs, rng, num_guid):
    """One worker of the reference: stream of [B,3] batches over `num_epochs` passes of one file."""
    B, n = self.batch_size, len(pairs)
    if n == 0:
      return
    total = n * self.num_epochs
    for start in range(0, total - B + 1, B):
      rows = np.arange(start, start + B) % n
      ap = pairs[rows]
      neg = rng.randint(0, num_guid, size=B)
      bad = (neg == ap[:, 0]) | (neg == ap[:, 1])
      while bad.any():
        neg[bad] = rng.randint(0, num_guid, size=int(bad.sum()))
        bad = (neg == ap[:, 0]) | (neg == ap[:, 1])
      yield np.concatenate([ap, neg[:, None]], axis=1)

  def _shard(self, ordered):
    """Rank r takes every world-th item of the round-robin order, in whole ROUNDS of `world` items: when fewer than
    `world` items remain they are dropped on every rank, so all ranks run the same number of steps (each step ends in a
    collective; a rank with one batch more than its peers would wait in the all-reduce forever)."""
    round_ = []
    for item in ordered:
      round_.append(item)
      if len(round_) == self.world:
        yield round_[self.rank]
        round_ = []

  @staticmethod
  def _round_robin(streams):
    while streams:
      alive = []
      for s in streams:
        item = next(s, None)
        if item is None:
          continue
        alive.append(s)
        yield item
      streams = alive

  def _index_batches(self):
    num_guid = len(FEATURES)
    streams = [self._file_batches(p, np.random.RandomState(self.seed + 7919 * i), num_guid)
               for i, p in enumerate(self._pairs)]
    return self._shard(self._round_robin(streams))

  def _position_batches(self):
    """(file, start position) of every batch, in the order `_index_batches` serves them (round-robin over the files, a
    file's stream is `num_epochs` passes back to back without its trailing partial batch, rank r takes every world-th of
    each complete round of `world` batches)."""
    B = self.batch_size
    def cursor(i, n):
      return ((i, start) for start in range(0, n * self.num_epochs - B + 1, B))
    cursors = [cursor(i, len(p)) for i, p in enumerate(self._pairs) if len(p)]
    return self._shard(self._round_robin(cursors))

  def get_batch_indices_device(self, out=None):
    """Next [B,3] int64 triplets as a DEVICE tensor produced by cdml_sample_triplets, or None at end of data."""
    if self._dev_iter is None:
      raise RuntimeError("call create_pipe() first")
    nxt = next(self._dev_iter, None)
    if nxt is None:
      return None
    i, start = nxt
    pairs = self._dev_pairs.get(i)
    if pairs is None:
      if self.device is None:
        raise RuntimeError("the device reader needs a CUDA device")
      pairs = self._dev_pairs[i] = torch.as_tensor(np.ascontiguousarray(self._pairs[i])).to(self.device)
    return ops.sample_triplets(pairs, start, self.batch_size, len(FEATURES), self.seed + 7919 * i, out=out)

  def get_batch_indices(self):
    """Next [B,3] int64 (anchor, positive, negative) guid-index triplets as a host array, or None at end of data."""
    if self._iter is None:
      raise RuntimeError("call create_pipe() first")
    return next(self._iter, None)

  def get_batch(self):
    """3-D float32 array [B,3,F] of gathered features, or None when the data is exhausted (inputs.py:144-166)."""
    idx = self.get_batch_indices()
    if idx is None:
      return None
    table = self.device_features()
    out = ops.gather_rows(table, torch.as_tensor(idx).to(table.device))
    return out.view(idx.shape[0], 3, table.shape[1]).cpu().numpy()

  def __del__(self):
    self._iter = self._dev_iter = None


class TripletPipe(BasePipe):
  """In-memory pipe over a ready-made [N,3] index array (inputs.py:32-59, the tf.data legacy reader): repeat,
  batch, then shuffle whole batches inside a `buffer_size` window."""

  def __init__(self, triplets):
    self.triplets = np.asarray(triplets)

  def create_pipe(self, batch_size=10, num_epochs=None, num_readers=1, buffer_size=1000, seed=0):
    rng = np.random.RandomState(seed)

    def gen():
      epoch, buf = 0, []
      carry = np.zeros((0,) + self.triplets.shape[1:], self.triplets.dtype)
      while num_epochs is None or epoch < num_epochs:
        data = np.concatenate([carry, self.triplets]) if len(carry) else self.triplets
        full = len(data) // batch_size * batch_size
        for s in range(0, full, batch_size):
          buf.append(data[s:s + batch_size])
          if len(buf) >= buffer_size:
            yield buf.pop(rng.randint(len(buf)))
        carry = data[full:]
        epoch += 1
      if len(carry):
        buf.append(carry)
      while buf:
        yield buf.pop(rng.randint(len(buf)))
    return gen()
