"""One shard of a W-way row-sharded search emulated on one GPU, per-kernel CUDA times:
   python tools/knn_shard_diag.py [N] [nq] [k] [W]   (bounds -> bounded search on N/W rows; merge of [W, nq/W, k])"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import cdml_b200  # noqa: F401
from cdml_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(4)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
W = int(sys.argv[4]) if len(sys.argv) > 4 else 8
X = torch.nn.functional.normalize(torch.randn((N, 256), generator=g, device=dev), dim=1)
Q = X[:nq].clone()
index = ops.FlatIndex(X[:N // W].contiguous(), "L2")


rec = torch.empty((nq, k), dtype=torch.int64, device=dev)


def one():
  """What one rank of faiss_knn.sharded_search runs (the all-reduces / the all-to-all replaced by nothing: one shard's
  own values stand in for the agreed ones, so the pruning is LOOSER than in the real W-way run)."""
  for s in range(0, nq, index.CHUNK):
    q = Q[s:s + index.CHUNK]
    pair = index.shard_bounds(q, k, -(-k // W))
    nom = index.shard_collect(q, k, -(-k // W), pair)
    index.shard_refine(q, k, nom, rec[s:s + index.CHUNK])
  return ops.knn_merge_packed(rec.view(W, nq // W, k), "L2")   # stand-in for the all-to-all output (same sizes, sorted lists)


for _ in range(2):
  one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
one()
e1.record()
torch.cuda.synchronize()
print("shard of %d rows (W=%d), %d queries, k=%d: %.3f ms wall on the stream; stats %s" % (N // W, W, nq, k, e0.elapsed_time(e1), index.last_stats()))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  one()
  torch.cuda.synchronize()
tot = 0.0
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
  if e.device_time_total > 0:
    tot += e.device_time_total
    print("%-90s n=%d %.3f ms" % (e.key[:90], e.count, e.device_time_total / 1e3), flush=True)
print("kernel total %.3f ms" % (tot / 1e3))
