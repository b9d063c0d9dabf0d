"""Per-kernel CUDA time of one KNN search (torch profiler / CUPTI): python tools/knn_diag.py [N] [nq] [k]"""
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
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
X = torch.nn.functional.normalize(torch.randn((N, 256), generator=g, device=dev), dim=1)
index = ops.FlatIndex(X, "L2")
for _ in range(2):
  index.search(X[:nq], k)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  index.search(X[:nq], k)
  torch.cuda.synchronize()
tot = 0.0
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
  if e.device_time_total > 0:
    tot += e.device_time_total
    print("%-90s n=%d %.3f ms" % (e.key[:90], e.count, e.device_time_total / 1e3), flush=True)
print("total %.3f ms for %d queries -> %.0f queries/s; stats %s" % (tot / 1e3, nq, nq / tot * 1e6, index.last_stats()))
