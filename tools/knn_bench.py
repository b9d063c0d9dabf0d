"""KNN-only micro benchmark: python tools/knn_bench.py [N] [nq] [k]   (one search after one warm-up search)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cdml_b200  # noqa: F401
from cdml_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(4)
X = torch.nn.functional.normalize(torch.randn((N, 256), generator=g, device=dev), dim=1)
index = ops.FlatIndex(X, "L2")
index.search(X[:1024], k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
D, I = index.search(X[:nq], k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("knn N=%d nq=%d k=%d: %.2f ms  %.0f queries/s  %.1f TFLOP/s  stats=%s" % (N, nq, k, ms, nq / ms * 1e3, 2.0 * nq * N * 256 / ms / 1e9, index.last_stats()))
if os.environ.get("KNN_PROF"):
  from torch.profiler import ProfilerActivity, profile
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    index.search(X[:nq], k)
    torch.cuda.synchronize()
  print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
