"""Raw-feature KNN (d > 256: the generic GEMM kernel, both operands streamed): python tools/knn_wide_bench.py [N] [nq] [d] [k]
(The CDML_KNN_WIDE_CHUNK sweep of profiles/r02_knn_wide_chunk_sweep.log used an experimental build; the committed library
ignores the variable and searches 32768 queries per pass.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cdml_b200  # noqa: F401
from cdml_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
d = int(sys.argv[3]) if len(sys.argv) > 3 else 1628
k = int(sys.argv[4]) if len(sys.argv) > 4 else 26
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(4)
X = torch.empty((N, d), dtype=torch.float32, device=dev)
for s in range(0, N, 100000):
  X[s:s + 100000] = torch.nn.functional.normalize(torch.rand((min(100000, N - s), d), generator=g, device=dev), dim=1)
index = ops.FlatIndex(X, "L2")
index.search(X[:2048], k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
D, I = index.search(X[:nq], k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("raw-feature knn N=%d nq=%d d=%d k=%d chunk=%s: %.2f ms  %.0f queries/s  %.1f TFLOP/s  self-first %.4f  stats=%s" % (
    N, nq, d, k, os.environ.get("CDML_KNN_WIDE_CHUNK", "32768 (library default)"), ms, nq / ms * 1e3, 2.0 * nq * N * d / ms / 1e9,
    float((I[:, 0] == torch.arange(nq, device=dev)).float().mean().item()), index.last_stats()))
