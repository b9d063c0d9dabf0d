"""KNN on tightly clustered embeddings (what a trained tower produces): exercises the overflow -> exact-fallback path at
scale and checks the result against a torch fp32 brute force on a sample of queries.
   python tools/knn_clustered_check.py [N] [clusters] [sigma] [nq]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cdml_b200  # noqa: F401
from cdml_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 65536
k = 100
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(5)
centres = torch.randn((C, 256), generator=g, device=dev)
cl = torch.randint(0, C, (N,), generator=g, device=dev)
X = torch.nn.functional.normalize(centres[cl] + sigma * torch.randn((N, 256), generator=g, device=dev), dim=1)
index = ops.FlatIndex(X, "L2")
index.search(X[:1024], k)
torch.cuda.synchronize()
t0 = time.time()
D, I = index.search(X[:nq], k)
torch.cuda.synchronize()
dt = time.time() - t0
st = index.last_stats()
bad = ((I < -1) | (I >= N)).sum().item()
print("N=%d clusters=%d sigma=%g nq=%d: %.1f ms, stats %s, out-of-range ids %d, self-first %.4f, same-cluster top-10 %.4f" %
      (N, C, sigma, nq, dt * 1e3, st, bad, (I[:, 0] == torch.arange(nq, device=dev)).float().mean().item(),
       (cl[I[:, :10].clamp(0, N - 1)] == cl[:nq, None]).float().mean().item()))
if bad:
  w = torch.nonzero((I < -1) | (I >= N))
  print("bad entries (query, column, id):", [(int(a), int(b), int(I[a, b])) for a, b in w[:12]])
  print("row of first bad query:", I[w[0, 0]].tolist()[:12], D[w[0, 0]].tolist()[:6])
pick = torch.randperm(nq, generator=g, device=dev)[:256]
Q = X[pick]
dist = (Q * Q).sum(1, keepdim=True) + (X * X).sum(1)[None, :] - 2.0 * Q @ X.T
Dw, Iw = torch.topk(dist, k, dim=1, largest=False)
Dg = D[pick]
print("max |D - brute force D| over 256 sampled queries: %.3e ; ids equal %.4f (ties aside)" %
      ((Dg - Dw.clamp(min=0)).abs().max().item(), (I[pick] == Iw).float().mean().item()))
if os.environ.get("KNN_PROF"):
  from torch.profiler import ProfilerActivity, profile
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    index.search(X[:nq], k)
    torch.cuda.synchronize()
  for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print("%-80s n=%d %.3f ms" % (e.key[:80], e.count, e.device_time_total / 1e3))
