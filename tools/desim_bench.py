"""De-similarity filter alone (cdml_desim): timing of the two kernels by CUDA events, for ncu captures.
  python tools/desim_bench.py [n] [ke] [kf]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft  # noqa: E402

graft.build()
from cdml_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
ke = int(sys.argv[2]) if len(sys.argv) > 2 else 81
kf = int(sys.argv[3]) if len(sys.argv) > 3 else 26
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev)
gen.manual_seed(6)
eI = torch.randint(0, n, (n, ke), generator=gen, device=dev, dtype=torch.int64)
fI = torch.randint(0, n, (n, kf), generator=gen, device=dev, dtype=torch.int64)
eI[:, 0] = fI[:, 0] = torch.arange(n, device=dev)
fD = torch.sort(torch.rand((n, kf), generator=gen, device=dev) * 2.0, dim=1).values
out = torch.empty_like(eI)
for _ in range(2):
  ops.desim(eI, fI, fD, 1.4, 31, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
  ops.desim(eI, fI, fD, 1.4, 31, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("desim n=%d ke=%d kf=%d: %.3f ms per call, %.1f M rows/s, dropped %.4f" %
      (n, ke, kf, ms, n / ms / 1e3, float((out < 0).float().mean().item())))
