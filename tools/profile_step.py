"""Per-op CUDA-event breakdown of one training step (bench workload).  python tools/profile_step.py [--batch B] [--no-mine]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cdml_b200  # noqa: F401
from cdml_b200 import engine, ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--guids", type=int, default=200000)
ap.add_argument("--no-mine", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
DIMS = [1500, 5000, 256]
eng = engine.TowerEngine(DIMS, device=dev)
B, G = args.batch, args.guids
table16 = torch.empty((G, eng.F_pad), dtype=torch.float16, device=dev)
for s in range(0, G, 65536):
  rows = min(65536, G - s)
  eng.prepare_table(torch.rand((rows, 1500), device=dev), out=table16[s:s + rows])
idx = torch.randint(0, G, (B, 3), device=dev)
records = []
names = ["gather_rows", "gemm16", "sum_partials", "colsum16", "triplet_hinge", "adam_prepare", "adam_apply", "mine_semihard"]
orig = {n: getattr(ops, n) for n in names}


def wrap(name):
  f = orig[name]

  def g(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = f(*a, **k)
    e1.record()
    tag = name
    if name == "gemm16":
      tag = "gemm16 M=%d N=%d K=%d epi=%d" % (a[2], a[3], a[4], a[7])
    records.append((tag, e0, e1))
    return r
  return g


for _ in range(3):
  eng.train_step_indices(table16, idx, mine=not args.no_mine)
torch.cuda.synchronize()
for n in names:
  setattr(ops, n, wrap(n))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(3):
  eng.train_step_indices(table16, idx, mine=not args.no_mine)
t1.record()
torch.cuda.synchronize()
agg = {}
for tag, e0, e1 in records:
  agg.setdefault(tag, []).append(e0.elapsed_time(e1))
total = t0.elapsed_time(t1) / 3
print("step %.3f ms (B=%d, mine=%s)" % (total, B, not args.no_mine))
for tag, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
  print("  %-48s calls/step=%d  %.3f ms/step" % (tag, len(v) // 3, sum(v) / 3))

# ---- e2e view: host wall time per step when the host reads the loss after every step
import time
for n in names:
  setattr(ops, n, orig[n])
host = torch.empty(4).pin_memory()
torch.cuda.synchronize()
walls, launches = [], []
for _ in range(8):
  t_a = time.perf_counter()
  st = eng.train_step_indices(table16, idx, mine=not args.no_mine)
  t_b = time.perf_counter()
  host.copy_(st, non_blocking=True)
  torch.cuda.current_stream().synchronize()
  t_c = time.perf_counter()
  walls.append((t_c - t_a) * 1e3)
  launches.append((t_b - t_a) * 1e3)
print("e2e per-step wall %.3f ms (host launch phase %.3f ms) over %s" % (sorted(walls)[len(walls) // 2], sorted(launches)[len(launches) // 2], ["%.2f" % w for w in walls]))
