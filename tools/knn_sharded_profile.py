"""Where a row-sharded search spends its time (run under torchrun): CUDA-event time of each phase of
faiss_knn.sharded_search on rank 0 and the CUDA-side kernel / NCCL totals from the torch profiler.
  torchrun --nproc-per-node W tools/knn_sharded_profile.py [N] [nq] [k]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import cdml_b200  # noqa: F401
from cdml_b200 import faiss_knn, ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda:%d" % local)
dist.init_process_group("nccl", device_id=dev)
pg = dist.group.WORLD
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
g = torch.Generator(device=dev)
g.manual_seed(4)
X = torch.nn.functional.normalize(torch.randn((N, 256), generator=g, device=dev), dim=1)
lo, hi = rank * N // world, (rank + 1) * N // world
Q = X[:nq].clone()
index = ops.FlatIndex(X[lo:hi].contiguous(), "L2")
del X
for _ in range(2):
  faiss_knn.sharded_search(index, Q, k, lo, "L2", pg)
torch.cuda.synchronize()
dist.barrier()


def ev():
  e = torch.cuda.Event(enable_timing=True)
  e.record()
  return e


import time
marks = []
k_part = -(-k // world)
t_host0 = time.perf_counter()
marks.append(("start", ev()))
pair = index.shard_bounds(Q, k, k_part); marks.append(("shard_bounds", ev()))
dist.all_reduce(pair, op=dist.ReduceOp.MAX); marks.append(("all_reduce pair", ev()))
nom = index.shard_collect(Q, k, k_part, pair); marks.append(("shard_collect", ev()))
dist.all_reduce(nom, op=dist.ReduceOp.MAX); marks.append(("all_reduce nom", ev()))
rec = torch.empty((nq, k), dtype=torch.int64, device=dev)
flag = torch.zeros((1,), dtype=torch.int32, device=dev)
index.shard_refine(Q, k, nom, rec, id_offset=lo, overflow_flag=flag); marks.append(("shard_refine (deferred check)", ev()))
got = torch.empty_like(rec)
dist.all_to_all_single(got, rec); marks.append(("all_to_all records", ev()))
mine = ops.knn_merge_packed(got.view(world, nq // world, k), "L2", as_records=True); marks.append(("merge", ev()))
allr = torch.empty((nq, k), dtype=torch.int64, device=dev)
dist.all_gather_into_tensor(allr, mine); marks.append(("all_gather", ev()))
D, I = ops.knn_unpack_records(allr, "L2"); marks.append(("unpack", ev()))
dist.all_reduce(flag, op=dist.ReduceOp.MAX); marks.append(("all_reduce overflow flag", ev()))
torch.cuda.synchronize()
t_host = (time.perf_counter() - t_host0) * 1e3
if rank == 0:
  print("world %d, N %d, nq %d, k %d: %.3f ms on the stream, %.3f ms host wall" % (world, N, nq, k, marks[0][1].elapsed_time(marks[-1][1]), t_host))
  for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
    print("  %-34s %.3f ms" % (name, a.elapsed_time(b)))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  faiss_knn.sharded_search(index, Q, k, lo, "L2", pg)
  torch.cuda.synchronize()
if rank == 0:
  tot = 0.0
  for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]:
    if e.device_time_total > 0:
      tot += e.device_time_total
      print("  %-86s n=%d %.3f ms" % (e.key[:86], e.count, e.device_time_total / 1e3), flush=True)
  print("  device total (top 14) %.3f ms" % (tot / 1e3))
dist.barrier()
dist.destroy_process_group()
