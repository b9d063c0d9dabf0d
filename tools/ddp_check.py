"""Multi-GPU protocol check (run under torchrun, NCCL):  data-parallel training == single-process training on the
concatenated batch (oracle), weights identical on all ranks; row-sharded KNN + all-gather + merge == flat oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import cdml_b200  # noqa: F401
from cdml_b200 import engine, faiss_knn
from oracle import cdml_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda:%d" % local)
dist.init_process_group("nccl", device_id=dev)
pg = dist.group.WORLD

G, F, Bl = 3000, 1500, 256
dims = [F, 5000, 256]
feats = O.synth_features(G, F, 0)
params = O.init_tower(dims, seed=2)
eng = engine.TowerEngine(dims, device=dev, base_lr=1e-3, margin=0.8, init_params=params, process_group=pg)
table16 = eng.prepare_table(torch.as_tensor(feats).to(dev))
tr = O.OracleTrainer(params, lr=1e-3, margin=0.8) if rank == 0 else None
losses = []
for step in range(4):
  trips = [O.synth_triplets(Bl, G, 100 + step * world + r) for r in range(world)]
  stats = eng.train_step_indices(table16, torch.as_tensor(trips[rank]).to(dev))
  local_loss = stats[0].clone()
  dist.all_reduce(local_loss)
  if rank == 0:
    x = O.flatten_triplets(O.gather_rows(feats, np.concatenate(trips)))
    lc, _, _ = tr.step(x)
    losses.append((float(local_loss.item()) / world, lc))
w = eng.w.clone()
wmax, wmin = w.clone(), w.clone()
dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
same = bool((wmax == wmin).all().item())
if rank == 0:
  rel = max(abs(a / b - 1) for a, b in losses)
  W1 = eng.get_params()[0][0]
  travel = np.linalg.norm((W1 - params[0][0]) - (tr.params[0][0] - params[0][0])) / np.linalg.norm(tr.params[0][0] - params[0][0])
  print("DDP world=%d: weights identical on all ranks: %s; loss curve max rel err vs global-batch oracle %.2e; W1 travel rel err %.3f"
        % (world, same, rel, travel), flush=True)
  assert same and rel < 1e-3 and travel < 0.2

emb = np.random.RandomState(6).standard_normal((20011, 256)).astype(np.float32)
D, I = faiss_knn.calc_knn(emb, nearest_num=51, process_group=pg)
if rank == 0:
  Dw, Iw = O.flat_knn(emb, k=51)
  frac = (I == Iw).mean()
  print("sharded KNN world=%d: ids equal %.5f, max |D-Dw| %.2e" % (world, frac, np.abs(D - Dw).max()), flush=True)
  assert frac > 0.999 and np.abs(D - Dw).max() < 2e-5
dist.barrier()
dist.destroy_process_group()
