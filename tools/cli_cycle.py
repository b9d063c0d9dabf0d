"""configs[4] through the CLIs, HDFS-free (cdml_run.sh:73-150): train.py -> (deploy: transend.signal) -> predict.py ->
faiss_knn.py --knn_mode strict, each as its own process (one per GPU under torchrun when --gpus > 1), on a synthetic data
set written to local disk in the reference's file formats.  Prints one JSON line with the stage wall times.

  python tools/cli_cycle.py [--gpus N] [--guids G] [--pairs P] [--batch B] [--epochs E] [--work DIR]

Sizes default to what a feature TEXT file allows (20 000 guids x 1500 floats = 0.27 GB of text for predict.py); the
full-size composition with device-generated features is tools/cycle.py."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "collaborative-deep-metric-learning_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--guids", type=int, default=20000)
ap.add_argument("--pairs", type=int, default=2000000)
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--epochs", type=int, default=1)
ap.add_argument("--files", type=int, default=4)
ap.add_argument("--work", default="")
ap.add_argument("--mine", action="store_true")
args = ap.parse_args()
work = args.work or tempfile.mkdtemp(prefix="cdml_cli_cycle_")
train_dir, ckpt_root, serving = os.path.join(work, "dataset"), os.path.join(work, "models"), os.path.join(work, "serving")
ckpt_dir = os.path.join(ckpt_root, "2019071001")
for d in (train_dir, ckpt_dir, serving, os.path.join(serving, "predict_result"), os.path.join(serving, "knn_result")):
  os.makedirs(d, exist_ok=True)
t = {}

# ---- stage 0: the files the reference's ETL leaves behind (online_data.py:256-295): features.npy, *.train, cowatches.eval/.test,
#      and the serving-side feature text file guid#f1,...,fF (online_data.py:66-77)
t0 = time.time()
rng = np.random.RandomState(0)
G, F = args.guids, 1500
clusters = 500
centres = rng.random_sample((clusters, F)).astype(np.float32)
cl = rng.randint(0, clusters, G)
feats = (centres[cl] + 0.25 * rng.random_sample((G, F))).astype(np.float32)
np.save(os.path.join(train_dir, "features.npy"), feats)
order = np.argsort(cl, kind="stable")
start = np.searchsorted(cl[order], np.arange(clusters + 1))
a = rng.randint(0, G, args.pairs)
span = np.maximum(start[cl[a] + 1] - start[cl[a]], 1)
p = order[start[cl[a]] + rng.randint(0, 1 << 30, args.pairs) % span]
per = args.pairs // args.files
for i in range(args.files):
  np.savetxt(os.path.join(train_dir, "cowatches_%d.train" % i), np.stack([a, p], 1)[i * per:(i + 1) * per], fmt="%d", delimiter=",")
for name in ("eval", "test"):
  ea = rng.randint(0, G, 2000)
  es = np.maximum(start[cl[ea] + 1] - start[cl[ea]], 1)
  ep = order[start[cl[ea]] + rng.randint(0, 1 << 30, 2000) % es]
  np.savetxt(os.path.join(train_dir, "cowatches." + name), np.stack([ea, ep], 1), fmt="%d", delimiter=",")
import io
with open(os.path.join(serving, "features"), "w") as f:
  for g0 in range(0, G, 2000):                      # C-level float formatting, the guid prefix added per line
    buf = io.StringIO()
    np.savetxt(buf, feats[g0:g0 + 2000], fmt="%.6f", delimiter=",")
    for j, line in enumerate(buf.getvalue().splitlines()):
      f.write("g%07d#%s\n" % (g0 + j, line))
t["write_dataset_s"] = time.time() - t0


def run(stage, module, flags):
  """One CLI as its own process (torchrun: one process per GPU)."""
  env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
  boot = "import cdml_b200, runpy, sys; sys.argv = sys.argv[1:]; runpy.run_module('cdml_b200.%s', run_name='__main__')" % module
  if args.gpus > 1:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
           "127.0.0.1", "--master-port", str(29540 + len(t)), "--no-python", sys.executable, "-c", boot, module] + flags
  else:
    cmd = [sys.executable, "-c", boot, module] + flags
  t0 = time.time()
  r = subprocess.run(cmd, env=env, capture_output=True, text=True)
  t[stage] = time.time() - t0
  if r.returncode != 0:
    sys.stderr.write(r.stdout[-3000:] + "\n" + r.stderr[-6000:])
    raise SystemExit("stage %s failed (rc %d)" % (stage, r.returncode))
  return r.stdout + r.stderr


out_train = run("train_py_s", "train", ["--train_dir", train_dir, "--checkpoint_dir", ckpt_dir, "--model", "VNet", "--batch_size", str(args.batch),
                                        "--num_epochs", str(args.epochs), "--learning_rate", "1e-4"] + (["--mine_semihard"] if args.mine else []))
open(os.path.join(ckpt_dir, "transend.signal"), "w").close()                   # the deploy convention of predict.py:119-132
out_pred = run("predict_py_s", "predict", ["--model_dir", ckpt_root, "--feature_file", os.path.join(serving, "features"),
                                          "--output_dir", os.path.join(serving, "predict_result"), "--pred_batch_size", "100000"])
pr = os.path.join(serving, "predict_result")
out_knn = run("faiss_knn_py_strict_s", "faiss_knn", ["--embedding_file", os.path.join(pr, "output.npy"), "--decode_map_file",
                                                    os.path.join(pr, "decode_map.json"), "--pred_feature_file", os.path.join(pr, "features.npy"),
                                                    "--knn_result", os.path.join(serving, "knn_result"), "--knn_mode", "strict",
                                                    "--nearest_num", "81", "--desim_nearest_num", "26"])
kr = os.path.join(serving, "knn_result")
files = sorted(f for f in os.listdir(kr) if f.startswith("strict_knn"))
emb = np.load(os.path.join(pr, "output.npy"))
I = np.load(os.path.join(kr, "strictI_desim.npy"))
lines = sum(1 for f in files for _ in open(os.path.join(kr, f)))
first = open(os.path.join(kr, files[0])).readline()[:120]
same = float((cl[np.where(I[:, 1:6] >= 0, I[:, 1:6], 0)] == cl[:, None])[I[:, 1:6] >= 0].mean())
summ = os.path.join(ckpt_dir, "summaries.jsonl")
hist = [json.loads(l) for l in open(summ)] if os.path.exists(summ) else []
print(json.dumps({"n_gpus": args.gpus, "guids": G, "cowatch_pairs": args.pairs, "batch": args.batch, "epochs": args.epochs, "mining": bool(args.mine),
                  "seconds": t, "embeddings": list(emb.shape), "unit_norm": bool(np.allclose(np.linalg.norm(emb, axis=1), 1, atol=1e-3)),
                  "strict_knn_files": len(files), "strict_knn_lines": lines, "first_line": first,
                  "kept_top5_same_cluster": same, "dropped_by_desim": float((I < 0).mean()),
                  "train_history_tail": hist[-2:], "checkpoints": sorted(os.listdir(ckpt_dir)), "work_dir": work}))
