"""Embedding inference -- host-side mirror of the reference's predict.py: `Prediction` (predict.py:45-96) and `main`
(predict.py:116-157) with the same names, flags and output files (output.npy, features.npy, decode_map.json).
The forward pass runs on `engine.TowerEngine` (normalise+cast kernel -> tcgen05 GEMMs with fused epilogues)."""
import json
import logging
import os
import time
import traceback

import numpy as np
import torch
from absl import app, flags

from . import models
from .engine import TowerEngine
from .online_data import read_features_txt
from .utils import get_latest_folder

FLAGS = flags.FLAGS
if "ckpt_dir" not in FLAGS:
  flags.DEFINE_string("ckpt_dir", "", "directory holding the checkpoint")
  flags.DEFINE_string("model_dir", "serving_dir/models", "serving model root; sub-folders are deployed checkpoints")
  flags.DEFINE_string("feature_file", "serving_dir/dataset/features", "feature text file to embed (guid#f1,f2,...)")
  flags.DEFINE_string("output_dir", "serving_dir/predict_result", "where output.npy / features.npy / decode_map.json go")
  flags.DEFINE_integer("pred_batch_size", 100000, "rows per forward batch")


def latest_checkpoint(ckpt_dir):
  """tf.train.latest_checkpoint: prefix named by the `checkpoint` index file, or None."""
  index = os.path.join(ckpt_dir, "checkpoint")
  if not os.path.exists(index):
    return None
  with open(index) as f:
    for line in f:
      if line.startswith("model_checkpoint_path:"):
        return os.path.join(ckpt_dir, line.split(":", 1)[1].strip().strip('"'))
  return None


def load_engine(ckpt, device=None, process_group=None):
  """Rebuild a TowerEngine from `<ckpt>.npz` (the counterpart of import_meta_graph + restore, predict.py:53-58) with the
  hyper-parameters saved next to the tensors: a bf16-trained tower is served in bf16, LARS / Momentum slots stay what
  they are.  Checkpoints without `hyper` (written before it existed) load with the defaults: fp16, Adam."""
  path = ckpt + ".npz"
  if not os.path.exists(path):
    raise IOError("Prediction __init__ Cannot find %s" % ckpt)
  z = np.load(path, allow_pickle=False)
  hyper = json.loads(str(z["hyper"])) if "hyper" in z.files else {}
  kw = {k: hyper[k] for k in ("dtype16", "alpha", "optimizer", "margin", "base_lr", "lr_decay_steps", "lr_decay", "beta1",
                              "beta2", "eps", "clip_norm", "momentum", "lars_weight_decay", "lars_eeta") if k in hyper}
  if "wd_reg" in hyper:
    kw["reg_penalty"], kw["l2_penalty"] = 1.0, hyper["wd_reg"]
  if "spec" in z.files:       # a fusion tower (fusion.GraphEngine)
    from .fusion import GraphEngine
    if "loss_scale" in hyper:
      kw["loss_scale"] = hyper["loss_scale"]
    eng = GraphEngine(json.loads(str(z["spec"])), feature_size=int(z["dims"][0]), device=device,
                      process_group=process_group, **kw)
  else:
    eng = TowerEngine([int(d) for d in z["dims"]], device=device, process_group=process_group, **kw)
    eng.loss_scale = float(hyper.get("loss_scale", eng.loss_scale))
  eng.load_state_dict({"w": z["w"], "m": z["m"], "v": z["v"], "step": int(z["step"]), "hyper": hyper or None})
  return eng


class Prediction():
  def __init__(self, sess=None, ckpt=None, config=None):
    """sess: a live TowerEngine (the trainer passes its own, train.py:282); otherwise `ckpt` is loaded."""
    self.sess, self.ckpt, self.config = sess, ckpt, config
    if not sess:
      logging.info(str(self.ckpt))
      self.sess = load_engine(self.ckpt)
    self.engine = self.sess.engine if hasattr(self.sess, "engine") else self.sess
    logging.info("Prediction __init__ Load predictor!")

  def predict(self, input_batch_np, as_device=False):
    """[n,F] float features -> [n,256] float32 embeddings (predict.py:67-69)."""
    x = input_batch_np if torch.is_tensor(input_batch_np) else torch.as_tensor(np.ascontiguousarray(input_batch_np, np.float32))
    e = self.engine.embed(x.to(self.engine.device, dtype=torch.float32))
    return e if as_device else e.cpu().numpy()

  def run_features(self, features, batch_size, output_dir='', suffix='', as_device=False, process_group=None,
                   gather=True):
    """Embed all rows in batches of `batch_size` (+ tail batch), optionally save output{suffix}.npy (predict.py:71-96).

    With a `process_group` (one process per GPU) the rows are split into `world` contiguous slices: rank r embeds rows
    [r*n//world, (r+1)*n//world) -- batches are independent (predict.py:74-84), so there is no data-path collective.
    `gather=True` joins the slices with one all-gather (every rank returns all n rows; rank 0 alone writes the file);
    `gather=False` returns this rank's slice only -- what a row-sharded KNN index consumes in-process."""
    n = features.shape[0]
    world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
    rank = torch.distributed.get_rank(process_group) if process_group is not None else 0
    lo, hi = rank * n // world, (rank + 1) * n // world
    D = self.engine.dims[-1]
    out = torch.empty((hi - lo, D), dtype=torch.float32, device=self.engine.device)
    for s in range(lo, hi, batch_size):
      e = min(s + batch_size, hi)
      out[s - lo:e - lo].copy_(self.predict(features[s:e], as_device=True))
    if world > 1 and gather:
      rows = (n + world - 1) // world                      # slices differ by at most one row: pad to the longest
      mine = torch.zeros((rows, D), dtype=torch.float32, device=out.device)
      mine[:hi - lo].copy_(out)
      parts = torch.empty((world * rows, D), dtype=torch.float32, device=out.device)
      torch.distributed.all_gather_into_tensor(parts, mine, group=process_group)
      parts = parts.view(world, rows, D)
      out = torch.cat([parts[r, :(r + 1) * n // world - r * n // world] for r in range(world)], dim=0)
    if as_device and not output_dir:
      return out
    output_np = out.cpu().numpy()
    if output_dir and (rank == 0 or not gather):
      try:
        shard = "" if (world == 1 or gather) else ".rank%d" % rank
        save_dir = os.path.join(output_dir, "output" + suffix + shard + ".npy")
        np.save(save_dir, output_np)
        logging.info("Saved to " + save_dir)
      except Exception as e:
        logging.error("Prediction.run_features save error" + str(e))
    return out if as_device else output_np


def _deployed_checkpoint(model_dir):
  """Newest sub-folder with a checkpoint AND transend.signal, else the second newest (predict.py:119-132)."""
  tried = []
  for nst in (1, 2):
    ckpt_dir = get_latest_folder(model_dir, nst_latest=nst)
    ckpt = latest_checkpoint(ckpt_dir)
    signal = os.path.join(ckpt_dir, "transend.signal")
    if ckpt is not None and os.path.exists(ckpt + ".npz") and os.path.exists(signal):
      return ckpt
    tried += [str(ckpt), signal]
    logging.warning("Prediction main Cannot find %s or %s", ckpt, signal)
  raise IOError("Prediction main Cannot find %s." % ", ".join(tried))


def main(args):
  try:
    if not FLAGS.ckpt_dir:
      ckpt = _deployed_checkpoint(FLAGS.model_dir)
    else:
      ckpt = latest_checkpoint(FLAGS.ckpt_dir)
      if ckpt is None or not os.path.exists(ckpt + ".npz"):
        raise IOError("Prediction main Cannot find %s" % ckpt)
    logging.info("ckpt is " + ckpt)
    begin = time.time()
    pg = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:          # torchrun: one process per GPU, rows split over the ranks
      torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
      torch.distributed.init_process_group("nccl")
      pg = torch.distributed.group.WORLD
    predictor = Prediction(ckpt=ckpt)
    features, _, decode_map = read_features_txt(FLAGS.feature_file, predictor.engine.F)
    logging.info("predict read_features_txt success! Cost: %fs", time.time() - begin)
    os.makedirs(FLAGS.output_dir, exist_ok=True)
    predictor.run_features(features=features, batch_size=FLAGS.pred_batch_size, output_dir=FLAGS.output_dir,
                           process_group=pg)
    if pg is None or torch.distributed.get_rank(pg) == 0:
      np.save(os.path.join(FLAGS.output_dir, "features.npy"), features)
      with open(os.path.join(FLAGS.output_dir, "decode_map.json"), "w") as f:
        json.dump(decode_map, f, ensure_ascii=False)
  except Exception:
    logging.error(traceback.format_exc())
    raise


if __name__ == "__main__":
  app.run(main)
