"""CPU: host-side logic of the package -- C-ABI exports, reader semantics, tower plug-in interface, result writer,
checkpoint index, and the N>1 protocol (reader sharding, gradient all-reduce, sharded-KNN merge) over gloo."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import __graft_entry__ as graft
from conftest import GOLDEN, ROOT
from oracle import cdml_oracle as O


@pytest.fixture(scope="session")
def lib():
  graft.build()
  import cdml_b200  # noqa: F401
  from cdml_b200 import _lib
  return _lib


def test_library_exports_every_symbol_of_the_header(lib):
  header = open(os.path.join(ROOT, "include", "cdml.h")).read()
  declared = sorted(set(re.findall(r"\b(cdml_[a-z0-9_]+)\s*\(", header)))
  assert len(declared) >= 20
  handle = lib.load()
  for name in declared:
    assert hasattr(handle, name), "libcdml.so does not export %s" % name
  assert sorted(lib.exported_symbols()) == declared          # the ctypes prototypes cover the header exactly
  assert handle.cdml_version() >= 100


def test_no_cpu_fallback_context_creation_fails_loudly(lib):
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  with pytest.raises(lib.CdmlError):
    lib.context(0)
  from cdml_b200 import engine
  with pytest.raises(RuntimeError):
    engine.TowerEngine([8, 16, 8])


def test_product_never_imports_the_oracle():
  pkg = os.path.join(ROOT, "collaborative-deep-metric-learning_b200")
  for root, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(".py"):
        src = open(os.path.join(root, f)).read()
        assert not re.search(r"^\s*(from|import)\s+\.*oracle|cdml_oracle\s*(import|as)|import_module\(.*oracle", src, re.M), \
            "%s imports the oracle" % f


def test_model_plugin_interface(lib):
  from cdml_b200 import models, utils
  cls = utils.find_class_by_name("VNet", [models])
  out = cls().create_model(models.placeholder(1500), output_size=256)
  assert set(out) >= {"layer_1", "layer_2", "l2_norm"} and out["l2_norm"].name == "model_output"
  spec = models.compile_chain(out["l2_norm"])
  assert spec["dims"] == [1500, 5000, 256] and spec["bias_init"] == [0.0, 0.0] and spec["alpha"] == 0.2
  wide = models.compile_chain(models.WideNet().create_model(models.placeholder(2048))["l2_norm"])
  assert wide["dims"] == [2048, 2048, 2048, 2048, 256]
  with pytest.raises(StopIteration):
    utils.find_class_by_name("VedeNet", [models])            # the reference's broken default
  with pytest.raises(NotImplementedError):                   # a fusion tower is a graph, not a chain
    models.compile_chain(models.ResNet().create_model(models.placeholder(1628))["l2_norm"])
  with pytest.raises(NotImplementedError):                   # topology outside the hot path
    models.compile_chain(models.fully_connected(models.placeholder(8), 4))


def _write_dataset(tmp_path, G=50, F=12, n_lines=(23, 10)):
  feats = O.synth_features(G, F, seed=0)
  np.save(tmp_path / "features.npy", feats)
  rng = np.random.RandomState(3)
  for i, n in enumerate(n_lines):
    pairs = O.synth_pairs(n, G, seed=10 + i)
    with open(tmp_path / ("cowatches_%d.train" % i), "w") as f:
      for a, p in pairs:
        f.write("%d,%d\n" % (a, p))
  return feats


def test_reader_semantics_and_rank_sharding(lib, tmp_path):
  from cdml_b200 import inputs
  feats = _write_dataset(tmp_path)
  pipe = inputs.MPTripletPipe(str(tmp_path / "*.train"), str(tmp_path / "features.npy"), seed=5)
  assert pipe.cowatch_num == 33 and len(pipe.cowatch_files) == 2
  assert np.array_equal(inputs.FEATURES, feats)
  pipe.create_pipe(num_epochs=2, batch_size=8)
  batches = []
  while True:
    b = pipe.get_batch_indices()
    if b is None:
      break
    batches.append(b)
  # file 0: 46 lines over 2 epochs -> 5 full batches; file 1: 20 lines -> 2 full batches (partial tail dropped)
  assert len(batches) == 7 and all(b.shape == (8, 3) and b.dtype == np.int64 for b in batches)
  for b in batches:
    assert ((b[:, 2] != b[:, 0]) & (b[:, 2] != b[:, 1])).all() and (b >= 0).all() and (b < 50).all()
  pairs0 = np.loadtxt(tmp_path / "cowatches_0.train", delimiter=",", dtype=np.int64)
  first_file_batches = [b for b in batches if np.array_equal(b[0, :2], pairs0[0]) or True]
  assert np.array_equal(batches[0][:, :2], pairs0[:8])
  # epoch wrap inside a batch: rows 16..23 of file 0's stream are lines 16..22 then line 0 again
  stream0 = np.concatenate([pairs0, pairs0])
  got0 = np.concatenate([b[:, :2] for b in (batches[0], batches[2], batches[4], batches[5], batches[6])])
  assert np.array_equal(got0, stream0[:40])
  # two ranks see disjoint halves of the same deterministic order
  shards = []
  for r in range(2):
    p = inputs.MPTripletPipe(str(tmp_path / "*.train"), str(tmp_path / "features.npy"), seed=5, rank=r, world=2)
    p.create_pipe(2, 8)
    s = []
    while (b := p.get_batch_indices()) is not None:
      s.append(b)
    shards.append(s)
  # 7 batches over 2 ranks: three complete rounds; the odd batch is dropped on BOTH ranks -- a rank with one step more
  # than its peer would wait forever in that step's all-reduce (every step ends in a collective)
  assert len(shards[0]) == 3 and len(shards[1]) == 3
  assert all(np.array_equal(shards[0][i], batches[2 * i]) for i in range(3))
  assert all(np.array_equal(shards[1][i], batches[2 * i + 1]) for i in range(3))


def test_feature_text_reader_and_cowatch_loader_match_reference(lib, golden):
  from cdml_b200 import online_data
  fe, enc, dec = online_data.read_features_txt(os.path.join(GOLDEN, "features_small.txt"), 12)
  assert np.array_equal(fe, golden["read_features_txt"])
  maps = json.load(open(os.path.join(GOLDEN, "features_small_maps.json")))
  assert enc == maps["encode"] and {str(k): v for k, v in dec.items()} == maps["decode"]
  cw = online_data.load_cowatches(os.path.join(GOLDEN, "cowatches_small.eval"))
  assert np.array_equal(np.asarray(cw), golden["load_cowatches"])


def test_knn_result_writer_matches_reference_bytes(lib, golden, tmp_path):
  from cdml_b200 import faiss_knn
  dm, enc = faiss_knn.load_decode_map(os.path.join(GOLDEN, "knn_decode_map.json"))
  assert dm[3] == "guid03" and enc["guid03"] == 3
  faiss_knn.write_process(str(tmp_path), 0, 0, golden["knn_D"], golden["knn_I"], "knn_split", decode_map=dm)
  want = open(os.path.join(GOLDEN, "knn_split0.txt")).read()
  assert open(tmp_path / "knn_split0").read() == want
  # write_knn: split_num patches, last takes the remainder, query ids continue across patches
  D = np.tile(golden["knn_D"], (4, 1))[:11]
  I = np.tile(golden["knn_I"], (4, 1))[:11]
  dm2 = {i: "g%d" % i for i in range(11)}
  faiss_knn.write_knn(str(tmp_path / "out"), split_num=3, D=D, I=I, prefix="knn_result", decode_map=dm2)
  sizes = [len(open(tmp_path / "out" / ("knn_result%d" % j)).readlines()) for j in range(3)]
  assert sizes == [3, 3, 5]
  assert open(tmp_path / "out" / "knn_result2").readline().startswith("g6,")
  assert "".join(open(tmp_path / "out" / ("knn_result%d" % j)).read() for j in range(3)) == \
      "".join(O.format_knn_rows(0, D, I, dm2))


def test_checkpoint_index_and_deploy_fallback(lib, tmp_path):
  from cdml_b200 import predict, utils
  d1, d2 = tmp_path / "2019071001", tmp_path / "2019071002"
  for d, step in ((d1, 10), (d2, 20)):
    d.mkdir()
    (d / "checkpoint").write_text('model_checkpoint_path: "model.ckpt-%d"\n' % step)
    np.savez(d / ("model.ckpt-%d.npz" % step), dims=np.array([4, 8, 4]))
  (d1 / "transend.signal").write_text("")
  os.utime(d1, (1, 1))
  assert predict.latest_checkpoint(str(d2)).endswith("model.ckpt-20")
  assert utils.get_latest_folder(str(tmp_path), 1) == str(d2)
  # newest folder lacks transend.signal -> fall back to the second newest (predict.py:119-132)
  assert predict._deployed_checkpoint(str(tmp_path)).endswith(os.path.join("2019071001", "model.ckpt-10"))
  (d1 / "transend.signal").unlink()
  os.utime(d1, (1, 1))
  with pytest.raises(IOError):
    predict._deployed_checkpoint(str(tmp_path))


class _FakeEngine(object):
  """Host stand-in with the three members save_checkpoint touches (the real engine needs a GPU)."""

  def __init__(self, step, fail=False):
    self.step, self.fail = step, fail

  def state_dict(self):
    if self.fail:
      raise IOError("disk full")
    return {"dims": [4, 8, 4], "w": np.full(7, self.step, np.float32), "m": np.zeros(7, np.float32),
            "v": np.zeros(7, np.float32), "step": self.step, "hyper": {"dtype16": 1, "optimizer": "lars", "alpha": 0.2}}

  def get_params(self):
    return [(np.zeros((4, 8), np.float32), np.zeros(8, np.float32)), (np.zeros((8, 4), np.float32), np.zeros(4, np.float32))]


def test_checkpoint_write_is_atomic_and_carries_hyper(lib, tmp_path, monkeypatch):
  """Saver(max_to_keep=1) order: new file complete -> index switched -> old file removed.  A failure while writing leaves
  the previous checkpoint and an index that names it; the npz carries the hyper-parameters load_engine honours."""
  from cdml_b200 import predict, train
  train.save_checkpoint(_FakeEngine(10), str(tmp_path), 10, "VNet")
  assert sorted(os.listdir(tmp_path)) == ["checkpoint", "model.ckpt-10.npz"]
  real_savez = np.savez

  def broken_savez(f, **kw):
    f.write(b"partial")
    raise IOError("disk full")
  monkeypatch.setattr(np, "savez", broken_savez)
  with pytest.raises(IOError):
    train.save_checkpoint(_FakeEngine(20), str(tmp_path), 20, "VNet")
  monkeypatch.setattr(np, "savez", real_savez)
  assert sorted(os.listdir(tmp_path)) == ["checkpoint", "model.ckpt-10.npz"]          # nothing lost, no debris
  assert predict.latest_checkpoint(str(tmp_path)).endswith("model.ckpt-10")
  assert float(np.load(tmp_path / "model.ckpt-10.npz")["w"][0]) == 10.0
  train.save_checkpoint(_FakeEngine(30), str(tmp_path), 30, "VNet")
  assert sorted(os.listdir(tmp_path)) == ["checkpoint", "model.ckpt-30.npz"]
  z = np.load(tmp_path / "model.ckpt-30.npz")
  assert json.loads(str(z["hyper"])) == {"dtype16": 1, "optimizer": "lars", "alpha": 0.2}
  assert "fully_connected/weights" in z.files and "fully_connected_1/biases" in z.files


def test_engine_buffer_cache_never_evicts_what_a_captured_graph_points_at(lib):
  """TowerEngine._store_buffers (host policy, no GPU needed): the (rows, train=True) buffers of a captured CUDA graph are
  pinned for the life of the engine; everything else is an LRU of four entries -- eval batch, eval tail, predict batch,
  predict tail with different row counts must not push the training buffers out (round 1 cleared the whole cache)."""
  from cdml_b200.engine import TowerEngine
  eng = object.__new__(TowerEngine)
  eng._bufs, eng._pinned_bufs = {}, set()
  eng._store_buffers((3 * 64, True), "train")
  eng._pinned_bufs.add((3 * 64, True))                      # what capture_step does
  for rows in (100, 7, 250, 50, 1000, 33, 10000):
    eng._store_buffers((rows, False), "infer%d" % rows)
  assert eng._bufs[(3 * 64, True)] == "train"
  loose = [k for k in eng._bufs if k not in eng._pinned_bufs]
  assert loose == [(50, False), (1000, False), (33, False), (10000, False)]       # the four most recent, oldest first
  eng._store_buffers((3 * 32, True), "train-unpinned")      # a training shape without a graph is evictable like the rest
  assert (50, False) not in eng._bufs and (3 * 64, True) in eng._bufs


def test_imitation_data_matches_reference_stream(lib, golden):
  from cdml_b200 import imitation_data
  np.random.seed(1234)
  assert np.array_equal(imitation_data.gen_features(64, 12), golden["gen_features_seed1234"])
  np.random.seed(7)
  assert np.array_equal(imitation_data.gen_triplets(5, 4), golden["gen_triplets_seed7"])
  assert imitation_data.gen_triplets(100, 256).shape == (100, 3, 256)       # tests/test_imitation_data.py:39-41


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import cdml_oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), 2
# --- data-parallel step: local sum-gradients, all-reduce(sum), scale 1/(B_local*world) == full-batch mean gradient
feats = O.synth_features(200, 24, 0); trip = O.synth_triplets(16, 200, 1)
params = O.init_tower([24, 32, 16], seed=2, dtype=np.float64)
x = O.flatten_triplets(O.gather_rows(feats, trip)).astype(np.float64)
full = O.OracleTrainer(params).loss_and_grads(x)[2]
Bl = 8
xl = x[rank * 3 * Bl:(rank + 1) * 3 * Bl]
local = O.OracleTrainer(params).loss_and_grads(xl)[2]          # mean over B_local
flat = torch.tensor(np.concatenate([np.concatenate([gW.ravel(), gb]) for gW, gb in local]) * Bl)   # SUM gradients
dist.all_reduce(flat)
flat = flat.numpy() / (Bl * world)
want = np.concatenate([np.concatenate([gW.ravel(), gb]) for gW, gb in full])
assert np.allclose(flat, want, atol=1e-12), np.abs(flat - want).max()
# --- row-sharded KNN through the PRODUCT's sharded_search (bound exchange -> bounded search -> all-to-all -> merge ->
#     all-gather), with a CPU stand-in for the device index / merge kernel built on the oracle
from cdml_b200 import faiss_knn
class ShardIndex:
  def __init__(self, xb): self.xb = xb
  def _scores(self, xq): return xq.numpy() @ self.xb.T - 0.5 * (self.xb ** 2).sum(1)[None, :]
  def bounds(self, xq, k_full, k_part):
    s = np.sort(self._scores(xq), axis=1)[:, ::-1]
    pick = lambda kk: torch.tensor(np.ascontiguousarray(s[:, kk - 1] if kk <= s.shape[1] else np.full(len(s), -np.inf), dtype=np.float32))
    return pick(k_full), pick(k_part)
  def search_bounded(self, xq, k, bf, bp, id_offset=0):
    s, thr = self._scores(xq), np.maximum(bf.numpy(), bp.numpy()) - 1e-5
    D = np.full((len(s), k), np.inf, np.float32); I = np.full((len(s), k), -1, np.int64)
    for i in range(len(s)):
      rows = np.nonzero(s[i] > thr[i])[0]
      if len(rows):
        d, ids = O.flat_knn(self.xb[rows], xq.numpy()[i:i + 1], k=min(k, len(rows)), l2_norm=False)
        D[i, :d.shape[1]], I[i, :d.shape[1]] = d[0], rows[ids[0]] + id_offset
    return torch.tensor(D), torch.tensor(I)
def cpu_merge(Dg, Ig, metric):
  Dm, Im = O.knn_merge([d.numpy() for d in Dg], [i.numpy() for i in Ig], Dg.shape[2])
  return torch.tensor(Dm), torch.tensor(Im)
X = O.knn_normalize(np.random.RandomState(4).standard_normal((301, 16)).astype(np.float32))
lo, hi = rank * 301 // world, (rank + 1) * 301 // world
Dw, Iw = O.flat_knn(X, k=7, l2_norm=False)
for nq in (300, 301):      # divisible by the world size (all-to-all path) and not (all-gather fallback)
  Dm, Im = faiss_knn.sharded_search(ShardIndex(X[lo:hi]), torch.tensor(X[:nq]), 7, lo, "L2", dist.group.WORLD, merge_fn=cpu_merge)
  assert np.array_equal(Im.numpy(), Iw[:nq]) and np.allclose(Dm.numpy(), Dw[:nq], atol=1e-6), nq
  assert (Im.numpy() >= 0).all()
# --- the same search through the record protocol (shard_bounds -> all-reduce -> shard_collect -> all-reduce -> shard_refine
#     -> ONE all-to-all of 64-bit records -> packed merge), chunked, with padding of the query count and with gather=False
class RecordShard(ShardIndex):
  CHUNK = 128
  def shard_bounds(self, xq, k, k_part):
    bf, bp = self.bounds(xq, k, k_part)
    return torch.stack([bf, -bp])
  def shard_collect(self, xq, k, k_part, pair):
    s = self._scores(xq)
    thr = np.maximum(pair[0].numpy(), -pair[1].numpy()) - 1e-5
    self.nom = [np.nonzero(s[i] > thr[i])[0] for i in range(len(s))]
    pick = lambda v, kk: np.sort(v)[::-1][kk - 1] if len(v) >= kk else -np.inf
    a = np.array([pick(s[i][self.nom[i]], k) for i in range(len(s))], np.float32)
    b = np.array([pick(s[i][self.nom[i]], k_part) for i in range(len(s))], np.float32)
    return torch.tensor(np.stack([a, -b]))
  def shard_refine(self, xq, k, nom, rec, id_offset=0):
    s = self._scores(xq)
    T = np.maximum(nom[0].numpy(), -nom[1].numpy()) - 1e-5
    D = np.full((len(s), k), np.inf, np.float32); I = np.full((len(s), k), -1, np.int64)
    self.reranked = 0
    for i in range(len(s)):
      rows = self.nom[i][s[i][self.nom[i]] >= T[i]]
      self.reranked += len(rows)
      if len(rows):
        d, ids = O.flat_knn(self.xb[rows], xq.numpy()[i:i + 1], k=min(k, len(rows)), l2_norm=False)
        D[i, :d.shape[1]], I[i, :d.shape[1]] = d[0], rows[ids[0]] + id_offset
    rec.copy_(torch.tensor(O.knn_pack_records(D, I).view(np.int64)))
def cpu_merge_packed(rec, metric):
  Dm, Im = O.knn_merge_records(rec.numpy().view(np.uint64), rec.shape[2], metric)
  return torch.tensor(Dm), torch.tensor(Im)
for nq in (300, 301):
  shard = RecordShard(X[lo:hi])
  Dm, Im = faiss_knn.sharded_search(shard, torch.tensor(X[:nq]), 7, lo, "L2", dist.group.WORLD, merge_packed_fn=cpu_merge_packed)
  assert np.array_equal(Im.numpy(), Iw[:nq]) and np.allclose(Dm.numpy(), Dw[:nq], atol=1e-6), nq
  assert shard.reranked < 7 * 2 * 128, shard.reranked      # the global bound keeps the re-rank near k rows per query over BOTH shards
  Ds, Is = faiss_knn.sharded_search(shard, torch.tensor(X[:nq]), 7, lo, "L2", dist.group.WORLD, merge_packed_fn=cpu_merge_packed, gather=False)
  per = (nq + 1) // 2
  assert np.array_equal(Is.numpy(), Iw[rank * per:min((rank + 1) * per, nq)])
# --- de-similarity filter with the rows split over the ranks (PRODUCT's sharded_desim, oracle-backed stand-in for cdml_desim)
rs = np.random.RandomState(3)
n, ke, kf = 101, 9, 6          # 101 rows: the last slice is short and padded for the all-gather
eI = np.stack([np.concatenate(([r], rs.choice(n, ke - 1, replace=False))) for r in range(n)]).astype(np.int64)
fI = np.stack([np.concatenate(([r], eI[r, 1:4], rs.choice(n, kf - 4, replace=False))) for r in range(n)]).astype(np.int64)
fD = np.sort(rs.rand(n, kf).astype(np.float32) * 2.0, axis=1)
def cpu_desim(e, f, d, thr, f_end, row_offset=0):
  return torch.tensor(O.iter_desim(e.numpy(), f.numpy(), d.numpy(), thr, f_end, row_offset=row_offset))
got = faiss_knn.sharded_desim(torch.tensor(eI), torch.tensor(fI), torch.tensor(fD), 1.4, 31, dist.group.WORLD, desim_fn=cpu_desim)
want = O.iter_desim(eI, fI, fD, 1.4, 31)
assert np.array_equal(got.numpy(), want) and (want != eI).sum() > n
# --- end of data with a batch count that is not a multiple of the world size: every step ends in a collective, so both
#     ranks must leave the reader loop after the same number of steps (7 batches -> 3 steps each; the old k % world
#     split gave rank 0 a 4th batch and it waited in its all-reduce until the timeout)
from cdml_b200 import inputs
pipe = inputs.MPTripletPipe(os.path.join(sys.argv[4], "*.train"), os.path.join(sys.argv[4], "features.npy"), seed=5, rank=rank, world=world)
pipe.create_pipe(num_epochs=2, batch_size=8)
steps, seen = 0, torch.zeros(1)
while (b := pipe.get_batch_indices()) is not None:
  t = torch.tensor([float(b[:, :2].sum())]); dist.all_reduce(t); seen += t; steps += 1
cnt = torch.tensor([steps, -steps]); dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
assert steps == 3 and cnt.tolist() == [3, -3], (steps, cnt)
# --- Prediction.run_features with the rows split over the ranks (predict.py:71-96 batches are independent): no collective
#     on the data path, one all-gather to join; 103 rows -> slices of 51 / 52, batch 20 with a tail batch in each slice
from cdml_b200 import predict
class HostTower:
  device, dims, calls = torch.device("cpu"), [6, 4], 0
  def embed(self, x):
    HostTower.calls += x.shape[0]
    return torch.tanh(x[:, :4] * 3.0 + x[:, 2:6])
feats = np.random.RandomState(9).rand(103, 6).astype(np.float32)
p = predict.Prediction(sess=HostTower())
whole = HostTower().embed(torch.tensor(feats)).numpy(); HostTower.calls = 0
got = p.run_features(feats, 20, process_group=dist.group.WORLD)
assert np.array_equal(got, whole) and HostTower.calls == (103 * (rank + 1) // 2 - 103 * rank // 2)
part = p.run_features(feats, 20, process_group=dist.group.WORLD, gather=False)
assert np.array_equal(part, whole[103 * rank // 2:103 * (rank + 1) // 2])
dist.barrier(); dist.destroy_process_group()
print("rank %d ok" % rank)
'''


def test_two_rank_protocol_over_gloo(tmp_path):
  script = tmp_path / "gloo_worker.py"
  script.write_text(_GLOO_WORKER)
  (tmp_path / "data").mkdir()
  _write_dataset(tmp_path / "data")
  port = str(29500 + os.getpid() % 2000)
  procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), str(tmp_path / "data")], stdout=subprocess.PIPE,
                            stderr=subprocess.STDOUT, text=True) for r in range(2)]
  outs = [p.communicate(timeout=240)[0] for p in procs]
  for r, (p, o) in enumerate(zip(procs, outs)):
    assert p.returncode == 0 and ("rank %d ok" % r) in o, o


def test_fusion_towers_compile_to_the_oracle_op_lists():
  """The product's recorder (models.MultiplyNet / MlpNet / ResNet / ResNetV2 -> compile_graph) must produce the op list
  the oracle restates from models.py:65-243; DenseNet is not constructible in the reference and says so."""
  import cdml_b200  # noqa: F401
  from cdml_b200 import models
  from oracle import cdml_oracle as O
  keys = {"input": ("lo", "hi"), "fc": ("src", "out", "bias_init", "alpha", "name"), "mul": ("src",), "add": ("src",),
          "l2norm": ("src",)}
  for name in ("MultiplyNet", "MlpNet", "ResNet", "ResNetV2"):
    g = models.compile_graph(getattr(models, name)().create_model(models.placeholder(1628))["l2_norm"])
    want = O.fusion_spec(name)
    assert g["F"] == 1628 and g["D"] == 256 and len(g["spec"]) == len(want)
    for got, w in zip(g["spec"], want):
      assert got["op"] == w["op"] and all(got[k] == w[k] for k in keys[w["op"]]), (name, got, w)
  with pytest.raises(NotImplementedError):
    models.DenseNet().create_model(models.placeholder(1628))
  # a plain chain still compiles as a graph (the GraphEngine then fuses its last layer with the output l2norm)
  g = models.compile_graph(models.VNet().create_model(models.placeholder(1500))["l2_norm"])
  assert [e["op"] for e in g["spec"]] == ["input", "fc", "fc", "l2norm"]


def test_device_reader_serves_positions_in_the_host_readers_order(lib, tmp_path):
  """`_position_batches` (what the device reader iterates) walks the files exactly as `_index_batches` does: same
  (anchor, positive) pairs per batch, same rank sharding, trailing partial batches dropped."""
  import cdml_b200  # noqa: F401
  from cdml_b200 import inputs
  _write_dataset(tmp_path)
  for world in (1, 2):
    for rank in range(world):
      pipe = inputs.MPTripletPipe(str(tmp_path / "*.train"), str(tmp_path / "features.npy"), rank=rank, world=world,
                                  device_reader=True)
      pipe.create_pipe(num_epochs=3, batch_size=4)
      host = []
      while True:
        b = pipe.get_batch_indices()
        if b is None:
          break
        host.append(b)
      pos = list(pipe._position_batches())
      assert len(host) == len(pos) > 0
      for batch, (i, start) in zip(host, pos):
        want = O.sample_triplets_device(pipe._pairs[i], start, 4, len(inputs.FEATURES), pipe.seed + 7919 * i)
        assert np.array_equal(batch[:, :2], want[:, :2])


# ---------------------------------------------------------------- native text formats
def test_native_float32_formatting_equals_numpy_str(lib):
  """cdml_format_f32 == str(np.float32(v)) (the distance format of knn_split*, faiss_knn.py:277) on 400k values: the KNN
  distance range, tiny values in scientific notation, both notation thresholds, signs and specials."""
  from cdml_b200 import _lib
  rng = np.random.RandomState(0)
  v = np.concatenate([rng.rand(250000).astype(np.float32) * 1.4, (rng.rand(50000) * 1e-3).astype(np.float32),
                      np.float32(10) ** rng.uniform(-12, 20, 100000).astype(np.float32),
                      np.array([1.0, 0.5, 1e-4, 9.9999e-5, 1e6, 999999.94, 123456.0, 0.1, 1e-5, 5e-5, 3e10, 0.0, -0.0, -2.5, 1.4,
                                1e-45, 3.4028235e38, np.inf, -np.inf, np.nan], np.float32)])
  out = np.empty(32 * len(v), np.uint8)
  n = _lib.load().cdml_format_f32(v.ctypes.data, len(v), out.ctypes.data, out.size)
  assert out[:n].tobytes().decode().split("\n")[:-1] == [str(x) for x in v]


def test_native_knn_row_formatter_equals_python_statement_and_golden(lib, tmp_path):
  from cdml_b200 import faiss_knn
  rng = np.random.RandomState(1)
  n, k = 500, 12
  decode = {i: ("g%d" % i if i % 7 else "视频_%d" % i) for i in range(n)}                    # non-ASCII guids too
  I = rng.randint(-1, n, (n, k)).astype(np.int64)
  D = (rng.rand(n, k) * 1.6).astype(np.float32)
  D[rng.rand(n, k) < 0.05] = 0.0
  D[3, 4], D[5, 6], D[7, 8] = np.float32(1.4), np.float32(1e-7), np.inf
  table = faiss_knn.GuidTable(decode)
  want = "".join(O.format_knn_rows(0, D, I, decode)).encode("utf-8")
  assert faiss_knn.format_rows_bytes(0, D, I, table) == want
  assert faiss_knn.format_rows_bytes(100, D[100:160], I[100:160], table) == "".join(O.format_knn_rows(100, D[100:160], I[100:160], decode)).encode("utf-8")
  assert "".join(faiss_knn.format_rows(0, D, I, decode)).encode("utf-8") == want
  faiss_knn.write_knn(str(tmp_path), split_num=3, D=D, I=I, prefix="knn_split", decode_map=decode)
  got = b"".join(open(tmp_path / ("knn_split%d" % j), "rb").read() for j in range(3))
  assert got == want
  bad = I.copy()
  bad[0, 1], D[0, 1] = n + 3, 0.5
  with pytest.raises(Exception):
    faiss_knn.format_rows_bytes(0, D, bad, table)


def test_native_feature_text_reader_follows_python_float_semantics(lib, tmp_path):
  from cdml_b200 import online_data
  lines = ["a#1,2.5,-3e-2", "b#1,2", "c#1,2,x", "no hash here", "d#4,5,6#7", "e# 7 ,+8.0,\t9e0", "f#inf,-Infinity,nan",
           "g#1e999,1e-999,.5", "", "h#0.1,0.2,0.30000001192092896\r", "视频#1,1,1", "a#9,9,9", "i#1,,3", "j#1_0,2,3",
           "k#0x10,2,3", "l#5.,6.,7."]
  path = tmp_path / "features.txt"
  path.write_text("\n".join(lines), encoding="utf-8")                                       # no newline after the last line
  feats, enc, dec = online_data.read_features_txt(str(path), width=3, num_threads=3)
  want_rows, want_guids = [], []
  for line in lines:                                                                          # online_data.py:66-77
    parts = line.split("#")
    if len(parts) != 2:
      continue
    try:
      if "_" in parts[1]:
        raise ValueError("digit separators are not accepted by the native reader (documented deviation)")
      vals = list(map(float, parts[1].split(",")))
    except ValueError:
      continue
    if len(vals) == 3:
      want_rows.append(np.asarray(vals, np.float64).astype(np.float32))
      want_guids.append(parts[0])
  assert [dec[i] for i in range(len(dec))] == want_guids and len(want_guids) == 8
  assert np.array_equal(feats, np.stack(want_rows), equal_nan=True) and feats.dtype == np.float32
  assert enc["a"] == want_guids.index("a", 1) and enc["视频"] == want_guids.index("视频")   # a later duplicate guid wins
  (tmp_path / "empty.txt").write_text("")
  assert online_data.read_features_txt(str(tmp_path / "empty.txt"), width=3)[0].shape == (0, 3)


def test_graph_recorder_rejects_what_the_engine_cannot_run(lib):
  """compile_graph's error paths: output must be an l2_normalize, slices only on the input placeholder, inner l2_normalize
  only on (a slice of) the input, one leaky slope, equal widths for joins; GuidTable needs every row index."""
  from cdml_b200 import faiss_knn, models
  x = models.placeholder(40)
  h = models.fully_connected(models.l2_normalize(x[:, :24]), 16)
  with pytest.raises(NotImplementedError):
    models.compile_graph(h)                                                   # not normalised
  with pytest.raises(NotImplementedError):
    models.compile_graph(models.l2_normalize(models.fully_connected(models.l2_normalize(h), 8)))   # inner l2norm on a layer
  with pytest.raises(NotImplementedError):
    models.compile_graph(models.l2_normalize(models.fully_connected(h[:, :8], 8)))                 # slice of a layer
  with pytest.raises(ValueError):
    models.multiply(h, models.fully_connected(models.l2_normalize(x[:, 24:]), 8))                  # widths differ
  with pytest.raises(ValueError):
    x[:, 30:30]
  with pytest.raises(NotImplementedError):
    x[::2]
  g = models.compile_graph(models.l2_normalize(h + h * h))                   # x + x*x: the same node twice in a join
  assert [e["op"] for e in g["spec"]] == ["input", "fc", "mul", "add", "l2norm"] and g["spec"][2]["src"] == [1, 1]
  with pytest.raises(KeyError):
    faiss_knn.GuidTable({0: "a", 2: "c"})
