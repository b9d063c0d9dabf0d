"""CPU: the oracle against the golden vectors produced by the reference's own Python (tests/golden/make_golden.py),
against the known answers of the reference's loss fixture, and against independent restatements (torch autograd,
finite differences, show_knn-style brute force)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cdml_oracle as O
from conftest import GOLDEN


def test_gather_matches_reference_numpy_fancy_index(golden):
  out = O.gather_rows(golden["gather_features"], golden["sampler_seed99_triplets"])
  assert out.dtype == np.float32 and out.shape == (8, 3, 12)
  assert np.array_equal(out, golden["gather_out"])
  assert np.array_equal(O.flatten_triplets(out)[4], golden["gather_out"][1, 1])


def test_negative_sampler_replays_reference_stream(golden):
  trip = O.sample_negatives(golden["sampler_seed99_pairs"], 64, np.random.RandomState(99))
  assert np.array_equal(trip, golden["sampler_seed99_triplets"])
  assert ((trip[:, 2] != trip[:, 0]) & (trip[:, 2] != trip[:, 1])).all()


def test_synthetic_features_match_imitation_data(golden):
  ref = golden["gen_features_seed1234"]
  assert np.array_equal(O.synth_features(64, 12, seed=1234), ref.astype(np.float32))
  assert golden["gen_triplets_seed7"].shape == (5, 3, 4)


def test_hinge_loss_known_answers_from_reference_fixture():
  with open(os.path.join(GOLDEN, "loss_fixture.json")) as f:
    fx = json.load(f)
  t = np.asarray(fx["triplets"], np.float32)
  r = O.hinge_loss(t, margin=0.1)
  assert r["pos_dist"].shape == (5, 1) and r["anchors"].shape == (5, 1, 2)
  assert np.allclose(r["pos_dist"][:, 0], fx["pos_dist"]) and np.allclose(r["neg_dist"][:, 0], fx["neg_dist"])
  assert np.allclose(r["hinge_dist"][:, 0], fx["hinge_dist@0.1"])
  assert abs(r["hinge_loss"] - fx["hinge_loss@0.1"]) < 1e-9
  assert abs(O.hinge_loss(t, margin=0.8)["hinge_loss"] - fx["hinge_loss@0.8"]) < 1e-9


def test_eval_mean_dist_and_rencode(golden):
  feats, cow = O.rencode_eval(golden["gather_features"], golden["eval_cowatches"])
  assert np.array_equal(feats, golden["eval_rencoded_features"])
  assert np.array_equal(np.asarray(cow), golden["eval_rencoded_cowatches"])
  assert abs(O.mean_dist(golden["eval_vectors"], golden["eval_cowatches"]) - float(golden["eval_mean_dist"])) < 1e-6


def test_knn_result_line_format(golden):
  with open(os.path.join(GOLDEN, "knn_decode_map.json")) as f:
    dm = {int(k): v for k, v in json.load(f).items()}
  with open(os.path.join(GOLDEN, "knn_split0.txt")) as f:
    want = f.read()
  assert "".join(O.format_knn_rows(0, golden["knn_D"], golden["knn_I"], dm)) == want
  assert O.split_ranges(103, 10)[-1] == (90, 103) and O.split_ranges(103, 10)[0] == (0, 10)


def _torch_tower(x, params, margin):
  x = torch.tensor(x, dtype=torch.float64)
  ps = [(torch.tensor(W, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True))
        for W, b in params]
  h = x * torch.rsqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=1e-12))
  for W, b in ps:
    h = torch.nn.functional.leaky_relu(h @ W + b, 0.2)
  e = h * torch.rsqrt(torch.clamp((h * h).sum(-1, keepdim=True), min=1e-12))
  E = e.view(-1, 3, e.shape[-1])
  pos = ((E[:, 0] - E[:, 1]) ** 2).sum(-1)
  neg = ((E[:, 0] - E[:, 2]) ** 2).sum(-1)
  loss = torch.clamp(pos - neg + margin, min=0).mean()
  loss.backward()
  return e.detach().numpy(), loss.item(), [(W.grad.numpy(), b.grad.numpy()) for W, b in ps]


def test_tower_forward_backward_against_torch_autograd():
  rng = np.random.RandomState(0)
  x = rng.random_sample((12, 40))
  params = O.init_tower([40, 64, 24, 16], seed=3, bias_init=0.05, dtype=np.float64)
  tr = O.OracleTrainer(params, margin=0.8)
  fwd, loss, grads = tr.loss_and_grads(x)
  e_t, loss_t, grads_t = _torch_tower(x, params, 0.8)
  assert np.allclose(fwd["l2_norm"], e_t, atol=1e-12)
  assert abs(loss["hinge_loss"] - loss_t) < 1e-12
  for (gW, gb), (tW, tb) in zip(grads, grads_t):
    assert np.allclose(gW, tW, atol=1e-12) and np.allclose(gb, tb, atol=1e-12)
  assert np.allclose(np.linalg.norm(fwd["l2_norm"], axis=1), 1.0)


def test_backward_against_finite_differences():
  rng = np.random.RandomState(1)
  x = rng.random_sample((6, 10))
  params = O.init_tower([10, 12, 8], seed=4, bias_init=0.1, dtype=np.float64)
  tr = O.OracleTrainer(params, margin=0.8)
  _, _, grads = tr.loss_and_grads(x)
  W0 = tr.params[0][0]
  for (i, j) in [(0, 0), (3, 5), (9, 11)]:
    eps = 1e-6
    W0[i, j] += eps
    lp = tr.loss_and_grads(x)[1]["hinge_loss"]
    W0[i, j] -= 2 * eps
    lm = tr.loss_and_grads(x)[1]["hinge_loss"]
    W0[i, j] += eps
    assert abs((lp - lm) / (2 * eps) - grads[0][0][i, j]) < 1e-6


def test_tf1_adam_first_step_closed_form_and_schedule():
  w, g = np.array([1.0, -2.0, 0.5]), np.array([0.3, -0.7, 1e-3])
  w1, m1, v1 = O.adam_step_tf1(w, np.zeros(3), np.zeros(3), g, lr=1e-3, t=1)
  # t=1: m=(1-b1)g, v=(1-b2)g^2, lr_t=lr*sqrt(1-b2)/(1-b1)  =>  w -= lr * g/(|g| + eps/sqrt(1-b2))
  assert np.allclose(w1, w - 1e-3 * g / (np.abs(g) + 1e-8 / np.sqrt(1 - 0.999)), rtol=1e-12)
  assert O.exponential_decay(0.01, 999999, 1000000, 0.96) == 0.01
  assert abs(O.exponential_decay(0.01, 2500000, 1000000, 0.96) - 0.01 * 0.96 ** 2) < 1e-15


def test_clip_momentum_lars_and_regularizer_statements():
  """The rest of build_graph's gradient path (train.py:47-64, :115-118, :133-146, :354), pinned to closed forms and to
  torch's independent SGD-with-Nesterov implementation (TF and torch share that formulation)."""
  rng = np.random.RandomState(3)
  w, g = rng.standard_normal((7, 5)), rng.standard_normal((7, 5))
  # clip_by_norm: untouched below the threshold, rescaled to exactly the threshold above it
  assert np.array_equal(O.clip_by_norm(g, 100.0), g)
  assert abs(np.linalg.norm(O.clip_by_norm(g, 1.0)) - 1.0) < 1e-12
  # Nesterov momentum vs torch.optim.SGD(momentum=0.9, nesterov=True)
  wt = torch.tensor(w.copy(), requires_grad=True)
  opt = torch.optim.SGD([wt], lr=0.1, momentum=0.9, nesterov=True)
  wn, acc = w.copy(), np.zeros_like(w)
  for i in range(4):
    wt.grad = torch.tensor(g * (i + 1))
    opt.step()
    wn, acc = O.momentum_step_tf1(wn, acc, g * (i + 1), 0.1)
  assert np.abs(wn - wt.detach().numpy()).max() < 1e-12
  # LARS as of TF r1.13 (weight decay in the trust ratio only; apply_momentum(var, mom, lr*trust, grad, momentum)).
  # First step (acc = 0): w1 = w - lr * eeta*|w|/(|g| + wd*|w|) * g; zero gradient -> trust ratio 1 and no movement
  w1, acc1 = O.lars_step_tf1(w, np.zeros_like(w), g, lr=1.0)
  wn_, gn_ = np.linalg.norm(w), np.linalg.norm(g)
  assert np.allclose(w1, w - 1e-3 * wn_ / (gn_ + 1e-4 * wn_) * g, rtol=1e-12) and np.array_equal(acc1, g)
  w2, _ = O.lars_step_tf1(w, np.zeros_like(w), np.zeros_like(w), lr=0.5)
  assert np.array_equal(w2, w)
  # second step: the accumulator carries the RAW gradient (the trust ratio of step 1 is not folded into it)
  g2 = rng.standard_normal((7, 5))
  w3, acc3 = O.lars_step_tf1(w1, acc1, g2, lr=1.0)
  t2 = 1e-3 * np.linalg.norm(w1) / (np.linalg.norm(g2) + 1e-4 * np.linalg.norm(w1))
  assert np.allclose(acc3, 0.9 * g + g2) and np.allclose(w3, w1 - t2 * (0.9 * g + g2), rtol=1e-12)
  # regularizer: d(reg_penalty * reg_loss)/dW by finite differences, and its use in step()
  params = O.init_tower([6, 8, 4], seed=2, dtype=np.float64)
  tr = O.OracleTrainer(params, optimizer="sgd", lr=0.5, reg_penalty=3.0, l2_penalty=1e-2)
  W0 = tr.params[0][0]
  base = tr.reg_loss()
  W0[1, 2] += 1e-6
  fd = (tr.reg_loss() - base) / 1e-6
  W0[1, 2] -= 1e-6
  assert abs(fd - 1e-2 * W0[1, 2]) < 1e-7
  x = np.random.RandomState(5).random_sample((12, 6))
  grads = tr.loss_and_grads(x)[2]
  want = W0 - 0.5 * (grads[0][0] + 3.0 * 1e-2 * W0)
  tr.step(x)
  assert np.allclose(tr.params[0][0], want, rtol=1e-12)


def test_flat_knn_against_independent_brute_force_and_merge():
  rng = np.random.RandomState(4)
  X = rng.standard_normal((600, 32)).astype(np.float32)
  D, I = O.flat_knn(X, k=10)
  Xn = O.knn_normalize(X)
  for q in (0, 17, 599):
    assert list(I[q]) == list(O.exact_ip_nn(Xn, q, 10))          # show_knn.py:63-68 statement
  assert (I[:, 0] == np.arange(600)).all() and np.allclose(D[:, 0], 0, atol=1e-6)
  assert (np.diff(D, axis=1) >= 0).all() and (D >= 0).all() and (D <= 4 + 1e-5).all()
  # sharded: 3 row shards with id offsets, merged == global
  parts = [O.flat_knn(Xn[s:e], Xn, k=10, l2_norm=False) for s, e in ((0, 200), (200, 400), (400, 600))]
  Dg = [p[0] for p in parts]
  Ig = [p[1] + off for p, off in zip(parts, (0, 200, 400))]
  Dm, Im = O.knn_merge(Dg, Ig, 10)
  assert np.array_equal(Im, I) and np.allclose(Dm, D, atol=1e-6)
  # fewer rows than k pads like faiss
  Ds, Is = O.flat_knn(X[:4], k=6)
  assert (Is[:, 4:] == -1).all() and np.isinf(Ds[:, 4:]).all()
  # inner product
  Dip, Iip = O.flat_knn(X, k=5, l2_norm=False, metric="IP")
  assert (np.diff(Dip, axis=1) <= 0).all() and Iip[3, 0] == np.argmax(X @ X[3])


def test_flat_knn_ids_pinned_to_reference_show_knn_and_sklearn(knn_golden):
  """Pin of the exact-KNN ids: (1) the ids the reference's OWN brute force (show_knn.calc_nn, executed under import shims
  by tests/golden/make_knn_golden.py) returned for seeded rows; (2) scikit-learn's brute-force NearestNeighbors as a
  third, independent statement.  The oracle must reproduce both except where the k-th / (k+1)-th inner products differ by
  less than fp32 summation noise (north_star's exact-tie exemption) -- and then only as a permutation inside the tie."""
  from sklearn.neighbors import NearestNeighbors
  for name, (seed, N, d, nq, k, clustered) in knn_golden["cases"].items():
    X = knn_golden["rows"](seed, N, d, clustered)
    q = knn_golden["ids"][name + "_queries"]
    want, gap = knn_golden["ids"][name + "_ids"], knn_golden["ids"][name + "_gap"]
    assert np.array_equal(want[:, 0], q)                                  # a row is its own nearest neighbour
    clear = gap > 2e-6
    assert clear.mean() > 0.9
    for metric in ("IP", "L2"):
      D, I = O.flat_knn(X, X[q], k=k, l2_norm=False, metric=metric)
      assert np.array_equal(I[clear], want[clear]), (name, metric)
      assert all(set(I[i]) ^ set(want[i]) == set() or gap[i] <= 2e-6 for i in range(nq))
      if metric == "L2":
        assert np.all(np.diff(D, axis=1) >= 0) and np.allclose(D[:, 0], 0, atol=2e-6)
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean").fit(X.astype(np.float64))
    dist, ids = nn.kneighbors(X[q].astype(np.float64))
    assert np.array_equal(ids[clear], want[clear]), name                   # third statement agrees with the reference's
    D, _ = O.flat_knn(X, X[q], k=k, l2_norm=False, metric="L2")
    assert np.allclose(D, dist ** 2, atol=5e-6)


def test_emulated16_selection_model_is_the_mining_definition_on_rounded_operands():
  """mine_semihard_emulated16 (the precision model the GPU test gates against) == mine_semihard evaluated on the rounded
  embeddings, wherever the pick is not a near-tie of the rounding itself."""
  rng = np.random.RandomState(3)
  B, D, G = 96, 32, 60
  trip = O.synth_triplets(B, G, seed=5)
  E = O.l2_normalize(O.l2_normalize(rng.standard_normal((G, D)))[trip.reshape(-1)] + 0.3 * rng.standard_normal((3 * B, D))).astype(np.float32)
  row16, d16, gap = O.mine_semihard_emulated16(E, trip, 0.8, "fp16")
  row, d = O.mine_semihard(O.round16(E, "fp16"), trip, 0.8)
  # the two differ only in HOW the rounded operands enter (d = 2 - 2 a16.c16 against the fp32 dp here, |a16 - c16|^2
  # against the rounded dp there: rounded rows are unit only to ~3e-4), which moves picks that sit at the d > dp boundary
  same = row16 == row
  assert same.mean() > 0.95
  assert np.allclose(d16[same], d[same], atol=2e-3)
  dp = ((E[0::3].astype(np.float64) - E[1::3]) ** 2).sum(-1)
  for i in np.flatnonzero(~same):
    assert min(abs(d16[i] - dp[i]), abs(d[i] - dp[i])) < 2e-3 or gap[i] < 2e-3, i


def test_knn_record_format_orders_like_distance_then_id():
  """The 64-bit shard records (cdml_knn_shard_refine / cdml_knn_merge_packed): f2key is an order-preserving bijection of
  fp32 onto uint32 (negative floats included: inner products), records sort by (distance, id), padding sorts last, and a
  merge of per-shard lists equals the (D, I)-list merge knn_merge."""
  rng = np.random.RandomState(8)
  x = np.concatenate([rng.standard_normal(2000).astype(np.float32) * 3, np.array([0.0, -0.0, 1e-30, -1e-30, np.inf, -np.inf], np.float32)])
  k = O.f2key(x)
  nz = ~((x == 0) & np.signbit(x))                                            # -0.0 == 0.0 as floats, adjacent as keys
  order = np.argsort(x[nz], kind="stable")
  assert np.all(np.diff(k[nz][order].astype(np.int64)) >= 0)                  # monotone
  z = O.f2key(np.array([0.0, -0.0], np.float32)).astype(np.int64)
  assert z[0] - z[1] == 1
  back = O.key2f(k)
  assert np.array_equal(back.view(np.uint32), x.view(np.uint32))              # exact inverse, bit for bit
  for metric in ("L2", "IP"):
    G, nq, kk = 3, 17, 5
    vals = rng.rand(G, nq, kk).astype(np.float32) * (1.0 if metric == "L2" else -1.0) + (0.0 if metric == "L2" else 0.5)
    vals = np.sort(vals, axis=2) if metric == "L2" else -np.sort(-vals, axis=2)
    ids = rng.permutation(G * nq * kk).reshape(G, nq, kk).astype(np.int64)
    ids[1, :, 3:] = -1                                                        # a short list: padded
    vals[1, :, 3:] = np.inf if metric == "L2" else -np.inf
    rec = np.stack([O.knn_pack_records(vals[g], ids[g], metric) for g in range(G)])
    assert np.all(np.diff(rec[0].astype(np.float64), axis=1) > 0)             # a sorted list gives ascending records
    assert np.all(rec[1, :, 3:] == np.uint64(0xFFFFFFFFFFFFFFFF))
    Dm, Im = O.knn_merge_records(rec, kk, metric)
    Dw, Iw = O.knn_merge([vals[g] for g in range(G)], [ids[g] for g in range(G)], kk, metric)
    assert np.array_equal(Im, Iw) and np.array_equal(Dm, Dw)


def test_emulated16_trainer_tracks_the_float64_trainer():
  """OracleTrainer(emulate16=...) -- the precision model the GPU trajectory test gates against -- is the float64 trainer with
  rounded operands: fp16 stays within a few 1e-4 of its loss, bf16 within a few 1e-3, and both train."""
  feats = O.synth_features(300, 24, 0)
  params = O.init_tower([24, 32, 16], seed=2, dtype=np.float64)
  tr = {name: O.OracleTrainer(params, lr=1e-2, margin=0.8, emulate16=e) for name, e in (("f64", None), ("fp16", "fp16"), ("bf16", "bf16"))}
  first = {}
  for t in range(12):
    x = O.flatten_triplets(O.gather_rows(feats, O.synth_triplets(64, 300, 5 + t)))
    for name, trainer in tr.items():
      loss = trainer.step(x)[0]
      first.setdefault(name, loss)
      last = loss if name != "f64" else loss
      tr[name].last = loss
  assert abs(tr["fp16"].last / tr["f64"].last - 1) < 2e-3 and abs(tr["bf16"].last / tr["f64"].last - 1) < 2e-2
  assert all(tr[n].last < first[n] for n in tr)


def test_semihard_mining_definition():
  rng = np.random.RandomState(5)
  B, D = 40, 16
  E = O.l2_normalize(rng.standard_normal((3 * B, D)))
  g = O.synth_triplets(B, 30, seed=2)
  neg_row, d_an = O.mine_semihard(E, g, margin=0.8)
  A, P = E[0::3], E[1::3]
  dp = ((A - P) ** 2).sum(-1)
  for i in range(B):
    r = neg_row[i]
    assert r % 3 != 0 or r == 3 * i + 2
    if r != 3 * i + 2 or True:
      guid_r = g[r // 3, r % 3]
      if r != 3 * i + 2:
        assert guid_r != g[i, 0] and guid_r != g[i, 1]
    assert abs(d_an[i] - ((A[i] - E[r]) ** 2).sum()) < 1e-9
    if dp[i] < d_an[i] < dp[i] + 0.8:
      # semi-hard: no candidate is closer while still beyond the positive
      cand = [3 * j + c for j in range(B) for c in (1, 2) if g[j, c] not in (g[i, 0], g[i, 1])]
      dc = ((A[i] - E[cand]) ** 2).sum(-1)
      assert d_an[i] <= dc[(dc > dp[i]) & (dc < dp[i] + 0.8)].min() + 1e-12


# ---------------------------------------------------------------- fusion towers (models.py:65-243)
def _shrunk(spec):
  """Same topology, toy widths (visual 30 / doc 12 columns) so the float64 checks run in milliseconds."""
  out = []
  for e in spec:
    e = dict(e)
    if e["op"] == "input":
      e["lo"], e["hi"] = (0, 30) if e["lo"] == 0 else (30, 42)
    elif e["op"] == "fc":
      e["out"] = {5000: 56, 400: 24, 600: 32, 256: 16}[e["out"]]
    out.append(e)
  return out


@pytest.mark.parametrize("name", ["MultiplyNet", "MlpNet", "ResNet", "ResNetV2"])
def test_fusion_graph_backward_matches_torch_autograd(name):
  """The oracle's op-list reverse sweep (graph_backward) against torch autograd of the same forward, float64."""
  import torch
  spec = _shrunk(O.fusion_spec(name))
  params = O.init_graph(spec, seed=3, dtype=np.float64)
  x = np.random.RandomState(0).rand(12, 42)
  fwd = O.graph_forward(x, spec, params)
  E = fwd["l2_norm"].reshape(-1, 3, 16)
  grads = O.graph_backward(fwd, spec, params, O.hinge_loss_grad(E, 0.8).reshape(-1, 16))
  tp = [(torch.tensor(W, requires_grad=True), torch.tensor(b, requires_grad=True)) for W, b in params]
  xt, vals, li = torch.tensor(x), [], 0
  l2n = lambda y: y * torch.rsqrt(torch.clamp((y * y).sum(-1, keepdim=True), min=1e-12))
  for e in spec:
    if e["op"] == "input":
      vals.append(l2n(xt[:, e["lo"]:e["hi"]]))
    elif e["op"] == "fc":
      W, b = tp[li]
      li += 1
      vals.append(torch.nn.functional.leaky_relu(vals[e["src"]] @ W + b, 0.2))
    elif e["op"] == "mul":
      vals.append(vals[e["src"][0]] * vals[e["src"][1]])
    elif e["op"] == "add":
      vals.append(sum(vals[s] for s in e["src"]))
    else:
      vals.append(l2n(vals[e["src"]]))
  Et = vals[-1].reshape(-1, 3, 16)
  loss = torch.clamp(((Et[:, 0] - Et[:, 1]) ** 2).sum(-1) - ((Et[:, 0] - Et[:, 2]) ** 2).sum(-1) + 0.8, min=0).mean()
  loss.backward()
  assert np.allclose(vals[-1].detach().numpy(), fwd["l2_norm"], atol=1e-12)
  assert abs(float(loss.detach()) - O.hinge_loss(E, 0.8)["hinge_loss"]) < 1e-12
  for (gW, gb), (W, b) in zip(grads, tp):
    assert np.allclose(gW, W.grad.numpy(), atol=1e-12) and np.allclose(gb, b.grad.numpy(), atol=1e-12)


def test_fusion_spec_shapes_follow_models_py():
  """Widths stated in models.py: visual 1500 -> 5000 -> 256 (:81-83), doc 128 -> 400 -> 256 (:86-88), MLP 600 (:118),
  residual layers 256 (:149-151); biases 0.1."""
  for name, shapes in (("MultiplyNet", [(1500, 5000), (5000, 256), (128, 400), (400, 256)]),
                       ("MlpNet", [(1500, 5000), (5000, 256), (128, 400), (400, 256), (256, 600), (600, 256)]),
                       ("ResNet", [(1500, 5000), (5000, 256), (128, 400), (400, 256), (256, 256), (256, 256)]),
                       ("ResNetV2", [(1500, 5000), (5000, 256), (1500, 256), (128, 400), (400, 256), (128, 256), (256, 256),
                                     (256, 256)])):
    params = O.init_graph(O.fusion_spec(name), seed=2)
    assert [W.shape for W, _ in params] == shapes
    assert all(np.all(b == np.float32(0.1)) for _, b in params)


# ---------------------------------------------------------------- de-similarity filter (faiss_knn.py:134-244)
def test_desim_oracle_equals_reference_iter_desim_mp():
  """oracle.iter_desim (row-wise restatement) against outputs of the reference's own iter_desim_mp / desim run in the
  build container (tests/golden/make_desim_golden.py): column sweep, f_end cut, -1 padding, more processes than rows."""
  g = np.load(os.path.join(GOLDEN, "desim_golden.npz"))
  for name in "abcd":
    f_end = int(g[name + "_args"][0])
    eI, fI, fD = g[name + "_eI"], g[name + "_fI"], g[name + "_fD"]
    keep = (eI.copy(), fI.copy(), fD.copy())
    assert np.array_equal(O.iter_desim(eI, fI, fD, 1.4, f_end), g[name + "_out"]), name
    assert np.array_equal(O.desim_simple(eI, fI), g[name + "_simple"]), name
    assert all(np.array_equal(a, b) for a, b in zip(keep, (eI, fI, fD)))          # the oracle does not mutate its inputs
  assert (g["b_out"] == -1).mean() > (g["b_eI"] == -1).mean()                     # something was actually filtered


# ---------------------------------------------------------------- device reader generator
def test_philox4x32_10_known_answers_and_sampler_properties():
  """Random123 known-answer vectors (kat_vectors: philox4x32 10) pin the generator the device reader is stated with;
  the sampler never returns the anchor or the positive and is uniform over the rest."""
  kat = (((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
         ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
         ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
          (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)))
  for ctr, key, want in kat:
    assert tuple(int(x) for x in O.philox4x32_10(ctr, key)) == want
  G = 1000
  pairs = np.stack([np.arange(G), (np.arange(G) + 1) % G], 1)
  t = O.sample_triplets_device(pairs, 5, 300000, G, seed=5)
  assert np.array_equal(t[:, :2], pairs[(5 + np.arange(300000)) % G])            # wrap-around over the file
  assert ((t[:, 2] != t[:, 0]) & (t[:, 2] != t[:, 1])).all() and t[:, 2].min() == 0 and t[:, 2].max() == G - 1
  h = np.bincount(t[:, 2], minlength=G)
  assert abs(h - 300.0).max() < 6 * np.sqrt(300.0)                               # uniform: 6 sigma of a Poisson(300) bin
  tiny = O.sample_triplets_device(np.array([[0, 1]]), 0, 64, 3, seed=1)          # only one id is admissible
  assert (tiny[:, 2] == 2).all()
  a = O.sample_triplets_device(pairs, 0, 100, G, seed=7)
  b = O.sample_triplets_device(pairs, 50, 50, G, seed=7)
  assert np.array_equal(a[50:], b)                                               # a position owns its stream
