"""CPU oracle for the CDML hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy, the arithmetic of the reference
(geekieo/collaborative-deep-metric-learning) hot path so that the CUDA kernels
can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package never does; it fails loudly when ``libcdml.so`` is missing.

Pinning status (see DESIGN.md "Oracle"):
  * gather / negative sampler / eval mean_dist / knn_result line format /
    feature-text reader: pinned against outputs of the reference's own Python
    run in the build container (tests/golden/make_golden.py, fixtures committed).
  * hinge loss: pinned against the known answers derived from the reference's
    fixture tests/test_losses.py:13-18 (the reference asserts nothing itself).
  * de-similarity filter (iter_desim, desim_simple, filter_fI): pinned against
    outputs of the reference's own faiss_knn.iter_desim_mp / desim / fliter_fI
    run in the build container (tests/golden/make_desim_golden.py).
  * device reader: the counter-based generator is pinned to the Random123
    known-answer vectors of philox4x32-10; the sampling rule follows
    inputs.py:112-130 (the reference's own stream is unseeded: only the
    distribution can be matched).
  * fusion towers (fusion_spec / graph_forward / graph_backward): **parity
    unpinned** like the chain tower (TensorFlow ops); cross-checked by torch
    autograd of the same forward to 1e-12.
  * exact flat KNN ids (flat_knn, L2 and IP): pinned against the ids returned
    by the reference's OWN brute force, show_knn.calc_nn (show_knn.py:63-79),
    executed in the build container under cv2/matplotlib import shims
    (tests/golden/make_knn_golden.py -> knn_ids_golden.npz), and against
    scikit-learn's brute-force NearestNeighbors as a third statement.  faiss
    itself (faiss-gpu==1.5.x, README.md:22) is not installable here: distances
    and the -1/inf padding follow its published IndexFlat semantics.
  * tower forward/backward, TF1 Adam / Momentum / LARS (tensorflow r1.13 form),
    clip_by_norm, l2_regularizer: **parity unpinned** -- the arithmetic lives
    in tensorflow-gpu==1.13.1 (README.md:19-21), neither vendored nor
    installable here.  The oracle restates the published semantics and is
    cross-checked by torch autograd, torch's Nesterov SGD, finite differences
    and closed forms.
  * precision models (tower_grads_emulated16, mine_semihard_emulated16,
    OracleTrainer(emulate16=...)): the same statements with operands rounded
    to fp16 / bf16 exactly where the kernels store them -- what the GPU tests
    gate the 16-bit kernels against next to the unrounded float64 oracle.

Every function cites the reference file:line it follows.  Default dtype is
float64 ("what the maths says"); pass ``dtype=np.float32`` for the bit-level
checks (gather, ids).
"""
from __future__ import annotations

import numpy as np

L2_EPS = 1e-12          # tf.nn.l2_normalize default epsilon (models.py:58,61)
LEAKY_ALPHA = 0.2       # tf.nn.leaky_relu default alpha (models.py:21)


# --------------------------------------------------------------------------- #
# reader / gather                                                             #
# --------------------------------------------------------------------------- #
def gather_rows(features: np.ndarray, guid_triplets) -> np.ndarray:
  """inputs.py:158 -- ``FEATURES[np.asarray(guid_triplets)]`` -> [B,3,F]."""
  return features[np.asarray(guid_triplets)]


def flatten_triplets(batch: np.ndarray) -> np.ndarray:
  """train.py:313 -- [B,3,F] -> [3B,F], row order a0,p0,n0,a1,..."""
  return np.reshape(batch, (-1, batch.shape[-1]))


def sample_negatives(pairs: np.ndarray, num_guid: int, rng: np.random.RandomState) -> np.ndarray:
  """inputs.py:123-129 + parse_data.py:292-298 -- for each (a,p) draw
  ``randint(0,num_guid)`` until it is not in {a,p}; returns [B,3] int64.
  Draw order is one scalar ``randint`` call per attempt, exactly as the
  reference's generator does, so a seeded RandomState reproduces its stream."""
  out = np.empty((len(pairs), 3), dtype=np.int64)
  for i, (a, p) in enumerate(np.asarray(pairs)):
    n = rng.randint(0, num_guid)
    while n == a or n == p:
      n = rng.randint(0, num_guid)
    out[i] = (a, p, n)
  return out


# --------------------------------------------------------------------------- #
# tower (models.py:19-30, 41-62)                                              #
# --------------------------------------------------------------------------- #
def l2_normalize(x: np.ndarray, eps: float = L2_EPS) -> np.ndarray:
  """tf.nn.l2_normalize(x, axis=-1): x * rsqrt(max(sum(x^2), eps))."""
  ss = np.sum(np.square(x), axis=-1, keepdims=True)
  return x / np.sqrt(np.maximum(ss, eps))


def leaky_relu(z: np.ndarray, alpha: float = LEAKY_ALPHA) -> np.ndarray:
  return np.where(z > 0, z, alpha * z)


def fully_connected(x, W, b, alpha: float = LEAKY_ALPHA):
  """models.py:19-30 -- slim.fully_connected: leaky_relu(x @ W + b); W is [in,out]."""
  return leaky_relu(x @ W + b, alpha)


def xavier_uniform(rng: np.random.RandomState, fan_in: int, fan_out: int, dtype=np.float32):
  """slim default weights_initializer: U(-sqrt(6/(in+out)), +sqrt(6/(in+out)))."""
  lim = np.sqrt(6.0 / (fan_in + fan_out))
  return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(dtype)


def init_tower(dims, seed: int = 2, bias_init: float = 0.0, dtype=np.float32):
  """dims = [F, H1, ..., D].  VNet (models.py:59-60): [1500, 5000, 256], bias 0."""
  rng = np.random.RandomState(seed)
  params = []
  for fi, fo in zip(dims[:-1], dims[1:]):
    params.append((xavier_uniform(rng, fi, fo, dtype), np.full((fo,), bias_init, dtype)))
  return params


def tower_forward(x, params, alpha: float = LEAKY_ALPHA, dtype=np.float64):
  """models.py:46-62 (VNet) generalised to a stack of fully_connected layers.

  Returns dict with ``xhat`` (normalised input), ``layers`` (post-activation
  output of every FC layer; layers[-1] is "layer_2" for VNet), ``rinv``
  (rsqrt(max(sum y^2,eps)) of the last layer) and ``l2_norm`` (the embedding).
  """
  x = np.asarray(x, dtype)
  xhat = l2_normalize(x)
  acts = []
  h = xhat
  for W, b in params:
    h = fully_connected(h, np.asarray(W, dtype), np.asarray(b, dtype), alpha)
    acts.append(h)
  ss = np.sum(np.square(h), axis=-1, keepdims=True)
  rinv = 1.0 / np.sqrt(np.maximum(ss, L2_EPS))
  return {"xhat": xhat, "layers": acts, "rinv": rinv[:, 0], "l2_norm": h * rinv}


# --------------------------------------------------------------------------- #
# loss (losses.py:21-49)                                                      #
# --------------------------------------------------------------------------- #
def hinge_loss(triplets, margin: float = 0.1, dtype=np.float64):
  """losses.py:33-38.  triplets [B,3,D] -> the 7-key dict of losses.py:43-49."""
  t = np.asarray(triplets, dtype)
  a, p, n = t[:, 0:1, :], t[:, 1:2, :], t[:, 2:3, :]
  pos = np.sum(np.square(a - p), axis=-1)
  neg = np.sum(np.square(a - n), axis=-1)
  hinge = np.maximum(pos - neg + margin, 0)
  return {"hinge_loss": np.mean(hinge), "anchors": a, "positives": p, "negatives": n,
          "pos_dist": pos, "neg_dist": neg, "hinge_dist": hinge}


def hinge_loss_grad(triplets, margin: float = 0.1, dtype=np.float64):
  """d(mean hinge)/d(triplets): 2(n-p)/B, 2(p-a)/B, 2(a-n)/B on active rows."""
  t = np.asarray(triplets, dtype)
  B = t.shape[0]
  a, p, n = t[:, 0], t[:, 1], t[:, 2]
  r = hinge_loss(t, margin, dtype)
  act = (r["hinge_dist"][:, 0] > 0)[:, None]
  g = np.zeros_like(t)
  g[:, 0] = np.where(act, 2 * (n - p) / B, 0)
  g[:, 1] = np.where(act, 2 * (p - a) / B, 0)
  g[:, 2] = np.where(act, 2 * (a - n) / B, 0)
  return g


def calc_var(triplets):
  """train.py:67-71 -- mean((E - mean_{B,3}(E))^2)."""
  mean = np.mean(triplets, axis=(0, 1))
  return np.mean((triplets - mean) ** 2)


# --------------------------------------------------------------------------- #
# backward (autodiff of train.py:141-142 written out)                          #
# --------------------------------------------------------------------------- #
def tower_backward(fwd, params, dE, alpha: float = LEAKY_ALPHA, dtype=np.float64):
  """Given forward cache and dL/d(l2_norm) [R,D] returns [(dW,db)] per layer.

  l2norm:  dy = (g - e*(e.g)) * rinv   (rows with sum y^2 >= eps)
  leaky :  dz = dy * (z>0 ? 1 : alpha)   (sign(z) == sign(leaky(z)))
  dense :  dW = in^T dz ; db = sum_rows dz ; d_in = dz W^T
  No gradient flows into xhat (placeholder input, train.py:265).
  """
  e = fwd["l2_norm"]
  g = np.asarray(dE, dtype)
  dy = (g - e * np.sum(e * g, axis=-1, keepdims=True)) * fwd["rinv"][:, None]
  grads = [None] * len(params)
  inputs = [fwd["xhat"]] + fwd["layers"][:-1]
  d_out = dy
  for li in range(len(params) - 1, -1, -1):
    y = fwd["layers"][li]
    dz = d_out * np.where(y > 0, 1.0, alpha)
    W = np.asarray(params[li][0], dtype)
    grads[li] = (inputs[li].T @ dz, np.sum(dz, axis=0))
    if li > 0:
      d_out = dz @ W.T
  return grads


def round16(a, kind="fp16"):
  """Round-to-nearest-even to the 16-bit storage type the kernels use, back to float64."""
  a = np.asarray(a, np.float64)
  if kind == "fp16":
    return a.astype(np.float16).astype(np.float64)
  u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
  u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
  return u.astype(np.uint32).view(np.float32).astype(np.float64)


def tower_grads_emulated16(x, params, margin, kind="fp16", alpha: float = LEAKY_ALPHA):
  """PRECISION MODEL of the GPU step (test infrastructure): the exact arithmetic of tower_forward / hinge /
  tower_backward in float64, but with operands rounded to 16 bits at exactly the places the kernels store them
  (normalised input, shadow weights, hidden activations, dz of every layer).  The kernels must match this model to
  ~1e-3; the distance between this model and the unrounded oracle is the inherent cost of 16-bit operands
  (differences of nearly equal embeddings amplify the ~5e-4 forward error to ~1-2e-2 in the gradients)."""
  r = lambda a: round16(a, kind)
  xh = r(l2_normalize(np.asarray(x, np.float64)))
  W16 = [r(W) for W, _ in params]
  acts, h = [], xh
  for li, (W, b) in enumerate(params):
    z = leaky_relu(h @ W16[li] + np.asarray(b, np.float64), alpha)
    h = z if li == len(params) - 1 else r(z)
    acts.append(h)
  y = acts[-1]
  rinv = 1.0 / np.sqrt(np.maximum(np.sum(y * y, -1, keepdims=True), L2_EPS))
  e = y * rinv
  E = e.reshape(-1, 3, e.shape[-1])
  B = E.shape[0]
  loss = hinge_loss(E, margin)
  g = hinge_loss_grad(E, margin).reshape(-1, e.shape[-1]) * B          # gradient of the SUM of hinges
  dz = r((g - e * np.sum(e * g, -1, keepdims=True)) * rinv * np.where(e > 0, 1.0, alpha))
  grads = [None] * len(params)
  inputs = [xh] + acts[:-1]
  for li in range(len(params) - 1, -1, -1):
    grads[li] = (inputs[li].T @ dz / B, np.sum(dz, 0) / B)
    if li > 0:
      dz = r((dz @ W16[li].T) * np.where(acts[li - 1] > 0, 1.0, alpha))
  return {"l2_norm": e, "loss": loss, "grads": grads}


# --------------------------------------------------------------------------- #
# fusion towers (models.py:65-243) as a small op list                           #
# --------------------------------------------------------------------------- #
def fusion_spec(name, feature_size=1628, visual=1500):
  """Op list of the reference's fusion towers, restated from models.py: MultiplyNet :65-91, MlpNet :93-122,
  ResNet :125-157, ResNetV2 :205-243.  Entries: input(lo,hi) = l2-normalised column slice (models.py:80-81, :85-86),
  fc(src,out,bias_init) = fully_connected (models.py:19-30), mul / add over earlier entries, l2norm = model output."""
  fc = lambda src, out, name: {"op": "fc", "src": src, "out": out, "bias_init": 0.1, "alpha": LEAKY_ALPHA, "name": name}
  vin = {"op": "input", "lo": 0, "hi": visual, "eps": L2_EPS}
  din = {"op": "input", "lo": visual, "hi": feature_size, "eps": L2_EPS}
  if name == "ResNetV2":
    spec = [vin, fc(0, 5000, "layer_visual_1_1"), fc(1, 256, "layer_visual_1_2"), fc(0, 256, "layer_visual_2_1"),
            din, fc(4, 400, "layer_doc_1_1"), fc(5, 256, "layer_doc_1_2"), fc(4, 256, "layer_doc_2_1"),
            {"op": "mul", "src": [2, 6]}, {"op": "mul", "src": [2, 7]}, {"op": "mul", "src": [3, 6]},
            {"op": "mul", "src": [3, 7]}, {"op": "add", "src": [8, 9, 10, 11, 2, 3, 6, 7]},
            fc(12, 256, "layer_fusion_1"), {"op": "add", "src": [12, 13]}, fc(14, 256, "layer_fusion_2"),
            {"op": "add", "src": [14, 15]}, {"op": "l2norm", "src": 16, "eps": L2_EPS}]
    return spec
  spec = [vin, fc(0, 5000, "layer_visual_1"), fc(1, 256, "layer_visual_2"),
          din, fc(3, 400, "layer_doc_1"), fc(4, 256, "layer_doc_2"), {"op": "mul", "src": [2, 5]}]
  if name == "MultiplyNet":
    spec += [{"op": "l2norm", "src": 6, "eps": L2_EPS}]
  elif name == "MlpNet":
    spec += [fc(6, 600, "layer_fusion_1"), fc(7, 256, "layer_fusion_2"), {"op": "l2norm", "src": 8, "eps": L2_EPS}]
  elif name == "ResNet":
    spec += [{"op": "add", "src": [6, 2, 5]}, fc(7, 256, "layer_fusion_1"), {"op": "add", "src": [7, 8]},
             fc(9, 256, "layer_fusion_2"), {"op": "add", "src": [9, 10]}, {"op": "l2norm", "src": 11, "eps": L2_EPS}]
  else:
    raise ValueError(name)
  return spec


def graph_widths(spec):
  w = []
  for e in spec:
    if e["op"] == "input":
      w.append(e["hi"] - e["lo"])
    elif e["op"] == "fc":
      w.append(e["out"])
    elif e["op"] in ("mul", "add"):
      w.append(w[e["src"][0]])
    else:
      w.append(w[e["src"]])
  return w


def init_graph(spec, seed: int = 2, dtype=np.float32):
  """Xavier-uniform weights / constant biases for every fc entry, in list order (= TF variable creation order)."""
  rng = np.random.RandomState(seed)
  w = graph_widths(spec)
  return [(xavier_uniform(rng, w[e["src"]], e["out"], dtype), np.full((e["out"],), e["bias_init"], dtype))
          for e in spec if e["op"] == "fc"]


def graph_forward(x, spec, params, dtype=np.float64, rnd=None):
  """Forward of an op list.  `rnd` (test infrastructure): rounding applied wherever the kernels store a 16-bit value
  (normalised inputs, shadow weights, every intermediate); None = exact arithmetic."""
  r = rnd or (lambda a: a)
  x = np.asarray(x, dtype)
  vals, li, out = [], 0, None
  for e in spec:
    if e["op"] == "input":
      vals.append(r(l2_normalize(x[:, e["lo"]:e["hi"]], e.get("eps", L2_EPS))))
    elif e["op"] == "fc":
      W, b = params[li]
      li += 1
      vals.append(r(fully_connected(vals[e["src"]], r(np.asarray(W, dtype)), np.asarray(b, dtype), e.get("alpha", LEAKY_ALPHA))))
    elif e["op"] == "mul":
      vals.append(r(vals[e["src"][0]] * vals[e["src"][1]]))
    elif e["op"] == "add":
      acc = vals[e["src"][0]]
      for s in e["src"][1:]:
        acc = r(acc + vals[s])
      vals.append(acc)
    else:
      y = vals[e["src"]]
      rinv = 1.0 / np.sqrt(np.maximum(np.sum(y * y, -1, keepdims=True), e.get("eps", L2_EPS)))
      vals.append(y * rinv)
      out = {"values": vals, "rinv": rinv[:, 0], "l2_norm": vals[-1]}
  return out


def graph_backward(fwd, spec, params, dE, dtype=np.float64, rnd=None):
  """Reverse sweep of the op list: [(dW, db)] per fc entry, from dL/d(l2_norm) [R,D].
  l2norm: dy = (g - e (e.g)) rinv ; fc: dz = dh leaky'(h), dW = in^T dz, db = sum dz, d_in = dz W^T ;
  mul: da = g b, db = g a ; add: every term receives g.  Nothing flows into the inputs."""
  r = rnd or (lambda a: a)
  vals = fwd["values"]
  g = [None] * len(spec)

  def acc(i, t):
    g[i] = t if g[i] is None else r(g[i] + t)

  fc_ids = [i for i, e in enumerate(spec) if e["op"] == "fc"]
  grads = [None] * len(fc_ids)
  e_out = fwd["l2_norm"]
  d = np.asarray(dE, dtype)
  acc(spec[-1]["src"], r((d - e_out * np.sum(e_out * d, -1, keepdims=True)) * fwd["rinv"][:, None]))
  for i in range(len(spec) - 2, -1, -1):
    e = spec[i]
    if e["op"] == "input" or g[i] is None:
      continue
    if e["op"] == "fc":
      li = fc_ids.index(i)
      dz = r(g[i] * np.where(vals[i] > 0, 1.0, e.get("alpha", LEAKY_ALPHA)))
      grads[li] = (vals[e["src"]].T @ dz, np.sum(dz, 0))
      if spec[e["src"]]["op"] != "input":
        acc(e["src"], r(dz @ r(np.asarray(params[li][0], dtype)).T))
    elif e["op"] == "mul":
      a, b = e["src"]
      acc(a, r(g[i] * vals[b]))
      acc(b, r(g[i] * vals[a]))
    else:
      for s in e["src"]:
        acc(s, g[i])
  return grads


def graph_grads_emulated16(x, spec, params, margin, kind="fp16", loss_scale=1.0):
  """PRECISION MODEL of the fusion-tower step (test infrastructure), the counterpart of tower_grads_emulated16:
  gradients of loss_scale * SUM of hinges flow backwards in 16 bits, 1/(B*loss_scale) is applied to the fp32 weight
  gradients."""
  rnd = lambda a: round16(a, kind)
  fwd = graph_forward(x, spec, params, np.float64, rnd)
  e = fwd["l2_norm"]
  E = e.reshape(-1, 3, e.shape[-1])
  B = E.shape[0]
  loss = hinge_loss(E, margin)
  g = hinge_loss_grad(E, margin).reshape(-1, e.shape[-1]) * (B * loss_scale)
  grads = graph_backward(fwd, spec, params, g, np.float64, rnd)
  return {"l2_norm": e, "loss": loss, "grads": [(gw / (B * loss_scale), gb / (B * loss_scale)) for gw, gb in grads]}


# --------------------------------------------------------------------------- #
# optimizer (train.py:82,108-113,146) -- TF1 AdamOptimizer                      #
# --------------------------------------------------------------------------- #
def exponential_decay(base_lr, global_step, decay_steps, decay_rate, staircase=True):
  """tf.train.exponential_decay (train.py:108-113)."""
  p = global_step / decay_steps
  if staircase:
    p = np.floor(p)
  return base_lr * decay_rate ** p


def adam_step_tf1(w, m, v, g, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
  """TF1 ApplyAdam; ``t`` is the 1-based step count.
  lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; w -= lr_t*m/(sqrt(v)+eps)."""
  lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
  m = beta1 * m + (1.0 - beta1) * g
  v = beta2 * v + (1.0 - beta2) * g * g
  w = w - lr_t * m / (np.sqrt(v) + eps)
  return w, m, v


def clip_by_norm(g, clip):
  """tf.clip_by_norm(t, clip) = t * clip / max(||t||_2, clip), applied per variable (train.py:47-64)."""
  n = float(np.sqrt(np.sum(np.asarray(g, np.float64) ** 2)))
  return g * (clip / max(n, clip))


def momentum_step_tf1(w, acc, g, lr, momentum=0.9, nesterov=True):
  """tf.train.MomentumOptimizer as build_graph constructs it (train.py:115-116; TF1 ApplyMomentum):
  acc <- momentum*acc + g ; nesterov: w <- w - lr*g - lr*momentum*acc ; plain: w <- w - lr*acc."""
  acc = momentum * acc + g
  w = w - lr * g - lr * momentum * acc if nesterov else w - lr * acc
  return w, acc


def lars_step_tf1(w, acc, g, lr, momentum=0.9, weight_decay=1e-4, eeta=1e-3, epsilon=0.0):
  """tf.contrib.opt.LARSOptimizer as of tensorflow r1.13 (the release README.md:20 pins; the optimizer of main(),
  train.py:354), restated from contrib/opt/python/training/lars_optimizer.py of that branch:
    compute_lr:   trust = eeta*|w| / (|g| + wd*|w| + eps) if |w| > 0 and |g| > 0 else 1 ; scaled_lr = lr*trust
    _apply_dense: training_ops.apply_momentum(var, mom, scaled_lr, grad, momentum, use_nesterov=False)
                  = acc <- momentum*acc + g ; w <- w - scaled_lr*acc
  weight_decay enters the trust ratio ONLY.  (Later releases add `grad + wd*var` and move scaled_lr inside the
  accumulator; that is not what 1.13.1 runs.)"""
  wn = float(np.sqrt(np.sum(np.asarray(w, np.float64) ** 2)))
  gn = float(np.sqrt(np.sum(np.asarray(g, np.float64) ** 2)))
  trust = eeta * wn / (gn + weight_decay * wn + epsilon) if (wn > 0 and gn > 0) else 1.0
  acc = momentum * acc + g
  return w - (lr * trust) * acc, acc


class OracleTrainer:
  """train.build_graph (train.py:74-146): final_loss = regularization_penalty * sum_l l2_penalty*|W_l|^2/2 + hinge loss,
  compute_gradients, optional per-variable clip_by_norm, apply_gradients with the staircase-decayed learning rate.
  Defaults = Trainer._build_model's constants (train.py:210-222): Adam, clip off, reg penalty 0, decay 0.96 / 1e6 steps."""

  def __init__(self, params, lr=1e-3, margin=0.8, decay_steps=1000000, decay=0.96, dtype=np.float64, optimizer="adam",
               clip_norm=0.0, reg_penalty=0.0, l2_penalty=1e-8, spec=None, emulate16=None):
    self.dtype = dtype
    self.spec = spec            # op list of a fusion tower (fusion_spec); None = the fully_connected chain
    self.emulate16 = emulate16  # "fp16" | "bf16": gradients from the 16-bit PRECISION MODEL (tower_grads_emulated16)
    self.params = [(np.asarray(W, dtype).copy(), np.asarray(b, dtype).copy()) for W, b in params]
    self.m = [(np.zeros_like(W), np.zeros_like(b)) for W, b in self.params]
    self.v = [(np.zeros_like(W), np.zeros_like(b)) for W, b in self.params]
    self.lr, self.margin, self.decay_steps, self.decay = lr, margin, decay_steps, decay
    self.optimizer, self.clip_norm, self.reg_penalty, self.l2_penalty = optimizer, clip_norm, reg_penalty, l2_penalty
    self.global_step = 0

  def loss_and_grads(self, x_rows):
    if self.emulate16 is not None and self.spec is None:
      # master weights stay float64/float32 here exactly as the kernels keep fp32 masters; only the step's operands round
      model = tower_grads_emulated16(x_rows, self.params, self.margin, self.emulate16)
      return {"l2_norm": model["l2_norm"]}, model["loss"], model["grads"]
    if self.spec is not None:
      fwd = graph_forward(x_rows, self.spec, self.params, dtype=self.dtype)
    else:
      fwd = tower_forward(x_rows, self.params, dtype=self.dtype)
    E = fwd["l2_norm"].reshape(-1, 3, fwd["l2_norm"].shape[-1])
    loss = hinge_loss(E, self.margin, self.dtype)
    dE = hinge_loss_grad(E, self.margin, self.dtype).reshape(-1, E.shape[-1])
    if self.spec is not None:
      return fwd, loss, graph_backward(fwd, self.spec, self.params, dE, dtype=self.dtype)
    return fwd, loss, tower_backward(fwd, self.params, dE, dtype=self.dtype)

  def reg_loss(self):
    """slim.l2_regularizer(l2_penalty)(W) = l2_penalty * sum(W^2) / 2, weights only (models.py:28)."""
    return float(sum(self.l2_penalty * np.sum(np.asarray(W, np.float64) ** 2) / 2 for W, _ in self.params))

  def _apply(self, w, m, v, g, lr, t):
    if self.clip_norm > 0:
      g = clip_by_norm(g, self.clip_norm)
    if self.optimizer == "adam":
      return adam_step_tf1(w, m, v, g, lr, t)
    if self.optimizer == "momentum":
      w2, m2 = momentum_step_tf1(w, m, g, lr)
      return w2, m2, v
    if self.optimizer == "lars":
      w2, m2 = lars_step_tf1(w, m, g, lr)
      return w2, m2, v
    if self.optimizer == "sgd":
      return w - lr * g, m, v
    raise ValueError(self.optimizer)

  def step(self, x_rows):
    fwd, loss, grads = self.loss_and_grads(x_rows)
    lr = exponential_decay(self.lr, self.global_step, self.decay_steps, self.decay)
    t = self.global_step + 1
    for li, ((W, b), (gW, gb)) in enumerate(zip(self.params, grads)):
      gW = gW + self.reg_penalty * self.l2_penalty * W          # d(reg_penalty * l2_penalty*|W|^2/2)/dW
      W2, mW, vW = self._apply(W, self.m[li][0], self.v[li][0], gW, lr, t)
      b2, mb, vb = self._apply(b, self.m[li][1], self.v[li][1], gb, lr, t)
      self.params[li], self.m[li], self.v[li] = (W2, b2), (mW, mb), (vW, vb)
    self.global_step = t
    return float(loss["hinge_loss"]), fwd, loss


# --------------------------------------------------------------------------- #
# in-batch semi-hard negative mining (SURVEY.md 8a row M -- build-defined)      #
# --------------------------------------------------------------------------- #
def mine_semihard(E, guid_triplets, margin, dtype=np.float64):
  """E [3B,D] embeddings (rows a0,p0,n0,...), guid_triplets [B,3].

  Candidates for triplet i: the positive and negative rows of every triplet
  (row index r = 3j+1, 3j+2) whose guid is not a_i or p_i.  With
  dp = |a_i-p_i|^2 and d_r = |a_i-E_r|^2 choose
    (1) argmin d_r with dp < d_r < dp+margin (semi-hard), else
    (2) argmin d_r with d_r >= dp+margin (easiest-to-violate beyond), else
    (3) keep the reader's own negative row 3i+2.
  Ties -> lowest row index.  Returns (neg_row [B] int32, d_an [B])."""
  E = np.asarray(E, dtype)
  g = np.asarray(guid_triplets)
  B = g.shape[0]
  cand_rows = np.stack([3 * np.arange(B) + 1, 3 * np.arange(B) + 2], 1).reshape(-1)
  cand_guid = g[:, 1:3].reshape(-1)
  C = E[cand_rows]
  A = E[0::3]
  P = E[1::3]
  dp = np.sum((A - P) ** 2, -1)
  neg_row = np.empty(B, np.int32)
  d_an = np.empty(B, dtype)
  cn = np.sum(C * C, -1)
  for i in range(B):
    d = np.sum(A[i] * A[i]) + cn - 2.0 * (C @ A[i])
    ok = (cand_guid != g[i, 0]) & (cand_guid != g[i, 1])
    semi = ok & (d > dp[i]) & (d < dp[i] + margin)
    beyond = ok & (d >= dp[i] + margin)
    if semi.any():
      j = np.flatnonzero(semi)[np.argmin(d[semi])]
    elif beyond.any():
      j = np.flatnonzero(beyond)[np.argmin(d[beyond])]
    else:
      neg_row[i] = 3 * i + 2
      d_an[i] = np.sum((A[i] - E[3 * i + 2]) ** 2)
      continue
    neg_row[i] = cand_rows[j]
    d_an[i] = d[j]
  return neg_row, d_an


def mine_semihard_emulated16(E, guid_triplets, margin, kind="fp16"):
  """PRECISION MODEL of cdml_mine_semihard (test infrastructure): the same definition as mine_semihard, with the
  selection arithmetic the kernel uses -- scores s = a16 . c16 on operands rounded to 16 bits (accumulated here in
  float64, rounded once to fp32; the tensor core accumulates in fp32), dp = |a-p|^2 from the fp32 embeddings,
  s_hi = 1 - dp/2, d = max(fma(-2, s, 2), 0) in fp32, valid = (s < s_hi) & (d > dp) & guid not in {a_i, p_i};
  pick = argmin over (d, row).  (1)/(2) of the definition collapse into that one criterion (csrc/mine.cu).
  Returns (neg_row [B] int32, d_sel [B] fp32 selection distances, runner_up_gap [B] fp32: distance between the two best
  valid selection distances -- picks whose gap is at fp32 accumulation noise may legitimately differ)."""
  E32 = np.asarray(E, np.float32)
  g = np.asarray(guid_triplets)
  B = g.shape[0]
  E16 = round16(E32, kind)
  A16 = E16[0::3]
  dpf = np.sum((E32[0::3] - E32[1::3]) ** 2, -1, dtype=np.float32)
  s_hi = (np.float32(1.0) - np.float32(0.5) * dpf).astype(np.float32)
  cand_rows = np.stack([3 * np.arange(B) + 1, 3 * np.arange(B) + 2], 1).reshape(-1)
  cand_guid = g[:, 1:3].reshape(-1)
  C16 = E16[cand_rows]
  neg_row = np.empty(B, np.int32)
  d_sel = np.full(B, np.inf, np.float32)
  gap = np.full(B, np.inf, np.float32)
  for s in range(0, B, 1024):
    S = (A16[s:s + 1024] @ C16.T).astype(np.float32)
    d = np.maximum(np.float32(2.0) - np.float32(2.0) * S, np.float32(0.0)).astype(np.float32)
    ok = (S < s_hi[s:s + 1024, None]) & (d > dpf[s:s + 1024, None])
    ok &= (cand_guid[None, :] != g[s:s + 1024, 0:1]) & (cand_guid[None, :] != g[s:s + 1024, 1:2])
    dm = np.where(ok, d, np.inf)
    for i in range(dm.shape[0]):
      row = dm[i]
      m = row.min()
      if not np.isfinite(m):
        neg_row[s + i] = 3 * (s + i) + 2
        continue
      j = np.flatnonzero(row == m)
      neg_row[s + i] = cand_rows[j].min()                     # ties -> lowest row
      d_sel[s + i] = m
      rest = row[row > m]
      gap[s + i] = 0.0 if len(j) > 1 else (rest.min() - m if len(rest) else np.inf)
  return neg_row, d_sel, gap


# --------------------------------------------------------------------------- #
# exact flat KNN (faiss_knn.py:98-131 with IndexFlatL2 / IndexFlatIP semantics) #
# --------------------------------------------------------------------------- #
def knn_normalize(x: np.ndarray) -> np.ndarray:
  """faiss_knn.py:99-104 -- rows /= ||row||_2 in float32 (no epsilon)."""
  x = x.astype(np.float32)
  return x / np.linalg.norm(x, axis=1, keepdims=True)


def flat_knn(xb, xq=None, k=51, l2_norm=True, metric="L2", block=4096, dtype=np.float32):
  """calc_knn with an exact flat index.  D [nq,k] squared L2 ascending (or inner
  product descending for metric="IP"), I [nq,k] int64.  Distances follow faiss's
  BLAS path ||q||^2 + ||x||^2 - 2 q.x clamped at 0; ties -> lower id first.
  Fewer than k database rows pads with +inf / -1 like faiss."""
  xb = np.asarray(xb, np.float32)
  if l2_norm:
    xb = knn_normalize(xb)
    xq = xb if xq is None else knn_normalize(np.asarray(xq, np.float32))
  elif xq is None:
    xq = xb
  xb = xb.astype(dtype)
  xq = np.asarray(xq, dtype)
  nq, N = xq.shape[0], xb.shape[0]
  kk = min(k, N)
  D = np.full((nq, k), np.inf if metric == "L2" else -np.inf, np.float32)
  I = np.full((nq, k), -1, np.int64)
  bn = np.sum(xb * xb, 1)
  for s in range(0, nq, block):
    q = xq[s:s + block]
    ip = q @ xb.T
    if metric == "L2":
      d = np.maximum(np.sum(q * q, 1)[:, None] + bn[None, :] - 2.0 * ip, 0)
      key = d
    else:
      d = ip
      key = -ip
    # stable: sort by (key, id)
    part = np.argpartition(key, kk - 1, axis=1)[:, :kk] if kk < N else np.tile(np.arange(N), (len(q), 1))
    pk = np.take_along_axis(key, part, 1)
    order = np.lexsort((part, pk), axis=1)
    idx = np.take_along_axis(part, order, 1)
    D[s:s + block, :kk] = np.take_along_axis(d, idx, 1)
    I[s:s + block, :kk] = idx
  return D, I


def exact_ip_nn(embeddings, query_row, k):
  """show_knn.py:63-68 -- independent statement: argsort(-(E @ e_q))[:k]."""
  sims = embeddings @ embeddings[query_row]
  return np.argsort(-sims, kind="stable")[:k]


def knn_merge(D_parts, I_parts, k, metric="L2"):
  """Sharded-index merge: parts [G,nq,k] -> global top-k,
  ties -> lower id.  Ids already carry their shard offset."""
  D = np.concatenate(list(D_parts), axis=1)
  I = np.concatenate(list(I_parts), axis=1)
  key = D if metric == "L2" else -D
  key = np.where(I < 0, np.inf, key)
  order = np.lexsort((I, key), axis=1)[:, :k]
  return np.take_along_axis(D, order, 1), np.take_along_axis(I, order, 1)


# --------------------------------------------------------------------------- #
# result writer (faiss_knn.py:267-283) and eval (evaluate.py:57-73)            #
# --------------------------------------------------------------------------- #
def f2key(x):
  """Order-preserving fp32 -> uint32 map of the KNN kernels (csrc/knn.cu f2key): negative floats complemented, others get
  the top bit."""
  u = np.ascontiguousarray(x, np.float32).view(np.uint32)
  return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)


def key2f(k):
  k = np.asarray(k, np.uint32)
  return np.where(k & np.uint32(0x80000000), k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32).view(np.float32)


def knn_pack_records(D, I, metric="L2"):
  """(D, I) lists -> the 64-bit records a shard sends (cdml_knn_shard_refine): (f2key(distance) << 32) | global id,
  padding (id < 0) = all ones.  Sorting records ascending sorts by (distance, id) -- descending inner product for IP."""
  D, I = np.asarray(D, np.float32), np.asarray(I, np.int64)
  key = f2key(D if metric == "L2" else -D).astype(np.uint64)
  rec = (key << np.uint64(32)) | (I.astype(np.uint64) & np.uint64(0xFFFFFFFF))
  return np.where(I < 0, np.uint64(0xFFFFFFFFFFFFFFFF), rec)


def knn_merge_records(rec, k, metric="L2"):
  """[G,nq,k] records -> global top-k (cdml_knn_merge_packed): ascending sort of the union, first k, unpacked."""
  rec = np.asarray(rec, np.uint64)
  G, nq, kk = rec.shape
  allr = np.sort(np.transpose(rec, (1, 0, 2)).reshape(nq, G * kk), axis=1)[:, :k]
  pad = allr == np.uint64(0xFFFFFFFFFFFFFFFF)
  d = key2f((allr >> np.uint64(32)).astype(np.uint32))
  D = np.where(pad, np.inf if metric == "L2" else -np.inf, d if metric == "L2" else -d).astype(np.float32)
  I = np.where(pad, -1, (allr & np.uint64(0xFFFFFFFF)).astype(np.int64))
  return D, I


def format_knn_rows(begin_index, D, I, decode_map):
  """write_process: '<query_guid>,<nbr_guid>#<dist><...\\n'; skips column 0 and
  keeps a neighbour only if idx > 0 and 0.0 < dist < 1.4; dist via str(np.float32)."""
  lines = []
  for i in range(I.shape[0]):
    topks = "".join(
      decode_map[int(j)] + "#" + str(np.float32(d)) + "<"
      if (j > 0 and d > 0.0 and d < 1.4) else ""
      for j, d in zip(I[i][1:], D[i][1:]))
    lines.append(decode_map[begin_index + i] + "," + topks + "\n")
  return lines


def split_ranges(total, split_num):
  """write_knn patching (faiss_knn.py:288-301): split_num contiguous patches,
  the last takes the remainder."""
  patch = total // split_num
  return [(i * patch, (i + 1) * patch if i < split_num - 1 else total) for i in range(split_num)]


def mean_dist(vectors, cowatches):
  """evaluate.py:57-73 -- mean squared L2 over cowatch pairs."""
  co = np.asarray(vectors)[np.asarray(cowatches)]
  return np.mean(np.sum((co[:, 0, :] - co[:, 1, :]) ** 2, axis=-1))


def rencode_eval(features, cowatches):
  """evaluate.py:34-55 -- compact eval guids to 0..U-1 in sorted order."""
  uniq = np.unique(np.asarray(cowatches).reshape(-1))
  remap = {int(o): i for i, o in enumerate(uniq)}
  return features[uniq], [[remap[int(a)], remap[int(b)]] for a, b in cowatches]


# --------------------------------------------------------------------------- #
# on-device triplet reader (inputs.py:102-142 with a counter-based generator)   #
# --------------------------------------------------------------------------- #
def philox4x32_10(ctr, key):
  """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).  ctr: 4 arrays
  of uint32 values, key: 2 ints.  Pinned in tests to the Random123 known-answer vectors."""
  c = [np.asarray(x, np.uint64) for x in ctr]
  k = [np.uint64(key[0]), np.uint64(key[1])]
  mask = np.uint64(0xFFFFFFFF)
  for _ in range(10):
    p0 = np.uint64(0xD2511F53) * c[0]
    p1 = np.uint64(0xCD9E8D57) * c[2]
    c = [((p1 >> np.uint64(32)) ^ c[1] ^ k[0]) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c[3] ^ k[1]) & mask, p0 & mask]
    k = [(k[0] + np.uint64(0x9E3779B9)) & mask, (k[1] + np.uint64(0xBB67AE85)) & mask]
  return c


def sample_triplets_device(pairs, start, B, num_guid, seed):
  """The device reader's statement of inputs.py:112-130: (a,p) = pairs[(start+i) % n]; n = randint(0,G) re-drawn while
  n in {a,p}.  randint = word 0 of Philox4x32-10(key=seed, counter=(position, attempt)) mapped to [0,G) by Lemire's
  multiply-shift with its rejection step (exactly uniform)."""
  pairs = np.asarray(pairs, np.int64)
  pos = np.arange(start, start + B, dtype=np.uint64)
  ap = pairs[(pos % np.uint64(len(pairs))).astype(np.int64)]
  G = np.uint64(num_guid)
  threshold = np.uint64((2 ** 32) % int(num_guid))
  neg = np.full(B, -1, np.int64)
  attempt = np.zeros(B, np.uint64)
  todo = np.arange(B)
  while len(todo):
    x = philox4x32_10((pos[todo] & np.uint64(0xFFFFFFFF), pos[todo] >> np.uint64(32), attempt[todo], np.zeros(len(todo), np.uint64)),
                      (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))[0]
    m = x * G
    cand = (m >> np.uint64(32)).astype(np.int64)
    ok = ((m & np.uint64(0xFFFFFFFF)) >= threshold) & (cand != ap[todo, 0]) & (cand != ap[todo, 1])
    neg[todo[ok]] = cand[ok]
    attempt[todo] += np.uint64(1)
    todo = todo[~ok]
  return np.concatenate([ap, neg[:, None]], axis=1)


# --------------------------------------------------------------------------- #
# de-similarity post-filter (faiss_knn.py:134-244)                              #
# --------------------------------------------------------------------------- #
def desim_simple(eI, fI):
  """faiss_knn.desim (faiss_knn.py:134-143): eI[i, j] = -1 wherever eI[i, j] occurs in row i of fI."""
  eI = np.array(eI, np.int64)
  for i in range(eI.shape[0]):
    eI[i][np.isin(eI[i], fI[i])] = -1
  return eI


def filter_fI(fI, fD, fD_threshold=1.4):
  """faiss_knn.fliter_fI (faiss_knn.py:146-155): feature neighbours farther than the threshold, and the row itself, -> -1."""
  fI = np.array(fI, np.int64)
  fI[np.asarray(fD) > fD_threshold] = -1
  fI[fI == np.arange(fI.shape[0])[:, None]] = -1
  return fI


def iter_desim(eI, fI, fD, fD_threshold=1.4, fI_end=31, row_offset=0):
  """faiss_knn.iter_desim_mp (faiss_knn.py:187-244) without its +1 index shift, zero row and process pool, which only
  serve the vectorised column sweep.  Row by row (rows are independent): walk the columns left to right; an entry that is
  still alive is a pivot v, and every LATER entry of the row that occurs among v's first fI_end filtered feature
  neighbours (filter_fI) is dropped (-1) and never becomes a pivot (faiss_knn.py:176-184, :203-233).  Finally the row's
  own id is dropped (faiss_knn.py:236-238).  -1 entries (short KNN lists) stay -1."""
  eI = np.array(eI, np.int64)
  F = filter_fI(fI, fD, fD_threshold)[:, :fI_end]
  for r in range(eI.shape[0]):
    row = eI[r]
    for c in range(row.shape[0]):
      v = row[c]
      if v < 0:
        continue
      near = F[v][F[v] >= 0]
      tail = row[c + 1:]
      tail[np.isin(tail, near)] = -1
    row[row == r + row_offset] = -1          # eI may be a slice of the rows starting at row_offset
  return eI


# --------------------------------------------------------------------------- #
# seeded synthetic inputs (imitation_data.py:41-53 but seeded)     #
# --------------------------------------------------------------------------- #
def synth_features(num_guid, feature_size, seed=0, decimals=8):
  """imitation_data.gen_features: around(U[0,1), 8) -- here with RandomState(seed), float32."""
  return np.around(np.random.RandomState(seed).random_sample((num_guid, feature_size)), decimals).astype(np.float32)


def synth_pairs(num_pairs, num_guid, seed=1):
  """cowatch pairs a != p, uniform over guids."""
  rng = np.random.RandomState(seed)
  a = rng.randint(0, num_guid, size=num_pairs)
  p = (a + 1 + rng.randint(0, num_guid - 1, size=num_pairs)) % num_guid
  return np.stack([a, p], 1).astype(np.int64)


def synth_triplets(num, num_guid, seed=1):
  """[num,3] int64 with n not in {a,p} (vectorised equivalent of inputs.py:123-129
  in distribution; use sample_negatives for stream-exact draws)."""
  rng = np.random.RandomState(seed)
  ap = synth_pairs(num, num_guid, seed)
  n = rng.randint(0, num_guid, size=num)
  bad = (n == ap[:, 0]) | (n == ap[:, 1])
  while bad.any():
    n[bad] = rng.randint(0, num_guid, size=int(bad.sum()))
    bad = (n == ap[:, 0]) | (n == ap[:, 1])
  return np.concatenate([ap, n[:, None]], 1).astype(np.int64)
