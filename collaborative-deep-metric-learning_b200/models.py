"""Tower definitions -- host-side mirror of the reference's models.py (same class names, same
`create_model(model_input, output_size=256) -> {"l2_norm": ...}` plug-in interface, models.py:33-62).

The reference builds TF graph ops; here the same calls record a small symbolic chain (placeholder -> l2_normalize ->
fully_connected* -> l2_normalize) that `train.build_graph` / `predict.Prediction` compile into a `TowerEngine`
running on libcdml's tcgen05 kernels.  Towers outside the hot-path scope (the visual+doc fusion nets)
keep their names so `find_class_by_name` resolves them, and say so when instantiated.
"""
import numpy as np

LEAKY_ALPHA = 0.2  # tf.nn.leaky_relu default (models.py:21)


class Node:
  """One recorded op of the tower chain."""

  def __init__(self, kind, src=None, name=None, **attrs):
    self.kind, self.src, self.name, self.attrs = kind, src, name, attrs

  @property
  def width(self):
    if self.kind == "input":
      return self.attrs["width"]
    if self.kind == "fc":
      return self.attrs["output_size"]
    return self.src.width

  def __repr__(self):
    return "Node(%s%s)" % (self.kind, "" if self.name is None else ":" + self.name)


def placeholder(width, name="input_batch"):
  """tf.placeholder(tf.float32, shape=(None, width), name=...) (train.py:265)."""
  return Node("input", name=name, width=int(width))


def leaky_relu(alpha=LEAKY_ALPHA):
  return ("leaky_relu", float(alpha))


def l2_normalize(x, axis=-1, name=None, epsilon=1e-12):
  """tf.nn.l2_normalize along the feature axis (models.py:58, :61)."""
  if axis not in (-1, 1):
    raise ValueError("only the feature axis can be normalised")
  return Node("l2norm", src=_as_node(x), name=name, epsilon=float(epsilon))


def fully_connected(input_tensor, output_size, activation_fn=None, l2_penalty=1e-8, bias_init=0.1, name=None):
  """slim.fully_connected with leaky-ReLU, constant bias initialiser and an (unused, penalty x0) L2 regulariser
  (models.py:19-30).  Weights are [in,out], Xavier-uniform."""
  act = leaky_relu() if activation_fn is None else activation_fn
  if not (isinstance(act, tuple) and act[0] == "leaky_relu"):
    raise NotImplementedError("the B200 epilogues implement leaky_relu only (models.py:21)")
  return Node("fc", src=_as_node(input_tensor), name=name, output_size=int(output_size), alpha=act[1],
              l2_penalty=float(l2_penalty), bias_init=float(bias_init))


def _as_node(x):
  if isinstance(x, Node):
    return x
  arr = np.asarray(x) if not hasattr(x, "shape") else x
  return Node("input", name="input_batch", width=int(arr.shape[-1]), value=x)


def compile_chain(out_node):
  """Walk back from the model output; returns dict(dims=[F,H..,D], bias_init=[..], alpha, value=bound input or None)."""
  chain = []
  n = out_node
  while n is not None:
    chain.append(n)
    n = n.src
  chain.reverse()
  kinds = [c.kind for c in chain]
  if len(chain) < 4 or kinds[0] != "input" or kinds[1] != "l2norm" or kinds[-1] != "l2norm" or \
     any(k != "fc" for k in kinds[2:-1]):
    raise NotImplementedError("tower topology %s is outside the B200 hot path (expected input -> l2norm -> fc+ -> l2norm)" % kinds)
  fcs = chain[2:-1]
  alphas = {f.attrs["alpha"] for f in fcs}
  if len(alphas) != 1:
    raise NotImplementedError("all layers must share one leaky slope")
  return {"dims": [chain[0].width] + [f.attrs["output_size"] for f in fcs],
          "bias_init": [f.attrs["bias_init"] for f in fcs], "alpha": alphas.pop(),
          "value": chain[0].attrs.get("value"), "names": [f.name for f in fcs],
          "l2_penalty": [f.attrs["l2_penalty"] for f in fcs]}


class BaseModel(object):
  """Inherit from this class when implementing new models (models.py:33-38)."""

  def create_model(self, unused_model_input, **unused_params):
    raise NotImplementedError()


class VNet(BaseModel):
  """Visual feature network: l2norm -> FC 5000 -> FC output_size -> l2norm (models.py:41-62)."""
  hidden = (5000,)
  bias_init = 0.0

  def create_model(self, model_input, output_size=256):
    x = l2_normalize(model_input, axis=-1, name="model_input")
    layers = {}
    h = x
    for i, width in enumerate(self.hidden):
      h = fully_connected(h, width, bias_init=self.bias_init)
      layers["layer_%d" % (i + 1)] = h
    h = fully_connected(h, output_size, bias_init=self.bias_init)
    layers["layer_%d" % (len(self.hidden) + 1)] = h
    out = l2_normalize(h, axis=-1, name="model_output")
    layers["l2_norm"] = out
    return layers


class WideNet(VNet):
  """BASELINE config 3: 2048-d input, 3 x 2048 hidden, 256-d output, same fully_connected semantics."""
  hidden = (2048, 2048, 2048)
  bias_init = 0.0


class _FusionTower(BaseModel):
  """Visual+doc fusion towers of the reference (models.py:65-243) -- SURVEY.md 8(f) 'next', not built this round."""
  _where = "models.py"

  def create_model(self, model_input, output_size=256):
    raise NotImplementedError("%s (%s) is a fusion tower outside this round's hot-path scope; use VNet"
                              % (type(self).__name__, self._where))


class MultiplyNet(_FusionTower):
  _where = "models.py:65-91"


class MlpNet(_FusionTower):
  _where = "models.py:93-122"


class ResNet(_FusionTower):
  _where = "models.py:125-157"


class DenseNet(_FusionTower):
  _where = "models.py:160-203"


class ResNetV2(_FusionTower):
  _where = "models.py:205-243"
