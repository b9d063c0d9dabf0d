"""Tower definitions -- host-side mirror of the reference's models.py (same class names, same
`create_model(model_input, output_size=256) -> {"l2_norm": ...}` plug-in interface, models.py:33-62).

The reference builds TF graph ops; here the same calls record a small symbolic chain (placeholder -> l2_normalize ->
fully_connected* -> l2_normalize) that `train.build_graph` / `predict.Prediction` compile into a `TowerEngine`
running on libcdml's tcgen05 kernels.  The visual+doc fusion towers (models.py:65-243) record a small
DAG (column slices, multiply, residual adds) that `compile_graph` lowers for `fusion.GraphEngine`.
"""
import numpy as np

LEAKY_ALPHA = 0.2  # tf.nn.leaky_relu default (models.py:21)


class Node:
  """One recorded op of the tower graph.  `src` is a Node (input -> slice / l2norm / fc chains) or a tuple of Nodes
  (mul, add).  Nodes are numbered in creation order, which is a topological order of the graph."""
  _serial = 0

  def __init__(self, kind, src=None, name=None, **attrs):
    self.kind, self.src, self.name, self.attrs = kind, src, name, attrs
    Node._serial += 1
    self.serial = Node._serial

  @property
  def width(self):
    if self.kind == "input":
      return self.attrs["width"]
    if self.kind == "fc":
      return self.attrs["output_size"]
    if self.kind == "slice":
      return self.attrs["hi"] - self.attrs["lo"]
    if self.kind in ("mul", "add"):
      return self.src[0].width
    return self.src.width

  def __getitem__(self, key):
    """model_input[:, lo:hi] -- the visual / doc column split of the fusion towers (models.py:80, :85)."""
    if not (isinstance(key, tuple) and len(key) == 2 and key[0] == slice(None) and isinstance(key[1], slice)
            and key[1].step in (None, 1)):
      raise NotImplementedError("only column slices x[:, lo:hi] are recorded")
    lo, hi, _ = key[1].indices(self.width)
    if hi <= lo:
      raise ValueError("empty column slice [%d:%d] of a %d-wide tensor" % (lo, hi, self.width))
    return Node("slice", src=self, lo=lo, hi=hi)

  def __add__(self, other):
    return add(self, other)

  def __mul__(self, other):
    return multiply(self, other)

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


def multiply(a, b, name=None):
  """tf.multiply of two equally wide tensors (models.py:89)."""
  a, b = _as_node(a), _as_node(b)
  if a.width != b.width:
    raise ValueError("multiply: widths differ (%d vs %d)" % (a.width, b.width))
  return Node("mul", src=(a, b), name=name)


def add(a, b, name=None):
  """Elementwise a + b (the residual connections, models.py:148-152)."""
  a, b = _as_node(a), _as_node(b)
  if a.width != b.width:
    raise ValueError("add: widths differ (%d vs %d)" % (a.width, b.width))
  return Node("add", src=(a, b), name=name)


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
    if isinstance(n.src, tuple) or n.kind == "slice":
      raise NotImplementedError("%r joins or slices tensors: not a plain chain (see compile_graph)" % n)
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


def compile_graph(out_node):
  """General form of compile_chain for the fusion towers: walks back from the model output and returns
  dict(F=input width, D=output width, value=bound input or None, spec=[...]) where `spec` lists the ops in creation
  (= topological) order, each a plain dict whose `src` fields index into the list:
    {"op":"input","lo","hi","eps"}       l2-normalised column slice of the raw input
    {"op":"fc","src","out","bias_init","alpha","l2_penalty","name"}
    {"op":"mul","src":[i,j]} | {"op":"add","src":[i,j,...]}       (nested single-use adds are flattened)
    {"op":"l2norm","src","eps"}          the model output; always last."""
  if out_node.kind != "l2norm":
    raise NotImplementedError("the model output must be an l2_normalize (models.py:61)")
  seen, order, stack = {}, [], [out_node]
  while stack:
    n = stack.pop()
    if id(n) in seen:
      continue
    seen[id(n)] = n
    order.append(n)
    if n.src is not None:
      stack.extend(n.src if isinstance(n.src, tuple) else (n.src,))
  order.sort(key=lambda n: n.serial)
  uses = {}
  for n in order:
    for s in (n.src if isinstance(n.src, tuple) else (n.src,) if n.src is not None else ()):
      uses[id(s)] = uses.get(id(s), 0) + 1
  roots = [n for n in order if n.kind == "input"]
  if len(roots) != 1:
    raise NotImplementedError("a tower has exactly one input placeholder")
  root = roots[0]
  index, spec = {}, []

  def emit(n, entry):
    index[id(n)] = len(spec)
    spec.append(entry)

  def add_terms(n):
    out = []
    for s in n.src:
      if s.kind == "add" and uses.get(id(s), 0) == 1:
        out.extend(add_terms(s))
      else:
        out.append(s)
    return out

  for n in order:
    if n.kind in ("input", "slice"):
      if n.kind == "slice" and n.src is not root:
        raise NotImplementedError("column slices are recorded on the input placeholder only")
      continue
    if n.kind == "l2norm" and n is not out_node:
      src = n.src
      if src is root:
        lo, hi = 0, root.width
      elif src.kind == "slice":
        lo, hi = src.attrs["lo"], src.attrs["hi"]
      else:
        raise NotImplementedError("an inner l2_normalize is only recorded on (a column slice of) the input")
      emit(n, {"op": "input", "lo": lo, "hi": hi, "eps": n.attrs["epsilon"]})
      continue
    if n.kind == "add" and uses.get(id(n), 0) == 1 and any(
        m.kind == "add" and n in m.src for m in order):
      continue   # folded into its only consumer, an add
    srcs = add_terms(n) if n.kind == "add" else (n.src if isinstance(n.src, tuple) else (n.src,))
    for s in srcs:
      if id(s) not in index:
        raise NotImplementedError("%r consumes %r, which is not l2-normalised input, fc, mul or add" % (n, s))
    ids = [index[id(s)] for s in srcs]
    if n.kind == "fc":
      emit(n, {"op": "fc", "src": ids[0], "out": n.attrs["output_size"], "bias_init": n.attrs["bias_init"],
               "alpha": n.attrs["alpha"], "l2_penalty": n.attrs["l2_penalty"], "name": n.name})
    elif n.kind in ("mul", "add"):
      emit(n, {"op": n.kind, "src": ids})
    else:
      emit(n, {"op": "l2norm", "src": ids[0], "eps": n.attrs["epsilon"]})
  if len({e["alpha"] for e in spec if e["op"] == "fc"}) > 1:
    raise NotImplementedError("all layers must share one leaky slope")
  return {"F": root.width, "D": out_node.width, "value": root.attrs.get("value"), "spec": spec}


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
  """Visual+doc fusion towers (models.py:65-243): the 1500-wide visual block and the doc block (the remaining columns,
  128 in production: feature_size 1628, online_data.py:38) are embedded by two fully_connected stacks and joined."""
  visual_width = 1500          # models.py:80 `model_input[:,:1500]`

  def _blocks(self, model_input, vname=("layer_visual_1", "layer_visual_2"), dname=("layer_doc_1", "layer_doc_2")):
    model_input = _as_node(model_input)
    visual_input = l2_normalize(model_input[:, :self.visual_width], axis=-1, name="visual_input")
    layer_visual_1 = fully_connected(visual_input, 5000, bias_init=0.1, name=vname[0])
    layer_visual_2 = fully_connected(layer_visual_1, 256, bias_init=0.1, name=vname[1])
    doc_input = l2_normalize(model_input[:, self.visual_width:], axis=-1, name="doc_input")
    layer_doc_1 = fully_connected(doc_input, 400, bias_init=0.1, name=dname[0])
    layer_doc_2 = fully_connected(layer_doc_1, 256, bias_init=0.1, name=dname[1])
    return visual_input, doc_input, layer_visual_2, layer_doc_2


class MultiplyNet(_FusionTower):
  """Fusion by multiply (models.py:65-91)."""

  def create_model(self, model_input, output_size=256):
    _, _, layer_visual_2, layer_doc_2 = self._blocks(model_input)
    layer_fusion = multiply(layer_visual_2, layer_doc_2, name="multiply_fusion")
    return {"l2_norm": l2_normalize(layer_fusion, axis=-1, name="model_output")}


class MlpNet(_FusionTower):
  """Fusion by multiply, then a 600 -> 256 MLP (models.py:93-122)."""

  def create_model(self, model_input, output_size=256):
    _, _, layer_visual_2, layer_doc_2 = self._blocks(model_input)
    layer_fusion = multiply(layer_visual_2, layer_doc_2, name="multiply_fusion")
    layer_fusion_1 = fully_connected(layer_fusion, 600, bias_init=0.1, name="layer_fusion_1")
    layer_fusion_2 = fully_connected(layer_fusion_1, 256, bias_init=0.1, name="layer_fusion_2")
    return {"l2_norm": l2_normalize(layer_fusion_2, axis=-1, name="model_output")}


class ResNet(_FusionTower):
  """The production tower of train.py main() (train.py:363): multiply fusion + two residual 256 -> 256 layers
  (models.py:125-157)."""

  def create_model(self, model_input, output_size=256):
    _, _, layer_visual_2, layer_doc_2 = self._blocks(model_input)
    layer_fusion = multiply(layer_visual_2, layer_doc_2, name="multiply_fusion")
    layer_res_1 = layer_fusion + layer_visual_2 + layer_doc_2
    layer_fusion_1 = fully_connected(layer_res_1, 256, bias_init=0.1, name="layer_fusion_1")
    layer_res_2 = layer_res_1 + layer_fusion_1
    layer_fusion_2 = fully_connected(layer_res_2, 256, bias_init=0.1, name="layer_fusion_2")
    layer_res_3 = layer_res_2 + layer_fusion_2
    return {"l2_norm": l2_normalize(layer_res_3, axis=-1, name="model_output")}


class DenseNet(_FusionTower):
  """models.py:160-203 cannot be constructed in the reference either: it asks tf.get_variable for a variable whose shape
  holds the tensor tf.shape(...) (models.py:196-198), and its docstring calls it a debugging network.  The name resolves
  for `find_class_by_name`; instantiating the graph says so."""

  def create_model(self, model_input, output_size=256):
    raise NotImplementedError("DenseNet (models.py:160-203) is not constructible in the reference (variable shape taken "
                              "from tf.shape, models.py:196-198); use ResNet")


class ResNetV2(_FusionTower):
  """Wider ResNet: two extra shallow branches, four cross products, residual MLP (models.py:205-243)."""

  def create_model(self, model_input, output_size=256):
    model_input = _as_node(model_input)
    visual_input = l2_normalize(model_input[:, :self.visual_width], axis=-1, name="visual_input")
    layer_visual_1_1 = fully_connected(visual_input, 5000, bias_init=.1, name="layer_visual_1_1")
    layer_visual_1_2 = fully_connected(layer_visual_1_1, 256, bias_init=.1, name="layer_visual_1_2")
    layer_visual_2_1 = fully_connected(visual_input, 256, bias_init=.1, name="layer_visual_2_1")
    doc_input = l2_normalize(model_input[:, self.visual_width:], axis=-1, name="doc_input")
    layer_doc_1_1 = fully_connected(doc_input, 400, bias_init=.1, name="layer_doc_1_1")
    layer_doc_1_2 = fully_connected(layer_doc_1_1, 256, bias_init=.1, name="layer_doc_1_2")
    layer_doc_2_1 = fully_connected(doc_input, 256, bias_init=.1, name="layer_doc_2_1")
    layer_fusion_1 = multiply(layer_visual_1_2, layer_doc_1_2, "multiply_fusion_1")
    layer_fusion_2 = multiply(layer_visual_1_2, layer_doc_2_1, "multiply_fusion_2")
    layer_fusion_3 = multiply(layer_visual_2_1, layer_doc_1_2, "multiply_fusion_3")
    layer_fusion_4 = multiply(layer_visual_2_1, layer_doc_2_1, "multiply_fusion_4")
    layer_res_1 = layer_fusion_1 + layer_fusion_2 + layer_fusion_3 + layer_fusion_4 + \
                  layer_visual_1_2 + layer_visual_2_1 + layer_doc_1_2 + layer_doc_2_1
    layer_fusion_1 = fully_connected(layer_res_1, 256, name="layer_fusion_1")
    layer_res_2 = layer_res_1 + layer_fusion_1
    layer_fusion_2 = fully_connected(layer_res_2, 256, name="layer_fusion_2")
    layer_res_3 = layer_res_2 + layer_fusion_2
    return {"l2_norm": l2_normalize(layer_res_3, axis=-1, name="model_output")}
