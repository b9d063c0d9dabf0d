"""Loss plug-ins -- host-side mirror of the reference's losses.py (BaseLoss / HingeLoss, losses.py:4-49).
`calculate_loss(triplets, margin)` takes [B,3,D] embeddings (numpy or CUDA torch tensor) and returns the same
7-key dict; the arithmetic runs in libcdml's fused hinge kernel (cdml_triplet_hinge)."""
import numpy as np
import torch

from . import ops


class BaseLoss(object):
  """Inherit from this class when implementing new losses (losses.py:4-18)."""

  def calculate_loss(self, unused_triplets, **unused_params):
    raise NotImplementedError()


class HingeLoss(BaseLoss):
  def calculate_loss(self, triplets, margin=0.1):
    """triplets: [batch, 3, embedding] = [anchor, positive, negative] (losses.py:21-49).
    Returns {'hinge_loss' scalar, 'anchors'/'positives'/'negatives' [B,1,D], 'pos_dist'/'neg_dist'/'hinge_dist' [B,1]}."""
    as_numpy = not torch.is_tensor(triplets)
    t = torch.as_tensor(np.asarray(triplets, np.float32) if as_numpy else triplets)
    if t.dim() != 3 or t.shape[1] != 3:
      raise ValueError("triplets must be [batch, 3, embedding]; got %s" % (tuple(t.shape),))
    if not t.is_cuda:
      if not torch.cuda.is_available():
        raise RuntimeError("HingeLoss runs on the CUDA device only (libcdml has no CPU path)")
      t = t.cuda()
    t = t.to(torch.float32).contiguous()
    B, _, D = t.shape
    r = ops.triplet_hinge(t.view(3 * B, D), B, margin)
    res = {"hinge_loss": r["stats"][0], "anchors": t[:, 0:1, :], "positives": t[:, 1:2, :], "negatives": t[:, 2:3, :],
           "pos_dist": r["pos_dist"].view(B, 1), "neg_dist": r["neg_dist"].view(B, 1),
           "hinge_dist": r["hinge_dist"].view(B, 1), "mean_pos_dist": r["stats"][1], "mean_neg_dist": r["stats"][2]}
    if as_numpy:
      res = {k: v.cpu().numpy() for k, v in res.items()}
    return res
