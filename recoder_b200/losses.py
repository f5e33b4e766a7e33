"""Loss modules under the names and constructor signatures the reference exposes (recoder/losses.py).

During training nothing here executes: `Recoder` reads the configuration of these objects (`confidence`, `reduction`)
and the loss together with dL/dlogits comes out of the fused decoder epilogue (`rcd_decoder_fwd_loss`,
`rcd_sddmm`, `rcd_loss_finish`).  `forward(input, target)` on dense tensors is kept for user code that calls the modules
directly, and it is what the generic path (`engine._custom_loss`) runs when a module is configured in a way the fused
epilogue does not cover (for instance a reduction other than 'sum').

Attribution: the public interface of this module (class / method names, argument lists and their documentation, log
messages, checkpoint keys) mirrors amoussawi/recoder (MIT License, Copyright (c) 2018 Abdallah Moussawi) so that it is
a drop-in for that library; see LICENSE.  The implementation underneath is original.
"""
import torch
from torch import nn

# reduction name -> how a tensor of per-element losses is collapsed (names as in recoder/losses.py:5-13)
_REDUCTIONS = {
  'none': lambda t: t,
  'elementwise_mean': torch.mean,
  'sum': torch.sum,
}


class _ElementwiseLoss(nn.Module):
  """Holds the `reduction` of a loss that is defined element by element and applies it."""

  def __init__(self, reduction):
    super().__init__()
    self.reduction = reduction

  def _collapse(self, per_element):
    try:
      collapse = _REDUCTIONS[self.reduction]
    except KeyError:
      raise ValueError('No such reduction {} defined'.format(self.reduction)) from None
    return collapse(per_element)


class MSELoss(_ElementwiseLoss):
  """Squared error in which observed (positive) targets can weigh more than the zeros
  (reference recoder/losses.py:16-47): element (u, j) contributes ``(1 + confidence * [t_uj > 0]) * (o_uj - t_uj)**2``.

  Args:
    confidence (float, optional): extra weight of the positive targets (0: plain squared error).
    reduction (string, optional): 'none', 'elementwise_mean' (default) or 'sum'.
  """

  def __init__(self, confidence=0, reduction='elementwise_mean'):
    super().__init__(reduction)
    self.confidence = confidence

  def forward(self, input, target):
    diff = input - target
    weight = torch.where(target > 0, 1.0 + self.confidence, 1.0).to(diff.dtype)
    return self._collapse(weight * diff * diff)


class MultinomialNLLLoss(_ElementwiseLoss):
  """Multinomial negative log-likelihood over the item axis (reference recoder/losses.py:50-71): element (u, j)
  contributes ``-t_uj * log_softmax(o_u)_j``.

  Args:
    reduction (string, optional): 'none', 'elementwise_mean' (default) or 'sum'.
  """

  def __init__(self, reduction='elementwise_mean'):
    super().__init__(reduction)

  def forward(self, input, target):
    log_prob = input - torch.logsumexp(input, dim=1, keepdim=True)
    return self._collapse(-(target * log_prob))
