"""Loss modules with the reference's interface (recoder/losses.py).

In training the loss and its gradient are computed by the fused `rcd_loss_grad` kernel straight from the
bf16 logits and the sparse target (engine.TrainEngine); these classes carry the configuration
(`confidence`, `reduction`) and keep a dense `forward(input, target)` for callers that hold dense CUDA tensors
(evaluation utilities, user code).  The dense forward is a convenience, not the training path.
"""
import torch
from torch import nn


def _reduce(x, reduction='elementwise_mean'):
  if reduction == 'none':
    return x
  elif reduction == 'elementwise_mean':
    return x.mean()
  elif reduction == 'sum':
    return x.sum()
  else:
    raise ValueError('No such reduction {} defined'.format(reduction))


class MSELoss(nn.Module):
  """
  Weighted mean squared error (reference recoder/losses.py:16-47): ``w = 1 + confidence * [target > 0]``,
  ``loss = w * (input - target)^2``.

  Args:
    confidence (float, optional): the weighting of positive observations.
    reduction (string, optional): 'none' | 'elementwise_mean' | 'sum'. Default: 'elementwise_mean'
  """

  def __init__(self, confidence=0, reduction='elementwise_mean'):
    super(MSELoss, self).__init__()
    self.reduction = reduction
    self.confidence = confidence

  def forward(self, input, target):
    weights = 1 + self.confidence * (target > 0).float()
    return _reduce(weights * (input - target) ** 2, reduction=self.reduction)


class MultinomialNLLLoss(nn.Module):
  """
  Negative log-likelihood of the multinomial distribution (reference recoder/losses.py:50-71):
  ``loss = - target * log_softmax(input, dim=1)``.

  Args:
    reduction (string, optional): 'none' | 'elementwise_mean' | 'sum'. Default: 'elementwise_mean'
  """

  def __init__(self, reduction='elementwise_mean'):
    super(MultinomialNLLLoss, self).__init__()
    self.reduction = reduction

  def forward(self, input, target):
    lse = torch.logsumexp(input, dim=1, keepdim=True)
    return _reduce(-target * (input - lse), reduction=self.reduction)
