"""Ranking metrics and the evaluator with the reference's interface (recoder/metrics.py).

`Recall`, `NDCG`, `AveragePrecision` (metrics.py:66-132) and `RecommenderEvaluator` (metrics.py:135-232) keep their
names, constructor arguments and per-user result lists, so `Recoder.train(..., metrics=[...])`, `Recoder.evaluate` and
the reference's golden-metric test (tests/test_model.py:14-84) run unchanged.  The recommendations themselves come
from the GPU (`Recoder.recommend`: CSR encoder -> full-width tcgen05 decoder -> `rcd_mask_seen` -> `rcd_topk_rows`);
the metric arithmetic on the k recommended ids per user is host-side NumPy, as in the reference.

Attribution: the public interface of this module (class / method names, argument lists and their documentation, log
messages, checkpoint keys) mirrors amoussawi/recoder (MIT License, Copyright (c) 2018 Abdallah Moussawi) so that it is
a drop-in for that library; see LICENSE.  The implementation underneath is original.
"""
import numpy as np

from .data import RecommendationDataLoader
from .recommender import InferenceRecommender, Recommender  # noqa: F401  (re-exported for convenience)


def _hits(x, y, k):
  x = np.asarray(x)[:k]
  return x, np.isin(x, y, assume_unique=True).astype(np.int64)


def average_precision(x, y, k, normalize=True):
  """AP@k of the ranked list `x` against the relevant set `y` (metrics.py:9-20)."""
  x, hit = _hits(x, y, k)
  precision_at = hit.cumsum() / (1 + np.arange(len(x)))
  norm = min(k, len(y)) if normalize else len(y)
  return (precision_at * hit).sum() / norm


def recall(x, y, k, normalize=True):
  """Recall@k (metrics.py:23-29); `normalize` divides by min(k, |y|) instead of |y|."""
  x, hit = _hits(x, y, k)
  norm = min(k, len(y)) if normalize else len(y)
  return hit.sum() / norm


def dcg(x, y, k):
  """DCG@k with binary gains and log2 discounts (metrics.py:32-38)."""
  x, hit = _hits(x, y, k)
  return (hit / np.log2(2 + np.arange(len(x)))).sum()


def ndcg(x, y, k):
  """NDCG@k (metrics.py:41-45): the ideal ranking lists the relevant items first."""
  return dcg(x, y, k) / dcg(y, y, k)


class Metric(object):
  """
  A Base class for metrics. All metrics should implement the ``evaluate`` function.

  Args:
    metric_name (str): metric name. useful for representing it as string (printing) and hashing.
  """

  def __init__(self, metric_name):
    self.metric_name = metric_name

  def __str__(self):
    return self.metric_name

  def __hash__(self):
    return self.metric_name.__hash__()

  def __eq__(self, other):
    return isinstance(other, Metric) and self.metric_name == other.metric_name

  def evaluate(self, x, y):
    """Evaluates the recommendations `x` (ranked item ids) with respect to the relevant items `y`."""
    raise NotImplementedError


class AveragePrecision(Metric):
  """Average Precision @ K (metrics.py:84-100)."""

  def __init__(self, k, normalize=True):
    super().__init__(metric_name='AveragePrecision@{}'.format(k))
    self.k = k
    self.normalize = normalize

  def evaluate(self, x, y):
    return average_precision(x, y, k=self.k, normalize=self.normalize)


class Recall(Metric):
  """Recall @ K (metrics.py:103-119)."""

  def __init__(self, k, normalize=True):
    super().__init__(metric_name='Recall@{}'.format(k))
    self.k = k
    self.normalize = normalize

  def evaluate(self, x, y):
    return recall(x, y, k=self.k, normalize=self.normalize)


class NDCG(Metric):
  """Normalized Discounted Cumulative Gain @ K (metrics.py:122-132)."""

  def __init__(self, k):
    super().__init__(metric_name='NDCG@{}'.format(k))
    self.k = k

  def evaluate(self, x, y):
    return ndcg(x, y, k=self.k)


class RecommenderEvaluator(object):
  """
  Evaluates a :class:`recoder_b200.recommender.Recommender` given a set of :class:`Metric` (metrics.py:135-232).

  Args:
    recommender (Recommender): the Recommender to evaluate
    metrics (list): list of metrics used to evaluate the recommender
  """

  def __init__(self, recommender, metrics):
    self.recommender = recommender
    self.metrics = metrics

  def evaluate(self, eval_dataset, batch_size=1, num_users=None, num_workers=0):
    """
    Evaluates the recommender with an evaluation dataset (sequential user order, like the reference's identity
    collate over an unshuffled... RandomSampler-ordered loader; the result lists are per user in visiting order).

    Returns:
      dict: A dict mapping each metric to the list of the metric values on each user in the dataset.
      ``num_workers`` is accepted for interface compatibility; the metric arithmetic is a few vector operations per
      user and runs in the calling process.
    """
    dataloader = RecommendationDataLoader(eval_dataset, batch_size=batch_size, collate_fn=lambda _: _)
    results = {metric: [] for metric in self.metrics}
    target_matrix = eval_dataset.target_interactions_matrix
    processed = 0
    for input, target in dataloader:
      recommendations = self.recommender.recommend(input)
      users = np.asarray(target.users)
      for x, u in zip(recommendations, users):
        lo, hi = target_matrix.indptr[u], target_matrix.indptr[u + 1]
        seg = target_matrix.indices[lo:hi]
        y = seg[target_matrix.data[lo:hi] != 0]        # `.nonzero()` drops stored zeros (metrics.py:206)
        for metric in self.metrics:
          results[metric].append(metric.evaluate(x, y))
      processed += len(users)
      if num_users is not None and processed >= num_users:
        break
    return results
