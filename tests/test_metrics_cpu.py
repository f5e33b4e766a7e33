"""Ranking metrics (recoder_b200/metrics.py) against the reference's own known-answer tests
(tests/test_metrics.py:12-54 there, restated with the same inputs and expected values, rtol 1e-9) and against the
per-user values the unmodified reference computed for tests/golden/eval/eval_golden.npz."""
import os

import numpy as np
import pytest

from recoder_b200.metrics import NDCG, AveragePrecision, Recall

RTOL, ATOL = 1e-9, 0.0
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'eval', 'eval_golden.npz')


@pytest.mark.parametrize('x, y, k, normalize, expected', [
  (np.arange(10), [0, 2, 5, 8, 9], 10, False, 1 / 5 * (1 + 2 / 3 + 3 / 6 + 4 / 9 + 5 / 10)),
  (np.arange(10), [1, 4, 5, 6, 12], 10, False, 1 / 5 * (1 / 2 + 2 / 5 + 3 / 6 + 4 / 7 + 0)),
  (np.arange(10), [0, 1, 2, 3, 4], 10, False, 1),
  (np.arange(10), [0, 2, 5, 8, 9], 3, True, 1 / 3 * (1 + 2 / 3)),
  (np.arange(10), [1, 4, 5, 6, 12], 3, True, 1 / 3 * (1 / 2)),
])
def test_average_precision_known_answers(x, y, k, normalize, expected):
  assert np.isclose(AveragePrecision(k=k, normalize=normalize).evaluate(x, y), expected, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('x, y, k, normalize, expected', [
  (np.arange(10), [0, 2, 5, 8, 9], 10, False, 1),
  (np.arange(10), [1, 4, 5, 6, 12], 10, False, 4 / 5),
  (np.arange(10), [0, 2, 5, 8, 9], 3, False, 2 / 5),
  (np.arange(10), [1, 4, 5, 6, 12], 3, False, 1 / 5),
  (np.arange(10), [0, 2, 5, 8, 9], 3, True, 2 / 3),
  (np.arange(10), [1, 4, 5, 6, 12], 3, True, 1 / 3),
])
def test_recall_known_answers(x, y, k, normalize, expected):
  assert np.isclose(Recall(k=k, normalize=normalize).evaluate(x, y), expected, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('x, y, k, expected', [
  (np.arange(10), [0, 2, 5, 8, 9], 10, 0.8296882915641869),
  (np.arange(10), [1, 4, 5, 6, 12], 10, 0.5790560467042355),
  (np.arange(10), [0, 2, 5, 8, 9], 3, 0.7039180890341347),
  (np.arange(10), [1, 4, 5, 6, 12], 3, 0.2960819109658652),
])
def test_ndcg_known_answers(x, y, k, expected):
  assert np.isclose(NDCG(k=k).evaluate(x, y), expected, rtol=RTOL, atol=ATOL)


def test_metrics_match_reference_on_recorded_recommendations():
  z = np.load(GOLDEN)
  U = int(z['shape'][0])
  tptr, tidx = z['tg_indptr'], z['tg_indices']
  metrics = [Recall(k=20, normalize=True), Recall(k=50, normalize=False), NDCG(k=50), AveragePrecision(k=10)]
  for kind in ('ae', 'ae2', 'mf'):
    recs = z[kind + '/recs']
    for m in metrics:
      want = z[kind + '/metric/' + str(m)]
      got = np.array([m.evaluate(recs[u], tidx[tptr[u]:tptr[u + 1]]) for u in range(U)])
      np.testing.assert_allclose(got, want, rtol=1e-12, atol=0, err_msg='%s %s' % (kind, m))
