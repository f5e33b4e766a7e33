"""Train-step parity (SURVEY.md §8d parity gates): loss, gradient norms and post-step parameters of the CUDA
step against (1) golden vectors recorded from the unmodified reference and (2) the CPU oracle on seeded
synthetic inputs.  Tolerance: 1e-3 relative on loss and gradient L2 norms (north_star); the SIMT engine, which
shares every kernel except the tensor-core GEMM, is held to the same bar so a failure can be localised."""
import numpy as np
import pytest
import torch

from oracle import recoder_oracle as O
from recoder_b200 import _native
from recoder_b200.data import collate_pool
from recoder_b200.synth import synthetic_csr
from tests.golden_util import Golden
from tests.gpu_util import compact_oracle_grads, device_dataset, inner_grads, make_engine, make_model, rel_err

pytestmark = pytest.mark.gpu

ENGINES = [pytest.param(_native.GEMM_SIMT, id='simt'), pytest.param(_native.GEMM_TCGEN05, id='tcgen05')]
GOLDEN_CASES = ['ae_mse_adam', 'ae_mse_conf_ratings', 'ae_nll_adam', 'ae_bce_adam', 'ae_mse_sgd',
                'ae_nll_sparseadam', 'ae_mse_noneg', 'ae_nll_pool', 'mf_mse_adam', 'mf_nll_sgd',
                # Adagrad / RMSprop, multi-layer and tied autoencoders, input noise and dropout
                # (the keep masks the reference drew are injected)
                'ae_mse_deep', 'ae_mse_adagrad', 'ae_nll_rmsprop', 'ae_nll_deep_tied', 'ae_bce_deep3',
                'ae_nll_noise_dropout', 'ae_mse_deep_dropout', 'mf_mse_dropout']

# bf16 operands (2^-9 relative rounding, zero mean) against an fp32 reference
TOL_LOSS = 1e-3
TOL_GRAD = 1e-3   # on L2 norms, realistic sizes (north_star)
# The golden cases are tiny (8-24 rows, ~60 items, H=16): bf16 quantisation noise of the logits does not average
# out over so few elements (a CPU emulation of the bf16 pipeline reproduces the GPU value to 6 digits), so their
# norm gate is 3e-3; every realistic-size case below is held to 1e-3.
TOL_GRAD_GOLDEN = 3e-3
TOL_GRAD_ELEM = 2e-2  # relative Frobenius distance of whole gradient blocks
TOL_PARAM = 5e-2  # Adam turns sign flips of near-zero gradients into +-lr moves; grads are the tight gate


@pytest.mark.parametrize('engine', ENGINES)
@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_step_matches_reference_golden(name, engine):
  g = Golden(name)
  m = g.meta
  kind = m['model']
  tied = bool(m.get('constrained', False))
  model = make_model(kind, m['num_items'], m['num_users'], m['hidden'], m['act'], g.init_params(), sparse=m['sparse'],
                     constrained=tied, noise=float(m.get('noise', 0.0)), dropout=float(m.get('dropout', 0.0)))
  eng = make_engine(model, m['loss'], m['loss_params'].get('confidence', 0.0), m['opt'], m['lr'], m['wd'], engine)
  ds = device_dataset(g.indptr, g.indices, g.data, m['num_items'])
  csr = ds.device_csr()
  named = dict(model.named_parameters())
  for users, steps in g.pools():
    pool = collate_pool(csr, users, m['neg'])
    for k, s in enumerate(steps):
      ref = g.step(s)
      if s > 0:
        # teacher forcing: start every step from the reference's own parameters, otherwise the comparison
        # measures how two Adam trajectories drift apart (sign flips of near-zero gradients move a weight by
        # 2*lr) instead of the parity of one step
        with torch.no_grad():
          for n2, v2 in g.step(s - 1)['params'].items():
            named[n2].copy_(torch.from_numpy(v2).to('cuda'))
      row0 = k * m['batch']
      rows = ref['size'][0]
      eng.debug_noise_keep = None if ref['noise_keep'] is None else torch.from_numpy(ref['noise_keep']).cuda()
      eng.debug_dropout_keep = None if ref['dropout_keep'] is None else \
        torch.from_numpy(np.ascontiguousarray(ref['dropout_keep'])).cuda()
      eng.train_step(pool, row0, rows)
      loss = float(eng.losses(1)[0])
      assert loss == pytest.approx(ref['loss'], rel=TOL_LOSS), 'loss step %d' % s
      items = ref['items'] if ref['items'] is not None else np.arange(m['num_items'])
      last = {k2: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k2, v in eng.last.items()}
      if kind == 'ae':
        pairs = [('dWe', ref['grads'][O.AE_EN_W][items]), ('dbe', ref['grads'][O.AE_EN_B]),
                 ('dbd', ref['grads'][O.AE_DE_B][items])]
        if not tied:   # tied: the single table's gradient (both contributions) is reported as dWe
          pairs.append(('dWd', ref['grads'][O.AE_DE_W][items]))
        for n2, gt in inner_grads(eng, model).items():
          last[n2] = gt.detach().cpu().numpy()
          pairs.append((n2, ref['grads'][n2]))
      else:
        pairs = [('dV', ref['grads'][O.MF_ITEM_W][items]), ('dbias', ref['grads'][O.MF_BIAS][items]),
                 ('dU', ref['grads'][O.MF_USER_W][ref['users']])]
      for key, want in pairs:
        got = last[key]
        nw = np.linalg.norm(want)
        assert np.linalg.norm(got) == pytest.approx(nw, rel=TOL_GRAD_GOLDEN, abs=1e-7), '%s norm step %d' % (key, s)
        assert rel_err(got, want) < TOL_GRAD_ELEM, '%s step %d' % (key, s)
      state = {n: p.detach().cpu().numpy() for n, p in model.named_parameters()}
      for n in g.param_names:
        assert rel_err(state[n], ref['params'][n]) < TOL_PARAM, '%s after step %d' % (n, s)


CONFIGS = [
  # kind, U, I, nnz, H, B, loss, act
  ('ae', 2000, 5000, 50, 128, 256, 'mse', 'tanh'),        # BASELINE config C1 shape
  ('ae', 3000, 26744, 144, 200, 500, 'mse', 'tanh'),      # C2 shape (H=200 is not a multiple of 16/64)
  ('ae', 4096, 20000, 100, 512, 1024, 'logloss', 'tanh'), # C3-like
  ('mf', 3000, 20000, 100, 256, 512, 'mse', 'none'),      # C4-like
  ('ae', 777, 3001, 30, 72, 333, 'logistic', 'sigmoid'),  # ragged everything
]


@pytest.mark.parametrize('engine', ENGINES)
@pytest.mark.parametrize('kind,U,I,nnz,H,B,loss,act', CONFIGS)
def test_step_matches_oracle(kind, U, I, nnz, H, B, loss, act, engine):
  if engine == _native.GEMM_SIMT and B * I > 3e7:
    pytest.skip('SIMT validation engine is too slow for this size')
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=11)
  torch.manual_seed(17)
  if kind == 'ae':
    params = O.init_ae_params(I, [H], seed=3)
    params[O.AE_EN_B] = torch.randn(H) * 0.05
    params[O.AE_DE_B] = torch.randn(I) * 0.05
  else:
    params = O.init_mf_params(I, U, H, seed=3)
    params[O.MF_BIAS] = torch.randn(I) * 0.05
  lr, wd = 1e-3, 2e-5
  tr = O.OracleTrainer(kind, params, loss=loss, confidence=0.0, optimizer='adam', lr=lr, weight_decay=wd,
                       activation=act)
  model = make_model(kind, I, U, [H] if kind == 'ae' else H, act, {k: v.numpy() for k, v in params.items()})
  eng = make_engine(model, loss, 0.0, 'adam', lr, wd, engine)
  ds = device_dataset(indptr, indices, data, I)
  order = np.random.default_rng(5).permutation(U)
  steps = 3
  for s in range(steps):
    users = order[s * B:(s + 1) * B]
    pool = collate_pool(ds.device_csr(), users, True)
    ob = O.collate(indptr, indices, data, I, users, B, True)[0]
    assert np.array_equal(pool.items.cpu().numpy(), ob.items)
    with torch.no_grad():  # teacher forcing: the oracle steps from the GPU's current parameters
      for n2, p2 in model.named_parameters():
        tr.params[n2].copy_(p2.detach().cpu())
    oloss, ograds = tr.step(ob)
    eng.train_step(pool, 0, len(users))
    loss_gpu = float(eng.losses(1)[0])
    assert loss_gpu == pytest.approx(oloss, rel=TOL_LOSS), 'loss step %d' % s
    want = compact_oracle_grads(kind, ograds, ob)
    for key, w in want.items():
      got = eng.last[key].detach().cpu().numpy()
      assert np.linalg.norm(got) == pytest.approx(np.linalg.norm(w), rel=TOL_GRAD), '%s norm step %d' % (key, s)
      assert rel_err(got, w) < TOL_GRAD_ELEM, '%s step %d' % (key, s)
  state = {n: p.detach().cpu().numpy() for n, p in model.named_parameters()}
  ost = tr.state()
  for n in ost:
    assert rel_err(state[n], ost[n]) < TOL_PARAM, n
