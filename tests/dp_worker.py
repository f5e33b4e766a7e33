"""Worker of tests/test_gpu_d_multigpu.py, launched with torch.distributed.run (one process per GPU, NCCL).

Trains the same model for a few steps three ways and compares the final parameters and losses:
  (1) rank 0 alone with batch_size = world * B           (the exact single-process reference of a DP step, §8e)
  (2) all ranks, dp_exchange='nccl'  (one all-reduce of the gradient slab, full Adam on every rank)
  (3) all ranks, dp_exchange='p2p'   (fused reduce-scatter -> Adam -> all-gather over CUDA-IPC peer memory)
and, for (1), against the CPU oracle's losses.  Prints DP_OK on rank 0 when everything matches.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def run(kind, loss, mode, pg, batch, steps, matrix, U, I, H, seed_params, resident=True, lazy='auto'):
  from recoder_b200.data import RecommendationDataset
  from recoder_b200.model import Recoder
  from recoder_b200.nn import DynamicAutoencoder, MatrixFactorization
  if kind == 'ae':
    model = DynamicAutoencoder(hidden_layers=[H], activation_type='tanh')
  else:
    model = MatrixFactorization(embedding_size=H, activation_type='none')
  torch.manual_seed(seed_params)
  if ':' in mode:   # 'p2p:ipc' = CUDA-IPC unicast ld/st, 'p2p:symm' = symmetric memory + NVLS multicast when available
    mode, backend = mode.split(':')
    os.environ['RCD_P2P_BACKEND'] = backend
  parallel = 'rows'
  os.environ.pop('RCD_IP_COLLECTIVES', None)
  os.environ.pop('RCD_P2P_MULTICAST', None)
  if mode.startswith('items'):   # item-parallel: every rank sees all rows, the item axis is sharded
    # 'items' = peer-memory collectives (unicast below 4 ranks), 'items-mc' = the same through NVLS multicast,
    # 'items-nccl' = NCCL all-reduces
    if mode == 'items-nccl':
      os.environ['RCD_IP_COLLECTIVES'] = 'nccl'
    if mode == 'items-mc':
      os.environ['RCD_P2P_MULTICAST'] = '1'
    mode, parallel = 'nccl', 'items'
  tr = Recoder(model=model, use_cuda=True, optimizer_type='adam', loss=loss, process_group=pg,
               dp_exchange=mode if mode != 'single' else 'nccl', parallel=parallel, lazy_adam=lazy)
  ds = RecommendationDataset(matrix, device_resident=resident)
  order = np.random.default_rng(5).permutation(U)
  tr.train(ds, lr=1e-2, weight_decay=1e-4, num_epochs=1, iters_per_epoch=steps, batch_size=batch,
           negative_sampling=True, user_order=lambda e: order)
  torch.cuda.synchronize()
  if tr._p2p is not None:
    tr.optimizer.gather_shards(tr._p2p)
  tr.sync_parameters()
  params = {n: p.detach().float().cpu().clone() for n, p in model.named_parameters()}
  state = {n: (s.m.cpu().clone(), s.v.cpu().clone()) for n, s in tr.optimizer.states.items()}
  if tr._ip is not None:
    assert parallel == 'items'
    full = tr._Recoder__full_optimizer_state(tr.optimizer.state_dict(dense=True))
    names = [n for n, _ in model.named_parameters()]
    state = {names[i]: (e['exp_avg'].clone(), e['exp_avg_sq'].clone()) for i, e in full['state'].items()}
  losses = tr.last_epoch_losses.copy()
  used_p2p = tr._p2p is not None
  if used_p2p and dist.get_rank() == 0:
    print('p2p backend %s, multicast %s' % (tr._p2p.backend, bool(tr._p2p.flags.mc_ptr and tr._p2p.multicast)), flush=True)
  return params, state, losses, used_p2p


def main():
  rank = int(os.environ['RANK'])
  world = int(os.environ['WORLD_SIZE'])
  torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
  dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
  from recoder_b200.synth import synthetic_csr, to_scipy
  U, I, nnz, H, B, steps = 8192, 3000, 40, 64, 128, 4
  indptr, indices, data = synthetic_csr(U, I, nnz, seed=11)
  matrix = to_scipy(indptr, indices, data, I)
  solo = dist.new_group([0])
  ok = True
  for kind, loss in (('ae', 'logloss'), ('ae', 'mse'), ('mf', 'mse')):
    ref = None
    if rank == 0:
      ref = run(kind, loss, 'single', solo, B * world, steps, matrix, U, I, H, 3)
    dist.barrier()
    nccl = run(kind, loss, 'nccl', None, B, steps, matrix, U, I, H, 3)
    p2p = run(kind, loss, 'p2p:ipc', None, B, steps, matrix, U, I, H, 3)
    p2p_mc = run(kind, loss, 'p2p:auto', None, B, steps, matrix, U, I, H, 3)
    assert p2p[3] and p2p_mc[3], 'peer-memory exchange was not used'
    assert not nccl[3]
    variants = [('nccl', nccl), ('p2p-ipc', p2p), ('p2p-auto', p2p_mc)]
    # host-resident matrix: every rank stages its own block of the pool, the blocks are all-gathered on the devices
    staged = run(kind, loss, 'p2p:auto', None, B, steps, matrix, U, I, H, 3, resident=False)
    variants.append(('p2p-host-staged', staged))
    # NCCL exchange with the deferred dense Adam forced on: every replica replays / updates the same rows
    lazy = run(kind, loss, 'nccl', None, B, steps, matrix, U, I, H, 3, lazy=True)
    assert not lazy[3]
    variants.append(('nccl-deferred-adam', lazy))
    if kind == 'ae':
      for tag in ('items', 'items-mc', 'items-nccl'):
        got = run(kind, loss, tag, None, B, steps, matrix, U, I, H, 3)
        assert not got[3]
        variants.append((tag, got))
    for tag, got in variants:
      # every rank holds the same replica
      for n, t in got[0].items():
        g = [torch.zeros_like(t, device='cuda') for _ in range(world)]
        dist.all_gather(g, t.cuda())
        for q in range(1, world):
          same = torch.equal(g[0], g[q])
          if not same:
            ok = False
            print('[rank %d] %s/%s %s: replica %d differs from replica 0 (max %.3e)' %
                  (rank, kind, loss, tag, q, float((g[0] - g[q]).abs().max())), flush=True)
      if rank == 0:
        for n, t in got[0].items():
          want = ref[0][n]
          err = float((t - want).norm() / max(float(want.norm()), 1e-30))
          if err > 2e-4:
            ok = False
            print('%s/%s %s: parameter %s differs from the single-process step: rel %.3e' % (kind, loss, tag, n, err),
                  flush=True)
        for n, (m, v) in got[1].items():
          wm, wv = ref[1][n]
          em = float((m - wm).norm() / max(float(wm.norm()), 1e-30))
          ev = float((v - wv).norm() / max(float(wv.norm()), 1e-30))
          if em > 2e-3 or ev > 2e-3:
            ok = False
            print('%s/%s %s: Adam state of %s differs: m %.3e v %.3e' % (kind, loss, tag, n, em, ev), flush=True)
        lerr = float(np.max(np.abs(got[2] - ref[2]) / np.abs(ref[2])))
        if lerr > 1e-3:
          ok = False
          print('%s/%s %s: losses differ: %s vs %s' % (kind, loss, tag, got[2], ref[2]), flush=True)
        print('%s/%s %s: losses %s (single %s)' % (kind, loss, tag, np.round(got[2], 4), np.round(ref[2], 4)),
              flush=True)
  flag = torch.tensor([1 if ok else 0], device='cuda')
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  if rank == 0:
    print('DP_OK' if int(flag.item()) else 'DP_FAIL', flush=True)
  dist.destroy_process_group()
  sys.exit(0 if int(flag.item()) else 1)


if __name__ == '__main__':
  main()
