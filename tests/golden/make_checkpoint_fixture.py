"""Generates the checkpoint-interchange fixture from the UNMODIFIED reference (build container only):

  tests/golden/eval/ref_checkpoint_epoch_2.model   written by the reference's `Recoder.save_state` (model.py:193-224)
                                                   after two epochs of `Recoder.train` (AE[16], tanh, Adam, logloss)
  tests/golden/eval/ref_checkpoint_resume.npz      what the reference computes after `init_from_model_file`
                                                   (model.py:166-191) + `__init_training` (optimizer state hand-over,
                                                   model.py:158-164) + ONE explicit step on users 0..23: loss, gradients,
                                                   post-step parameters and Adam state

    python tests/golden/make_checkpoint_fixture.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_shims  # noqa: E402
import make_golden  # noqa: E402

LR, WD, BATCH = 1e-2, 1e-4, 24


def main():
  rdata, rnn, rlosses, rmodel = ref_shims.import_reference()
  import warnings
  warnings.simplefilter('ignore')
  torch.manual_seed(99)
  csr = make_golden.make_matrix(7, False)
  dataset = rdata.RecommendationDataset(csr)
  model = rnn.DynamicAutoencoder(hidden_layers=[16], activation_type='tanh')
  trainer = rmodel.Recoder(model=model, use_cuda=False, optimizer_type='adam', loss='logloss')
  prefix = os.path.join(HERE, 'eval', 'ref_checkpoint')
  trainer.train(train_dataset=dataset, batch_size=BATCH, lr=LR, weight_decay=WD, num_epochs=2, negative_sampling=True,
                model_checkpoint_prefix=prefix)
  ckpt = prefix + '_epoch_2.model'
  assert os.path.isfile(ckpt)

  model2 = rnn.DynamicAutoencoder()
  trainer2 = rmodel.Recoder(model=model2, use_cuda=False, optimizer_type='adam', loss='logloss')
  trainer2.init_from_model_file(ckpt)
  trainer2._Recoder__init_training(train_dataset=dataset, lr=LR, weight_decay=WD)
  users = np.arange(BATCH)
  ui, _ = dataset[users]
  batch = rdata.BatchCollator(batch_size=BATCH, negative_sampling=True).collate(ui)[0]
  trainer2.optimizer.zero_grad()
  loss = trainer2._Recoder__compute_loss(batch, None)
  loss.backward()
  out = {'users': users.astype(np.int64), 'loss': np.array(loss.item()), 'current_epoch': np.array(trainer2.current_epoch),
         'csr_indptr': csr.indptr.astype(np.int64), 'csr_indices': csr.indices.astype(np.int32),
         'csr_data': csr.data.astype(np.float32), 'items': batch.items.numpy().astype(np.int64),
         'hyper': np.array([LR, WD, BATCH])}
  names = [n for n, _ in model2.named_parameters()]
  out['param_names'] = np.array(names)
  for n, p in model2.named_parameters():
    out['grad/' + n] = p.grad.detach().numpy().copy()
  trainer2.optimizer.step()
  for i, (n, p) in enumerate(model2.named_parameters()):
    out['param/' + n] = p.detach().numpy().copy()
    st = trainer2.optimizer.state[p]
    out['exp_avg/' + n] = st['exp_avg'].numpy().copy()
    out['exp_avg_sq/' + n] = st['exp_avg_sq'].numpy().copy()
    out['step/' + n] = np.array(float(st['step']))
  np.savez_compressed(os.path.join(HERE, 'eval', 'ref_checkpoint_resume.npz'), **out)
  print('checkpoint %.1f KB, resume fixture %.1f KB; current_epoch=%d loss=%.6f' % (
    os.path.getsize(ckpt) / 1024, os.path.getsize(os.path.join(HERE, 'eval', 'ref_checkpoint_resume.npz')) / 1024,
    trainer2.current_epoch, loss.item()))


if __name__ == '__main__':
  main()
