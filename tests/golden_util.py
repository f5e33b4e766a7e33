"""Helpers to read tests/golden/*.npz (made by tests/golden/make_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def case_names():
  return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


class Golden:
  def __init__(self, name):
    self.z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    self.meta = json.loads(str(self.z['meta']))
    self.param_names = [str(s) for s in self.z['param_names']]
    self.num_steps = int(self.z['num_steps'])
    self.indptr = self.z['csr_indptr']
    self.indices = self.z['csr_indices']
    self.data = self.z['csr_data']
    self.user_order = self.z['user_order']

  def init_params(self):
    return {n: self.z['init/' + n] for n in self.param_names}

  def step(self, i):
    pre = 'step%d/' % i
    z = self.z
    return dict(users=z[pre + 'users'], items=(z[pre + 'items'] if bool(z[pre + 'has_items']) else None),
                indices=z[pre + 'indices'], values=z[pre + 'values'], size=tuple(int(v) for v in z[pre + 'size']),
                loss=float(z[pre + 'loss']),
                noise_keep=(z[pre + 'noise_keep'] if (pre + 'noise_keep') in z.files else None),
                dropout_keep=(z[pre + 'dropout_keep'] if (pre + 'dropout_keep') in z.files else None),
                grads={n: z[pre + 'grad/' + n] for n in self.param_names},
                params={n: z[pre + 'param/' + n] for n in self.param_names})

  def pools(self):
    """Yields (pool_users, [global step index for each slice])."""
    pool, batch = self.meta['pool'], self.meta['batch']
    step = 0
    for off in range(0, len(self.user_order), pool):
      users = self.user_order[off:off + pool]
      k = (len(users) + batch - 1) // batch
      yield users, list(range(step, step + k))
      step += k
