"""Probe (run under torchrun on >=2 GPUs): does torch symmetric memory rendezvous work here, and is NVLS multicast available?"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ['RANK']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
import torch.distributed._symmetric_memory as symm_mem
from torch._C._distributed_c10d import _SymmetricMemory
try:
  print(rank, 'has_multicast_support', _SymmetricMemory.has_multicast_support(torch.device('cuda').type, lr), flush=True)
except Exception as e:
  print(rank, 'has_multicast_support failed', repr(e), flush=True)
try:
  t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=torch.device('cuda', lr))
  hdl = symm_mem.rendezvous(t, dist.group.WORLD)
  print(rank, 'ptr', hex(t.data_ptr()), 'buffer_ptrs', [hex(p) for p in hdl.buffer_ptrs], 'mc', hex(hdl.multicast_ptr),
        'offset', getattr(hdl, 'offset', None), 'size', hdl.buffer_size, flush=True)
  big = symm_mem.empty(1 << 30, dtype=torch.uint8, device=torch.device('cuda', lr))
  h2 = symm_mem.rendezvous(big, dist.group.WORLD)
  print(rank, 'big mc', hex(h2.multicast_ptr), 'ptrs', [hex(p) for p in h2.buffer_ptrs], flush=True)
except Exception as e:
  import traceback; traceback.print_exc()
  print(rank, 'symm failed', repr(e), flush=True)
dist.barrier()
dist.destroy_process_group()
