"""ctypes binding of librecoder_b200.so (the C ABI declared in include/recoder_b200.h).

There is NO CPU fallback: if the library is missing or a call fails, a RuntimeError is raised.
PyTorch tensors are only containers here — every call passes `tensor.data_ptr()`, sizes and the raw handle
of the current CUDA stream.
"""
import ctypes
import os
from ctypes import c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# RCD_LIB: an alternative build of the same library (kernel experiments: tools/gpu_*.sh)
LIB_PATH = os.environ.get('RCD_LIB') or os.path.join(_HERE, 'csrc', 'librecoder_b200.so')

ACT_IDS = {'none': 0, 'tanh': 1, 'sigmoid': 2, 'relu': 3, 'selu': 4, 'celu': 5, 'hardshrink': 6, 'atan': 7, 'sinh': 8,
           'asinh': 9, 'expm1': 10}
LOSS_IDS = {'mse': 0, 'logloss': 1, 'logistic': 2}
GEMM_TCGEN05, GEMM_SIMT = 0, 1
DEC_MODE_LOSS, DEC_MODE_ROWMAX = 0, 1

_P = c_void_p
# name -> (restype, argtypes); mirrors include/recoder_b200.h one to one
_SIGNATURES = {
  'rcd_abi_version': (c_int, []),
  'rcd_last_error': (ctypes.c_char_p, []),
  'rcd_device_sms': (c_int, []),
  'rcd_launch_count': (c_longlong, []),
  'rcd_host_stage_rows': (c_longlong, [_P, _P, _P, _P, c_int, c_longlong, c_longlong, _P, _P, _P]),
  'rcd_collate_scratch_bytes': (c_size_t, [c_int, c_int]),
  'rcd_collate': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                          c_size_t, _P]),
  'rcd_collate_coo': (c_int, [_P, _P, c_int, c_int, _P, _P]),
  'rcd_slice_csc_scratch_bytes': (c_size_t, [c_int, c_int]),
  'rcd_slice_csc': (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
  'rcd_dense_to_csr_scratch_bytes': (c_size_t, [c_int]),
  'rcd_dense_to_csr': (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
  'rcd_gather_rows': (c_int, [_P, c_int, _P, c_int, c_int, _P, c_int, _P, _P]),
  'rcd_gather_vec': (c_int, [_P, _P, c_int, _P, _P]),
  'rcd_ae_encoder_fwd': (c_int, [_P, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, c_int, _P]),
  'rcd_decoder_tile_n': (c_int, []),
  'rcd_decoder_fwd': (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, _P]),
  'rcd_decoder_stat_cols': (c_int, [c_int]),
  'rcd_decoder_fwd_loss': (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, c_float, _P, _P, c_int, _P,
                                   c_int, c_int, _P, _P]),
  'rcd_sddmm': (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, _P, _P, _P,
                        _P]),
  'rcd_loss_finish': (c_int, [_P, c_int, c_int, c_int, c_int, c_float, c_float, _P, _P, _P, _P, _P, c_int, _P, _P,
                              c_int, _P, c_int, _P, _P, c_int, _P, _P, _P, _P, _P]),
  'rcd_loss_finish_blocks': (c_int, [c_int]),
  'rcd_nll_ref_fix': (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P]),
  'rcd_loss_sum': (c_int, [_P, c_int, _P, _P, _P]),
  'rcd_sparse_dgrad': (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, _P, c_int, _P]),
  'rcd_csc_rows_accumulate': (c_int, [_P, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, c_size_t, c_longlong, _P]),
  'rcd_csc_heavy_scratch_bytes': (c_size_t, [c_int, c_longlong, c_int]),
  'rcd_decoder_dgrad_splits': (c_int, [c_int, c_int, c_int]),
  'rcd_decoder_dgrad': (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P]),
  'rcd_decoder_wgrad': (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_int, _P]),
  'rcd_dz_act': (c_int, [_P, c_int, c_int, _P, c_int, _P, c_int, c_int, c_int, _P, _P, _P]),
  'rcd_ae_encoder_wgrad': (c_int, [_P, c_int, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, c_size_t, c_longlong,
                                   _P]),
  'rcd_adam_step': (c_int, [_P, _P, _P, c_longlong, c_int, _P, c_int, _P, c_double, c_double, c_double, c_double,
                            c_double, c_longlong, _P]),
  'rcd_sgd_step': (c_int, [_P, _P, c_longlong, c_int, _P, c_int, _P, c_double, c_double, c_double, _P]),
  'rcd_adam_lazy_catchup': (c_int, [_P, _P, _P, c_int, _P, c_longlong, _P, c_longlong, _P, c_longlong, c_longlong,
                                    c_double, c_double, c_double, c_double, c_int, _P, _P, _P]),
  'rcd_adam_lazy_update': (c_int, [_P, _P, _P, c_int, _P, c_longlong, _P, c_int, _P, c_double, c_double, c_double,
                                   c_double, c_double, c_longlong, _P]),
  'rcd_adam_scalars': (c_int, [c_double, c_double, c_double, c_longlong, c_int, _P]),
  'rcd_sparse_adam_step': (c_int, [_P, _P, _P, c_int, _P, c_int, _P, c_int, c_double, c_double, c_double, c_double,
                                   c_longlong, _P]),
  'rcd_scatter_pos': (c_int, [_P, c_int, _P, c_int, _P]),
  'rcd_p2p_alloc': (c_int, [c_size_t, _P]),
  'rcd_p2p_free': (c_int, [_P]),
  'rcd_p2p_export': (c_int, [_P, _P]),
  'rcd_p2p_open': (c_int, [_P, _P]),
  'rcd_p2p_close': (c_int, [_P]),
  'rcd_p2p_barrier': (c_int, [_P, c_int, c_int, ctypes.c_uint, _P, c_double, _P]),
  'rcd_p2p_reduce': (c_int, [_P, c_int, c_longlong, c_longlong, _P, c_int, _P]),
  'rcd_p2p_allreduce': (c_int, [_P, _P, c_longlong, c_int, c_int, _P]),
  'rcd_adam_step_p2p': (c_int, [_P, _P, _P, c_longlong, c_longlong, c_int, _P, c_int, _P, c_int, c_int, c_int,
                                c_double, c_double, c_double, c_double, c_double, c_longlong, _P, _P, _P]),
  'rcd_adagrad_step': (c_int, [_P, _P, c_longlong, c_int, _P, c_int, _P, c_double, c_double, c_double, _P]),
  'rcd_rmsprop_step': (c_int, [_P, _P, _P, c_longlong, c_int, _P, c_int, _P, c_double, c_double, c_double, c_double,
                               c_double, _P]),
  'rcd_sgemm': (c_int, [c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, _P]),
  'rcd_dropout': (c_int, [_P, c_longlong, c_float, ctypes.c_ulonglong, ctypes.c_uint, c_longlong, _P, _P, _P]),
  'rcd_act_grad': (c_int, [_P, _P, c_longlong, c_int, _P, _P]),
  'rcd_colsum': (c_int, [_P, c_int, c_int, c_int, _P, _P]),
  'rcd_f32_to_bf16_rows': (c_int, [_P, c_int, c_int, _P, c_int, _P]),
  'rcd_bias_act': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, _P]),
  'rcd_rowsum': (c_int, [_P, c_int, c_int, c_int, _P, _P]),
  'rcd_mask_seen': (c_int, [_P, _P, c_int, c_int, _P, c_longlong, _P]),
  'rcd_topk_rows': (c_int, [_P, c_longlong, c_int, c_int, c_int, _P, _P, _P]),
  'rcd_sumsq': (c_int, [_P, c_longlong, c_int, c_int, _P, _P]),
  'rcd_gemm_bf16': (c_int, [c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P]),
}



# ---- K12 native step executor: C structs of include/recoder_b200.h ("K12") ------------------------------------------
class RcdParam(ctypes.Structure):
  _fields_ = [('p', _P), ('s1', _P), ('s2', _P), ('rows', c_longlong), ('cols', c_int), ('pad_', c_int),
              ('weight_decay', c_double), ('t', c_longlong), ('last', _P)]


class RcdPoolView(ctypes.Structure):
  _fields_ = [('row_ptr', _P), ('raw_items', _P), ('cols', _P), ('vals', _P), ('row_inv_norm', _P), ('row_sum', _P),
              ('pos', _P), ('items', _P), ('users', _P), ('nnz_slice', c_longlong), ('n', c_int), ('pad_', c_int)]


class RcdStepIp(ctypes.Structure):
  _fields_ = [('enabled', c_int), ('rank', c_int), ('world', c_int), ('use_nccl_', c_int), ('flags_host', _P),
              ('seq_host', _P), ('barrier_timeout_s', c_double), ('shared_local', _P), ('shared_host', _P),
              ('shared_mc', _P), ('off_z', c_longlong), ('off_dz', c_longlong), ('off_ref', c_longlong),
              ('off_sum', c_longlong)]


class RcdStepArgs(ctypes.Structure):
  _fields_ = [('abi', c_int), ('kind', c_int), ('H', c_int), ('act', c_int), ('loss', c_int), ('optimizer', c_int),
              ('train', c_int), ('overlap', c_int), ('confidence', c_float), ('inv_b', c_float), ('lr', c_double),
              ('table_in', RcdParam), ('bias_in', RcdParam), ('table_out', RcdParam), ('bias_out', RcdParam),
              ('pool_in', RcdPoolView), ('pool_tgt', RcdPoolView), ('same_pool', c_int), ('row0', c_int),
              ('rows', c_int), ('cap_rows', c_int), ('cap_n', c_int), ('cap_n_in', c_int), ('cap_nnz', c_longlong),
              ('cap_tnnz', c_longlong), ('ws', _P), ('ws_bytes', c_size_t), ('loss_acc', _P), ('bad_flag', _P),
              ('redo_flag', _P), ('user_pos', _P), ('scal', _P), ('scal_base', c_longlong), ('scal_len', c_longlong),
              ('next_items_in', _P), ('next_n_in', _P), ('next_cap_in', c_longlong), ('next_items_out', _P),
              ('next_n_out', _P), ('next_cap_out', c_longlong), ('stream_main', _P), ('stream_side', _P), ('stream_aux', _P),
              ('ip', RcdStepIp), ('out_dW_in', c_longlong), ('out_db_in', c_longlong), ('out_dW_out', c_longlong),
              ('out_db_out', c_longlong)]


STEP_ABI = 4
MODEL_IDS = {'ae': 0, 'mf': 1}
OPT_IDS = {'adam': 0, 'sgd': 1, 'adagrad': 2, 'rmsprop': 3}

_SIGNATURES.update({
  'rcd_step_args_size': (c_size_t, []),
  'rcd_step_create': (c_int, [ctypes.POINTER(_P)]),
  'rcd_step_destroy': (c_int, [_P]),
  'rcd_step_workspace_bytes': (c_size_t, [ctypes.POINTER(RcdStepArgs)]),
  'rcd_step_run': (c_int, [_P, ctypes.POINTER(RcdStepArgs)]),
  'rcd_step_join': (c_int, [_P, _P]),
  'rcd_step_profile': (c_int, [_P, c_int, ctypes.c_char_p]),
  'rcd_step_profile_read': (c_int, [_P, ctypes.c_char_p, c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_int), c_int]),
})

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())

_lib = None


def load():
  """Loads the shared library (no GPU needed to load it)."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.isfile(LIB_PATH):
    raise RuntimeError('recoder_b200: CUDA extension %s is missing — build it with '
                       '`python -m recoder_b200.csrc.build` (there is no CPU fallback)' % LIB_PATH)
  lib = ctypes.CDLL(LIB_PATH)
  for name, (res, args) in _SIGNATURES.items():
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args
  if lib.rcd_abi_version() != 1:
    raise RuntimeError('recoder_b200: ABI version mismatch')
  if lib.rcd_step_args_size() != ctypes.sizeof(RcdStepArgs):
    raise RuntimeError('recoder_b200: rcd_step_args layout mismatch (%d vs %d bytes)' %
                       (lib.rcd_step_args_size(), ctypes.sizeof(RcdStepArgs)))
  _lib = lib
  return lib


def require_cuda():
  if not torch.cuda.is_available():
    raise RuntimeError('recoder_b200 needs a CUDA device (sm_100a); there is no CPU path')


def ptr(t):
  """Device pointer of a tensor (None -> NULL)."""
  if t is None:
    return None
  return c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream_ptr():
  """Raw cudaStream_t of torch's current stream on the current device."""
  if _raw_stream is not None:
    return c_void_p(_raw_stream(torch.cuda.current_device()))
  return c_void_p(torch.cuda.current_stream().cuda_stream)


def check(status, what):
  if status != 0:
    msg = load().rcd_last_error().decode('utf-8', 'replace')
    raise RuntimeError('recoder_b200: %s failed (status %d): %s' % (what, status, msg))


# Optional CUDA-event timing of entry points (bench.py): PROFILE is None, 'all', or a set of entry-point names;
# TIMINGS maps entry-point name -> list of (start_event, end_event) recorded on the current stream.
PROFILE = None
TIMINGS = {}
_FN = {}


def call(name, *args):
  """Calls an `int`-status entry point on the current stream (stream appended automatically)."""
  fn = _FN.get(name)
  if fn is None:
    fn = _FN[name] = getattr(load(), name)
  if PROFILE is not None and (PROFILE == 'all' or name in PROFILE):
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    status = fn(*args, stream_ptr())
    end.record()
    TIMINGS.setdefault(name, []).append((start, end))
  else:
    status = fn(*args, stream_ptr())
  if status != 0:
    check(status, name)
