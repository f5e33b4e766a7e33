"""Builds librecoder_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m recoder_b200.csrc.build [--force] [--verbose]

The library links cudart statically and does NOT link libcuda (the one driver symbol it needs,
cuTensorMapEncodeTiled, is resolved at run time with cudaGetDriverEntryPoint), so it loads on a CPU-only box.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ['core.cu', 'collate.cu', 'embed.cu', 'loss.cu', 'optim.cu', 'gemm.cu', 'gemm_simt.cu', 'gemm_tc.cu', 'decoder_tc.cu', 'sparse.cu', 'p2p.cu', 'dense.cu', 'eval.cu', 'step.cu']
HEADERS = ['common.cuh', 'scan.cuh', 'gemm_internal.cuh', 'tc_ptx.cuh', os.path.join(ROOT, 'include', 'recoder_b200.h')]
LIB = os.path.join(HERE, 'librecoder_b200.so')
OBJ_DIR = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
# fused decoder kernel: operand ring depth.  4 stages of 48 KB with single-buffered epilogue staging measured 0.231 ms
# for the C3 forward against 0.268 ms with 3 stages + double-buffered staging (profiles/README.md r02i)
DEC_STAGES = os.environ.get('RCD_DEC_STAGES', '4')
FLAGS = ['-DRCD_DEC_STAGES=' + DEC_STAGES, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '--expt-relaxed-constexpr',
         '--extended-lambda', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def _digest():
  h = hashlib.sha256()
  for f in SOURCES + HEADERS:
    p = f if os.path.isabs(f) else os.path.join(HERE, f)
    with open(p, 'rb') as fh:
      h.update(fh.read())
  h.update(' '.join(FLAGS).encode())
  return h.hexdigest()


def build(force=False, verbose=False):
  stamp = os.path.join(OBJ_DIR, 'stamp')
  digest = _digest()
  if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
    return LIB
  os.makedirs(OBJ_DIR, exist_ok=True)
  extra = ['-Xptxas', '-v'] if verbose else []

  def compile_one(src):
    obj = os.path.join(OBJ_DIR, src.replace('.cu', '.o'))
    cmd = [NVCC] + FLAGS + extra + ['-c', os.path.join(HERE, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, r

  with ThreadPoolExecutor(max_workers=8) as ex:
    results = list(ex.map(compile_one, SOURCES))
  objs = []
  for src, obj, r in results:
    if verbose or r.returncode != 0:
      sys.stderr.write('== %s\n%s%s' % (src, r.stdout, r.stderr))
    if r.returncode != 0:
      raise RuntimeError('nvcc failed on %s' % src)
    objs.append(obj)
  cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static']
  r = subprocess.run(cmd, capture_output=True, text=True)
  if r.returncode != 0:
    sys.stderr.write(r.stdout + r.stderr)
    raise RuntimeError('link failed')
  with open(stamp, 'w') as fh:
    fh.write(digest)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
