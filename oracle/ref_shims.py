"""Compat shims that let the UNMODIFIED reference (`/root/reference/recoder`) import on this image.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` and by CPU tests that pin the oracle against the
live reference when `/root/reference` is mounted (it is not on the GPU box).  Nothing under `recoder_b200/`
imports this.  No reference file is edited or copied; the five shims are the ones listed in SURVEY.md §8c:

1. stub module `glog`            (imported at recoder/model.py:3, recoder/embedding.py:5)
2. stub module `annoy`           (imported at recoder/embedding.py:1)
3. `numpy.int = int`             (recoder/metrics.py:11,25,34)
4. `scipy.sparse.sputils`        (recoder/data.py:6,51,66 use issequence / isintlike)
5. `torch.load(weights_only=False)` default (recoder/model.py:176 loads numpy arrays)
"""
import logging
import os
import sys
import types

# /root/reference exists in the build container only; `baseline/_ref` (pip install --target of the unmodified
# reference, git-ignored, travels with gpurun) is what bench.py's reference arm finds on the GPU box.
_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("RECODER_REFERENCE_ROOT"), "/root/reference",
               os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isfile(os.path.join(c, "recoder", "model.py"))),
                      "/root/reference")


def reference_available() -> bool:
  return os.path.isfile(os.path.join(REFERENCE_ROOT, "recoder", "model.py"))


def install() -> None:
  """Idempotently install the shims and put the reference on sys.path."""
  if not reference_available():
    raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)

  if "glog" not in sys.modules:
    glog = types.ModuleType("glog")
    _logger = logging.getLogger("glog")
    for name in ("debug", "info", "warning", "warn", "error", "critical", "exception"):
      setattr(glog, name, getattr(_logger, name if name != "warn" else "warning"))
    glog.setLevel = _logger.setLevel
    sys.modules["glog"] = glog

  if "annoy" not in sys.modules:
    annoy = types.ModuleType("annoy")

    class AnnoyIndex:  # never exercised on the training path
      def __init__(self, *a, **k):
        raise RuntimeError("annoy is not installed; the ANN index is out of scope")

    annoy.AnnoyIndex = AnnoyIndex
    sys.modules["annoy"] = annoy

  import numpy as np
  if not hasattr(np, "int"):
    np.int = int  # noqa

  import scipy.sparse
  import scipy.sparse._sputils as _sputils
  shim = types.ModuleType("scipy.sparse.sputils")
  shim.issequence = _sputils.issequence
  shim.isintlike = _sputils.isintlike
  sys.modules["scipy.sparse.sputils"] = shim
  scipy.sparse.sputils = shim

  import torch
  if not getattr(torch.load, "_recoder_shim", False):
    _orig_load = torch.load

    def _load(*args, **kwargs):
      kwargs.setdefault("weights_only", False)
      return _orig_load(*args, **kwargs)

    _load._recoder_shim = True
    torch.load = _load

  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)


def import_reference():
  """Returns the reference's (data, nn, losses, model) modules."""
  install()
  import warnings
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    import recoder.data as rdata
    import recoder.nn as rnn
    import recoder.losses as rlosses
    import recoder.model as rmodel
  return rdata, rnn, rlosses, rmodel
