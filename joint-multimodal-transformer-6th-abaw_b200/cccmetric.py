"""EvaluationMetrics/cccmetric.py on the device: ccc, cccva, ccc_numpy, CCCMetric -- same names and
argument meaning.  Predictions/labels may be CUDA tensors (stay on device, one reduction kernel) or
host arrays/lists (copied once).  Results are Python floats like the reference's numpy scalars."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .engine import _ptr, _stream
from .losses import six_sums


def _dev(a, device=None):
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))
    if not t.is_cuda:
        t = t.to(device or torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
    return t.reshape(-1).float()


def _finalize(sums: torch.Tensor, kind: int) -> torch.Tensor:
    npairs = sums.shape[0]
    val = torch.empty((npairs,), dtype=torch.float32, device=sums.device)
    L.check(L.lib().jmt_ccc_finalize(_ptr(sums), npairs, kind, 0.0, 1e-8, _ptr(val), None, _stream()), "jmt_ccc_finalize")
    return val


def ccc_from_sums(sums: torch.Tensor, kind: int = L.CCC_METRIC) -> torch.Tensor:
    """Device finaliser over (npairs, 6) fp64 sums (e.g. after an all-reduce of per-rank partials)."""
    return _finalize(sums, kind)


def ccc(x, y, ignore=-5.0, group=None):
    """cccmetric.py:4-21 (population std; `ignore` is unused there and here).  len <= 1: the reference
    calls sys.exit(); here a ValueError is raised."""
    xt = _dev(x)
    yt = _dev(y, xt.device)
    if yt.numel() <= 1:
        raise ValueError("ccc needs more than one sample (the reference calls sys.exit())")
    return float(_finalize(six_sums(xt, yt, None, group=group), L.CCC_METRIC)[0])


def cccva(y_true, y_pred, ignore=-5.0):
    """cccmetric.py:24-38: columns 0/1 = valence/arousal."""
    yt = y_true if isinstance(y_true, torch.Tensor) else torch.from_numpy(np.asarray(y_true, dtype=np.float32))
    yp = y_pred if isinstance(y_pred, torch.Tensor) else torch.from_numpy(np.asarray(y_pred, dtype=np.float32))
    ccc_v = ccc(yt[:, 0].contiguous(), yp[:, 0].contiguous(), ignore)
    ccc_a = ccc(yt[:, 1].contiguous(), yp[:, 1].contiguous(), ignore)
    return ccc_v, ccc_a, (ccc_v + ccc_a) / 2


def ccc_numpy(y_true, y_pred):
    """cccmetric.py:41-56 (np.cov N-1 over np.var N, +1e-8)."""
    xt = _dev(y_true)
    yt = _dev(y_pred, xt.device)
    return float(_finalize(six_sums(xt, yt, None), L.CCC_NUMPY)[0])


class CCCMetric(object):
    """cccmetric.py:59-79 accumulate/get API.  The reference's get() prints shapes and sys.exit()s
    (SURVEY Q11); this one returns the CCC of the stacked arrays.  Partial sums are accumulated on the
    device per update, so get() is O(1) and data-parallel ranks combine with one 48-byte all-reduce."""

    def __init__(self, ignore_index=-5.0):
        self.ignore = ignore_index
        self.sums = None

    def update(self, y_pred, y_true):
        xt = _dev(y_pred)
        yt = _dev(y_true, xt.device)
        self.sums = six_sums(xt, yt, None, out=self.sums)

    def clear(self):
        self.sums = None

    def get(self, group=None):
        """CCC of everything accumulated.  group=None (default): this process's sums only -- safe to call on one rank for
        logging.  group=True / a ProcessGroup: all-reduce the partial sums first (a COLLECTIVE: every rank of the group must
        call it), same convention as losses.six_sums and valpost.ValPostprocessor."""
        if self.sums is None:
            raise ValueError("CCCMetric.get() before update()")
        s = self.sums.clone()
        if group is not None:
            torch.distributed.all_reduce(s, group=None if group is True else group)
        return float(_finalize(s, L.CCC_METRIC)[0])
