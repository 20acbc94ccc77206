"""CCC losses on the device: one single-pass six-sum reduction kernel + a device-side finaliser and a
closed-form backward.  No host synchronisation: the loss stays a device scalar.

  LiveCCCLoss / CCCLoss(digitize_num=1) : losses/loss.py:9-32  (the criterion main.py:794 builds)
  CCCLossMasked(ignore=-5.0)            : losses/CCCLoss.py:8-43
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import nn

from . import _lib as L
from .engine import _ptr, _stream, cuda_memset0, require_cuda


def six_sums(x: torch.Tensor, y: torch.Tensor, ignore: Optional[float] = None, out: Optional[torch.Tensor] = None,
             group=None) -> torch.Tensor:
    """(npairs, 6) fp64 sums (N, Sx, Sy, Sxy, Sxx, Syy) of row-paired fp32 matrices x, y (npairs, n);
    `out` accumulates (partial shards).  group=True / a ProcessGroup: all-reduce the sums."""
    require_cuda(x, y)
    x2 = x.reshape(1, -1) if x.dim() == 1 else x.reshape(x.shape[0], -1)
    y2 = y.reshape(1, -1) if y.dim() == 1 else y.reshape(y.shape[0], -1)
    if x2.dtype != torch.float32 or y2.dtype != torch.float32:
        raise RuntimeError("six_sums expects fp32 predictions and labels")
    x2, y2 = x2.contiguous(), y2.contiguous()
    assert x2.shape == y2.shape
    npairs, n = x2.shape
    if out is None:
        out = torch.empty((npairs, 6), dtype=torch.float64, device=x.device)
        cuda_memset0(out)
    lib = L.lib()
    L.check(lib.jmt_ccc_sums(_ptr(x2), _ptr(y2), n, npairs, n, 0 if ignore is None else 1,
                             0.0 if ignore is None else float(ignore), _ptr(out), _stream()), "jmt_ccc_sums")
    if group is not None:                     # True = default process group
        torch.distributed.all_reduce(out, group=None if group is True else group)
    return out


class _CCCLossFn(torch.autograd.Function):
    @staticmethod
    def forward(actx, x, y, kind, ignore, eps, group):
        xf = x.reshape(1, -1).contiguous()
        yf = y.reshape(1, -1).contiguous()
        if xf.dtype != torch.float32:
            xf = xf.float()
        if yf.dtype != torch.float32:
            yf = yf.float()
        n = xf.shape[1]
        sums = six_sums(xf, yf, ignore, group=group)
        lib = L.lib()
        val = torch.empty((1,), dtype=torch.float32, device=x.device)
        coef = torch.empty((1, 4), dtype=torch.float64, device=x.device)
        n_all = float(y.shape[0]) if kind == L.CCC_LOSS_MASKED else float(n)
        L.check(lib.jmt_ccc_finalize(_ptr(sums), 1, kind, n_all, eps, _ptr(val), _ptr(coef), _stream()), "jmt_ccc_finalize")
        actx.save_for_backward(xf, yf, coef)
        actx.ignore, actx.xshape = ignore, x.shape
        return val.reshape(())

    @staticmethod
    def backward(actx, g):
        xf, yf, coef = actx.saved_tensors
        lib = L.lib()
        dx = torch.empty_like(xf)
        gg = g.reshape(1).contiguous().float()
        ign = actx.ignore
        L.check(lib.jmt_ccc_bwd(_ptr(xf), _ptr(yf), xf.shape[1], 1, xf.shape[1], _ptr(coef), _ptr(gg), 0,
                                0 if ign is None else 1, 0.0 if ign is None else float(ign), _ptr(dx), _stream()),
                "jmt_ccc_bwd")
        return dx.view(actx.xshape), None, None, None, None, None


class _CCCLossPairFn(torch.autograd.Function):
    """loss(v, lv) + loss(a, la) with ONE (2, 6) fp64 sums buffer: two reduction launches fill its rows, one all-reduce
    (96 bytes) combines the ranks, one finaliser emits both values and both coefficient sets."""

    @staticmethod
    def forward(actx, v, lv, a, la, kind, eps, group):
        flat = lambda t: (t.reshape(1, -1).contiguous().float() if t.dtype != torch.float32 else t.reshape(1, -1).contiguous())  # noqa: E731
        xs, ys = (flat(v), flat(a)), (flat(lv), flat(la))
        n = xs[0].shape[1]
        assert xs[1].shape[1] == n and ys[0].shape[1] == n and ys[1].shape[1] == n
        lib = L.lib()
        sums = torch.empty((2, 6), dtype=torch.float64, device=v.device)
        cuda_memset0(sums)
        for i in range(2):
            L.check(lib.jmt_ccc_sums(_ptr(xs[i]), _ptr(ys[i]), n, 1, n, 0, 0.0, _ptr(sums[i]), _stream()), "jmt_ccc_sums")
        if group is not None:
            torch.distributed.all_reduce(sums, group=None if group is True else group)
        val = torch.empty((2,), dtype=torch.float32, device=v.device)
        coef = torch.empty((2, 4), dtype=torch.float64, device=v.device)
        L.check(lib.jmt_ccc_finalize(_ptr(sums), 2, kind, float(n), eps, _ptr(val), _ptr(coef), _stream()), "jmt_ccc_finalize")
        total = torch.empty((1,), dtype=torch.float32, device=v.device)
        cuda_memset0(total)
        L.check(lib.jmt_axpy(_ptr(val[0:1]), _ptr(total), 1.0, 1, L.F32, _stream()), "jmt_axpy")
        L.check(lib.jmt_axpy(_ptr(val[1:2]), _ptr(total), 1.0, 1, L.F32, _stream()), "jmt_axpy")
        actx.save_for_backward(xs[0], ys[0], xs[1], ys[1], coef)
        actx.shapes = (v.shape, a.shape)
        return total.reshape(())

    @staticmethod
    def backward(actx, g):
        x0, y0, x1, y1, coef = actx.saved_tensors
        lib = L.lib()
        gg = g.reshape(1).contiguous().float()
        outs = []
        for i, (x, y) in enumerate(((x0, y0), (x1, y1))):
            dx = torch.empty_like(x)
            L.check(lib.jmt_ccc_bwd(_ptr(x), _ptr(y), x.shape[1], 1, x.shape[1], _ptr(coef[i]), _ptr(gg), 0, 0, 0.0, _ptr(dx),
                                    _stream()), "jmt_ccc_bwd")
            outs.append(dx.view(actx.shapes[i]))
        return outs[0], None, outs[1], None, None, None, None


class CCCLoss(nn.Module):
    """losses/loss.py::CCCLoss(digitize_num, range=[-1,1], eps=1e-8) -- the LIVE training criterion
    (main.py:794 uses digitize_num=1).  Only digitize_num == 1 is on the hot path; other values
    (softmax over bins, unused by main.py) raise NotImplementedError.  Unbiased std, eps only in rho's
    denominator, NO -5 masking.  `global_stats=True` all-reduces the six sums across the
    default process group inside forward so every rank optimises the global-batch CCC."""

    def __init__(self, digitize_num, range=[-1, 1], eps=1e-8, global_stats: bool = False):
        super().__init__()
        if digitize_num != 1:
            raise NotImplementedError("only digitize_num=1 (main.py:794) is on the hot path")
        self.digitize_num, self.range, self.eps, self.global_stats = digitize_num, range, eps, global_stats

    def forward(self, x, y):
        group = True if (self.global_stats and torch.distributed.is_initialized()) else None
        return _CCCLossFn.apply(x, y, L.CCC_LOSS_LIVE, None, float(self.eps), group)

    def forward_va(self, v, labels_v, a, labels_a):
        """`self(v, labels_v) + self(a, labels_a)` (train.py:309-311) as one call: with global_stats the valence and arousal
        sums travel in ONE 96-byte all-reduce instead of two stream-serialised 48-byte ones (an addition to the reference's
        criterion API; same value and gradients)."""
        group = True if (self.global_stats and torch.distributed.is_initialized()) else None
        return _CCCLossPairFn.apply(v, labels_v, a, labels_a, L.CCC_LOSS_LIVE, float(self.eps), group)


LiveCCCLoss = CCCLoss


class CCCLossMasked(nn.Module):
    """losses/CCCLoss.py::CCCLoss(ignore=-5.0): y_true != ignore mask, unbiased variances, Sxy divided by
    the PRE-mask length, 0 when <= 1 valid element.  forward(y_pred (N,), y_true (N,))."""

    def __init__(self, ignore=-5.0):
        super().__init__()
        self.ignore = ignore

    def forward(self, y_pred, y_true):
        return _CCCLossFn.apply(y_pred, y_true, L.CCC_LOSS_MASKED, float(self.ignore), 1e-8, None)


def label_mask(y_true: torch.Tensor, ignore: float = -5.0) -> torch.Tensor:
    """y_true != ignore as a bool tensor (losses/CCCLoss.py:19) -- the reference's only padding mask."""
    require_cuda(y_true)
    y = y_true.contiguous().float() if y_true.dtype != torch.float32 else y_true.contiguous()
    mask = torch.empty(y.shape, dtype=torch.uint8, device=y.device)
    L.check(L.lib().jmt_label_mask(_ptr(y), y.numel(), float(ignore), _ptr(mask), _stream()), "jmt_label_mask")
    return mask.view(torch.bool)
