"""Validation post-processing on the device -- the step the reference runs right after the hot
path as a Python loop over `.cpu()` copies of every batch (val.py:313-382): scatter per-frame predictions into
per-video arrays by frame id (skipping -5 labels), clip to [-1, 1], `uniform_filter1d(size=20 / 50, mode='constant')`,
then the global CCC.  Here every batch costs two tiny kernels and nothing is copied to the host until `finalize()`.

    acc = ValPostprocessor(video_lengths)                  # list of frames per video, in first-seen order
    for batch: acc.update(vouts, aouts, labelsV, labelsA, frame_ids, video_idx)      # all (B, T) CUDA tensors
    ccc_v, ccc_a = acc.finalize()                          # val.py:381-382 accV, accA

Data-parallel use: shard by VIDEO (every window of a video on one rank); `finalize(group=...)` all-reduces the
(2, 6) fp64 sums.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib as L
from .engine import _ptr, _stream, cuda_memset0, require_cuda


class ValPostprocessor:
    def __init__(self, video_lengths: Sequence[int], device=None, ignore: float = -5.0):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lens = torch.as_tensor(list(video_lengths), dtype=torch.int64)
        assert lens.numel() >= 1 and int(lens.min()) >= 0
        off = torch.zeros(lens.numel() + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(lens, 0)
        self.videos = int(lens.numel())
        self.total = int(off[-1])
        self.offsets = off.to(self.device)
        self.ignore = float(ignore)
        z = lambda dt: torch.empty(max(self.total, 1), dtype=dt, device=self.device)      # noqa: E731
        self.pred_v, self.pred_a, self.label_v, self.label_a = z(torch.float32), z(torch.float32), z(torch.float32), z(torch.float32)
        self.stamp = z(torch.int64)
        self.clear()

    def clear(self):
        for t in (self.pred_v, self.pred_a, self.label_v, self.label_a, self.stamp):
            cuda_memset0(t)                     # pred / label start at 0 (val.py:326-329), stamps at "never written"
        self.seq = 0

    def update(self, vouts, aouts, labels_v, labels_a, frame_ids, video_idx):
        """One validation batch; every argument is a CUDA tensor with B*T elements in the same (b, t) order
        (val.py:313-321 iterates batch-major, time-minor).  frame_ids are 1-based; video_idx indexes video_lengths."""
        require_cuda(vouts, aouts, labels_v, labels_a, frame_ids, video_idx)
        f32 = lambda t: t.reshape(-1).contiguous().float()          # noqa: E731
        i32 = lambda t: t.reshape(-1).contiguous().to(torch.int32)  # noqa: E731
        v, a, lv, la = f32(vouts), f32(aouts), f32(labels_v), f32(labels_a)
        fid, vid = i32(frame_ids), i32(video_idx)
        n = v.numel()
        assert a.numel() == n and lv.numel() == n and la.numel() == n and fid.numel() == n and vid.numel() == n
        self.seq += 1
        L.check(L.lib().jmt_valpost_scatter(_ptr(v), _ptr(a), _ptr(lv), _ptr(la), _ptr(fid), _ptr(vid), n, _ptr(self.offsets),
                                            self.ignore, self.seq, _ptr(self.stamp), _ptr(self.pred_v), _ptr(self.pred_a),
                                            _ptr(self.label_v), _ptr(self.label_a), _stream()), "jmt_valpost_scatter")

    def sums(self, size_v: int = 20, size_a: int = 50, smooth_out: Optional[tuple] = None, group=None) -> torch.Tensor:
        """(2, 6) fp64 CCC sums of the clipped + smoothed predictions vs the labels (row 0 valence, row 1 arousal)."""
        s = torch.empty((2, 6), dtype=torch.float64, device=self.device)
        cuda_memset0(s)
        sv, sa = smooth_out if smooth_out is not None else (None, None)
        L.check(L.lib().jmt_valpost_finalize(_ptr(self.pred_v), _ptr(self.pred_a), _ptr(self.label_v), _ptr(self.label_a),
                                             _ptr(self.offsets), self.videos, self.total, size_v, size_a, _ptr(sv), _ptr(sa),
                                             _ptr(s), _stream()), "jmt_valpost_finalize")
        if group is not None:
            torch.distributed.all_reduce(s, group=None if group is True else group)
        return s

    def finalize(self, size_v: int = 20, size_a: int = 50, return_smoothed: bool = False, group=None):
        """(accV, accA) of val.py:381-382; with return_smoothed also the per-frame smoothed predictions (val.py:362-367)."""
        from .cccmetric import ccc_from_sums
        sm = None
        if return_smoothed:
            sm = (torch.empty(max(self.total, 1), dtype=torch.float32, device=self.device),
                  torch.empty(max(self.total, 1), dtype=torch.float32, device=self.device))
        val = ccc_from_sums(self.sums(size_v, size_a, sm, group)).cpu()
        out = (float(val[0]), float(val[1]))
        if return_smoothed:
            return out + (sm[0][:self.total], sm[1][:self.total])
        return out
