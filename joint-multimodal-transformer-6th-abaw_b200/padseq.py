"""Collate zero-fill semantics of padSequence.py:4-101 (Train/Val/TestPadSequence) for spectrogram
batches already on the device: a zero tensor (B,16,1,64,maxW) with each item RIGHT-aligned.  The
reference's branch condition compares the mel-bin dim (`shape[2]` = 64) instead of the width
(padSequence.py:16,46,87): that behaviour is kept -- when 64 >= maxW the item is copied
whole, which (as in the reference) requires equal widths."""
from __future__ import annotations

from typing import Sequence

import torch

from . import _lib as L
from .engine import _ptr, _stream, require_cuda


def pad_spectrograms(specs: Sequence[torch.Tensor]) -> torch.Tensor:
    require_cuda(*specs)
    widths = [int(s.shape[3]) for s in specs]
    max_w = max(widths)
    out = torch.empty((len(specs), 16, 1, 64, max_w), dtype=torch.float32, device=specs[0].device)
    lib = L.lib()
    for i, s in enumerate(specs):
        if not (s.shape[2] < max_w) and s.shape[3] != max_w:
            raise RuntimeError(
                f"The expanded size of the tensor ({max_w}) must match the existing size ({s.shape[3]}) "
                "(reference padSequence.py:21 behaviour)")
        sc = s.contiguous().float()
        rows = sc.numel() // sc.shape[3]
        L.check(lib.jmt_pad_right_align(_ptr(sc), rows, sc.shape[3], _ptr(out[i]), max_w, _stream()), "jmt_pad_right_align")
    return out


class TrainPadSequence:
    """padSequence.py:4-31 for device-resident samples (clip, spectrogram, labelV, labelA, wavfile)."""

    def __call__(self, sorted_batch):
        audio = pad_spectrograms([x[1] for x in sorted_batch])
        visual = torch.stack([x[0] for x in sorted_batch])
        return (visual, audio, torch.stack([x[2] for x in sorted_batch]), torch.stack([x[3] for x in sorted_batch]),
                [x[4] for x in sorted_batch])


class ValPadSequence:
    """padSequence.py:34-72: samples (clip, spectrogram, frameids, v_ids, v_lengths, labelV, labelA, wavfile)."""

    def __call__(self, sorted_batch):
        audio = pad_spectrograms([x[1] for x in sorted_batch])
        visual = torch.stack([x[0] for x in sorted_batch])
        return (visual, audio, [x[2] for x in sorted_batch], [x[3] for x in sorted_batch], [x[4] for x in sorted_batch],
                torch.stack([x[5] for x in sorted_batch]), torch.stack([x[6] for x in sorted_batch]),
                [x[7] for x in sorted_batch])


class TestPadSequence:
    """padSequence.py:75-101: samples (clip, spectrogram, frameids, v_ids, v_lengths, wavfile); no labels."""
    __test__ = False          # not a pytest class

    def __call__(self, sorted_batch):
        audio = pad_spectrograms([x[1] for x in sorted_batch])
        visual = torch.stack([x[0] for x in sorted_batch])
        return (visual, audio, [x[2] for x in sorted_batch], [x[3] for x in sorted_batch], [x[4] for x in sorted_batch],
                [x[5] for x in sorted_batch])
