"""CPU tests of the validation post-processing oracle (val.py:313-382 restatement): pinned to what the reference's own
lines computed (tests/golden/valpost_seed*.npz, written by tests/golden/make_valpost_golden.py, which exec()s the fragment
of /root/reference/val.py) and to an independent explicit implementation (box filter written out, last writer wins)."""
import os

import numpy as np
import pytest

from oracle import val_post_oracle as VO


def make_case(seed, videos=5, B=6, T=16, nbatches=7):
    rng = np.random.RandomState(seed)
    lengths = rng.randint(30, 140, size=videos)
    lengths[1] = 7                                  # shorter than both filter windows
    batches = []
    for _ in range(nbatches):
        vid = rng.randint(0, videos, size=(B, 1)).repeat(T, 1)
        start = np.array([[rng.randint(1, lengths[v] + 1)] for v in vid[:, 0]])
        fid = start + np.arange(T)[None, :] * rng.randint(1, 3)          # some run past the end of the video
        v = rng.randn(B, T).astype(np.float32) * 0.9                      # some |x| > 1: clip matters
        a = rng.randn(B, T).astype(np.float32) * 0.9
        lv = rng.uniform(-1, 1, (B, T)).astype(np.float32)
        la = rng.uniform(-1, 1, (B, T)).astype(np.float32)
        lv[rng.rand(B, T) < 0.1] = -5.0
        la[rng.rand(B, T) < 0.1] = -5.0
        batches.append((v, a, lv, la, fid.astype(np.int32), vid.astype(np.int32)))
    return batches, lengths


def brute(batches, lengths, size_v, size_a):
    off = np.concatenate([[0], np.cumsum(lengths)])
    total = off[-1]
    pv, pa, lv_, la_ = (np.zeros(total) for _ in range(4))
    for (v, a, lv, la, fid, vid) in batches:
        for b in range(v.shape[0]):
            for t in range(v.shape[1]):
                L = lengths[vid[b, t]]
                if not (1 <= fid[b, t] <= L) or lv[b, t] == -5.0 or la[b, t] == -5.0:
                    continue
                i = off[vid[b, t]] + fid[b, t] - 1
                pv[i], pa[i], lv_[i], la_[i] = v[b, t], a[b, t], lv[b, t], la[b, t]

    def box(x, size):
        out = np.zeros_like(x)
        for k in range(len(lengths)):
            lo, hi = off[k], off[k + 1]
            c = np.clip(x[lo:hi], -1, 1)
            for i in range(hi - lo):
                j0, j1 = max(0, i - size // 2), min(hi - lo, i - size // 2 + size)
                out[lo + i] = c[j0:j1].sum() / size
        return out
    return box(pv, size_v), box(pa, size_a), lv_, la_


def test_oracle_matches_explicit_box_filter():
    for seed in (0, 1, 2):
        batches, lengths = make_case(seed)
        accv, acca, vout, aout = VO.val_ccc(batches, lengths, 20, 50)
        bv, ba, lv, la = brute(batches, lengths, 20, 50)
        assert np.abs(vout - bv).max() < 1e-12 and np.abs(aout - ba).max() < 1e-12
        assert abs(accv - VO.ccc(bv, lv)) < 1e-12 and abs(acca - VO.ccc(ba, la)) < 1e-12


def test_ignored_and_out_of_range_frames_stay_zero():
    batches, lengths = make_case(3)
    pv, pa, lv, la, sv, sa = VO.val_postprocess(batches, lengths)
    for (v, a, labv, laba, fid, vid) in batches[-1:]:
        for b in range(v.shape[0]):
            for t in range(v.shape[1]):
                k = int(vid[b, t])
                if fid[b, t] > lengths[k]:
                    continue
                if labv[b, t] == -5.0 or laba[b, t] == -5.0:
                    continue
                # the last batch is the last writer of every frame it touches (later t / b within it may overwrite)
                assert lv[k][fid[b, t] - 1] in labv[vid == k]
    assert all(len(sv[k]) == lengths[k] for k in sv)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_reference_val_py_fragment(seed, golden_dir):
    """The restatement vs the REFERENCE's lines val.py:313-382 executed on the same synthetic batches."""
    g = np.load(os.path.join(golden_dir, f"valpost_seed{seed}.npz"))
    batches = [(g["v"][i], g["a"][i], g["lv"][i], g["la"][i], g["fid"][i], g["vid"][i]) for i in range(g["v"].shape[0])]
    lengths = g["lengths"]
    _, _, label_v, label_a, sm_v, sm_a = VO.val_postprocess(batches, lengths, 20, 50)
    order = g["first_seen"].tolist()              # the reference concatenates videos in first-seen order (dict order)
    vout = np.concatenate([sm_v[k] for k in order])
    aout = np.concatenate([sm_a[k] for k in order])
    vtar = np.concatenate([np.asarray(label_v[k], dtype=np.float64) for k in order])
    atar = np.concatenate([np.asarray(label_a[k], dtype=np.float64) for k in order])
    assert np.abs(vout - g["vout"]).max() < 1e-6 and np.abs(aout - g["aout"]).max() < 1e-6
    assert np.array_equal(vtar, g["vtar"]) and np.array_equal(atar, g["atar"])
    accv, acca, _, _ = VO.val_ccc(batches, lengths, 20, 50)
    assert abs(accv - float(g["accV"])) < 1e-7 and abs(acca - float(g["accA"])) < 1e-7
