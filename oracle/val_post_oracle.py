"""TEST INFRASTRUCTURE ONLY (CPU oracle) -- restatement of the reference's validation post-processing,
val.py:313-382, for checking jmt_b200.valpost.  Nothing in the product imports this file.

Parity pinning: PINNED to the reference's own lines.  `validate()` is monolithic (data loaders, backbones), so the fragment
cannot be imported -- tests/golden/make_valpost_golden.py slices val.py:69-79, 313-357 and 359-382 out of the reference source
at run time, exec()s them on synthetic batches and stores inputs + results as tests/golden/valpost_seed*.npz;
tests/test_valpost_cpu.py checks this restatement against those files (and against an explicit box-filter loop), and
tests/test_valpost_gpu.py checks the device implementation against the same files.
"""
import numpy as np
from scipy.ndimage import uniform_filter1d


def val_postprocess(batches, video_lengths, size_v=20, size_a=50, ignore=-5.0):
    """batches: iterable of (vouts, aouts, labelsV, labelsA, frame_ids, video_idx), each (B, T) array-like.
    Returns (pred_v, pred_a, label_v, label_a, smooth_v, smooth_a) as per-video lists (dict order = video index)."""
    pred_v = {v: [0] * int(n) for v, n in enumerate(video_lengths)}          # val.py:326-329 (allocated on first sight)
    pred_a = {v: [0] * int(n) for v, n in enumerate(video_lengths)}
    label_v = {v: [0] * int(n) for v, n in enumerate(video_lengths)}
    label_a = {v: [0] * int(n) for v, n in enumerate(video_lengths)}
    for vouts, aouts, labV, labA, fids, vids in batches:                    # val.py:313-321
        for b in range(len(vouts)):
            for t in range(len(vouts[b])):
                vid, frameid = int(vids[b][t]), int(fids[b][t])
                length = int(video_lengths[vid])
                if frameid < 1 or frameid > length:                          # val.py:346 (`frameid <= length`)
                    continue
                if labA[b][t] == ignore or labV[b][t] == ignore:             # val.py:335-339, 348-352
                    continue
                pred_a[vid][frameid - 1] = float(aouts[b][t])                # val.py:354-357: later windows overwrite
                pred_v[vid][frameid - 1] = float(vouts[b][t])
                label_a[vid][frameid - 1] = float(labA[b][t])
                label_v[vid][frameid - 1] = float(labV[b][t])
    smooth_v, smooth_a = {}, {}
    for key in pred_a:                                                       # val.py:362-367
        smooth_v[key] = uniform_filter1d(np.clip(pred_v[key], -1.0, 1.0).astype(np.float64), size=size_v, mode='constant') \
            if len(pred_v[key]) else np.zeros(0)
        smooth_a[key] = uniform_filter1d(np.clip(pred_a[key], -1.0, 1.0).astype(np.float64), size=size_a, mode='constant') \
            if len(pred_a[key]) else np.zeros(0)
    return pred_v, pred_a, label_v, label_a, smooth_v, smooth_a


def ccc(x, y):
    """EvaluationMetrics/cccmetric.py:4-21 (population std): the restatement pinned to the reference-generated golden
    values in oracle/jmt_oracle.py."""
    from oracle.jmt_oracle import ccc_metric
    return ccc_metric(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))


def val_ccc(batches, video_lengths, size_v=20, size_a=50, ignore=-5.0):
    """(accV, accA) of val.py:369-382: CCC over the concatenation of all videos (untouched frames are (0, 0) pairs)."""
    _, _, label_v, label_a, sm_v, sm_a = val_postprocess(batches, video_lengths, size_v, size_a, ignore)
    vout = np.concatenate([sm_v[k] for k in sm_v])
    aout = np.concatenate([sm_a[k] for k in sm_a])
    vtar = np.concatenate([np.asarray(label_v[k], dtype=np.float64) for k in label_v])
    atar = np.concatenate([np.asarray(label_a[k], dtype=np.float64) for k in label_a])
    return ccc(vout, vtar), ccc(aout, atar), vout, aout
