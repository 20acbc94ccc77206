"""Pin the validation post-processing to the REFERENCE'S OWN LINES.

`validate()` in /root/reference/val.py is monolithic (data loaders, backbones, CUDA), so the fragment that follows the hot
path cannot be imported -- but it can be executed: this script slices val.py:69-79 (accumulator initialisation),
val.py:313-357 (the per-batch scatter loop) and val.py:359-382 (clip, uniform_filter1d, global CCC) out of the reference
source AT RUN TIME (nothing is copied into the repo), dedents them and exec()s them on synthetic per-batch inputs named as
in the reference (`audiovisual_vouts`, `audiovisual_aouts`, `labelsV`, `labelsA`, `frame_ids`, `videos`, `vid_lengths`).
The inputs and what the reference's code computed from them are written to tests/golden/valpost_*.npz; the tests compare
oracle/val_post_oracle.py (CPU) and jmt_b200.valpost (GPU) with them.

Run in the build container only (needs /root/reference):   python tests/golden/make_valpost_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("JMT_REFERENCE", "/root/reference")


def _fragment(lines, first, last, tabs, must_start, must_end):
    """Lines first..last (1-based, inclusive) of val.py with `tabs` leading tabs removed; whitespace-only lines blanked."""
    out = []
    for ln in lines[first - 1:last]:
        if not ln.strip():
            out.append("")
            continue
        assert ln.startswith("\t" * tabs), (first, last, repr(ln))
        out.append(ln[tabs:])
    assert must_start in out[0], (must_start, out[0])
    assert must_end in [l for l in out if l][-1], (must_end, out[-1])
    return "\n".join(out) + "\n"


def reference_fragments():
    lines = open(os.path.join(REF, "val.py")).read().split("\n")
    init = _fragment(lines, 69, 79, 1, "vout = []", "count = 0")
    loop = _fragment(lines, 313, 357, 3, "for voutputs, aoutputs, labelV, labelA, frameids, video, vid_length in zip(", "label_v[vid][frameid-1] = labV")
    post = _fragment(lines, 359, 382, 1, "_smooth_pred_v = {}", "accA = ccc(np.array(aout), np.array(atar))")
    return init, loop, post


def make_case(seed, videos=5, B=6, T=16, nbatches=7):
    """Windows of T consecutive (or strided) frames; the FIRST time a video appears it starts at frame 1 (the reference
    exits otherwise, val.py:324-329); some windows run past the end of their video, some labels are the -5 sentinel, some
    predictions exceed [-1, 1]."""
    rng = np.random.RandomState(seed)
    lengths = rng.randint(30, 140, size=videos)
    lengths[1] = 7                                   # shorter than both filter windows and than one window
    order = rng.permutation(videos)                  # first-seen order != index order
    seen = set()
    batches = []
    for bi in range(nbatches):
        vid = np.zeros((B, T), np.int32)
        fid = np.zeros((B, T), np.int32)
        for b in range(B):
            k = int(order[(bi * B + b) % videos]) if len(seen) < videos else int(rng.randint(0, videos))
            start = 1 if k not in seen else int(rng.randint(1, lengths[k] + 1))
            seen.add(k)
            vid[b] = k
            fid[b] = start + np.arange(T) * int(rng.randint(1, 3))
        v = (rng.randn(B, T) * 0.9).astype(np.float32)
        a = (rng.randn(B, T) * 0.9).astype(np.float32)
        lv = rng.uniform(-1, 1, (B, T)).astype(np.float32)
        la = rng.uniform(-1, 1, (B, T)).astype(np.float32)
        lv[rng.rand(B, T) < 0.1] = -5.0
        la[rng.rand(B, T) < 0.1] = -5.0
        batches.append((v, a, lv, la, fid, vid))
    return batches, lengths


def run_reference(batches, lengths):
    sys.path.insert(0, REF)
    from scipy.ndimage import uniform_filter1d          # val.py:7
    from EvaluationMetrics.cccmetric import ccc         # val.py:11
    init, loop, post = reference_fragments()
    ns = {"np": np, "sys": sys, "uniform_filter1d": uniform_filter1d, "ccc": ccc, "store_results_pkl": ""}
    exec(init, ns)
    for (v, a, lv, la, fid, vid) in batches:
        B, T = v.shape
        ns.update(audiovisual_vouts=v, audiovisual_aouts=a, labelsV=lv, labelsA=la,
                  frame_ids=[[int(x) for x in fid[b]] for b in range(B)],
                  videos=[[f"video{int(x)}" for x in vid[b]] for b in range(B)],
                  vid_lengths=[[int(lengths[int(x)])] * T for x in vid[:, 0]])
        exec(loop, ns)
    exec(post, ns)
    keys = list(ns["pred_a"].keys())                     # first-seen order = order of the concatenated outputs
    first_seen = [int(k[len("video"):]) for k in keys]
    return dict(accV=float(ns["accV"]), accA=float(ns["accA"]), vout=np.asarray(ns["vout"], np.float64),
                aout=np.asarray(ns["aout"], np.float64), vtar=np.asarray(ns["vtar"], np.float64),
                atar=np.asarray(ns["atar"], np.float64), first_seen=np.asarray(first_seen, np.int32))


def main():
    for seed in (0, 1, 2):
        batches, lengths = make_case(seed)
        r = run_reference(batches, lengths)
        assert sorted(r["first_seen"].tolist()) == list(range(len(lengths)))
        np.savez_compressed(os.path.join(HERE, f"valpost_seed{seed}.npz"), lengths=lengths.astype(np.int64),
                            v=np.stack([b[0] for b in batches]), a=np.stack([b[1] for b in batches]),
                            lv=np.stack([b[2] for b in batches]), la=np.stack([b[3] for b in batches]),
                            fid=np.stack([b[4] for b in batches]), vid=np.stack([b[5] for b in batches]), **r)
        print(f"valpost_seed{seed}: accV={r['accV']:.9f} accA={r['accA']:.9f} frames={len(r['vout'])} first_seen={r['first_seen'].tolist()}")


if __name__ == "__main__":
    main()
