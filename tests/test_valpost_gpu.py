"""GPU parity of jmt_b200.valpost against the CPU oracle of val.py:313-382."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import jmt_b200  # noqa: E402
from jmt_b200.valpost import ValPostprocessor  # noqa: E402
from oracle import val_post_oracle as VO  # noqa: E402
from test_valpost_cpu import make_case  # noqa: E402


@pytest.mark.parametrize("seed,sizes", [(0, (20, 50)), (1, (20, 50)), (2, (3, 8)), (4, (1, 1))])
def test_valpost_matches_oracle(seed, sizes):
    batches, lengths = make_case(seed, videos=6, B=8, T=16, nbatches=9)
    accv, acca, vout, aout = VO.val_ccc(batches, lengths, *sizes)
    dev = torch.device("cuda")
    acc = ValPostprocessor(lengths.tolist(), dev)
    for (v, a, lv, la, fid, vid) in batches:
        acc.update(*(torch.from_numpy(x).to(dev) for x in (v, a, lv, la, fid, vid)))
    gv, ga, sv, sa = acc.finalize(*sizes, return_smoothed=True)
    assert np.abs(sv.cpu().numpy() - vout).max() < 1e-6 and np.abs(sa.cpu().numpy() - aout).max() < 1e-6
    assert abs(gv - accv) < 1e-6 and abs(ga - acca) < 1e-6, (gv, accv, ga, acca)      # north-star: CCC within 1e-4
    # scatter state is bit-exact (fp32 copies of the winning elements; untouched frames 0)
    pv, pa, lvv, laa, _, _ = VO.val_postprocess(batches, lengths, *sizes)
    want = np.concatenate([np.asarray(pv[k], dtype=np.float32) for k in pv])
    assert np.array_equal(acc.pred_v[:acc.total].cpu().numpy(), want)
    want_l = np.concatenate([np.asarray(laa[k], dtype=np.float32) for k in laa])
    assert np.array_equal(acc.label_a[:acc.total].cpu().numpy(), want_l)
    # clear() resets the accumulator
    acc.clear()
    assert float(acc.pred_v.abs().sum()) == 0.0


def test_valpost_on_engine_predictions():
    """End of the path: Two_transformers predictions (B, T) -> valpost -> CCC, vs the oracle on the same predictions."""
    from oracle import jmt_oracle as O
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = jmt_b200.Two_transformers(0.0, 0.0, 2, 1, "FC", "FC", 512, precision="fp32").to(dev).eval()
    B, T = 4, 16
    lengths = [40, 64, 25]
    batches = []
    acc = ValPostprocessor(lengths, dev)
    rng = np.random.RandomState(0)
    for i in range(3):
        aud, vis = (t.to(dev) for t in O.synth_features(B, T, [512, 512], 10 + i))
        with torch.no_grad():
            v, a = model(aud, vis)                                  # (B, T) for joint_modalities='FC'
        vid = rng.randint(0, 3, size=(B, 1)).repeat(T, 1).astype(np.int32)
        fid = (rng.randint(1, 20, size=(B, 1)) + np.arange(T)[None]).astype(np.int32)
        lv = rng.uniform(-1, 1, (B, T)).astype(np.float32)
        la = rng.uniform(-1, 1, (B, T)).astype(np.float32)
        la[rng.rand(B, T) < 0.15] = -5.0
        acc.update(v, a, torch.from_numpy(lv).to(dev), torch.from_numpy(la).to(dev), torch.from_numpy(fid).to(dev),
                   torch.from_numpy(vid).to(dev))
        batches.append((v.cpu().numpy(), a.cpu().numpy(), lv, la, fid, vid))
    accv, acca, _, _ = VO.val_ccc(batches, lengths)
    gv, ga = acc.finalize()
    assert abs(gv - accv) < 1e-6 and abs(ga - acca) < 1e-6


def test_feature_shard_loader(tmp_path):
    """the asynchronous loader yields every window of its rank's share once, in order, bit-exact."""
    from jmt_b200 import features as F
    rng = np.random.RandomState(1)
    W, T, B = 22, 10, 4
    paths = []
    allv, alll = [], []
    for k in range(2):
        vis = rng.randn(W, 16, T).astype(np.float32)
        aud = rng.randn(W, T, 8).astype(np.float32)
        lv = rng.uniform(-1, 1, (W, T)).astype(np.float32)
        la = rng.uniform(-1, 1, (W, T)).astype(np.float32)
        p = str(tmp_path / f"s{k}.jmtshard")
        F.write_shard(p, vis, aud, lv, la)
        paths.append(p)
        allv.append(torch.from_numpy(vis).to(torch.bfloat16))
        alll.append(torch.from_numpy(lv))
    for rank, world in ((0, 1), (1, 2)):
        got_v, got_l = [], []
        for aud_d, vis_d, lv_d, la_d in F.FeatureShardLoader(paths, B, "cuda", rank=rank, world=world):
            assert vis_d.is_cuda and vis_d.dtype == torch.bfloat16 and tuple(vis_d.shape) == (B, 16, T)
            got_v.append(vis_d.cpu().clone())
            got_l.append(lv_d.cpu().clone())
        want_v, want_l = [], []
        for k in range(2):
            lo, hi = (rank * W) // world, ((rank + 1) * W) // world
            nfull = (hi - lo) // B * B
            want_v.append(allv[k][lo:lo + nfull])
            want_l.append(alll[k][lo:lo + nfull])
        assert torch.equal(torch.cat(got_v), torch.cat(want_v))
        assert torch.equal(torch.cat(got_l), torch.cat(want_l))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_valpost_matches_reference_val_py_fragment(seed, golden_dir):
    """jmt_b200.valpost on the device vs what the reference's own lines (val.py:313-382, exec()'d by
    tests/golden/make_valpost_golden.py) computed from the same batches: smoothed predictions, labels, accV / accA."""
    import os
    g = np.load(os.path.join(golden_dir, f"valpost_seed{seed}.npz"))
    lengths = g["lengths"]
    acc = jmt_b200.valpost.ValPostprocessor(lengths.tolist())
    for i in range(g["v"].shape[0]):
        acc.update(*(torch.from_numpy(g[k][i]).cuda() for k in ("v", "a", "lv", "la", "fid", "vid")))
    accv, acca, sv, sa = acc.finalize(return_smoothed=True)
    assert abs(accv - float(g["accV"])) < 1e-5 and abs(acca - float(g["accA"])) < 1e-5
    off = np.concatenate([[0], np.cumsum(lengths)])
    order = g["first_seen"].tolist()              # reference output order = first-seen order of the videos
    sv, sa = sv.cpu().numpy(), sa.cpu().numpy()
    got_v = np.concatenate([sv[off[k]:off[k + 1]] for k in order])
    got_a = np.concatenate([sa[off[k]:off[k + 1]] for k in order])
    assert np.abs(got_v - g["vout"]).max() < 2e-6 and np.abs(got_a - g["aout"]).max() < 2e-6
    lab_v = acc.label_v.cpu().numpy()
    assert np.array_equal(np.concatenate([lab_v[off[k]:off[k + 1]] for k in order]).astype(np.float64), g["vtar"])
