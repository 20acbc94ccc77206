"""Parity at the benchmarked configurations with DEFAULT-INITIALISED weights (GPU).

The reference modules were constructed under torch.manual_seed(seed) by tests/golden/make_golden.py; the drop-in modules
constructed under the same seed come out bit-identical (checked there with torch.equal and here through `param_sums`), so
no weights travel.  Cases: Two_transformers variants incl. the batch-dimension attention of NONE over several
128-row tiles (B = 300) and beyond the fused kernel's key limit (B = 400, B = 600); the TCN at L = 300 (all dilations cross
tile borders of the flat layout); the whole benchmarked pipeline TCN -> FcLayer -> Two_transformers -> live CCC loss at
B = 4, T = 300, forward AND backward; intra-modal fusion at T = 1024.

Every precision is compared with the REFERENCE's outputs and gradients.  Gradients are checked for the live CCC loss and for
a well-conditioned random cotangent (the CCC cotangent of a small batch is nearly constant: cancelling terms amplify operand
rounding, which says nothing about the operators).  Bounds are <= 1.5x the values measured on B200 (listed next to them);
'bf16x3' is the tensor-core mode that must meet the north-star gate (predictions 1e-3, CCC 1e-4)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import jmt_b200  # noqa: E402
from oracle import jmt_oracle as O  # noqa: E402

DEV = "cuda"
PRECISIONS = ["bf16x3", "fp32", "bf16"]
MEASURED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_measured.jsonl")


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def _rl2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def _record(name, precision, vals):
    try:
        os.makedirs(os.path.dirname(MEASURED), exist_ok=True)
        with open(MEASURED, "a") as f:
            f.write(json.dumps({"case": name, "precision": precision, **{k: float(v) for k, v in vals.items()}}) + "\n")
    except OSError:
        pass


def _check(name, precision, vals, bounds):
    """vals: measured deltas; bounds: {metric: (bf16x3, fp32, bf16)}.  Everything is recorded, then asserted."""
    _record(name, precision, vals)
    i = PRECISIONS.index(precision)
    bad = {k: (v, bounds[k][i]) for k, v in vals.items() if k in bounds and not v < bounds[k][i]}
    assert not bad, (name, precision, bad)


def _check_sums(module, sums, prefix=""):
    sd = module.state_dict()
    for k, s in sums.items():
        if not k.startswith(prefix):
            continue
        kk = k[len(prefix):]
        assert abs(float(sd[kk].double().sum()) - s) <= 1e-6 * max(1.0, abs(s)), f"seeded init differs from the reference: {k}"


def _grad_l2(module, names):
    """per-parameter gradient L2 norms, first 8 entries and element counts"""
    named = dict(module.named_parameters())
    out, head, numel = [], [], []
    for n in names:
        g = named[n].grad
        assert g is not None, n
        g = g.detach().double().reshape(-1).cpu()
        out.append(float(g.norm()))
        numel.append(g.numel())
        h = np.zeros(8)
        h[: min(8, g.numel())] = g[:8].numpy()
        head.append(h)
    return np.array(out), (np.stack(head), np.array(numel))


def _grad_metrics(l2, head_numel, want_l2, want_head):
    """worst relative deviation of the per-parameter gradient norms, and of the first 8 entries of every gradient relative
    to the larger of their own magnitude and the tensor's rms (a head of ~zeros must not blow the ratio up)."""
    head, numel = head_numel
    rel_l2 = float((np.abs(l2 - want_l2) / (want_l2 + 1e-30)).max())
    worst = 0.0
    for i in range(len(want_l2)):
        sc = max(np.abs(want_head[i]).max(), want_l2[i] / np.sqrt(numel[i]), 1e-30)
        worst = max(worst, float(np.abs(head[i] - want_head[i]).max() / sc))
    return rel_l2, worst


# Bounds per precision (bf16x3, fp32, bf16), each <= ~1.5-2x the worst value measured on B200 in round 2 (in the comment);
# pred_rel / ccc_delta of the gate precisions stay at the north-star numbers (1e-3 / 1e-4) they have to meet.
# Gradient deltas of the gate precisions are dominated by single ReLU / LeakyReLU sign decisions on pre-activations that are ~0
# (an O(1) change of a few entries), input gradients of bf16 by the cancellation in the F.normalize backward
# (dx = r (dy - y (y.dy)) on bf16-rounded dy).
TT_BOUNDS = {
    "pred_rel": (1e-3, 1e-3, 3.5e-2),        # 1.9e-5   2.2e-6   2.3e-2 (SELF_ATTEN h2; 1.6e-3 .. 9.6e-3 elsewhere)
    "pred_l2": (5e-5, 5e-6, 2.1e-2),         # 1.6e-5   1.5e-6   1.4e-2
    "loss": (1e-5, 1e-5, 2.5e-4),            # 4.8e-7   1.2e-7   1.5e-4
    "ccc_delta": (1e-4, 1e-4, 1e-4),         # 3.5e-7   7.9e-9   6.1e-5   <- plain bf16 meets the 1e-4 CCC gate on default init
    "c_din_l2": (2e-3, 1e-3, 1.5e-1),        # 1.2e-3   3.8e-4   9.8e-2
    "c_grad_l2": (2e-4, 2e-4, 3.2e-2),       # 8.5e-5   6.8e-5   2.1e-2
    "live_din_l2": (1e-3, 1e-3, 2e-1),       # 2.5e-4   2.5e-4   1.3e-1
    "live_grad_l2": (3e-3, 3e-3, 1e-1),      # 1.1e-3   9.2e-4   6.1e-2
}
TT_DEFAULT = ["ttd_transformer_fc_h1", "ttd_transformer_sa_h2", "ttd_fc_fc", "ttd_none_fc_h2", "ttd_none_fc_b300_t2",
              "ttd_none_fc_b400_t2", "ttd_none_fc_b600_t1_h2"]


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", TT_DEFAULT)
def test_two_transformers_default_init(name, precision, golden_meta, golden_dir):
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    torch.manual_seed(m["init_seed"])
    model = jmt_b200.Two_transformers(0.0, 0.0, m["heads"], m["layers"], m["joint"], m["fmt"], m["vin"], precision=precision)
    _check_sums(model, m["param_sums"])
    model = model.to(DEV).eval()
    aud, vis = O.synth_features(m["B"], m["T"], [512, m["vin"]], m["feat_seed"])
    lv, la = O.synth_labels(m["B"], m["T"], m["label_seed"])
    st = m["sample_stride"]
    vals = {}
    aud_d, vis_d = aud.to(DEV).requires_grad_(True), vis.to(DEV).requires_grad_(True)
    v, a = model(aud_d, vis_d)
    assert list(v.shape) == m["out_shape"]
    vals["pred_rel"] = max(_rel(v.detach().cpu(), g["vout"]), _rel(a.detach().cpu(), g["aout"]))
    vals["pred_l2"] = max(_rl2(v.detach().cpu(), g["vout"]), _rl2(a.detach().cpu(), g["aout"]))
    crit = jmt_b200.CCCLoss(digitize_num=1)
    n = v.shape[0] * v.shape[1]
    loss = crit(v.view(-1, n), lv.to(DEV).view(-1, n)) + crit(a.view(-1, n), la.to(DEV).view(-1, n))
    vals["loss"] = abs(loss.item() - float(g["loss"]))
    # the metric on our predictions vs on the reference's predictions, same labels (flattened like train.py:303-307)
    lab = lv.numpy().reshape(-1).astype(np.float64)
    vals["ccc_delta"] = abs(O.ccc_metric(g["vout"].reshape(-1).astype(np.float64), lab) -
                            jmt_b200.cccmetric.ccc(v.detach().reshape(-1), lv.to(DEV).reshape(-1)))
    # (1) well-conditioned random cotangent: every precision against the REFERENCE gradients
    gen_v = torch.Generator().manual_seed(m["cot_seeds"][0])
    gen_a = torch.Generator().manual_seed(m["cot_seeds"][1])
    cv, ca = torch.randn(v.shape, generator=gen_v).to(DEV), torch.randn(a.shape, generator=gen_a).to(DEV)
    torch.autograd.backward([v, a], [cv, ca], retain_graph=False)
    vals["c_din_l2"] = max(_rl2(aud_d.grad.cpu()[:, :, ::st], g["c_d_aud"]), _rl2(vis_d.grad.cpu()[:, :, ::st], g["c_d_vis"]))
    l2, head = _grad_l2(model, m["grad_names"])
    vals["c_grad_l2"], vals["c_grad_head"] = _grad_metrics(l2, head, g["c_grad_l2"], g["c_grad_head"])
    # (2) the live CCC loss (a second forward: the tape is consumed by backward)
    model.zero_grad(set_to_none=True)
    aud_d2, vis_d2 = aud.to(DEV).requires_grad_(True), vis.to(DEV).requires_grad_(True)
    v2, a2 = model(aud_d2, vis_d2)
    loss2 = crit(v2.view(-1, n), lv.to(DEV).view(-1, n)) + crit(a2.view(-1, n), la.to(DEV).view(-1, n))
    loss2.backward()
    vals["live_din_l2"] = max(_rl2(aud_d2.grad.cpu()[:, :, ::st], g["d_aud"]), _rl2(vis_d2.grad.cpu()[:, :, ::st], g["d_vis"]))
    l2, head = _grad_l2(model, m["grad_names"])
    vals["live_grad_l2"], vals["live_grad_head"] = _grad_metrics(l2, head, g["grad_l2"], g["grad_head"])
    for nme, p in model.named_parameters():
        if "final_encoder" in nme or "gated_attention" in nme:
            assert p.grad is None, nme
    _check(name, precision, vals, TT_BOUNDS)


@pytest.mark.parametrize("name", ["ttd_none_fc_b600_t1_h2", "ttd_none_fc_b400_t2"])
def test_long_key_attention_forward_only_chunked(name, golden_meta, golden_dir):
    """Evaluation (no_grad) of the NONE variant whose encoders attend across the batch (L = S = B) beyond the fused
    kernel's key limit: the engine runs the fused kernel over key chunks and merges with the running log-sum-exp
    (jmt_attn_merge) -- no (L, S) score / probability tensor is materialised.  Against the REFERENCE's predictions, and against
    the composed path (GEMM -> fp32 scores -> softmax -> GEMM) the training graph uses at these lengths."""
    from jmt_b200 import engine as E
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    torch.manual_seed(m["init_seed"])
    model = jmt_b200.Two_transformers(0.0, 0.0, m["heads"], m["layers"], m["joint"], m["fmt"], m["vin"], precision="bf16").to(DEV).eval()
    aud, vis = (t.to(DEV) for t in O.synth_features(m["B"], m["T"], [512, m["vin"]], m["feat_seed"]))
    E.LONG_S_CHUNK_MIN_BYTES = 0           # (the engine keeps the composed path while the fp32 scores are small)
    with torch.no_grad():
        model(aud, vis)                   # warm the bf16 operand cache so that the launch counts below are comparable
    n0 = jmt_b200.launch_count()
    with torch.no_grad():
        v, a = model(aud, vis)
    n_chunked = jmt_b200.launch_count() - n0
    E.LONG_S_CHUNKED = False
    try:
        n0 = jmt_b200.launch_count()
        with torch.no_grad():
            v2, a2 = model(aud, vis)
        n_composed = jmt_b200.launch_count() - n0
    finally:
        E.LONG_S_CHUNKED = True
        E.LONG_S_CHUNK_MIN_BYTES = 2 << 30
    vals = {"pred_rel": max(_rel(v.cpu(), g["vout"]), _rel(a.cpu(), g["aout"])),
            "pred_l2": max(_rl2(v.cpu(), g["vout"]), _rl2(a.cpu(), g["aout"])),
            "vs_composed": max(_rel(v.cpu(), v2.cpu()), _rel(a.cpu(), a2.cpu()))}
    _record(name + "/chunked_eval", "bf16", vals)
    if m["B"] > 512:                      # S = B = 600 > 512 keys: beyond one fused tile -> chunks + merge launches
        assert n_chunked != n_composed, "the chunked path was not taken"
    else:                                 # S = 400 at dh = 512 still fits one fused tile (fewer operand slots)
        assert n_chunked == n_composed
    assert vals["pred_rel"] < TT_BOUNDS["pred_rel"][2] and vals["pred_l2"] < TT_BOUNDS["pred_l2"][2], vals
    assert vals["vs_composed"] < 2e-2, vals


TCN_BOUNDS = {
    "out_rel": (1e-3, 1e-3, 1e-2),           # 8.4e-6   1.1e-6   6.8e-3
    "out_l2": (3e-5, 3e-6, 6e-3),            # 9.4e-6   6.1e-7   3.9e-3
    "dx_l2": (3e-3, 2e-3, 6.5e-2),           # 1.7e-3   1.0e-3   4.3e-2
    "grad_l2": (1e-3, 6e-4, 1e-2),           # 4.4e-4   2.6e-4   6.5e-3
}


@pytest.mark.parametrize("precision", PRECISIONS)
def test_tcn_L300_default_init(precision, golden_meta, golden_dir):
    name = "tcnd_1024_512x4_k5_L300"
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    torch.manual_seed(m["init_seed"])
    model = jmt_b200.TemporalConvNet(1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1, precision=precision)
    _check_sums(model, m["param_sums"])
    model = model.to(DEV).eval()
    x = torch.randn(m["N"], 1024, m["L"], generator=torch.Generator().manual_seed(m["x_seed"]))
    xd = x.to(DEV).requires_grad_(True)
    out = model(xd)
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(m["cot_seed"])).to(DEV)
    (out * cot).sum().backward()
    vals = {"out_rel": _rel(out.detach().cpu()[:, ::m["out_stride"]], g["out"]),
            "out_l2": _rl2(out.detach().cpu()[:, ::m["out_stride"]], g["out"]),
            "out_norm": abs(float(out.detach().double().norm()) - float(g["out_l2"])) / float(g["out_l2"]),
            "dx_l2": _rl2(xd.grad.cpu()[:, ::m["dx_stride"]], g["d_x"])}
    l2, head = _grad_l2(model, m["grad_names"])
    vals["grad_l2"], vals["grad_head"] = _grad_metrics(l2, head, g["grad_l2"], g["grad_head"])
    _check(name, precision, vals, TCN_BOUNDS)


PIPE_BOUNDS = {
    "pred_rel": (1e-3, 1e-3, 1.8e-2),        # 2.4e-5   2.6e-6   1.2e-2
    "pred_l2": (3e-5, 5e-6, 8.5e-3),         # 8.5e-6   1.0e-6   5.5e-3
    "loss": (1e-5, 1e-5, 1e-5),              # 0        0        2.0e-6
    "ccc_delta": (1e-4, 1e-4, 1e-5),         # 5.8e-9   3.4e-9   1.8e-6
    "c_din_l2": (2e-3, 5e-4, 1.2e-1),        # 1.1e-3   2.0e-4   7.6e-2
    "c_grad_l2": (5e-4, 1e-4, 2.6e-2),       # 2.4e-4   4.3e-5   1.7e-2
    "live_din_l2": (2e-3, 2e-3, 1.5e-1),     # 8.7e-4   7.3e-4   1.0e-1
    "live_grad_l2": (1e-2, 3e-3, 5.5e-1),    # 5.9e-3   1.2e-3   3.6e-1  (CCC cotangent: nearly constant, cancelling terms)
}


@pytest.mark.parametrize("mode", ["gemm_colsum", "no_ext"])
def test_pipeline_b4_t300_backward_epilogue_modes(mode, golden_meta, golden_dir, monkeypatch):
    """The non-default backward-epilogue modes of the bf16 engine (JMT_EPI_EXT = 2: the GEMMs that write a gradient also emit
    its column sums = the bias gradients; 0: no epilogue extensions at all) against the same reference gradients and bounds."""
    from jmt_b200 import engine as E
    monkeypatch.setattr(E, "EPI_EXT", mode != "no_ext")
    monkeypatch.setattr(E, "EPI_COLSUM", mode == "gemm_colsum")
    test_pipeline_b4_t300_default_init("bf16", golden_meta, golden_dir)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_pipeline_b4_t300_default_init(precision, golden_meta, golden_dir):
    """The benchmarked pipeline (BASELINE.json configs[1]) at B = 4, T = 300 through jmt_b200.JMTPipeline -- TCN on the flat
    padded layout (3 tiles per sequence), unpad, l2norm, FcLayer, fusion, live loss -- forward and backward vs the reference."""
    name = "piped_b4_t300"
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B, T = m["B"], m["T"]
    torch.manual_seed(m["init_seed"])
    model = jmt_b200.JMTPipeline(jmt_b200.Two_transformers(0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512, precision=precision),
                                 jmt_b200.FcLayer(768, 512, precision=precision),
                                 jmt_b200.TemporalConvNet(1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1, precision=precision))
    for pre, mod in (("fusion.", model.fusion), ("fc_audio.", model.fc_audio), ("tcn.", model.tcn)):
        _check_sums(mod, m["param_sums"], pre)
    model = model.to(DEV).eval()
    gen = torch.Generator().manual_seed(m["data_seed"])
    vis = torch.randn(B, 1024, T, generator=gen)
    aud = torch.randn(B, T, 768, generator=gen)
    lv, la = O.synth_labels(B, T, m["label_seed"])
    crit = jmt_b200.CCCLoss(digitize_num=1)
    n = B * T
    vals = {}
    aud_d, vis_d = aud.to(DEV).requires_grad_(True), vis.to(DEV).requires_grad_(True)
    v, a = model(aud_d, vis_d)
    assert list(v.shape) == m["out_shape"]
    vals["pred_rel"] = max(_rel(v.detach().cpu(), g["vout"]), _rel(a.detach().cpu(), g["aout"]))
    vals["pred_l2"] = max(_rl2(v.detach().cpu(), g["vout"]), _rl2(a.detach().cpu(), g["aout"]))
    loss = crit(v.view(-1, n), lv.to(DEV).view(-1, n)) + crit(a.view(-1, n), la.to(DEV).view(-1, n))
    vals["loss"] = abs(loss.item() - float(g["loss"]))
    lab = lv.numpy().reshape(-1).astype(np.float64)
    vals["ccc_delta"] = abs(O.ccc_metric(g["vout"].reshape(-1).astype(np.float64), lab) -
                            jmt_b200.cccmetric.ccc(v.detach().reshape(-1), lv.to(DEV).reshape(-1)))
    cv = torch.randn(v.shape, generator=torch.Generator().manual_seed(m["cot_seeds"][0])).to(DEV)
    ca = torch.randn(a.shape, generator=torch.Generator().manual_seed(m["cot_seeds"][1])).to(DEV)
    torch.autograd.backward([v, a], [cv, ca])
    vals["c_din_l2"] = max(_rl2(aud_d.grad.cpu()[:, :, ::m["aud_stride"]], g["c_d_aud"]),
                           _rl2(vis_d.grad.cpu()[:, ::m["vis_stride"]], g["c_d_vis"]))
    l2, head = _grad_l2(model, m["grad_names"])
    vals["c_grad_l2"], vals["c_grad_head"] = _grad_metrics(l2, head, g["c_grad_l2"], g["c_grad_head"])
    model.zero_grad(set_to_none=True)
    aud_d2, vis_d2 = aud.to(DEV).requires_grad_(True), vis.to(DEV).requires_grad_(True)
    v2, a2 = model(aud_d2, vis_d2)
    (crit(v2.view(-1, n), lv.to(DEV).view(-1, n)) + crit(a2.view(-1, n), la.to(DEV).view(-1, n))).backward()
    vals["live_din_l2"] = max(_rl2(aud_d2.grad.cpu()[:, :, ::m["aud_stride"]], g["d_aud"]),
                              _rl2(vis_d2.grad.cpu()[:, ::m["vis_stride"]], g["d_vis"]))
    l2, head = _grad_l2(model, m["grad_names"])
    vals["live_grad_l2"], vals["live_grad_head"] = _grad_metrics(l2, head, g["grad_l2"], g["grad_head"])
    _check(name, precision, vals, PIPE_BOUNDS)


INTRA_BOUNDS = {
    "out_rel": (1e-3, 1e-3, 7.5e-3),         # 8.1e-6   8.7e-7   4.8e-3
    "out_l2": (3e-5, 3e-6, 8e-3),            # 8.7e-6   8.1e-7   5.4e-3
    "din_l2": (1e-4, 5e-6, 2e-2),            # 3.0e-5   1.0e-6   1.3e-2
    "grad_l2": (3e-5, 5e-7, 2.2e-3),         # 8.3e-6   7.9e-8   1.4e-3
}


@pytest.mark.parametrize("precision", PRECISIONS)
def test_intra_modal_t1024_default_init(precision, golden_meta, golden_dir):
    """BASELINE.json configs[3]: Intra_modal_transformer_fusion at T = 1024 (length-2 'sequences')."""
    name = "intrad_b2_t1024"
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    torch.manual_seed(m["init_seed"])
    model = jmt_b200.Intra_modal_transformer_fusion(512, m["heads"], 512, m["layers"], precision=precision)
    _check_sums(model, m["param_sums"])
    model = model.to(DEV).eval()
    fa, fb = O.synth_features(m["B"], m["T"], [512, 768], m["feat_seed"])
    fa_d, fb_d = fa.to(DEV).requires_grad_(True), fb.to(DEV).requires_grad_(True)
    out = model(fa_d, fb_d)
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(m["cot_seed"])).to(DEV)
    (out * cot).sum().backward()
    so, sd = m["out_stride"], m["d_stride"]
    vals = {"out_rel": _rel(out.detach().cpu()[:, :, ::so], g["out"]), "out_l2": _rl2(out.detach().cpu()[:, :, ::so], g["out"]),
            "out_norm": abs(float(out.detach().double().norm()) - float(g["out_l2"])) / float(g["out_l2"]),
            "din_l2": max(_rl2(fa_d.grad.cpu()[:, :, ::sd], g["d_a"]), _rl2(fb_d.grad.cpu()[:, :, ::sd], g["d_b"]))}
    l2, head = _grad_l2(model, m["grad_names"])
    vals["grad_l2"], vals["grad_head"] = _grad_metrics(l2, head, g["grad_l2"], g["grad_head"])
    _check(name, precision, vals, INTRA_BOUNDS)
