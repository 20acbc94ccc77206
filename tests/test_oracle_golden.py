"""Pin the CPU oracle (oracle/jmt_oracle.py) against golden vectors produced by the reference
itself (tests/golden/make_golden.py).  CPU-only; runs in seconds."""
import os

import numpy as np
import pytest
import torch

from oracle import jmt_oracle as O

TT_NAMES = ["tt_transformer_fc_h1_l1", "tt_transformer_fc_h4_l2", "tt_transformer_fc_h8_vin1024",
            "tt_transformer_sa_h2_l1", "tt_none_fc_h2_l1", "tt_fc_fc"]


def _grad_summary(params, names):
    l2, s, head = [], [], []
    for n in names:
        g = params[n].grad.detach().double().reshape(-1)
        l2.append(float(g.norm()))
        s.append(float(g.sum()))
        h = np.zeros(8)
        h[: min(8, g.numel())] = g[:8].numpy()
        head.append(h)
    return np.array(l2), np.array(s), np.stack(head)


@pytest.mark.parametrize("name", TT_NAMES)
def test_two_transformers_forward_backward(name, golden_meta, golden_dir):
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    shapes = O.two_transformers_shapes(m["layers"], m["joint"], m["fmt"], m["vin"], include_dead=False)
    # dead params consume RNG draws in the golden generator: build with them, then drop
    full = O.synth_params(O.two_transformers_shapes(m["layers"], m["joint"], m["fmt"], m["vin"]), m["param_seed"])
    params = {k: full[k].clone().requires_grad_(True) for k, _ in shapes}
    aud, vis = O.synth_features(m["B"], m["T"], [512, m["vin"]], m["feat_seed"])
    lv, la = O.synth_labels(m["B"], m["T"], m["label_seed"])
    aud.requires_grad_(True)
    vis.requires_grad_(True)
    v, a = O.two_transformers_forward(aud, vis, params, m["heads"], m["layers"], m["joint"], m["fmt"])
    assert list(v.shape) == m["out_shape"]
    np.testing.assert_allclose(v.detach().numpy(), g["vout"], rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(a.detach().numpy(), g["aout"], rtol=2e-4, atol=2e-6)
    loss = O.ccc_loss_live(v, lv) + O.ccc_loss_live(a, la)
    np.testing.assert_allclose(float(loss.detach()), float(g["loss"]), rtol=1e-5)
    loss.backward()
    np.testing.assert_allclose(aud.grad.numpy(), g["d_aud"], rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(vis.grad.numpy(), g["d_vis"], rtol=2e-3, atol=1e-7)
    live_names = [n for n in m["grad_names"]]
    for n in live_names:
        assert params[n].grad is not None, n
    l2, s, head = _grad_summary(params, live_names)
    np.testing.assert_allclose(l2, g["grad_l2"], rtol=1e-3, atol=1e-9)
    np.testing.assert_allclose(head, g["grad_head"], rtol=5e-3, atol=1e-7)
    # dead parameters never receive a gradient
    dead = [k for k, _ in shapes if k not in live_names]
    for k in dead:
        assert params[k].grad is None, k


def test_c1_forward(golden_meta, golden_dir):
    m = golden_meta["c1_b8_t300"]
    g = np.load(os.path.join(golden_dir, "c1_b8_t300.npz"))
    params = O.synth_params(O.two_transformers_shapes(1, "TRANSFORMER", "FC", 512), m["param_seed"])
    fcp = O.synth_params([("fc_layer.weight", (512, 768)), ("fc_layer.bias", (512,))], m["fc_seed"])
    vis, aud768 = O.synth_features(m["B"], m["T"], [512, 768], m["feat_seed"])
    with torch.no_grad():
        v, a = O.two_transformers_forward(O.fc_layer_forward(aud768, fcp), vis, params, 1, 1, "TRANSFORMER", "FC")
    assert tuple(v.shape) == (300, 8)          # (T, B): SURVEY Q1
    np.testing.assert_allclose(v.numpy(), g["vout"], rtol=5e-4, atol=3e-5)   # fp32 summation-order noise at T=300
    np.testing.assert_allclose(a.numpy(), g["aout"], rtol=5e-4, atol=3e-5)


def test_inventory_shapes(golden_meta):
    inv = golden_meta["inventory"]
    for joint, fmt in [("TRANSFORMER", "FC"), ("TRANSFORMER", "SELF_ATTEN"), ("NONE", "FC"), ("FC", "FC")]:
        ref = inv[f"Two_transformers/{joint}/{fmt}"]
        mine = O.two_transformers_shapes(1, joint, fmt, 512)
        assert [[k, list(s)] for k, s in mine] == ref["keys"]
        assert sum(int(np.prod(s)) for _, s in mine) == ref["n_params"]
    assert [[k, list(s)] for k, s in O.intra_modal_shapes(1)] == inv["Intra_modal_transformer_fusion"]["keys"]
    tcn = O.synth_params(O.tcn_shapes(1024, [512] * 4, 5), 0)
    assert [[k, list(v.shape)] for k, v in O.tcn_state_dict(tcn).items()] == inv["TemporalConvNet"]["keys"]
    assert sum(v.numel() for v in tcn.values()) == inv["TemporalConvNet"]["n_params"] == 12329472
    assert inv["Two_transformers/TRANSFORMER/FC"]["n_params"] == 52742658


@pytest.mark.parametrize("name", ["intra_512_768_h2", "intra_512_512_h1_l2"])
def test_intra_modal(name, golden_meta, golden_dir):
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    params = {k: v.requires_grad_(True) for k, v in O.synth_params(O.intra_modal_shapes(m["layers"]), m["param_seed"]).items()}
    fa, fb = O.synth_features(m["B"], m["T"], [m["da"], m["db"]], m["feat_seed"])
    fa.requires_grad_(True)
    fb.requires_grad_(True)
    out = O.intra_modal_forward(fa, fb, params, m["heads"], m["layers"])
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=2e-4, atol=2e-5)
    w = torch.linspace(-1, 1, out.numel()).reshape(out.shape)
    (out * w).sum().backward()
    np.testing.assert_allclose(fa.grad.numpy(), g["d_a"], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(fb.grad.numpy(), g["d_b"], rtol=2e-3, atol=1e-5)
    l2, s, head = _grad_summary(params, m["grad_names"])
    np.testing.assert_allclose(l2, g["grad_l2"], rtol=1e-3)


@pytest.mark.parametrize("name", ["tcn_1024_512x4_k5_L7", "tcn_1024_512x4_k5_L40", "tcn_16_8x2_k3_L19"])
def test_tcn(name, golden_meta, golden_dir):
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    params = {k: v.requires_grad_(True) for k, v in
              O.synth_params(O.tcn_shapes(m["cin"], m["chans"], m["k"]), m["param_seed"]).items()}
    gen = torch.Generator().manual_seed(m["x_seed"])
    x = torch.randn(m["N"], m["cin"], m["L"], generator=gen, requires_grad=True)
    out = O.tcn_forward(x, params, len(m["chans"]))
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=3e-4, atol=3e-5)
    w = torch.linspace(-1, 1, out.numel()).reshape(out.shape)
    (out * w).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), g["d_x"], rtol=3e-3, atol=3e-5)
    l2, s, head = _grad_summary(params, m["grad_names"])
    np.testing.assert_allclose(l2, g["grad_l2"], rtol=1e-3)


def test_single_backbone(golden_meta, golden_dir):
    m = golden_meta["single_backbone"]
    g = np.load(os.path.join(golden_dir, "single_backbone.npz"))
    params = O.synth_params(O._regressor_shapes("regressor.", 512, 2), m["param_seed"])
    (x,) = O.synth_features(m["B"], m["T"], [512], m["feat_seed"])
    v, a = O.single_backbone_pretrainer_forward(x, params)
    np.testing.assert_allclose(v.numpy(), g["v"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(a.numpy(), g["a"], rtol=1e-4, atol=1e-6)


def test_ccc_known_answers(golden_meta, golden_dir):
    c = golden_meta["ccc"]
    g = np.load(os.path.join(golden_dir, "ccc.npz"))
    x, y, ym = g["x"], g["y"], g["ym"]
    # SURVEY section 4 known answers (computed with the reference functions)
    assert abs(c["metric"] - 0.6565824151) < 1e-9 and abs(c["loss_masked"] - 0.4134925604) < 1e-7
    assert abs(O.ccc_metric(x, y) - c["metric"]) < 1e-6
    assert abs(O.ccc_numpy(y, x) - c["ccc_numpy"]) < 1e-6
    xt = torch.tensor(x, requires_grad=True)
    l_live = O.ccc_loss_live(xt, torch.tensor(y))
    assert abs(float(l_live) - c["loss_live"]) < 1e-6
    l_live.backward()
    np.testing.assert_allclose(xt.grad.numpy(), g["g_live"], rtol=1e-3, atol=1e-8)
    xt2 = torch.tensor(x, requires_grad=True)
    l_m = O.ccc_loss_masked(xt2, torch.tensor(ym))
    assert abs(float(l_m) - c["loss_masked"]) < 1e-6
    l_m.backward()
    np.testing.assert_allclose(xt2.grad.numpy(), g["g_masked"], rtol=1e-3, atol=1e-9)
    assert float(O.ccc_loss_masked(torch.tensor([0.1, 0.2, 0.3]), torch.tensor([-5.0, 0.5, -5.0]))) == 0.0
    cv, ca, cm = O.cccva(np.stack([y, ym], 1), np.stack([x, x * 0.5], 1))
    np.testing.assert_allclose([cv, ca, cm], c["cccva"], rtol=1e-5)
    # closed forms from the six fp64 sums (what the CUDA reduction produces)
    s = O.six_sums(x, y)
    assert abs(O.ccc_from_sums(s, "metric") - c["metric"]) < 1e-6
    assert abs(O.ccc_from_sums(s, "loss_live") - c["loss_live"]) < 1e-6
    assert abs(O.ccc_from_sums(O.six_sums(y, x), "ccc_numpy") - c["ccc_numpy"]) < 1e-6
    sm = O.six_sums(x, ym, ignore=-5.0)
    assert sm[0] == 900
    assert abs(O.ccc_from_sums(sm, "loss_masked", n_all=1000) - c["loss_masked"]) < 1e-6
    # large-N: regenerate the generator's big arrays (same RandomState(0) draw order)
    rs = np.random.RandomState(0)
    rs.randn(1000), rs.randn(1000)
    big_x = (rs.randn(200000) * 0.3 + 0.2).astype(np.float32)
    big_y = np.clip(big_x * 0.7 + rs.randn(200000).astype(np.float32) * 0.2, -1, 1).astype(np.float32)
    assert abs(O.ccc_from_sums(O.six_sums(big_x, big_y), "metric") - c["metric_big"]) < 1e-9
    with pytest.raises(ValueError):
        O.ccc_metric(np.array([1.0]), np.array([1.0]))


def test_label_mask_and_padseq(golden_meta, golden_dir):
    y = torch.tensor([0.3, -5.0, -4.9999995, -5.0000005, 1.0, -5.0])
    assert O.label_mask(y).tolist() == [True, False, True, True, True, False]
    m = golden_meta["padseq"]
    g = np.load(os.path.join(golden_dir, "padseq.npz"))
    gen = torch.Generator().manual_seed(m["seed"])
    specs = []
    for w in m["widths"]:
        torch.randn(16, 3, 2, 4, 4, generator=gen)
        specs.append(torch.randn(16, 1, 64, w, generator=gen) + 3.0)
        torch.randn(16, generator=gen), torch.randn(16, generator=gen)
    out = O.pad_spectrograms(specs)
    assert np.array_equal((out == 0).numpy(), g["zero_mask"])      # zero-fill positions bit-exact
    np.testing.assert_allclose(out.double().sum().numpy(), g["audio_sum"], rtol=1e-12)
