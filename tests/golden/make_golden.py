"""Generate the golden fixtures in this directory from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference modules (with sys.modules stubs for comet_ml / matplotlib, which
models/mm_transformers.py:2-6 imports at module top and which are not installed), loads
deterministic synthetic parameters (oracle.synth_params) with load_state_dict(strict=True),
runs forward (+ backward through the live CCC loss) on seeded synthetic inputs under the
installed torch in fp32 and stores outputs, input-gradients and per-parameter gradient
checksums as small .npz files.  The GPU box has no /root/reference: tests only read the
.npz/.json written here.
"""
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("JMT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import jmt_oracle as O  # noqa: E402


def _install_stubs():
    for name in ["comet_ml", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["comet_ml"].Experiment = object
    sys.modules["mpl_toolkits.axes_grid1"].ImageGrid = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def _import_reference():
    _install_stubs()
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    from models.two_transformers import Two_transformers, SingleBackbonePretrainer
    from models.intra_modal_transformer_fusion import Intra_modal_transformer_fusion
    from models.temporal_convolutional_model import TemporalConvNet
    from models.fc_layer import FcLayer
    from losses.loss import CCCLoss as LiveCCCLoss
    from losses.CCCLoss import CCCLoss as MaskedCCCLoss
    from EvaluationMetrics import cccmetric
    import padSequence
    return dict(Two_transformers=Two_transformers, SingleBackbonePretrainer=SingleBackbonePretrainer,
                Intra=Intra_modal_transformer_fusion, TCN=TemporalConvNet, FcLayer=FcLayer,
                LiveCCCLoss=LiveCCCLoss, MaskedCCCLoss=MaskedCCCLoss, cccmetric=cccmetric,
                padSequence=padSequence)


def _live_loss(R):
    # losses/loss.py:16 calls .cuda() in __init__; bins are unused when digitize_num == 1
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        return R["LiveCCCLoss"](digitize_num=1)
    finally:
        torch.Tensor.cuda = orig


def _grad_summary(module):
    names, l2, s, head = [], [], [], []
    for n, p in module.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().double().reshape(-1)
        names.append(n)
        l2.append(float(g.norm()))
        s.append(float(g.sum()))
        h = np.zeros(8)
        h[: min(8, g.numel())] = g[:8].numpy()
        head.append(h)
    return names, np.array(l2), np.array(s), np.stack(head)


TT_CASES = [
    # name, B, T, heads, layers, joint, out_format, vision_in_ft
    ("tt_transformer_fc_h1_l1", 2, 24, 1, 1, "TRANSFORMER", "FC", 512),
    ("tt_transformer_fc_h4_l2", 3, 17, 4, 2, "TRANSFORMER", "FC", 512),
    ("tt_transformer_fc_h8_vin1024", 2, 13, 8, 1, "TRANSFORMER", "FC", 1024),
    ("tt_transformer_sa_h2_l1", 2, 9, 2, 1, "TRANSFORMER", "SELF_ATTEN", 512),
    ("tt_none_fc_h2_l1", 5, 7, 2, 1, "NONE", "FC", 512),
    ("tt_fc_fc", 3, 10, 1, 1, "FC", "FC", 512),
]


def gen_two_transformers(R, meta):
    live = _live_loss(R)
    for (name, B, T, h, L, joint, fmt, vin) in TT_CASES:
        shapes = O.two_transformers_shapes(L, joint, fmt, vin, include_dead=True)
        params = O.synth_params(shapes, seed=100 + len(name))
        model = R["Two_transformers"](0.0, 0.0, h, L, joint, fmt, vin)
        ref_shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
        assert ref_shapes == [(k, tuple(s)) for k, s in shapes], name
        model.load_state_dict(params, strict=True)
        model.eval()
        aud, vis = O.synth_features(B, T, [512, vin], seed=7)
        lv, la = O.synth_labels(B, T, seed=11)
        aud.requires_grad_(True)
        vis.requires_grad_(True)
        v, a = model(aud, vis)
        # train.py:303-311: flatten preds and labels independently to (1, B*T) and sum the losses
        n = v.shape[0] * v.shape[1]
        loss = live(v.reshape(-1, n), lv.reshape(-1, n)) + live(a.reshape(-1, n), la.reshape(-1, n))
        loss.backward()
        names, l2, s, head = _grad_summary(model)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            vout=v.detach().numpy(), aout=a.detach().numpy(), loss=loss.detach().numpy(),
                            d_aud=aud.grad.numpy(), d_vis=vis.grad.numpy(),
                            grad_l2=l2, grad_sum=s, grad_head=head)
        meta[name] = dict(B=B, T=T, heads=h, layers=L, joint=joint, fmt=fmt, vin=vin,
                          param_seed=100 + len(name), feat_seed=7, label_seed=11,
                          grad_names=names, out_shape=list(v.shape))
        print(name, tuple(v.shape), float(loss))


def gen_c1(R, meta):
    """BASELINE.json configs[0]: FcLayer(768,512) on WavLM audio -> Two_transformers(TRANSFORMER, FC,
    h=1, L=1), B=8, T=300, forward only."""
    B, T = 8, 300
    shapes = O.two_transformers_shapes(1, "TRANSFORMER", "FC", 512, include_dead=True)
    params = O.synth_params(shapes, seed=1)
    fc_params = O.synth_params([("fc_layer.weight", (512, 768)), ("fc_layer.bias", (512,))], seed=2)
    model = R["Two_transformers"](0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512)
    model.load_state_dict(params, strict=True)
    model.eval()
    fc = R["FcLayer"](768, 512)
    fc.load_state_dict(fc_params, strict=True)
    vis, aud768 = O.synth_features(B, T, [512, 768], seed=3)
    with torch.no_grad():
        v, a = model(fc(aud768), vis)
    np.savez_compressed(os.path.join(HERE, "c1_b8_t300.npz"), vout=v.numpy(), aout=a.numpy())
    meta["c1_b8_t300"] = dict(B=B, T=T, param_seed=1, fc_seed=2, feat_seed=3, out_shape=list(v.shape))
    print("c1", tuple(v.shape))


def gen_default_inventory(R, meta):
    """Key/shape inventory and parameter counts of default-constructed reference modules."""
    inv = {}
    for joint, fmt in [("TRANSFORMER", "FC"), ("TRANSFORMER", "SELF_ATTEN"), ("NONE", "FC"), ("FC", "FC")]:
        m = R["Two_transformers"](0.0, 0.0, 1, 1, joint, fmt, 512)
        inv[f"Two_transformers/{joint}/{fmt}"] = dict(
            keys=[[k, list(v.shape)] for k, v in m.state_dict().items()],
            n_params=sum(p.numel() for p in m.parameters()))
    m = R["Intra"](512, 1, 512, 1)
    inv["Intra_modal_transformer_fusion"] = dict(keys=[[k, list(v.shape)] for k, v in m.state_dict().items()],
                                                 n_params=sum(p.numel() for p in m.parameters()))
    m = R["TCN"](1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1)
    inv["TemporalConvNet"] = dict(keys=[[k, list(v.shape)] for k, v in m.state_dict().items()],
                                  n_params=sum(p.numel() for p in m.parameters()))
    m = R["SingleBackbonePretrainer"](0.0, 0.0)
    inv["SingleBackbonePretrainer"] = dict(keys=[[k, list(v.shape)] for k, v in m.state_dict().items()],
                                           n_params=sum(p.numel() for p in m.parameters()))
    m = R["FcLayer"](768, 512)
    inv["FcLayer"] = dict(keys=[[k, list(v.shape)] for k, v in m.state_dict().items()],
                          n_params=sum(p.numel() for p in m.parameters()))
    # seeded default construction: checksum per tensor, to pin the constructors' init order
    torch.manual_seed(0)
    m = R["Two_transformers"](0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512)
    inv["seed0_init_sums"] = {k: float(v.double().sum()) for k, v in m.state_dict().items()
                              if "final_encoder" not in k}
    meta["inventory"] = inv


def gen_intra(R, meta):
    for name, da, db, h, L, B, T in [("intra_512_768_h2", 512, 768, 2, 1, 2, 11),
                                     ("intra_512_512_h1_l2", 512, 512, 1, 2, 3, 5)]:
        shapes = O.intra_modal_shapes(L)
        params = O.synth_params(shapes, seed=41)
        m = R["Intra"](512, h, 512, L)
        assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in shapes]
        m.load_state_dict(params, strict=True)
        m.eval()
        fa, fb = O.synth_features(B, T, [da, db], seed=42)
        fa.requires_grad_(True)
        fb.requires_grad_(True)
        out = m(fa, fb)
        w = torch.linspace(-1, 1, out.numel()).reshape(out.shape)
        (out * w).sum().backward()
        names, l2, s, head = _grad_summary(m)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), out=out.detach().numpy(),
                            d_a=fa.grad.numpy(), d_b=fb.grad.numpy(), grad_l2=l2, grad_sum=s, grad_head=head)
        meta[name] = dict(B=B, T=T, da=da, db=db, heads=h, layers=L, param_seed=41, feat_seed=42, grad_names=names)
        print(name, tuple(out.shape))


def gen_tcn(R, meta):
    for name, cin, chans, k, N, L in [("tcn_1024_512x4_k5_L7", 1024, [512] * 4, 5, 2, 7),
                                      ("tcn_1024_512x4_k5_L40", 1024, [512] * 4, 5, 2, 40),
                                      ("tcn_16_8x2_k3_L19", 16, [8, 8], 3, 3, 19)]:
        shapes = O.tcn_shapes(cin, chans, k)
        params = O.synth_params(shapes, seed=51)
        m = R["TCN"](cin, chans, kernel_size=k, attention=0, dropout=0.1)
        full = O.tcn_state_dict(params)
        assert [(kk, tuple(v.shape)) for kk, v in m.state_dict().items()] == \
            [(kk, tuple(v.shape)) for kk, v in full.items()]
        assert [n for n, _ in m.named_parameters()] == [kk for kk, _ in shapes]
        m.load_state_dict(full, strict=True)
        m.eval()
        gen = torch.Generator().manual_seed(52)
        x = torch.randn(N, cin, L, generator=gen, requires_grad=True)
        out = m(x)
        w = torch.linspace(-1, 1, out.numel()).reshape(out.shape)
        (out * w).sum().backward()
        names, l2, s, head = _grad_summary(m)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), out=out.detach().numpy(), d_x=x.grad.numpy(),
                            grad_l2=l2, grad_sum=s, grad_head=head)
        meta[name] = dict(cin=cin, chans=chans, k=k, N=N, L=L, param_seed=51, x_seed=52, grad_names=names)
        print(name, tuple(out.shape))


def gen_misc(R, meta):
    # SingleBackbonePretrainer + FcLayer
    shapes = O._regressor_shapes("regressor.", 512, 2)
    params = O.synth_params(shapes, seed=61)
    m = R["SingleBackbonePretrainer"](0.0, 0.0)
    m.load_state_dict(params, strict=True)
    m.eval()
    (x,) = O.synth_features(3, 6, [512], seed=62)
    with torch.no_grad():
        v, a = m(x)
    np.savez_compressed(os.path.join(HERE, "single_backbone.npz"), v=v.numpy(), a=a.numpy())
    meta["single_backbone"] = dict(B=3, T=6, param_seed=61, feat_seed=62)

    # CCC known answers + extra cases
    rs = np.random.RandomState(0)
    x = rs.randn(1000).astype(np.float32)
    y = (0.5 * x + 0.5 * rs.randn(1000)).astype(np.float32)
    cm = R["cccmetric"]
    live = _live_loss(R)
    masked = R["MaskedCCCLoss"]()
    ym = y.copy()
    ym[::10] = -5.0
    xt = torch.tensor(x, requires_grad=True)
    l_live = live(xt[None], torch.tensor(y)[None])
    l_live.backward()
    g_live = xt.grad.clone()
    xt2 = torch.tensor(x, requires_grad=True)
    l_mask = masked(xt2, torch.tensor(ym))
    l_mask.backward()
    big_x = (rs.randn(200000) * 0.3 + 0.2).astype(np.float32)
    big_y = np.clip(big_x * 0.7 + rs.randn(200000).astype(np.float32) * 0.2, -1, 1).astype(np.float32)
    ccc = dict(metric=float(cm.ccc(x, y)), ccc_numpy=float(cm.ccc_numpy(y, x)),
               loss_live=float(l_live), loss_masked=float(l_mask),
               metric_big=float(cm.ccc(big_x.astype(np.float64), big_y.astype(np.float64))),
               loss_masked_one_valid=float(masked(torch.tensor([0.1, 0.2, 0.3]), torch.tensor([-5.0, 0.5, -5.0]))),
               cccva=[float(t) for t in cm.cccva(np.stack([y, ym], 1), np.stack([x, x * 0.5], 1))])
    np.savez_compressed(os.path.join(HERE, "ccc.npz"), x=x, y=y, ym=ym, g_live=g_live.numpy(),
                        g_masked=xt2.grad.numpy())   # big_x/big_y: regenerated by the tests (RandomState(0) order)
    meta["ccc"] = ccc
    print("ccc", ccc)

    # padSequence zero-fill: equal mel dim 64 < maxW triggers the right-aligned copy for every item
    gen = torch.Generator().manual_seed(71)
    widths = [70, 90, 90, 66]   # maxW > 64 mel bins -> right-aligned branch (padSequence.py:16-19)
    batch = []
    for w in widths:
        clip = torch.randn(16, 3, 2, 4, 4, generator=gen)
        spec = torch.randn(16, 1, 64, w, generator=gen) + 3.0   # strictly non-zero content
        batch.append((clip, spec, torch.randn(16, generator=gen), torch.randn(16, generator=gen), "w"))
    _, audio, lv, la, _ = R["padSequence"].TrainPadSequence()(batch)
    np.savez_compressed(os.path.join(HERE, "padseq.npz"), zero_mask=(audio == 0).numpy(),
                        audio_sum=audio.double().sum().numpy())
    meta["padseq"] = dict(widths=widths, seed=71)


# ----------------------------------------------------------------------------------------------------------------------
# Default-init, benchmark-size cases (round 2).  Weights are NOT synthetic here: reference modules are constructed under
# torch.manual_seed(seed) and the drop-in modules, constructed under the same seed, must come out bit-identical (the
# constructors mirror the reference's initialisation order; pinned by `param_sums` + tests/test_host_cpu.py).  Large
# tensors are stored as strided channel samples (every row = every (n, t) position is kept: tile borders and dilations
# live on the row axis) plus whole-tensor L2 norms.
# ----------------------------------------------------------------------------------------------------------------------
def _param_sums(module):
    return {k: float(v.double().sum()) for k, v in module.state_dict().items()
            if "final_encoder" not in k and not k.split(".")[-2].startswith("net") and ".net." not in k}


def _check_same_init(ours, ref, what):
    """The drop-in constructor must reproduce the reference's seeded initialisation bit for bit."""
    so, sr = ours.state_dict(), ref.state_dict()
    assert list(so.keys()) == list(sr.keys()), what
    for k in sr:
        assert torch.equal(so[k], sr[k].detach()), (what, k)


def _cot(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def gen_default_init(R, meta):
    import jmt_b200
    live = _live_loss(R)

    # ---- Two_transformers variants with default init; gradients for BOTH the live CCC loss and a well-conditioned random
    #      cotangent (the latter is what the bf16 backward is compared with: reference gradients, not the repo's own fp32 engine)
    cases = [("ttd_transformer_fc_h1", 2, 24, 1, 1, "TRANSFORMER", "FC", 512, 201),
             ("ttd_transformer_sa_h2", 2, 9, 2, 1, "TRANSFORMER", "SELF_ATTEN", 512, 202),
             ("ttd_fc_fc", 3, 10, 1, 1, "FC", "FC", 512, 203),
             ("ttd_none_fc_h2", 5, 7, 2, 1, "NONE", "FC", 512, 204),
             # batch-dimension attention over more than one 128-row tile, and beyond the fused kernel's S <= 320
             # limit at dh = 512 (the long-S path): L = S = B
             ("ttd_none_fc_b300_t2", 300, 2, 1, 1, "NONE", "FC", 512, 205),
             ("ttd_none_fc_b400_t2", 400, 2, 1, 1, "NONE", "FC", 512, 206),
             ("ttd_none_fc_b600_t1_h2", 600, 1, 2, 1, "NONE", "FC", 512, 207)]
    for (name, B, T, h, L, joint, fmt, vin, seed) in cases:
        torch.manual_seed(seed)
        model = R["Two_transformers"](0.0, 0.0, h, L, joint, fmt, vin).eval()
        torch.manual_seed(seed)
        _check_same_init(jmt_b200.Two_transformers(0.0, 0.0, h, L, joint, fmt, vin), model, name)
        aud, vis = O.synth_features(B, T, [512, vin], seed=seed + 1000)
        lv, la = O.synth_labels(B, T, seed=seed + 2000)
        aud.requires_grad_(True)
        vis.requires_grad_(True)
        v, a = model(aud, vis)
        n = v.shape[0] * v.shape[1]
        loss = live(v.reshape(-1, n), lv.reshape(-1, n)) + live(a.reshape(-1, n), la.reshape(-1, n))
        loss.backward(retain_graph=True)
        names, l2, s, head = _grad_summary(model)
        d_aud, d_vis = aud.grad.clone(), vis.grad.clone()
        model.zero_grad(set_to_none=True)
        aud.grad = vis.grad = None
        cv, ca = _cot(v.shape, seed + 3000), _cot(a.shape, seed + 3001)
        torch.autograd.backward([v, a], [cv, ca])
        names2, l2c, sc, headc = _grad_summary(model)
        assert names2 == names
        big = B * T * 512 > 40000
        sl = (slice(None), slice(None), slice(None, None, 16)) if big else (slice(None),) * 3
        np.savez_compressed(os.path.join(HERE, name + ".npz"), vout=v.detach().numpy(), aout=a.detach().numpy(),
                            loss=loss.detach().numpy(), d_aud=d_aud.numpy()[sl], d_vis=d_vis.numpy()[sl],
                            d_aud_l2=float(d_aud.norm()), d_vis_l2=float(d_vis.norm()),
                            grad_l2=l2, grad_sum=s, grad_head=head,
                            c_d_aud=aud.grad.numpy()[sl], c_d_vis=vis.grad.numpy()[sl],
                            c_d_aud_l2=float(aud.grad.norm()), c_d_vis_l2=float(vis.grad.norm()),
                            c_grad_l2=l2c, c_grad_sum=sc, c_grad_head=headc)
        meta[name] = dict(B=B, T=T, heads=h, layers=L, joint=joint, fmt=fmt, vin=vin, init_seed=seed, feat_seed=seed + 1000,
                          label_seed=seed + 2000, cot_seeds=[seed + 3000, seed + 3001], grad_names=names,
                          out_shape=list(v.shape), sample_stride=16 if big else 1, param_sums=_param_sums(model))
        print(name, tuple(v.shape), float(loss))

    # ---- TCN at the benchmarked length: (N=4, 1024, L=300), all four dilations cross 128-row tile borders of the flat layout
    name, N, L, seed = "tcnd_1024_512x4_k5_L300", 4, 300, 211
    torch.manual_seed(seed)
    m = R["TCN"](1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1).eval()
    torch.manual_seed(seed)
    _check_same_init(jmt_b200.TemporalConvNet(1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1), m, name)
    x = torch.randn(N, 1024, L, generator=torch.Generator().manual_seed(seed + 1), requires_grad=True)
    out = m(x)
    cot = _cot(out.shape, seed + 2)
    (out * cot).sum().backward()
    names, l2, s, head = _grad_summary(m)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), out=out.detach().numpy()[:, ::16], out_l2=float(out.norm()),
                        d_x=x.grad.numpy()[:, ::32], d_x_l2=float(x.grad.norm()), grad_l2=l2, grad_sum=s, grad_head=head)
    meta[name] = dict(N=N, L=L, init_seed=seed, x_seed=seed + 1, cot_seed=seed + 2, grad_names=names,
                      out_stride=16, dx_stride=32, param_sums=_param_sums(m))
    print(name, tuple(out.shape))

    # ---- the benchmarked pipeline (BASELINE.json configs[1] at B=4): TCN -> transpose (I3DWSDDA.py:44) | FcLayer(768,512)
    #      -> Two_transformers(TRANSFORMER, FC, h=1, L=1) -> live CCC loss (train.py:283-311), forward + backward
    name, B, T, seed = "piped_b4_t300", 4, 300, 221
    torch.manual_seed(seed)
    fusion = R["Two_transformers"](0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512).eval()
    fc = R["FcLayer"](768, 512).eval()
    tcn = R["TCN"](1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1).eval()
    torch.manual_seed(seed)
    ours = jmt_b200.JMTPipeline(jmt_b200.Two_transformers(0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512), jmt_b200.FcLayer(768, 512),
                                jmt_b200.TemporalConvNet(1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1))
    _check_same_init(ours.fusion, fusion, name)
    _check_same_init(ours.fc_audio, fc, name)
    _check_same_init(ours.tcn, tcn, name)
    g = torch.Generator().manual_seed(seed + 1)
    vis = torch.randn(B, 1024, T, generator=g, requires_grad=True)
    aud = torch.randn(B, T, 768, generator=g, requires_grad=True)
    lv, la = O.synth_labels(B, T, seed=seed + 2)
    v, a = fusion(fc(aud), tcn(vis).transpose(1, 2).contiguous())
    n = v.shape[0] * v.shape[1]
    loss = live(v.reshape(-1, n), lv.reshape(-1, n)) + live(a.reshape(-1, n), la.reshape(-1, n))
    mods = {"fusion": fusion, "fc_audio": fc, "tcn": tcn}

    def summary():
        names, l2, s, head = [], [], [], []
        for pre, mod in mods.items():
            nm, a_, b_, c_ = _grad_summary(mod)
            names += [pre + "." + k for k in nm]
            l2.append(a_); s.append(b_); head.append(c_)
        return names, np.concatenate(l2), np.concatenate(s), np.concatenate(head)
    loss.backward(retain_graph=True)
    names, l2, s, head = summary()
    d_aud, d_vis = aud.grad.clone(), vis.grad.clone()
    for mod in mods.values():
        mod.zero_grad(set_to_none=True)
    aud.grad = vis.grad = None
    cv, ca = _cot(v.shape, seed + 3), _cot(a.shape, seed + 4)
    torch.autograd.backward([v, a], [cv, ca])
    names2, l2c, sc, headc = summary()
    assert names2 == names
    np.savez_compressed(os.path.join(HERE, name + ".npz"), vout=v.detach().numpy(), aout=a.detach().numpy(),
                        lv=lv.numpy(), la=la.numpy(),
                        loss=loss.detach().numpy(), d_aud=d_aud.numpy()[:, :, ::32], d_vis=d_vis.numpy()[:, ::64],
                        d_aud_l2=float(d_aud.norm()), d_vis_l2=float(d_vis.norm()), grad_l2=l2, grad_sum=s, grad_head=head,
                        c_d_aud=aud.grad.numpy()[:, :, ::32], c_d_vis=vis.grad.numpy()[:, ::64],
                        c_d_aud_l2=float(aud.grad.norm()), c_d_vis_l2=float(vis.grad.norm()),
                        c_grad_l2=l2c, c_grad_sum=sc, c_grad_head=headc)
    sums = {}
    for pre, mod in mods.items():
        sums.update({pre + "." + k: vv for k, vv in _param_sums(mod).items()})
    meta[name] = dict(B=B, T=T, init_seed=seed, data_seed=seed + 1, label_seed=seed + 2, cot_seeds=[seed + 3, seed + 4],
                      grad_names=names, out_shape=list(v.shape), aud_stride=32, vis_stride=64, param_sums=sums)
    print(name, tuple(v.shape), float(loss))

    # ---- intra-modal fusion at BASELINE.json configs[3]'s length: B=2, T=1024 (M = 2*B*T = 4096 rows of L=2 'sequences')
    name, B, T, seed = "intrad_b2_t1024", 2, 1024, 231
    torch.manual_seed(seed)
    m = R["Intra"](512, 1, 512, 1).eval()
    torch.manual_seed(seed)
    _check_same_init(jmt_b200.Intra_modal_transformer_fusion(512, 1, 512, 1), m, name)
    fa, fb = O.synth_features(B, T, [512, 768], seed=seed + 1)
    fa.requires_grad_(True)
    fb.requires_grad_(True)
    out = m(fa, fb)
    (out * _cot(out.shape, seed + 2)).sum().backward()
    names, l2, s, head = _grad_summary(m)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), out=out.detach().numpy()[:, :, ::16], out_l2=float(out.norm()),
                        d_a=fa.grad.numpy()[:, :, ::32], d_b=fb.grad.numpy()[:, :, ::32], d_a_l2=float(fa.grad.norm()),
                        d_b_l2=float(fb.grad.norm()), grad_l2=l2, grad_sum=s, grad_head=head)
    meta[name] = dict(B=B, T=T, heads=1, layers=1, init_seed=seed, feat_seed=seed + 1, cot_seed=seed + 2, grad_names=names,
                      out_stride=16, d_stride=32, param_sums=_param_sums(m))
    print(name, tuple(out.shape))


def reference_checkpoint_functions():
    """The reference's own checkpoint writer / reader, main.py:54-70 (`load_clean_weights`) and main.py:105-177
    (`dump_models_into_disk`), compiled from the reference source at run time (main.py as a whole imports the data loaders,
    dllogger configuration etc. and cannot be imported here); their helpers come from the importable tools.py."""
    import ast
    from collections import OrderedDict
    from os.path import join
    _install_stubs()
    sys.modules["matplotlib.ticker"] = types.ModuleType("matplotlib.ticker")
    sys.modules["matplotlib.ticker"].MaxNLocator = object
    sys.path.insert(0, REF)
    import tools as ref_tools
    src = open(os.path.join(REF, "main.py")).read()
    tree = ast.parse(src)
    want = {"load_clean_weights", "dump_models_into_disk", "get_state_dict"}
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {f.name for f in fns} == want, [f.name for f in fns]

    class _Log:
        @staticmethod
        def log(*a, **k):
            pass
    ns = {"torch": torch, "OrderedDict": OrderedDict, "join": join, "DLLogger": _Log, "MyDataParallel": ref_tools.MyDataParallel,
          "state_dict_to_cpu": ref_tools.state_dict_to_cpu, "state_dict_to_gpu": ref_tools.state_dict_to_gpu}
    exec(compile(ast.Module(body=fns, type_ignores=[]), os.path.join(REF, "main.py"), "exec"), ns)
    return ns["dump_models_into_disk"], ns["load_clean_weights"]


def gen_checkpoint(R, meta):
    """A checkpoint written by the REFERENCE's dump_models_into_disk (main.py:105-177) from a reference module: the small
    `backbone_pretrainer_w.pt` (SingleBackbonePretrainer with the weights of the `single_backbone` golden) is committed so that
    the GPU box can load it strict=True into the drop-in and reproduce the golden outputs."""
    dump, _ = reference_checkpoint_functions()
    shapes = O._regressor_shapes("regressor.", 512, 2)
    m = R["SingleBackbonePretrainer"](0.0, 0.0)
    m.load_state_dict(O.synth_params(shapes, seed=61), strict=True)
    out = os.path.join(HERE, "ckpt_ref")
    os.makedirs(out, exist_ok=True)
    dump({}, None, m, None, None, None, None, 3, out)
    assert os.listdir(out) == ["backbone_pretrainer_w.pt"], os.listdir(out)
    meta["ckpt_ref"] = dict(file="ckpt_ref/backbone_pretrainer_w.pt", param_seed=61, golden="single_backbone")


def main():
    torch.set_num_threads(8)
    R = _import_reference()
    meta = {}
    gen_default_inventory(R, meta)
    gen_two_transformers(R, meta)
    gen_c1(R, meta)
    gen_intra(R, meta)
    gen_tcn(R, meta)
    gen_misc(R, meta)
    gen_default_init(R, meta)
    gen_checkpoint(R, meta)
    meta["_generator"] = dict(torch=torch.__version__, numpy=np.__version__, reference=REF)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", os.path.join(HERE, "golden_meta.json"))


if __name__ == "__main__":
    main()
