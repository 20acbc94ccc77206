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
    meta["_generator"] = dict(torch=torch.__version__, numpy=np.__version__, reference=REF)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", os.path.join(HERE, "golden_meta.json"))


if __name__ == "__main__":
    main()
