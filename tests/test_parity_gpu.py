"""Module-level parity (GPU): the drop-in modules, driven through the C-ABI kernels, against
(a) golden vectors produced by the reference itself and (b) the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): predictions within 1e-3 relative in the fp32-operand mode
('fp32'); the bf16-operand mode is reported separately with a looser bound; CCC within 1e-4."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import jmt_b200  # noqa: E402
from oracle import jmt_oracle as O  # noqa: E402

TT_NAMES = ["tt_transformer_fc_h1_l1", "tt_transformer_fc_h4_l2", "tt_transformer_fc_h8_vin1024",
            "tt_transformer_sa_h2_l1", "tt_none_fc_h2_l1", "tt_fc_fc"]
DEV = "cuda"


def _rel(a, b):
    """max |a-b| / max |b|"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-12)


def _rl2(a, b):
    """||a-b|| / ||b||: used for gradients, where a single LeakyReLU/ReLU sign decision on a pre-activation
    that is ~0 legitimately changes a few entries by O(1) (observed: one flip in the seed-51 TCN)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


# tolerances: fp32-operand mode is the north-star 1e-3 gate; bf16-operand mode is stated separately
# (bf16 bounds are ~1.5x the worst deltas measured on B200 with the deliberately "hot" synthetic weights of
#  oracle.synth_params -- 1.5/sqrt(fan_in) -- see DESIGN.md "Numerics"; default-init weights give ~1e-2.)
# 'bf16x3' = the tensor-core mode that must meet the north-star gate (three tcgen05.mma per k-step on hi/lo split operands);
# 'fp32' (FFMA kernel) is kept as an independent cross-check with the same bounds
PRED_TOL = {"bf16x3": 1e-3, "fp32": 1e-3, "bf16": 1.5e-1}     # max-abs relative on predictions / features
PRED_L2 = {"bf16x3": 5e-4, "fp32": 5e-4, "bf16": 8e-2}
# gradients through the CCC loss of the tiny golden cases (B*T = 18..80 predictions) are ill-conditioned (nearly constant
# cotangent -> cancelling terms amplify operand rounding ~200x): bf16x3 (2^-16 operands) measured 4.0e-3 worst
# (tt_transformer_fc_h4_l2), the fp32 FFMA cross-check 3e-5; well-conditioned cotangents are tested on the *_default cases
GRAD_L2 = {"bf16x3": 6e-3, "fp32": 2e-3, "bf16": 3e-1}
GATE = ("bf16x3", "fp32")


def _load(module, params):
    sd = module.state_dict()
    missing = [k for k in sd if k not in params]
    assert not missing, missing
    module.load_state_dict({k: params[k] for k in sd}, strict=True)
    return module.to(DEV)


def _grad_summary(module, names):
    named = dict(module.named_parameters())
    l2, head = [], []
    for n in names:
        g = named[n].grad
        assert g is not None, n
        g = g.detach().double().reshape(-1).cpu()
        l2.append(float(g.norm()))
        h = np.zeros(8)
        h[: min(8, g.numel())] = g[:8].numpy()
        head.append(h)
    return np.array(l2), np.stack(head)


@pytest.mark.parametrize("precision", ["bf16x3", "fp32", "bf16"])
@pytest.mark.parametrize("name", TT_NAMES)
def test_two_transformers(name, precision, golden_meta, golden_dir):
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    params = O.synth_params(O.two_transformers_shapes(m["layers"], m["joint"], m["fmt"], m["vin"]), m["param_seed"])
    model = _load(jmt_b200.Two_transformers(0.0, 0.0, m["heads"], m["layers"], m["joint"], m["fmt"], m["vin"],
                                            precision=precision), params)
    model.eval()
    aud, vis = O.synth_features(m["B"], m["T"], [512, m["vin"]], m["feat_seed"])
    lv, la = O.synth_labels(m["B"], m["T"], m["label_seed"])
    aud_d = aud.to(DEV).requires_grad_(True)
    vis_d = vis.to(DEV).requires_grad_(True)
    v, a = model(aud_d, vis_d)
    assert list(v.shape) == m["out_shape"] and v.is_contiguous() and v.dtype == torch.float32
    for got, want, nm in ((v, g["vout"], "vout"), (a, g["aout"], "aout")):
        assert _rel(got.detach().cpu(), want) < PRED_TOL[precision], (nm, _rel(got.detach().cpu(), want))
        assert _rl2(got.detach().cpu(), want) < PRED_L2[precision], (nm, _rl2(got.detach().cpu(), want))
    # live loss exactly as train.py:303-311 (independent flatten of preds and labels)
    crit = jmt_b200.CCCLoss(digitize_num=1)
    n = v.shape[0] * v.shape[1]
    loss = crit(v.view(-1, n), lv.to(DEV).view(-1, n)) + crit(a.view(-1, n), la.to(DEV).view(-1, n))
    ltol = 1e-4 if precision in GATE else 3e-2
    assert abs(loss.item() - float(g["loss"])) < ltol, (loss.item(), float(g["loss"]))
    gtol = GRAD_L2[precision]
    if precision == "bf16":
        # The CCC cotangent of a tiny case (B*T = 18..80 predictions) is nearly constant across elements, so the
        # parameter gradients are differences of large cancelling terms: a 2 % bf16 prediction delta became a 40 %
        # gradient delta on tt_transformer_sa_h2_l1 (and shrinks 4x per halving of the deliberately hot weights --
        # a round-1 weight-scale sweep), which says nothing about the operators.  So the bf16 backward is checked
        # operator-for-operator: the SAME well-conditioned random cotangent goes through the bf16 engine and the
        # fp32 engine (whose CCC-loss gradients are pinned to the golden vectors by the fp32 leg of this test).
        ref = _load(jmt_b200.Two_transformers(0.0, 0.0, m["heads"], m["layers"], m["joint"], m["fmt"], m["vin"],
                                              precision="fp32"), params).eval()
        aud_r = aud.to(DEV).requires_grad_(True)
        vis_r = vis.to(DEV).requires_grad_(True)
        vr, ar = ref(aud_r, vis_r)
        gen = torch.Generator().manual_seed(7)
        cv = torch.randn(v.shape, generator=gen).to(DEV)
        ca = torch.randn(a.shape, generator=gen).to(DEV)
        torch.autograd.backward([vr, ar], [cv, ca])
        torch.autograd.backward([v, a], [cv, ca])
        want_aud, want_vis = aud_r.grad.cpu().numpy(), vis_r.grad.cpu().numpy()
        want_l2, _ = _grad_summary(ref, m["grad_names"])
    else:
        loss.backward()
        want_aud, want_vis, want_l2 = g["d_aud"], g["d_vis"], g["grad_l2"]
    assert _rl2(aud_d.grad.cpu(), want_aud) < gtol, ("d_aud", _rl2(aud_d.grad.cpu(), want_aud))
    assert _rl2(vis_d.grad.cpu(), want_vis) < gtol, ("d_vis", _rl2(vis_d.grad.cpu(), want_vis))
    l2, head = _grad_summary(model, m["grad_names"])
    rel_l2 = np.abs(l2 - want_l2) / (want_l2 + 1e-12)
    assert rel_l2.max() < gtol * 2, (m["grad_names"][int(rel_l2.argmax())], rel_l2.max())
    if precision in GATE:
        for i, nme in enumerate(m["grad_names"]):
            sc = np.abs(g["grad_head"][i]).max() + 1e-9
            htol = 5e-3 if precision == "fp32" else 2e-2          # bf16x3: 2^-16 operands, ~200x cancellation (see GRAD_L2)
            assert np.abs(head[i] - g["grad_head"][i]).max() < htol * max(sc, g["grad_l2"][i] / 50), nme
    # dead parameters never get a gradient
    for nme, p in model.named_parameters():
        if "final_encoder" in nme or "gated_attention" in nme:
            assert p.grad is None, nme


@pytest.mark.parametrize("precision", ["bf16x3", "fp32", "bf16"])
def test_c1_config(precision, golden_meta, golden_dir):
    """BASELINE.json configs[0]: FcLayer(768,512) + Two_transformers(TRANSFORMER, FC), B=8, T=300."""
    m = golden_meta["c1_b8_t300"]
    g = np.load(os.path.join(golden_dir, "c1_b8_t300.npz"))
    params = O.synth_params(O.two_transformers_shapes(1, "TRANSFORMER", "FC", 512), m["param_seed"])
    fcp = O.synth_params([("fc_layer.weight", (512, 768)), ("fc_layer.bias", (512,))], m["fc_seed"])
    model = _load(jmt_b200.Two_transformers(0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512, precision=precision), params).eval()
    fc = _load(jmt_b200.FcLayer(768, 512, precision=precision), fcp).eval()
    vis, aud768 = O.synth_features(m["B"], m["T"], [512, 768], m["feat_seed"])
    with torch.no_grad():
        v, a = model(fc(aud768.to(DEV)), vis.to(DEV))
    assert tuple(v.shape) == (300, 8)
    for got, want in ((v, g["vout"]), (a, g["aout"])):
        assert _rel(got.cpu(), want) < PRED_TOL[precision] and _rl2(got.cpu(), want) < PRED_L2[precision], \
            (_rel(got.cpu(), want), _rl2(got.cpu(), want))
    # CCC of the engine's predictions vs CCC of the reference's predictions on the same labels: within 1e-4 (fp32 mode)
    lv, _ = O.synth_labels(m["B"], m["T"], 5)
    c_ref = O.ccc_metric(g["vout"].reshape(-1).astype(np.float64), lv.numpy().reshape(-1).astype(np.float64))
    c_new = jmt_b200.cccmetric.ccc(v.reshape(-1), lv.to(DEV).reshape(-1))
    assert abs(c_ref - c_new) < (1e-4 if precision in GATE else 1e-2), (c_ref, c_new)


@pytest.mark.parametrize("precision", ["bf16x3", "fp32", "bf16"])
@pytest.mark.parametrize("name", ["intra_512_768_h2", "intra_512_512_h1_l2"])
def test_intra_modal(name, precision, golden_meta, golden_dir):
    m = golden_meta[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    params = O.synth_params(O.intra_modal_shapes(m["layers"]), m["param_seed"])
    model = _load(jmt_b200.Intra_modal_transformer_fusion(512, m["heads"], 512, m["layers"], precision=precision), params).eval()
    fa, fb = O.synth_features(m["B"], m["T"], [m["da"], m["db"]], m["feat_seed"])
    fa_d, fb_d = fa.to(DEV).requires_grad_(True), fb.to(DEV).requires_grad_(True)
    out = model(fa_d, fb_d)
    assert _rel(out.detach().cpu(), g["out"]) < PRED_TOL[precision] and _rl2(out.detach().cpu(), g["out"]) < PRED_L2[precision]
    w = torch.linspace(-1, 1, out.numel()).reshape(out.shape).to(DEV)
    (out * w).sum().backward()
    gtol = GRAD_L2[precision]
    assert _rl2(fa_d.grad.cpu(), g["d_a"]) < gtol and _rl2(fb_d.grad.cpu(), g["d_b"]) < gtol
    l2, _ = _grad_summary(model, m["grad_names"])
    assert (np.abs(l2 - g["grad_l2"]) / (g["grad_l2"] + 1e-12)).max() < gtol * 2


@pytest.mark.parametrize("precision", ["bf16x3", "fp32", "bf16"])
@pytest.mark.parametrize("name", ["tcn_1024_512x4_k5_L7", "tcn_1024_512x4_k5_L40", "tcn_16_8x2_k3_L19"])
def test_tcn(name, precision, golden_meta, golden_dir):
    m = golden_meta[name]
    if precision == "bf16" and m["cin"] % 8 != 0:
        pytest.skip("tcgen05 path needs channel counts that are multiples of 8")
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    params = O.synth_params(O.tcn_shapes(m["cin"], m["chans"], m["k"]), m["param_seed"])
    model = jmt_b200.TemporalConvNet(m["cin"], m["chans"], kernel_size=m["k"], attention=0, dropout=0.1, precision=precision)
    model.load_state_dict(O.tcn_state_dict(params), strict=True)
    model = model.to(DEV).eval()
    gen = torch.Generator().manual_seed(m["x_seed"])
    x = torch.randn(m["N"], m["cin"], m["L"], generator=gen)
    xd = x.to(DEV).requires_grad_(True)
    out = model(xd)
    assert tuple(out.shape) == g["out"].shape
    assert _rel(out.detach().cpu(), g["out"]) < PRED_TOL[precision], _rel(out.detach().cpu(), g["out"])
    assert _rl2(out.detach().cpu(), g["out"]) < PRED_L2[precision], _rl2(out.detach().cpu(), g["out"])
    w = torch.linspace(-1, 1, out.numel()).reshape(out.shape).to(DEV)
    (out * w).sum().backward()
    gtol = GRAD_L2[precision]
    assert _rl2(xd.grad.cpu(), g["d_x"]) < gtol, _rl2(xd.grad.cpu(), g["d_x"])
    l2, _ = _grad_summary(model, m["grad_names"])
    rel = np.abs(l2 - g["grad_l2"]) / (g["grad_l2"] + 1e-12)
    assert rel.max() < gtol * 2, (m["grad_names"][int(rel.argmax())], rel.max())


@pytest.mark.parametrize("precision", ["bf16x3", "fp32", "bf16"])
def test_reference_written_checkpoint_loads_into_dropin(precision, golden_meta, golden_dir, tmp_path):
    """SURVEY 8f N3 on the device: `backbone_pretrainer_w.pt` as written by the REFERENCE's dump_models_into_disk
    (main.py:105-177; tests/golden/make_golden.py::gen_checkpoint) loads strict=True into the drop-in through
    jmt_b200.checkpoint (main.py:54-70 conventions) and reproduces the reference's outputs; the drop-in's own dump of the
    same weights is byte-for-byte the same state_dict."""
    from jmt_b200 import checkpoint as CK
    c = golden_meta["ckpt_ref"]
    m = golden_meta[c["golden"]]
    g = np.load(os.path.join(golden_dir, c["golden"] + ".npz"))
    model = jmt_b200.SingleBackbonePretrainer(0.0, 0.0, precision=precision).to(DEV).eval()
    CK.load_models_from_disk(os.path.join(golden_dir, "ckpt_ref"), {"backbone_pretrainer": model}, map_location=DEV)
    (x,) = O.synth_features(m["B"], m["T"], [512], m["feat_seed"])
    with torch.no_grad():
        v, a = model(x.to(DEV))
    tol = 1e-3 if precision != "bf16" else 2e-2
    assert _rel(v.cpu(), g["v"]) < tol and _rel(a.cpu(), g["a"]) < tol
    CK.dump_models_into_disk(str(tmp_path), {"backbone_pretrainer": model})
    ref_sd = torch.load(os.path.join(golden_dir, "ckpt_ref", "backbone_pretrainer_w.pt"), weights_only=True)
    my_sd = torch.load(os.path.join(tmp_path, "backbone_pretrainer_w.pt"), weights_only=True)
    assert list(ref_sd.keys()) == list(my_sd.keys())
    for k in ref_sd:
        assert my_sd[k].device.type == "cpu" and torch.equal(ref_sd[k], my_sd[k]), k


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("fmt", ["FC", "SELF_ATTEN"])
def test_w_jr_standalone_layouts(fmt, precision):
    """MultimodalTransformer_w_JR used on its own (mm_multi_transformers.py:118-214): the FC head returns (T, B, 1024) -- the
    reference never permutes back -- the SELF_ATTEN head (B, T, 512); forward and input gradients vs the oracle."""
    B, T, heads = 3, 11, 2
    shapes = [(k[len("mm_transformer."):], sh) for k, sh in O.two_transformers_shapes(1, "TRANSFORMER", fmt, 512) if k.startswith("mm_transformer.")]
    params = O.synth_params(shapes, 77)
    model = _load(jmt_b200.MultimodalTransformer_w_JR(512, 512, heads, 512, 1, fmt, precision=precision), params).eval()
    vis, aud = O.synth_features(B, T, [512, 512], 78)
    vis_d, aud_d = vis.to(DEV).requires_grad_(True), aud.to(DEV).requires_grad_(True)
    out = model(vis_d, aud_d)
    vo, ao = vis.clone().requires_grad_(True), aud.clone().requires_grad_(True)
    po = {k: t.clone() for k, t in params.items()}
    want = O.w_jr_forward(vo, ao, po, "", heads, 1, fmt)
    assert tuple(out.shape) == tuple(want.shape) == ((T, B, 1024) if fmt == "FC" else (B, T, 512))
    tol = 1e-3 if precision == "bf16x3" else PRED_TOL["bf16"]
    assert _rel(out.detach().cpu(), want.detach()) < tol, _rel(out.detach().cpu(), want.detach())
    cot = torch.randn(want.shape, generator=torch.Generator().manual_seed(79))
    (want * cot).sum().backward()
    (out * cot.to(DEV)).sum().backward()
    gt = 8e-3 if precision == "bf16x3" else GRAD_L2["bf16"]      # hot synthetic weights; SELF_ATTEN measured 4.6e-3 (bf16x3)
    assert _rl2(vis_d.grad.cpu(), vo.grad) < gt and _rl2(aud_d.grad.cpu(), ao.grad) < gt


def test_single_backbone_and_fc(golden_meta, golden_dir):
    m = golden_meta["single_backbone"]
    g = np.load(os.path.join(golden_dir, "single_backbone.npz"))
    params = O.synth_params(O._regressor_shapes("regressor.", 512, 2), m["param_seed"])
    model = _load(jmt_b200.SingleBackbonePretrainer(0.0, 0.0, precision="fp32"), params).eval()
    (x,) = O.synth_features(m["B"], m["T"], [512], m["feat_seed"])
    xd = x.to(DEV).requires_grad_(True)
    v, a = model(xd)
    assert _rel(v.detach().cpu(), g["v"]) < 1e-3 and _rel(a.detach().cpu(), g["a"]) < 1e-3
    (v.sum() + 2 * a.sum()).backward()
    xo = x.clone().requires_grad_(True)
    po = {k: t.clone().requires_grad_(True) for k, t in params.items()}
    vo, ao = O.single_backbone_pretrainer_forward(xo, po)
    (vo.sum() + 2 * ao.sum()).backward()
    assert _rel(xd.grad.cpu(), xo.grad) < 2e-3
    for k, p in model.named_parameters():
        assert _rel(p.grad.cpu(), po[k].grad) < 2e-3, k


def test_training_semantics_dropout_and_modes():
    """train()/eval() toggle dropout only; p>0 training changes outputs, eval does not; deepcopy works."""
    import copy
    torch.manual_seed(0)
    model = jmt_b200.Two_transformers(0.5, 0.5, 2, 1, "FC", "FC", 512, precision="fp32").to(DEV)
    aud, vis = (t.to(DEV) for t in O.synth_features(2, 5, [512, 512], 1))
    model.eval()
    v0, _ = model(aud, vis)
    v1, _ = model(aud, vis)
    assert torch.equal(v0, v1)
    model.train()
    v2, _ = model(aud, vis)
    assert not torch.equal(v0, v2)
    m2 = copy.deepcopy(model).eval()
    v3, _ = m2(aud, vis)
    assert torch.equal(v0, v3)
    # a step of SGD under autocast + GradScaler (train.py:89,101,314-316) runs and changes the live params only
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    scaler = torch.amp.GradScaler("cuda")
    lv, la = (t.to(DEV) for t in O.synth_labels(2, 5, 2))
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    with torch.autocast("cuda", dtype=torch.float16):
        v, a = model(aud, vis)
        crit = jmt_b200.CCCLoss(1)
        loss = crit(v.view(1, -1), lv.view(1, -1)) + crit(a.view(1, -1), la.view(1, -1))
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    changed = [k for k, p in model.named_parameters() if not torch.equal(p, before[k])]
    assert "vregressor.3.weight" in changed and "mm_transformer.fc.weight" in changed


def test_replicas_from_python_threads_are_reentrant():
    """MyDataParallel (main.py:487-489) calls forward of replicated modules from one Python thread per replica: the
    library keeps no global mutable state (thread-local error text, per-call tensor maps, stream taken from the caller),
    so concurrent forward+backward on two streams gives exactly what the same calls give one after the other."""
    import copy, threading
    torch.manual_seed(0)
    base = jmt_b200.Two_transformers(0.0, 0.0, 2, 1, "TRANSFORMER", "FC", 512, precision="bf16").to(DEV).train()
    feats = [tuple(t.to(DEV) for t in O.synth_features(3, 40, [512, 512], 10 + i)) for i in range(2)]

    def run(model, aud, vis, out, idx, stream=None):
        for _ in range(4):
            model.zero_grad(set_to_none=True)
            with torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext():
                v, a = model(aud, vis)
                (v.float().sum() + 2 * a.float().sum()).backward()
            if stream is not None:
                stream.synchronize()
        out[idx] = (v.detach().clone(), a.detach().clone(),
                    {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})

    import contextlib
    serial, threaded = [None, None], [None, None]
    for i in range(2):
        run(copy.deepcopy(base), *feats[i], serial, i)
    torch.cuda.synchronize()
    errors = []

    def guarded(*args):
        try:
            run(*args)
        except Exception as e:       # surface worker exceptions in the main thread
            errors.append(e)

    threads = [threading.Thread(target=guarded, args=(copy.deepcopy(base), *feats[i], threaded, i, torch.cuda.Stream()))
               for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    assert not errors, errors
    for i in range(2):
        # split-K / reduce-add GEMMs accumulate with fp32 atomics, so the order (not the operands) may differ between runs
        for j in range(2):
            assert _rel(threaded[i][j].float().cpu(), serial[i][j].float().cpu()) < 1e-2
        assert serial[i][2].keys() == threaded[i][2].keys()
        for k in serial[i][2]:
            assert _rel(threaded[i][2][k].float().cpu(), serial[i][2][k].float().cpu()) < 1e-2, k


def test_cpu_tensor_is_refused():
    model = jmt_b200.FcLayer(768, 512).to(DEV)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.randn(2, 3, 768))


def test_graphed_step_matches_eager():
    """jmt_b200.GraphedStep (whole training step as one CUDA graph per buffer set) produces the same losses and
    parameters as launching every kernel from Python."""
    import copy
    torch.manual_seed(0)
    B, T = 4, 24
    base = jmt_b200.JMTPipeline(jmt_b200.Two_transformers(0.0, 0.0, 2, 1, "TRANSFORMER", "FC", 512, precision="bf16"),
                                jmt_b200.FcLayer(768, 512, precision="bf16"),
                                jmt_b200.TemporalConvNet(64, [512] * 2, kernel_size=3, attention=0, dropout=0.0, precision="bf16")).to(DEV).train()
    gen = torch.Generator().manual_seed(3)
    sets = []
    for _ in range(2):
        sets.append((torch.randn(B, T, 768, generator=gen).to(DEV), torch.randn(B, 64, T, generator=gen).to(DEV),
                     (torch.rand(B, T, generator=gen) * 2 - 1).to(DEV), (torch.rand(B, T, generator=gen) * 2 - 1).to(DEV)))
    n = B * T
    crit = jmt_b200.CCCLoss(digitize_num=1)
    results = []
    for graphed in (False, True):
        model = copy.deepcopy(base)
        opt = torch.optim.SGD(model.live_parameters(), lr=1e-2)

        def step(aud, vis, lv, la):
            v, a = model(aud, vis)
            loss = crit(v.view(-1, n), lv.view(-1, n)) + crit(a.view(-1, n), la.view(-1, n))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss
        losses = []
        if graphed:
            g = jmt_b200.GraphedStep(step, sets, warmup=2)           # 2 warm-up steps happen before capture ...
            assert g.launches_per_step > 50
            for i in range(4):
                losses.append(float(g.replay(i % 2).item()))
        else:
            for i in range(2):                                       # ... so the eager run takes the same 2 first
                step(*sets[i % 2])
            for i in range(4):
                losses.append(float(step(*sets[i % 2]).item()))
        results.append((losses, [p.detach().float().cpu().clone() for p in model.live_parameters()]))
    (le, pe), (lg, pg) = results
    # fp32 atomics in split-K / bias-gradient reductions make bf16 steps reproducible only to rounding noise
    assert np.allclose(le, lg, atol=2e-3), (le, lg)
    num = sum(float((a - b).norm() ** 2) for a, b in zip(pe, pg)) ** 0.5
    den = sum(float(a.norm() ** 2) for a in pe) ** 0.5
    assert num / den < 1e-3, num / den


@pytest.mark.parametrize("precision", ["bf16x3", "fp32", "bf16"])
def test_tcn_sequence_features_and_time_max(precision, golden_meta, golden_dir):
    """temporal(x).transpose(1,2) (I3DWSDDA.py:44) and the fused torch.max(ft, 1) (tsav.py:216) on the
    reference-faithful placement (many clips of L = 7), against the golden TCN output of the reference."""
    m = golden_meta["tcn_1024_512x4_k5_L7"]
    g = np.load(os.path.join(golden_dir, "tcn_1024_512x4_k5_L7.npz"))
    params = O.synth_params(O.tcn_shapes(m["cin"], m["chans"], m["k"]), m["param_seed"])
    model = jmt_b200.TemporalConvNet(m["cin"], m["chans"], kernel_size=m["k"], attention=0, dropout=0.1, precision=precision)
    model.load_state_dict(O.tcn_state_dict(params), strict=True)
    model = model.to(DEV).eval()
    gen = torch.Generator().manual_seed(m["x_seed"])
    x = torch.randn(m["N"], m["cin"], m["L"], generator=gen)
    ref = torch.from_numpy(g["out"]).transpose(1, 2)                     # (N, L, 512)
    xd = x.to(DEV).requires_grad_(True)
    seq = model.forward_sequence_features(xd)
    assert tuple(seq.shape) == tuple(ref.shape)
    assert _rel(seq.detach().cpu(), ref) < PRED_TOL[precision]
    pooled = model.forward_sequence_features(xd, max_over_time=True)
    want, idx = ref.max(1)
    assert _rel(pooled.detach().cpu(), want) < PRED_TOL[precision]
    # backward of the pooled path = backward of gathering the arg-max positions of the sequence path
    w = torch.linspace(-1, 1, pooled.numel()).reshape(pooled.shape).to(DEV)
    (pooled * w).sum().backward()
    g_pool = xd.grad.clone()
    xd.grad = None
    seq2 = model.forward_sequence_features(xd)
    am = seq2.detach().argmax(1, keepdim=True)
    (seq2.gather(1, am).squeeze(1) * w).sum().backward()
    assert _rl2(g_pool.cpu(), xd.grad.cpu()) < (1e-4 if precision in GATE else 3e-2)


def test_full_size_properties():
    """BASELINE.json configs[1] size (B=256 windows, T=300): the oracle cannot run it in seconds, so parity is carried by
    size-independent properties of the domain:
      (1) windows are independent (the path shards along the batch): the first 16 windows give the same
          predictions alone as inside the full batch of 256;
      (2) the two independent implementations (tcgen05 bf16 vs FFMA fp32 kernels) agree on 64 full-length windows;
      (3) the six CCC sums are additive over shards (the quantity the ranks all-reduce); ccc(x, x) = 1 and, for centred x,
          ccc(x, -x) = -1."""
    torch.manual_seed(0)
    B, T = 256, 300
    mods = {}
    for prec in ("bf16", "fp32"):
        torch.manual_seed(1)
        mods[prec] = jmt_b200.JMTPipeline(jmt_b200.Two_transformers(0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512, precision=prec),
                                          jmt_b200.FcLayer(768, 512, precision=prec),
                                          jmt_b200.TemporalConvNet(1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1,
                                                                   precision=prec)).to(DEV).eval()
    mods["fp32"].load_state_dict(mods["bf16"].state_dict(), strict=True)
    gen = torch.Generator().manual_seed(2)
    aud = torch.randn(B, T, 768, generator=gen).to(DEV)
    vis = torch.randn(B, 1024, T, generator=gen).to(DEV)
    with torch.no_grad():
        v_full, a_full = mods["bf16"](aud, vis)                       # (T, B): SURVEY Q1
        v_16, a_16 = mods["bf16"](aud[:16].contiguous(), vis[:16].contiguous())
        v_32f, a_32f = mods["fp32"](aud[:64].contiguous(), vis[:64].contiguous())
    assert tuple(v_full.shape) == (T, B)
    # (1) every kernel reduces each output element in a batch-independent order: bit-exact
    assert torch.equal(v_full[:, :16], v_16) and torch.equal(a_full[:, :16], a_16)
    # (2) bf16-operand tcgen05 path vs fp32 FFMA path, default-init weights
    for got, want in ((v_full[:, :64], v_32f), (a_full[:, :64], a_32f)):
        assert _rl2(got.cpu(), want.cpu()) < 3e-2, _rl2(got.cpu(), want.cpu())
    # (3) CCC: shard additivity of the sums and the metric's fixed points, at the full 76 800 predictions
    from jmt_b200.losses import six_sums
    x = v_full.t().contiguous()                                       # (B, T)
    y = torch.rand(B, T, generator=gen).to(DEV) * 2 - 1
    whole = six_sums(x.reshape(1, -1), y.reshape(1, -1))
    parts = sum(six_sums(x[i::8].reshape(1, -1).contiguous(), y[i::8].reshape(1, -1).contiguous()) for i in range(8))
    assert torch.allclose(whole, parts, rtol=1e-12, atol=1e-9)
    assert abs(jmt_b200.cccmetric.ccc(x.reshape(-1), x.reshape(-1)) - 1.0) < 1e-6
    xc = (x - x.mean()).reshape(-1)
    assert abs(jmt_b200.cccmetric.ccc(xc, -xc) + 1.0) < 1e-4


@pytest.mark.parametrize("precision", ["bf16x3", "fp32", "bf16"])
@pytest.mark.parametrize("B,T,heads,joint,fmt", [(1, 1, 8, "TRANSFORMER", "FC"), (1, 2, 1, "TRANSFORMER", "SELF_ATTEN"),
                                                 (3, 1, 2, "NONE", "FC"), (1, 5, 4, "FC", "FC")])
def test_degenerate_shapes_against_oracle(B, T, heads, joint, fmt, precision):
    """Smallest inputs the reference accepts (one window, one time step: attention over a single key; a batch-dimension
    attention of length 3): forward and input gradients against the CPU oracle on the same seeded weights."""
    params = O.synth_params(O.two_transformers_shapes(1, joint, fmt, 512), 77)
    model = _load(jmt_b200.Two_transformers(0.0, 0.0, heads, 1, joint, fmt, 512, precision=precision), params).eval()
    aud, vis = O.synth_features(B, T, [512, 512], 78)
    po = {k: v.clone() for k, v in params.items()}
    ao, vo_in = aud.clone().requires_grad_(True), vis.clone().requires_grad_(True)
    vo, aout = O.two_transformers_forward(ao, vo_in, po, heads, 1, joint, fmt)
    (vo.sum() + 2 * aout.sum()).backward()
    ad, vd = aud.to(DEV).requires_grad_(True), vis.to(DEV).requires_grad_(True)
    v, a = model(ad, vd)
    assert tuple(v.shape) == tuple(vo.shape)
    tol = 1e-3 if precision in GATE else PRED_TOL["bf16"]
    assert _rel(v.detach().cpu(), vo.detach()) < tol and _rel(a.detach().cpu(), aout.detach()) < tol
    (v.sum() + 2 * a.sum()).backward()
    gt = 2e-3 if precision in GATE else GRAD_L2["bf16"]
    assert _rl2(ad.grad.cpu(), ao.grad) < gt and _rl2(vd.grad.cpu(), vo_in.grad) < gt


def test_training_reduces_ccc_loss():
    """End-to-end sanity beyond single-step parity: a few dozen Adam steps of the bf16 pipeline (TCN + FcLayer +
    Two_transformers, CCC loss) on a learnable synthetic target lower the loss substantially and stay finite."""
    torch.manual_seed(0)
    B, T = 8, 40
    model = jmt_b200.JMTPipeline(jmt_b200.Two_transformers(0.0, 0.0, 2, 1, "TRANSFORMER", "FC", 512, precision="bf16"),
                                 jmt_b200.FcLayer(768, 512, precision="bf16"),
                                 jmt_b200.TemporalConvNet(64, [512] * 2, kernel_size=3, attention=0, dropout=0.0,
                                                          precision="bf16")).to(DEV).train()
    gen = torch.Generator().manual_seed(1)
    aud = torch.randn(B, T, 768, generator=gen).to(DEV)
    vis = torch.randn(B, 64, T, generator=gen).to(DEV)
    # targets the model can express: smooth functions of a few input channels, laid out like the (T, B) predictions (Q1)
    tv = torch.tanh(aud[:, :, :4].sum(-1)).t().contiguous()
    ta = torch.tanh(vis[:, :4, :].sum(1)).t().contiguous()
    crit = jmt_b200.CCCLoss(digitize_num=1)
    opt = torch.optim.Adam(model.live_parameters(), lr=3e-4)
    n = B * T
    losses = []
    for _ in range(60):
        v, a = model(aud, vis)
        loss = crit(v.view(-1, n), tv.view(-1, n)) + crit(a.view(-1, n), ta.view(-1, n))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(float(loss.item()))
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < 0.5 * losses[0], (losses[0], losses[-1])


def test_make_graphed_callables_trains_like_eager():
    """The one-line route for unmodified train.py: torch.cuda.make_graphed_callables(module, sample_args,
    allow_unused_input=True) (dead parameters never get gradients).  The captured forward must re-cast
    the fp32 parameters on every replay -- otherwise it would keep using the bf16 copies of the capture-time weights."""
    import copy
    torch.manual_seed(0)
    B, T = 4, 32
    base = jmt_b200.Two_transformers(0.0, 0.0, 2, 1, "TRANSFORMER", "FC", 512, precision="bf16").to(DEV).train()
    gen = torch.Generator().manual_seed(5)
    aud = torch.randn(B, T, 512, generator=gen).to(DEV)
    vis = torch.randn(B, T, 512, generator=gen).to(DEV)
    tv = torch.tanh(aud[:, :, :4].sum(-1)).t().contiguous()
    crit = jmt_b200.CCCLoss(digitize_num=1)
    out = []
    for graphed in (False, True):
        model = copy.deepcopy(base)
        opt = torch.optim.SGD(model.live_parameters(), lr=0.05)
        call = torch.cuda.make_graphed_callables(model, (aud, vis), allow_unused_input=True) if graphed else model
        losses = []
        for _ in range(12):
            v, a = call(aud, vis)
            loss = crit(v.reshape(1, -1), tv.reshape(1, -1)) + crit(a.reshape(1, -1), tv.reshape(1, -1))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            losses.append(float(loss.item()))
        out.append(losses)
    eager, graph = out
    assert eager[-1] < eager[0] - 0.02, eager                 # the toy problem is being learned (weights really change)
    assert np.allclose(eager, graph, atol=5e-3), (eager, graph)  # ... identically through the graphed callable


def test_dropout_masks_vary_across_graph_replays():
    """Dropout under CUDA-graph replay: the Philox counter lives on the device and every forward advances it, so two
    replays of the same captured step draw different masks (host-side seeds are frozen into the graph at capture)."""
    torch.manual_seed(0)
    model = jmt_b200.Two_transformers(0.5, 0.5, 2, 1, "FC", "FC", 512, precision="bf16").to(DEV).train()
    aud, vis = (t.to(DEV) for t in O.synth_features(2, 16, [512, 512], 3))
    outs = []

    def fwd(a_, v_):
        with torch.no_grad():
            v, a = model(a_, v_)
        outs.append(v)
        return v.sum()
    g = jmt_b200.GraphedStep(fwd, [(aud, vis)], warmup=2)
    static_v = outs[-1]                      # the tensor the captured graph writes
    g.replay(0)
    v1 = static_v.clone()
    g.replay(0)
    v2 = static_v.clone()
    assert not torch.equal(v1, v2)
    model.eval()                             # eval: no dropout, deterministic
    with torch.no_grad():
        e1, _ = model(aud, vis)
        e2, _ = model(aud, vis)
    assert torch.equal(e1, e2)
