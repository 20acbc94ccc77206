"""Kernel-level parity tests (GPU): every C-ABI kernel against a plain PyTorch fp32/fp64 reference of the
same op on the same seeded inputs.  GEMM cases cover K-/MN-major operands, batches, ragged tails, taps with
zero-fill shifts, reduce_batch, split-K, every epilogue -- for both the tcgen05 (bf16) and the FFMA (fp32) path."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import jmt_b200  # noqa: E402
from jmt_b200 import _lib as L  # noqa: E402
from jmt_b200 import engine as E  # noqa: E402


def _ctx(precision="fp32"):
    return E.Ctx({}, precision, False, False)


def _ref_gemm(A, B, *, a_major, b_major, M, N, K, ntaps, a_shift, b_shift, nb, reduce_batch, alpha, bias, act, slope):
    """fp64 reference of the jmt_gemm_desc semantics.  A: (nb, rows, cols) per-batch 2-D operand as stored."""
    outs = []
    for b in range(nb):
        acc = torch.zeros(M, N, dtype=torch.float64)
        for j in range(ntaps):
            ash = a_shift[0] + j * a_shift[1]
            bsh = b_shift[0] + j * b_shift[1]
            a2 = A[b].double()
            if a_major == L.MAJOR_K:        # stored (a_rows, K): row m+ash
                Am = torch.zeros(M, K, dtype=torch.float64)
                rows = torch.arange(M) + ash
                ok = (rows >= 0) & (rows < a2.shape[0])
                Am[ok] = a2[rows[ok], :K]
            else:                            # stored (k rows, M): row k+ash
                Am = torch.zeros(M, K, dtype=torch.float64)
                for k in range(K):
                    r = k + ash
                    if 0 <= r < a2.shape[0]:
                        Am[:, k] = a2[r, :M]
            b2 = B[b if B.shape[0] > 1 else 0].double()
            if b_major == L.MAJOR_K:        # stored (N, ntaps*K)
                Bm = b2[:N, j * K:(j + 1) * K]
            else:                            # stored (k rows, N)
                Bm = torch.zeros(N, K, dtype=torch.float64)
                for k in range(K):
                    r = k + bsh
                    if 0 <= r < b2.shape[0]:
                        Bm[:, k] = b2[r, :N]
            acc += Am @ Bm.t()
        outs.append(acc)
    if reduce_batch:
        outs = [sum(outs)]
    res = []
    for o in outs:
        o = alpha * o
        if bias is not None:
            o = o + bias.double()
        if act == L.ACT_RELU:
            o = torch.relu(o)
        elif act == L.ACT_LEAKY:
            o = torch.where(o >= 0, o, o * slope)
        res.append(o)
    return torch.stack(res)


GEMM_CASES = [
    # name, M, N, K, a_major, b_major, nb, ntaps, a_shift, b_shift, reduce, split, act, bias, d_dtype, store
    ("tn_basic", 256, 256, 128, 0, 0, 1, 1, (0, 0), (0, 0), False, 1, 0, False, "f32", 0),
    ("tn_tails", 300, 300, 72, 0, 0, 3, 1, (0, 0), (0, 0), False, 1, 0, True, "f32", 0),
    ("tn_bf16_relu", 200, 512, 512, 0, 0, 1, 1, (0, 0), (0, 0), False, 1, 1, True, "act", 0),
    ("tn_n1536", 384, 1536, 512, 0, 0, 1, 1, (0, 0), (0, 0), False, 1, 0, True, "act", 0),
    ("tn_accumulate", 130, 96, 64, 0, 0, 2, 1, (0, 0), (0, 0), False, 1, 0, False, "f32", 1),
    ("nn_bmn", 300, 64, 300, 0, 1, 4, 1, (0, 0), (0, 0), False, 1, 0, False, "act", 0),       # P @ V
    ("tt_amn", 200, 128, 136, 1, 0, 2, 1, (0, 0), (0, 0), False, 1, 0, False, "f32", 0),
    ("wgrad_mnmn_split", 512, 512, 1000, 1, 1, 1, 1, (0, 0), (0, 0), False, 4, 0, False, "f32", 2),
    ("dk_mnmn_batched", 300, 64, 300, 1, 1, 3, 1, (0, 0), (0, 0), False, 1, 0, False, "act", 1),
    ("conv_fwd_taps", 40, 96, 64, 0, 0, 3, 5, (-8, 2), (0, 0), False, 1, 2, True, "act", 0),
    ("conv_dgrad_taps", 40, 64, 96, 0, 0, 3, 5, (8, -2), (0, 0), False, 1, 0, False, "act", 0),
    ("conv_wgrad_tap", 96, 64, 40, 1, 1, 3, 1, (0, 0), (-4, 0), True, 2, 0, False, "f32", 2),
    ("n_small", 150, 8, 128, 0, 0, 1, 1, (0, 0), (0, 0), False, 1, 0, True, "f32", 0),
    ("k_small", 150, 300, 16, 0, 0, 2, 1, (0, 0), (0, 0), False, 1, 0, False, "f32", 0),
    # CTA-pair (cta_group::2) paths: ghost tile of an odd M tail, MN-major B halves that are not 64-multiples,
    # pairs across two batch entries sharing B (conv), many tiles per CTA pair (persistent loop + both TMEM stages)
    ("pair_m_ghost", 1100, 160, 192, 0, 0, 1, 1, (0, 0), (0, 0), False, 1, 0, True, "act", 0),
    ("pair_bmn_n192", 512, 192, 256, 0, 1, 2, 1, (0, 0), (0, 0), False, 1, 0, False, "f32", 0),
    ("conv_fwd_taps_nb4", 40, 96, 64, 0, 0, 4, 5, (-8, 2), (0, 0), False, 1, 2, True, "act", 0),
    ("conv_dgrad_taps_nb6", 300, 64, 96, 0, 0, 6, 5, (8, -2), (0, 0), False, 1, 0, False, "act", 1),
    ("pair_persistent", 128 * 40, 512, 128, 0, 0, 8, 1, (0, 0), (0, 0), False, 1, 1, True, "act", 0),
    # wide pair tiles (256 x 512, two MMAs per k-step, one accumulator stage): >= 24 k-iterations per tile and N % 512 == 0
    ("wide_tn_relu", 1024, 1024, 1600, 0, 0, 1, 1, (0, 0), (0, 0), False, 1, 1, True, "act", 0),
    ("wide_tn_ghost_f32_acc", 1100, 512, 1544, 0, 0, 2, 1, (0, 0), (0, 0), False, 1, 0, True, "f32", 1),
    ("wide_wgrad_mnmn", 512, 512, 3200, 1, 1, 1, 1, (0, 0), (0, 0), False, 2, 0, False, "f32", 2),
    ("wide_conv_dgrad_taps", 1100, 512, 320, 0, 0, 1, 5, (8, -2), (0, 0), False, 1, 0, False, "act", 0),
    ("wide_nn_bmn", 512, 512, 1600, 0, 1, 2, 1, (0, 0), (0, 0), False, 1, 0, False, "act", 0),
    # tail split: the tiles of a nearly empty last round (78 pair tiles on 74 CTA pairs, 156 tiles on 148 CTAs) are cut along N
    # into pieces with their own UMMA descriptor and B tensor map
    ("tail_tn_bias_relu", 256 * 78, 256, 128, 0, 0, 1, 1, (0, 0), (0, 0), False, 1, 1, True, "act", 0),
    ("tail_wide_nn_bmn_acc", 256 * 78, 512, 512, 0, 1, 1, 1, (0, 0), (0, 0), False, 1, 0, False, "act", 1),
    ("tail_wide_tn_f32", 256 * 78 - 40, 512, 576, 0, 0, 1, 1, (0, 0), (0, 0), False, 1, 2, True, "f32", 0),
    ("tail_1cta_batched", 300, 256, 64, 0, 0, 52, 1, (0, 0), (0, 0), False, 1, 0, True, "act", 0),
]


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("case", GEMM_CASES, ids=[c[0] for c in GEMM_CASES])
def test_gemm(case, precision):
    (name, M, N, K, a_major, b_major, nb, ntaps, a_shift, b_shift, reduce, split, act, use_bias, d_kind, store) = case
    torch.manual_seed(hash(name) % 1000)
    dev = torch.device("cuda")
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    pad = lambda v: (v + 7) // 8 * 8  # noqa: E731
    # stored operand shapes (per batch), leading dims padded to 8 with garbage beyond the valid extent
    if a_major == L.MAJOR_K:
        a_rows, a_cols = M, K
    else:
        a_rows, a_cols = K, M
    if b_major == L.MAJOR_K:
        b_rows, b_cols = N, ntaps * K
    else:
        b_rows, b_cols = K, N
    a_ld, b_ld = pad(a_cols) + 8, pad(b_cols)
    A_store = torch.full((nb, a_rows, a_ld), float("nan"))
    A_val = (torch.randn(nb, a_rows, a_cols) * 0.5).to(dt).float()
    A_store[:, :, :a_cols] = A_val
    share_b = name.startswith("conv_fwd") or name.startswith("conv_dgrad") or name == "tn_tails"
    nbb = 1 if share_b else nb
    B_store = torch.full((nbb, b_rows, b_ld), float("nan"))
    B_val = (torch.randn(nbb, b_rows, b_cols) * 0.5).to(dt).float()
    B_store[:, :, :b_cols] = B_val
    bias = torch.randn(N) if use_bias else None
    alpha, slope = 0.75, 0.01
    ref = _ref_gemm(A_val, B_val, a_major=a_major, b_major=b_major, M=M, N=N, K=K, ntaps=ntaps, a_shift=a_shift,
                    b_shift=b_shift, nb=nb, reduce_batch=reduce, alpha=alpha, bias=bias, act=act, slope=slope)
    d_dt = torch.float32 if d_kind == "f32" else dt
    nbd = 1 if reduce else nb
    d_ld = pad(N) + 8
    D0 = torch.randn(nbd, M, d_ld) * 0.1
    if store == L.STORE:
        expect = ref
    else:
        expect = ref + D0[:, :, :N].to(d_dt).double()
    A_d, B_d, D_d = A_store.to(dev, dt), B_store.to(dev, dt), D0.to(dev, d_dt)
    ctx = _ctx(precision)
    E.gemm(ctx, A_d, B_d, D_d, M=M, N=N, K=K, a_major=a_major, b_major=b_major, a_rows=a_rows, b_rows=b_rows,
           a_ld=a_ld, b_ld=b_ld, d_ld=d_ld, nb0=1, nb1=nb, a_bs=(0, a_rows * a_ld),
           b_bs=(0, 0 if share_b else b_rows * b_ld), d_bs=(0, 0 if reduce else M * d_ld),
           bias=bias.to(dev) if use_bias else None, act=act, slope=slope, alpha=alpha, store=store, ntaps=ntaps,
           a_shift=a_shift, b_shift=b_shift, reduce_batch=reduce, split_k=split)
    torch.cuda.synchronize()
    got = D_d.float().cpu()[:, :, :N].double()
    pad_after = D_d.float().cpu()[:, :, N:]
    assert torch.equal(pad_after, D0[:, :, N:].to(d_dt).float()), "GEMM wrote outside its N columns"
    # bf16x3 (full fp32 operand values, 16 operand mantissa bits on the tensor cores): dropped lo*lo terms ~2^-16 per product
    tol = (6e-5 if precision == "bf16x3" else 2e-4) if d_dt == torch.float32 else 1.2e-2
    err = (got - expect).abs()
    scale = expect.abs().max().item() + 1e-6
    bad = err > tol * scale
    if bad.any():
        idx = bad.nonzero()[:8].tolist()
        raise AssertionError(f"{name}/{precision}: max err {err.max().item():.4g} (scale {scale:.3g}); first bad {idx}; "
                             f"got {[got[tuple(i)].item() for i in idx[:4]]} want {[expect[tuple(i)].item() for i in idx[:4]]}; "
                             f"bad rows {sorted(set(i[1] for i in bad.nonzero().tolist()))[:16]} "
                             f"bad cols {sorted(set(i[2] for i in bad.nonzero().tolist()))[:16]} frac {bad.float().mean().item():.3f}")


BRES_CASES = [
    # name, M, N, K, b_major, act, bias, store, aux
    ("lin_relu", 1100, 512, 512, 0, 1, True, 0, False),
    ("dgrad_mn_acc", 1024, 1024, 448, 1, 0, False, 1, False),
    ("k256_n256", 2048, 256, 256, 0, 0, True, 0, False),
    ("n1536_ragged_m", 1333, 1536, 512, 0, 0, True, 0, False),
    ("dgrad_fold", 1280, 512, 512, 1, 0, False, 0, True),
]


@pytest.mark.parametrize("case", BRES_CASES, ids=[c[0] for c in BRES_CASES])
def test_gemm_b_stationary(case):
    """B-stationary CTA pairs (TcParams::bres: the pair's 256-column slice of B resident in shared memory, A streamed) forced at
    small M: against fp64 on the same bf16 operands, and bit-identical to the regular schedule."""
    name, M, N, K, b_major, act, has_bias, store, aux = case
    torch.manual_seed(len(name))
    dev = torch.device("cuda")
    ctx = _ctx("bf16")
    a = (torch.randn(M, K) * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K) * 0.5).to(torch.bfloat16) if b_major == 0 else (torch.randn(K, N) * 0.5).to(torch.bfloat16)
    bias = torch.randn(N) if has_bias else None
    y = (torch.randn(M, N)).to(torch.bfloat16) if aux else None
    d0 = (torch.randn(M, N) * 0.5).to(torch.bfloat16)
    bm = b.double() if b_major == 0 else b.double().t()
    ref = a.double() @ bm.t()
    if bias is not None:
        ref = ref + bias.double()
    if act == 1:
        ref = torch.relu(ref)
    if aux:
        ref = ref * torch.where(y.double() > 0, 1.0, 0.25)
    if store == 1:
        ref = ref + d0.double()
    outs = []
    lib = L.lib()
    for mode in (2, 0):
        prev = lib.jmt_gemm_set_bres_mode(mode)
        try:
            d = d0.clone().to(dev)
            E.gemm(ctx, a.to(dev), b.to(dev), d, M=M, N=N, K=K, b_major=b_major, b_rows=N if b_major == 0 else K,
                   bias=bias.to(dev) if bias is not None else None, act=act, store=store,
                   epi_aux=y.to(dev) if aux else None, aux_slope=0.25)
            torch.cuda.synchronize()
        finally:
            lib.jmt_gemm_set_bres_mode(prev)
        outs.append(d.cpu())
    err = (outs[0].double() - ref).abs().max()
    assert err < 1.5e-2 * ref.abs().max(), (name, float(err))
    assert torch.equal(outs[0], outs[1]), name


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("nb", [3, 4])
def test_gemm_colmask(precision, nb):
    """Channel dropout fused into the epilogue: D(m, n) = act(...) * (mask[b, n] ? scale : 0), both kernels,
    1-CTA (odd batch) and CTA-pair-across-batch (even batch) tile schedules."""
    torch.manual_seed(11)
    dev = torch.device("cuda")
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    M, N, K, taps = 70, 160, 64, 3
    x = (torch.randn(nb, M, K) * 0.5).to(dt)
    w = (torch.randn(N, taps * K) * 0.5).to(dt)
    bias = torch.randn(N)
    mask = (torch.rand(nb, N) > 0.4).to(torch.uint8)
    ref = _ref_gemm(x.float(), w.float()[None], a_major=0, b_major=0, M=M, N=N, K=K, ntaps=taps, a_shift=(-4, 2), b_shift=(0, 0),
                    nb=nb, reduce_batch=False, alpha=1.0, bias=bias, act=L.ACT_LEAKY, slope=0.01)
    ref = ref * (mask.double() * 1.6)[:, None, :]
    xd, wd, md = x.to(dev), w.to(dev), mask.to(dev)
    d = torch.empty(nb, M, N, device=dev, dtype=dt)
    E.gemm(_ctx(precision), xd, wd, d, M=M, N=N, K=K, a_rows=M, b_rows=N, a_ld=K, b_ld=taps * K, d_ld=N, nb0=1, nb1=nb,
           a_bs=(0, M * K), d_bs=(0, M * N), bias=bias.to(dev), act=L.ACT_LEAKY, slope=0.01, ntaps=taps, a_shift=(-4, 2),
           colmask=md, colmask_scale=1.6)
    torch.cuda.synchronize()
    tol = {"fp32": 2e-4, "bf16x3": 6e-5, "bf16": 1.2e-2}[precision]
    assert (d.float().cpu().double() - ref).abs().max() < tol * (ref.abs().max() + 1e-6)
    assert (d.float().cpu()[mask[:, None, :].expand(nb, M, N) == 0] == 0).all()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("geom", [(5, 130, 8, 64, 96, 3, 4), (6, 130, 8, 192, 512, 3, 4), (146, 130, 8, 192, 512, 3, 4)],
                         ids=["narrow", "wide512", "wide512_tail"])
def test_gemm_flat_tcn_layout(precision, geom):
    """Flat padded TCN layout: one GEMM over N*(pad+L) rows with row shifts, padding rows written as zeros
    (zero_row_period) and the channel-dropout mask row taken from m / period -- against per-sequence causal convs."""
    torch.manual_seed(5)
    dev = torch.device("cuda")
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    Nn, Ls, pad, cin, cout, taps, dil = geom      # wide512: 256 x 512 pair tiles with the channel-dropout keep-flag words
    Lp = Ls + pad
    x = torch.zeros(Nn, Lp, cin)
    x[:, pad:] = (torch.randn(Nn, Ls, cin) * 0.5).to(dt).float()
    w = (torch.randn(cout, taps * cin) * 0.5).to(dt).float()
    bias = torch.randn(cout)
    mask = (torch.rand(Nn, cout) > 0.4).to(torch.uint8)
    ref = torch.zeros(Nn, Lp, cout, dtype=torch.float64)
    xs = x[:, pad:].double()                                           # (Nn, Ls, cin)
    acc = bias.double().expand(Nn, Ls, cout).clone()
    for j in range(taps):
        sh = (taps - 1 - j) * dil                                      # out[t] += W_j x[t - sh], zeros for t < sh
        if sh < Ls:
            acc[:, sh:] += xs[:, :Ls - sh] @ w[:, j * cin:(j + 1) * cin].double().t()
    acc = torch.where(acc >= 0, acc, acc * 0.01)
    ref[:, pad:] = acc * (mask.double() * 1.5)[:, None, :]
    xd, wd, md = x.reshape(Nn * Lp, cin).to(dev, dt), w.to(dev, dt), mask.to(dev)
    d = torch.full((Nn * Lp, cout), float("nan"), device=dev, dtype=dt)
    E.gemm(_ctx(precision), xd, wd, d, M=Nn * Lp, N=cout, K=cin, a_rows=Nn * Lp, b_rows=cout, a_ld=cin, b_ld=taps * cin, d_ld=cout,
           bias=bias.to(dev), act=L.ACT_LEAKY, slope=0.01, ntaps=taps, a_shift=(-(taps - 1) * dil, dil),
           colmask=md, colmask_scale=1.5, colmask_row_period=Lp, zero_rows=(Lp, pad))
    torch.cuda.synchronize()
    got = d.float().cpu().double().reshape(Nn, Lp, cout)
    tol = {"fp32": 2e-4, "bf16x3": 6e-5, "bf16": 1.2e-2}[precision]
    assert (got - ref).abs().max() < tol * (ref.abs().max() + 1e-6), (got - ref).abs().max()
    assert (got[:, :pad] == 0).all(), "padding rows must be written as zeros"
    if precision != "fp32":
        # sequences shorter than a tile would need the keep-flags of more than two samples per tile: refused, not mis-computed
        with pytest.raises(RuntimeError, match="colmask_row_period"):
            E.gemm(_ctx(precision), xd, wd, d, M=Nn * Lp, N=cout, K=cin, a_rows=Nn * Lp, b_rows=cout, a_ld=cin, b_ld=taps * cin,
                   d_ld=cout, ntaps=taps, a_shift=(-(taps - 1) * dil, dil), colmask=md, colmask_scale=1.5, colmask_row_period=64)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("Ls", [7, 200])
def test_causal_conv_channel_dropout_short_and_long_sequences(precision, Ls):
    """Dropout2d of the TCN in training mode (whole channels per sample) through engine.causal_conv: fused into the
    GEMM epilogue for sequences of >= 127 flat rows, a separate mask pass for the reference's 7-frame clips (tsav.py:214-216)
    -- both against the no-dropout convolution times the regenerated Philox keep-mask."""
    torch.manual_seed(2)
    Nn, cin, cout, k, dil, pad, p_drop = 37, 64, 96, 3, 2, 4, 0.4
    Lp = Ls + pad
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    params = {"c.weight_g": (torch.rand(cout, 1, 1) + 0.5).cuda(), "c.weight_v": (torch.randn(cout, cin, k) * 0.2).cuda(),
              "c.bias": (torch.randn(cout) * 0.1).cuda()}
    x = torch.zeros(Nn, Lp, cin)
    x[:, pad:] = torch.randn(Nn, Ls, cin)
    xd = x.reshape(Nn * Lp, cin).to("cuda", dt)
    seed = 1234
    ctx_t = E.Ctx(params, precision, False, True, seed=seed)          # training: channel dropout on
    y_t = E.causal_conv(ctx_t, E.Var(xd), "c.", Nn, Ls, cin, cout, k, dil, L.ACT_LEAKY, drop_p=p_drop, pad=pad).data.float()
    ctx_e = E.Ctx(params, precision, False, False, seed=seed)         # eval: same conv, no dropout
    y_e = E.causal_conv(ctx_e, E.Var(xd), "c.", Nn, Ls, cin, cout, k, dil, L.ACT_LEAKY, drop_p=p_drop, pad=pad).data.float()
    mask = torch.empty(Nn * cout, dtype=torch.uint8, device="cuda")
    L.check(L.lib().jmt_dropout_mask(E._ptr(mask), Nn * cout, p_drop, seed, 0, None, E._stream()), "jmt_dropout_mask")
    torch.cuda.synchronize()
    keep = mask.view(Nn, 1, cout).float()
    assert 0.3 < 1 - keep.mean().item() < 0.5
    want = y_e.view(Nn, Lp, cout) * keep / (1 - p_drop)
    got = y_t.view(Nn, Lp, cout)
    assert (got - want).abs().max() < (1e-5 if precision == "bf16x3" else 2e-2) * want.abs().max()
    assert (got[:, :pad] == 0).all()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("geom", [(5, 77, 8, 64, 96, 3, 4, 1), (6, 300, 32, 128, 512, 5, 8, 3), (3, 50, 16, 256, 128, 5, 2, 2)],
                         ids=["small", "c2like", "wide_in"])
def test_gemm_wgrad_taps_batched(precision, geom):
    """Conv weight gradient on the flat padded TCN layout with all taps in ONE launch: tap j = batch entry j, A = dy shared
    (batch stride 0), B = x at an overlapping batch stride of dil rows under one common shift, D = column block j of dW,
    split-K with fp32 reduce-add -- against per-tap dy^T shift_j(x) in fp64."""
    Nn, Ls, pad, cin, cout, taps, dil, split = geom
    assert pad >= (taps - 1) * dil
    torch.manual_seed(9)
    dev = torch.device("cuda")
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    Lp = Ls + pad
    R = Nn * Lp
    x = torch.zeros(Nn, Lp, cin)
    x[:, pad:] = (torch.randn(Nn, Ls, cin) * 0.5).to(dt).float()
    dy = torch.zeros(Nn, Lp, cout)
    dy[:, pad:] = (torch.randn(Nn, Ls, cout) * 0.5).to(dt).float()
    xf, dyf = x.reshape(R, cin).double(), dy.reshape(R, cout).double()
    ref = torch.zeros(cout, taps * cin, dtype=torch.float64)
    for j in range(taps):
        sh = (taps - 1 - j) * dil                      # dW_j = sum_r dy[r]^T x[r - sh]
        ref[:, j * cin:(j + 1) * cin] = dyf[sh:].t() @ xf[:R - sh] if sh else dyf.t() @ xf
    init = torch.randn(cout, taps * cin) * 0.1
    dw = init.clone().to(dev)
    # NaN rows right behind both operands: the last 64-row k-block must not pick them up (0 * NaN)
    xbuf = torch.full((R + 64, cin), float("nan"), device=dev, dtype=dt)
    dybuf = torch.full((R + 64, cout), float("nan"), device=dev, dtype=dt)
    xbuf[:R], dybuf[:R] = xf.to(dev, dt), dyf.to(dev, dt)
    xd, dyd = xbuf[:R], dybuf[:R]
    E.gemm(_ctx(precision), dyd, xd, dw, M=cout, N=cin, K=R, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, a_rows=R,
           b_rows=R - (taps - 1) * dil, a_ld=cout, b_ld=cin, d_ld=taps * cin, nb1=taps, a_bs=(0, 0), b_bs=(0, dil * cin), d_bs=(0, cin),
           b_shift=(-(taps - 1) * dil, 0), store=L.ATOMIC_ADD, split_k=split)
    torch.cuda.synchronize()
    want = ref + init.double()
    tol = {"fp32": 2e-4, "bf16x3": 2e-4, "bf16": 2e-3}[precision]   # bf16 operands are exact here (inputs pre-rounded); fp32 accumulation
    assert (dw.cpu().double() - want).abs().max() < tol * want.abs().max()


@pytest.mark.parametrize("case", ["wide_fold_colsum", "narrow_colsum_acc", "heads_colsum", "ragged_fold", "n1024_colsum_global_atomics"])
def test_gemm_backward_epilogue_extensions(case):
    """jmt_gemm_desc.epi_aux / d_colsum: the dgrad epilogue multiplies by act'(aux) and adds the column sums of what it stores
    (the producer's bias gradient) -- against the plain GEMM followed by the separate passes."""
    torch.manual_seed(21)
    dev = torch.device("cuda")
    ctx = _ctx("bf16")
    if case == "heads_colsum":
        B, T, E_, h = 5, 150, 256, 4
        dh = E_ // h
        ds = (torch.randn(B, h, T, 152) * 0.3).to(torch.bfloat16).to(dev)
        kk = (torch.randn(B * T, E_) * 0.5).to(torch.bfloat16).to(dev)
        dq = torch.zeros(B * T, E_, dtype=torch.bfloat16, device=dev)
        cs = torch.zeros(E_, dtype=torch.float32, device=dev)
        E.gemm(ctx, ds, kk, dq, M=T, N=dh, K=T, a_rows=T, b_major=L.MAJOR_MN, b_rows=T, a_ld=152, b_ld=E_, d_ld=E_,
               nb0=h, nb1=B, a_bs=(T * 152, h * T * 152), b_bs=(dh, T * E_), d_bs=(dh, T * E_), d_colsum=cs, colsum_bs0=dh)
        torch.cuda.synchronize()
        ref = torch.einsum("bhts,bshd->bthd", ds.float()[..., :T], kk.float().view(B, T, h, dh)).reshape(B * T, E_)
        assert (dq.float() - ref).abs().max() < 2e-2 * ref.abs().max()
        want = ref.double().sum(0)
        assert (cs.double() - want).abs().max() < 2e-3 * want.abs().max() + 1e-3, (cs.double() - want).abs().max()
        return
    M, N, K, store, slope, use_aux = {"wide_fold_colsum": (256 * 78, 512, 512, L.STORE, 0.0, True),
                                      "narrow_colsum_acc": (1000, 384, 192, L.ACCUMULATE, 0.0, False),
                                      "ragged_fold": (333, 200, 136, L.STORE, 0.01, True),
                                      "n1024_colsum_global_atomics": (2100, 1024, 256, L.STORE, 0.0, True)}[case]
    a = (torch.randn(M, K) * 0.5).to(torch.bfloat16).to(dev)
    w = (torch.randn(K, N) * 0.5).to(torch.bfloat16).to(dev)               # B MN-major (dgrad of a Linear)
    aux = torch.randn(M, N).to(torch.bfloat16).to(dev)
    d0 = (torch.randn(M, N) * 0.1).to(torch.bfloat16).to(dev)
    d = d0.clone()
    cs = torch.full((N,), 0.5, dtype=torch.float32, device=dev)             # accumulates on top of what is there
    E.gemm(ctx, a, w, d, M=M, N=N, K=K, b_major=L.MAJOR_MN, store=store, epi_aux=aux if use_aux else None, aux_slope=slope,
           d_colsum=cs)
    torch.cuda.synchronize()
    val = a.double() @ w.double()
    if use_aux:
        val = torch.where(aux.double() > 0, val, val * slope)
    want_d = val if store == L.STORE else val + d0.double()
    assert (d.double() - want_d).abs().max() < 1.2e-2 * want_d.abs().max()
    want_cs = val.sum(0) + 0.5
    assert (cs.double() - want_cs).abs().max() < 2e-3 * want_cs.abs().max() + 1e-2, (cs.double() - want_cs).abs().max()


def test_gemm_heads_geometry():
    """(b, head) batching through two batch dims with non-monotonic strides (Q of shape (B*T, 3E))."""
    torch.manual_seed(3)
    dev = torch.device("cuda")
    B, T, E_, h = 3, 70, 128, 4
    dh = E_ // h
    for precision, dt, tol in [("fp32", torch.float32, 1e-4), ("bf16x3", torch.float32, 6e-5), ("bf16", torch.bfloat16, 2e-2)]:
        qkv = (torch.randn(B * T, 3 * E_) * 0.5).to(dt)
        q, k = qkv[:, :E_].float(), qkv[:, E_:2 * E_].float()
        ref = torch.einsum("bthd,bshd->bhts", q.view(B, T, h, dh), k.view(B, T, h, dh)).double()
        s_ld = (T + 7) // 8 * 8
        S = torch.zeros(B, h, T, s_ld, device=dev)
        qd = qkv.to(dev)
        E.gemm(_ctx(precision), qd[:, :E_], qd[:, E_:2 * E_], S, M=T, N=T, K=dh, a_rows=T, b_rows=T, a_ld=3 * E_,
               b_ld=3 * E_, d_ld=s_ld, nb0=h, nb1=B, a_bs=(dh, T * 3 * E_), b_bs=(dh, T * 3 * E_),
               d_bs=(T * s_ld, h * T * s_ld))
        torch.cuda.synchronize()
        got = S.cpu()[..., :T].double()
        assert (got - ref).abs().max() < tol * ref.abs().max(), precision


ATTN_CASES = [
    # name, NB, heads, Lq, S, E, cross (separate kv tensor), batch_major (attention across the batch dim)
    ("t300_h1", 3, 1, 300, 300, 512, False, False),
    ("t70_s40_h4", 2, 4, 70, 40, 512, True, False),
    ("t9_h2", 2, 2, 9, 9, 512, False, False),
    ("t130_s77_h8", 2, 8, 130, 77, 512, True, False),
    ("t260_h2_e256", 2, 2, 260, 260, 256, False, False),
    ("batchdim_h2", 5, 2, 37, 37, 512, False, True),
    ("t300_s123_h1_x", 3, 1, 300, 123, 512, True, False),
    ("t200_s300_h2_x", 2, 2, 200, 300, 512, True, False),
    ("t300_h1_nb80", 80, 1, 300, 300, 512, False, False),      # 480 pair tiles of the dQ/dK/dV kernel: several per CTA pair, ring wrap-around
]


@pytest.mark.parametrize("fused", [("full", "full"), ("full", "ds"), ("full", "ds", "gemms"), ("p", "ds"), (False, False)],
                         ids=["full-full", "full-ds", "full-ds-gemms", "p-ds", "composed"])
@pytest.mark.parametrize("case", ATTN_CASES, ids=[c[0] for c in ATTN_CASES])
def test_attention_core_bf16(case, fused):
    """softmax(QK^T/sqrt(dh))V forward + dQ/dK/dV through engine.attention_core in bf16 mode: the fused tcgen05
    kernel (jmt_attn_chain_bf16) and the GEMM + softmax composition against an fp64 reference on the same
    bf16-rounded operands."""
    name, NB, h, Lq, S, E_, cross, batch_major = case
    torch.manual_seed(len(name) * 7 + NB)
    dev = torch.device("cuda")
    dh = E_ // h
    ctx = _ctx("bf16")
    ctx.record = True
    old = (E.FUSED_ATTENTION, E.FUSED_ATTENTION_BWD, E.ATTN_BWD_DQKV)
    E.FUSED_ATTENTION, E.FUSED_ATTENTION_BWD = (True if fused[0] == "full" else fused[0]), fused[1]
    E.ATTN_BWD_DQKV = len(fused) < 3           # "gemms": dQ / dK / dV by three batched jmt_gemm_bf16 instead of jmt_attn_bwd_dqkv_bf16
    if NB > 8 and fused[:2] != ("full", "ds"):
        pytest.skip("large-batch case only for the default path")
    try:
        # projection matrices as the engine holds them: rows = batch*seq (or seq*batch for the batch-major geometry)
        qw = E_ if cross else 3 * E_          # a cross-attention Q projection is (rows, E); self-attention packs Q | K | V
        qm = (torch.randn(NB * Lq, qw) * 0.7).to(torch.bfloat16)
        km = (torch.randn(NB * S, 2 * E_) * 0.7).to(torch.bfloat16) if cross else None
        if batch_major:       # row = s * NB + n: "sequence" index has stride NB, batch stride 1
            gq = E.AttnGeom(Lq, NB, NB, 1)
            gk = E.AttnGeom(S, NB, NB, 1)
            to4 = lambda m, L_, w: m.float().view(L_, NB, w).permute(1, 0, 2)     # noqa: E731  (NB, L, w)
        else:
            gq = E.AttnGeom(Lq, NB, 1, Lq)
            gk = E.AttnGeom(S, NB, 1, S)
            to4 = lambda m, L_, w: m.float().view(NB, L_, w)                      # noqa: E731
        qv = E.Var(qm.to(dev))
        if cross:
            kv = E.Var(km.to(dev))
            out = E.attention_core(ctx, qv, 0, kv, 0, kv, E_, E_, h, gq, gk)
            q4, k4, v4 = to4(qm, Lq, E_), to4(km, S, 2 * E_)[..., :E_], to4(km, S, 2 * E_)[..., E_:]
        else:
            out = E.attention_core(ctx, qv, 0, qv, E_, qv, 2 * E_, E_, h, gq, gk)
            a4 = to4(qm, Lq, 3 * E_)
            q4, k4, v4 = a4[..., :E_], a4[..., E_:2 * E_], a4[..., 2 * E_:]
        q64 = q4.double().reshape(NB, Lq, h, dh).permute(0, 2, 1, 3).clone().requires_grad_(True)
        k64 = k4.double().reshape(NB, S, h, dh).permute(0, 2, 1, 3).clone().requires_grad_(True)
        v64 = v4.double().reshape(NB, S, h, dh).permute(0, 2, 1, 3).clone().requires_grad_(True)
        ref = torch.softmax(q64 @ k64.transpose(-1, -2) / math.sqrt(dh), -1) @ v64          # (NB, h, Lq, dh)
        do = (torch.randn(NB * Lq, E_) * 0.5).to(torch.bfloat16)
        do4 = to4(do, Lq, E_).double().reshape(NB, Lq, h, dh).permute(0, 2, 1, 3)
        ref.backward(do4)
        torch.cuda.synchronize()
        got = to4(out.data.cpu(), Lq, E_).double().reshape(NB, Lq, h, dh).permute(0, 2, 1, 3)
        assert (got - ref.detach()).abs().max() < 2e-2 * ref.detach().abs().max(), (got - ref.detach()).abs().max()
        out.gbuf = E.GradBuf(do.to(dev))
        ctx.backward()
        torch.cuda.synchronize()

        def back(g4, L_):      # (NB, h, L, dh) -> (NB, L, E)
            return g4.permute(0, 2, 1, 3).reshape(NB, L_, E_)
        gqm = to4(qv.grad.cpu(), Lq, qw).double()
        want_q = back(q64.grad, Lq)
        assert (gqm[..., :E_] - want_q).abs().max() < 3e-2 * want_q.abs().max(), "dQ"
        if cross:
            gkm = to4(kv.grad.cpu(), S, 2 * E_).double()
            gk_got, gv_got = gkm[..., :E_], gkm[..., E_:]
        else:
            gk_got, gv_got = gqm[..., E_:2 * E_], gqm[..., 2 * E_:]
        want_k, want_v = back(k64.grad, S), back(v64.grad, S)
        assert (gk_got - want_k).abs().max() < 3e-2 * want_k.abs().max(), "dK"
        assert (gv_got - want_v).abs().max() < 3e-2 * want_v.abs().max(), "dV"
    finally:
        E.FUSED_ATTENTION, E.FUSED_ATTENTION_BWD, E.ATTN_BWD_DQKV = old


DQKV_CASES = [
    # name, NB, heads, Lq, S, dh
    ("w300_dh512", 5, 1, 300, 300, 512),
    ("w300_dh512_nb160", 160, 1, 300, 300, 512),
    ("lq77_s290_dh256_h2", 3, 2, 77, 290, 256),
    ("lq320_s17_dh512", 4, 1, 320, 17, 512),
    ("lq1_s1_dh256", 2, 1, 1, 1, 256),
    ("lq130_s129_dh256_h4", 2, 4, 130, 129, 256),
]


@pytest.mark.parametrize("case", DQKV_CASES, ids=[c[0] for c in DQKV_CASES])
def test_attn_bwd_dqkv_kernel(case):
    """jmt_attn_bwd_dqkv_bf16 (dQ = dS K, dK = dS^T Q, dV = P^T dO in one launch, transposed CTA-pair tiles) against fp64 einsums on
    the same bf16 operands: STORE, then ACCUMULATE on top, bias-gradient column sums, strided (packed QKV) operand / output
    geometry, padded x_ld with NaN in the pad columns (they must never be read)."""
    name, NB, h, Lq, S, dh = case
    torch.manual_seed(len(name) + NB)
    dev = torch.device("cuda")
    E_ = h * dh
    ctx = _ctx("bf16")
    x_ld = (S + 7) // 8 * 8
    ds = torch.full((NB, h, Lq, x_ld), float("nan")).to(torch.bfloat16)
    pr = torch.full((NB, h, Lq, x_ld), float("nan")).to(torch.bfloat16)
    ds[..., :S] = (torch.randn(NB, h, Lq, S) * 0.3).to(torch.bfloat16)
    pr[..., :S] = torch.rand(NB, h, Lq, S).to(torch.bfloat16)
    qkv = (torch.randn(NB * max(Lq, S), 3 * E_) * 0.5).to(torch.bfloat16)      # packed projections: Q | K | V column slices, ld = 3E
    do = (torch.randn(NB * Lq, E_) * 0.5).to(torch.bfloat16)
    rows = max(Lq, S)
    q4 = qkv.view(NB, rows, 3 * E_)[:, :Lq, :E_].double().reshape(NB, Lq, h, dh)
    k4 = qkv.view(NB, rows, 3 * E_)[:, :S, E_:2 * E_].double().reshape(NB, S, h, dh)
    do4 = do.view(NB, Lq, E_).double().reshape(NB, Lq, h, dh)
    want_q = torch.einsum("bhqs,bshd->bqhd", ds[..., :S].double(), k4).reshape(NB, Lq, E_)
    want_k = torch.einsum("bhqs,bqhd->bshd", ds[..., :S].double(), q4).reshape(NB, S, E_) * 0.5
    want_v = torch.einsum("bhqs,bqhd->bshd", pr[..., :S].double(), do4).reshape(NB, S, E_)
    dsd, prd, qkvd, dod = ds.to(dev), pr.to(dev), qkv.to(dev), do.to(dev)
    g = torch.full((NB * rows, 3 * E_), float("nan"), dtype=torch.bfloat16, device=dev)     # gradient of the packed projection
    cs = torch.zeros(3, E_, dtype=torch.float32, device=dev)
    ld = 3 * E_
    geo = (ld, dh, rows * ld)
    o_geo = (E_, dh, Lq * E_)

    def run(store):
        E._attn_bwd_dqkv(ctx, [
            (qkvd[:, E_:2 * E_], geo, dsd, 0, g[:, :E_], geo, store, 1.0, cs[0]),
            (qkvd[:, :E_], geo, dsd, 1, g[:, E_:2 * E_], geo, store, 0.5, cs[1]),
            (dod, o_geo, prd, 1, g[:, 2 * E_:], geo, store, 1.0, cs[2]),
        ], Lq, S, dh, h, NB, x_ld)
        torch.cuda.synchronize()

    def check(mult):
        g4 = g.cpu().double().view(NB, rows, 3 * E_)
        for nm, got, want in (("dQ", g4[:, :Lq, :E_], want_q), ("dK", g4[:, :S, E_:2 * E_], want_k), ("dV", g4[:, :S, 2 * E_:], want_v)):
            err = (got - mult * want).abs().max()
            assert err < 1.2e-2 * mult * want.abs().max() + 1e-6, (nm, float(err), float(want.abs().max()))
        # rows beyond the valid extent of a part are never written
        if Lq < rows:
            assert torch.isnan(g4[:, Lq:, :E_]).all()
        if S < rows:
            assert torch.isnan(g4[:, S:, E_:]).all()
        for i, want in enumerate((want_q, want_k, want_v)):
            wsum = mult * want.sum(dim=(0, 1))
            err = (cs[i].cpu().double() - wsum).abs().max()
            assert err < 2e-2 * mult * want.abs().sum(dim=(0, 1)).max() / math.sqrt(want.shape[0] * want.shape[1]) + 1e-3, ("colsum", i, float(err))

    assert L.lib().jmt_attn_bwd_dqkv_supported is not None
    run(L.STORE)
    check(1.0)
    run(L.ACCUMULATE)
    check(2.0)


def test_attn_bwd_dqkv_rejects_unsupported_geometry():
    g = L.AttnBwdDesc()
    g.Lq, g.S, g.dh, g.heads, g.NB, g.x_ld = 300, 300, 64, 8, 2, 304
    assert L.lib().jmt_attn_bwd_dqkv_supported(C.byref(g)) == 0
    assert L.lib().jmt_attn_bwd_dqkv_bf16(C.byref(g), None) == -3
    g.dh, g.heads, g.S = 512, 1, 400
    assert L.lib().jmt_attn_bwd_dqkv_supported(C.byref(g)) == 0
    g.S = 300
    assert L.lib().jmt_attn_bwd_dqkv_supported(C.byref(g)) == 1


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("with_mask", [False, True])
def test_add_act_bwd_fused(dt, with_mask):
    """jmt_add_act_bwd_fused (residual LeakyReLU backward + conv2's LeakyReLU / Dropout2d backward + bias gradient in one pass) is
    bit-identical to jmt_act_bwd followed by jmt_act_bwd_fused."""
    torch.manual_seed(11)
    dev = torch.device("cuda")
    lib = L.lib()
    st = E._stream()
    code = L.BF16 if dt == torch.bfloat16 else L.F32
    N, Lp, C = 3, 45, 264
    rows = N * Lp
    dy = torch.randn(rows, C).to(dt).to(dev)
    out = torch.randn(rows, C).to(dt).to(dev)
    a = torch.randn(rows, C).to(dt).to(dev)
    mask = (torch.rand(N, C) > 0.3).to(torch.uint8).to(dev) if with_mask else None
    scale, slope, slope2 = 1.0 / 0.7, 0.01, 0.02
    dz_ref = torch.empty_like(dy); dz2_ref = torch.empty_like(dy); cs_ref = torch.zeros(C, device=dev)
    L.check(lib.jmt_act_bwd(E._ptr(dy), E._ptr(out), E._ptr(dz_ref), dy.numel(), slope, code, st), "act_bwd")
    L.check(lib.jmt_act_bwd_fused(E._ptr(dz_ref), E._ptr(a), E._ptr(mask), E._ptr(dz2_ref), rows, C, Lp, scale, slope2, E._ptr(cs_ref), code, st), "fused")
    dz = torch.empty_like(dy); dz2 = torch.empty_like(dy); cs = torch.zeros(C, device=dev)
    L.check(lib.jmt_add_act_bwd_fused(E._ptr(dy), E._ptr(out), E._ptr(a), E._ptr(mask), E._ptr(dz), E._ptr(dz2), rows, C, Lp, scale, slope,
                                      slope2, E._ptr(cs), code, st), "add_act_bwd_fused")
    torch.cuda.synchronize()
    assert torch.equal(dz, dz_ref) and torch.equal(dz2, dz2_ref)
    assert (cs - cs_ref).abs().max() <= 1e-4 * (cs_ref.abs().max() + 1)          # (different summation order of the fp32 column sums)


def test_l2norm_on_padded_tcn_layout():
    """jmt_l2norm_fwd_seq / jmt_l2norm_bwd_seq (F.normalize reading from / differentiating into the TCN's flat padded layout) against
    the compact kernels on the un-padded copy: bit-identical values, zero padding rows in dx, nothing written elsewhere."""
    torch.manual_seed(3)
    dev = torch.device("cuda")
    lib = L.lib()
    st = E._stream()
    N, Ls, pad, D = 5, 37, 32, 512
    Lp = Ls + pad
    xp = torch.randn(N, Lp, D).to(torch.bfloat16)
    xp[:, :pad] = 0
    xc = xp[:, pad:].reshape(N * Ls, D).contiguous()
    dy = torch.randn(N * Ls, D).to(torch.bfloat16)
    xp_d, xc_d, dy_d = xp.to(dev), xc.to(dev), dy.to(dev)
    out = [torch.empty(N * Ls, D, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    inv = [torch.empty(N * Ls, device=dev) for _ in range(2)]
    L.check(lib.jmt_l2norm_fwd_seq(E._ptr(xp_d), L.BF16, D, E._ptr(out[0]), L.BF16, N, Ls, Lp, pad, D, 1e-12, E._ptr(inv[0]), st), "fwd_seq")
    L.check(lib.jmt_l2norm_fwd(E._ptr(xc_d), L.BF16, D, E._ptr(out[1]), L.BF16, N * Ls, D, 1e-12, E._ptr(inv[1]), st), "fwd")
    torch.cuda.synchronize()
    assert torch.equal(out[0], out[1]) and torch.equal(inv[0], inv[1])
    dxp = torch.full((N, Lp, D), 9.0, device=dev, dtype=torch.bfloat16)
    dxc = torch.empty(N * Ls, D, device=dev, dtype=torch.bfloat16)
    L.check(lib.jmt_l2norm_bwd_seq(E._ptr(dy_d), E._ptr(out[0]), L.BF16, E._ptr(inv[0]), 1e-12, E._ptr(dxp), L.BF16, N, Ls, Lp, pad, D, st), "bwd_seq")
    L.check(lib.jmt_l2norm_bwd(E._ptr(dy_d), E._ptr(out[1]), L.BF16, E._ptr(inv[1]), 1e-12, E._ptr(dxc), L.BF16, N * Ls, D, st), "bwd")
    torch.cuda.synchronize()
    assert torch.equal(dxp[:, pad:].reshape(N * Ls, D), dxc)
    assert (dxp[:, :pad] == 0).all()


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layernorm_l2norm_softmax(dt):
    torch.manual_seed(0)
    dev = torch.device("cuda")
    lib = L.lib()
    code = E._DT[dt]
    rows, D = 333, 512
    st = E._stream()
    x = torch.randn(rows, D).to(dt)
    r = torch.randn(rows, D).to(dt)
    g, b = 1 + 0.1 * torch.randn(D), 0.1 * torch.randn(D)
    xd, rd, gd, bd = x.to(dev), r.to(dev), g.to(dev), b.to(dev)
    y = torch.empty(rows, D, device=dev, dtype=dt)
    mean = torch.empty(rows, device=dev)
    rstd = torch.empty(rows, device=dev)
    L.check(lib.jmt_add_layernorm_fwd(E._ptr(xd), E._ptr(rd), E._ptr(gd), E._ptr(bd), 1e-5, E._ptr(y), E._ptr(mean), E._ptr(rstd), rows, D, code, st), "ln")
    z = (x.float() + r.float()).double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(z, (D,), g.double(), b.double(), 1e-5)
    tol = 1e-5 if dt == torch.float32 else 2e-2
    assert (y.float().cpu().double() - ref).abs().max() < tol * 4
    dy = torch.randn(rows, D).to(dt)
    gg = g.double().requires_grad_(True)
    bb = b.double().requires_grad_(True)
    ref2 = torch.nn.functional.layer_norm(z, (D,), gg, bb, 1e-5)
    ref2.backward(dy.double())
    dz = torch.empty(rows, D, device=dev, dtype=dt)
    dgam = torch.zeros(D, device=dev)
    dbet = torch.zeros(D, device=dev)
    dzs = torch.zeros(D, device=dev)
    dy_d = dy.to(dev)
    L.check(lib.jmt_add_layernorm_bwd(E._ptr(dy_d), E._ptr(xd), E._ptr(rd), E._ptr(gd), E._ptr(mean), E._ptr(rstd), E._ptr(dz), 0,
                                      E._ptr(dgam), E._ptr(dbet), E._ptr(dzs), rows, D, code, st), "lnb")
    assert (dz.float().cpu().double() - z.grad).abs().max() < tol * 8
    assert (dzs.cpu().double() - z.grad.sum(0)).abs().max() < tol * 40       # fused bias gradient of the residual Linear
    assert (dgam.cpu().double() - gg.grad).abs().max() < tol * 40
    assert (dbet.cpu().double() - bb.grad).abs().max() < tol * 40
    # l2norm
    x768 = torch.randn(rows, 768)
    x768[5] = 0.0                     # zero row: clamp path
    out = torch.empty(rows, 768, device=dev, dtype=dt)
    inv = torch.empty(rows, device=dev)
    x768_d = x768.to(dev)
    L.check(lib.jmt_l2norm_fwd(E._ptr(x768_d), L.F32, 768, E._ptr(out), code, rows, 768, 1e-12, E._ptr(inv), st), "l2")
    xr = x768.double().requires_grad_(True)
    refn = torch.nn.functional.normalize(xr, dim=-1)
    assert (out.float().cpu().double() - refn).abs().max() < tol
    dyn = torch.randn(rows, 768).to(dt)
    refn.backward(dyn.double())
    dx = torch.empty(rows, 768, device=dev)
    dyn_d = dyn.to(dev)
    L.check(lib.jmt_l2norm_bwd(E._ptr(dyn_d), E._ptr(out), code, E._ptr(inv), 1e-12, E._ptr(dx), L.F32, rows, 768, st), "l2b")
    m = torch.ones(rows, dtype=torch.bool)
    m[5] = False
    assert (dx.cpu().double()[m] - xr.grad[m]).abs().max() < tol * 4
    # softmax fwd / bwd
    R, Ccols, ld = 257, 300, 304
    s = torch.randn(R, ld) * 3
    p = torch.full((R, ld), 7.0, device=dev, dtype=dt)
    s_d = s.to(dev)
    L.check(lib.jmt_softmax_fwd(E._ptr(s_d), ld, E._ptr(p), code, ld, R, Ccols, st), "sm")
    refp = torch.softmax(s[:, :Ccols].double(), -1)
    assert (p.float().cpu()[:, :Ccols].double() - refp).abs().max() < tol
    assert (p.float().cpu()[:, Ccols:] == 0).all()
    dp = torch.randn(R, ld)
    ds = torch.empty(R, ld, device=dev, dtype=dt)
    dp_d = dp.to(dev)
    L.check(lib.jmt_softmax_bwd(E._ptr(p), code, ld, E._ptr(dp_d), ld, E._ptr(ds), code, ld, R, Ccols, st), "smb")
    pp = p.float().cpu()[:, :Ccols].double()
    refds = pp * (dp[:, :Ccols].double() - (pp * dp[:, :Ccols].double()).sum(-1, keepdim=True))
    assert (ds.float().cpu()[:, :Ccols].double() - refds).abs().max() < tol * 2
    torch.cuda.synchronize()


def test_ccc_kernels_vs_closed_form():
    from oracle import jmt_oracle as O
    torch.manual_seed(1)
    dev = torch.device("cuda")
    n = 76800
    x = torch.randn(2, n) * 0.3
    y = (0.6 * x + 0.3 * torch.randn(2, n)).clamp(-1, 1)
    y[0, ::17] = -5.0
    xd, yd = x.to(dev), y.to(dev)
    from jmt_b200.losses import six_sums
    s_all = six_sums(xd, yd, None).cpu().numpy()
    s_msk = six_sums(xd, yd, -5.0).cpu().numpy()
    for i in range(2):
        np.testing.assert_allclose(s_all[i], O.six_sums(x[i].numpy(), y[i].numpy()), rtol=1e-12)
        np.testing.assert_allclose(s_msk[i], O.six_sums(x[i].numpy(), y[i].numpy(), ignore=-5.0), rtol=1e-12)
    # the three formulas + gradients against the oracle (autograd)
    crit = jmt_b200.CCCLoss(digitize_num=1)
    xg = xd[0].clone().requires_grad_(True)
    loss = crit(xg.view(1, -1), yd[0].view(1, -1))
    loss.backward()
    xo = x[0].clone().double().requires_grad_(True)
    lo = O.ccc_loss_live(xo, y[0].double())
    lo.backward()
    assert abs(loss.item() - lo.item()) < 1e-6
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xo.grad.numpy(), rtol=1e-4, atol=1e-10)
    critm = jmt_b200.CCCLossMasked()
    xg2 = xd[0].clone().requires_grad_(True)
    lm = critm(xg2, yd[0])
    lm.backward()
    xo2 = x[0].clone().double().requires_grad_(True)
    lmo = O.ccc_loss_masked(xo2, y[0].double())
    lmo.backward()
    assert abs(lm.item() - lmo.item()) < 1e-6
    np.testing.assert_allclose(xg2.grad.cpu().numpy(), xo2.grad.numpy(), rtol=1e-4, atol=1e-12)
    assert abs(jmt_b200.cccmetric.ccc(xd[1], yd[1]) - O.ccc_metric(x[1].double().numpy(), y[1].double().numpy())) < 1e-6
    assert abs(jmt_b200.cccmetric.ccc_numpy(yd[1], xd[1]) - O.ccc_numpy(y[1].numpy(), x[1].numpy())) < 1e-6
    # masks bit-exact
    from jmt_b200.losses import label_mask
    assert torch.equal(label_mask(yd).cpu(), O.label_mask(y))
    # <= 1 valid element -> loss 0, grad 0
    yp = torch.tensor([0.1, 0.2, 0.3], device=dev, requires_grad=True)
    l0 = critm(yp, torch.tensor([-5.0, 0.5, -5.0], device=dev))
    l0.backward()
    assert l0.item() == 0.0 and (yp.grad == 0).all()


def test_ccc_golden(golden_meta, golden_dir):
    import os
    c = golden_meta["ccc"]
    g = np.load(os.path.join(golden_dir, "ccc.npz"))
    dev = torch.device("cuda")
    x, y, ym = (torch.tensor(g[k], device=dev) for k in ("x", "y", "ym"))
    assert abs(jmt_b200.cccmetric.ccc(x, y) - c["metric"]) < 1e-6
    assert abs(jmt_b200.cccmetric.ccc_numpy(y, x) - c["ccc_numpy"]) < 1e-6
    xg = x.clone().requires_grad_(True)
    l = jmt_b200.CCCLoss(1)(xg[None], y[None])
    l.backward()
    assert abs(l.item() - c["loss_live"]) < 1e-6
    np.testing.assert_allclose(xg.grad.cpu().numpy(), g["g_live"], rtol=2e-3, atol=1e-8)
    xg2 = x.clone().requires_grad_(True)
    l2 = jmt_b200.CCCLossMasked()(xg2, ym)
    l2.backward()
    assert abs(l2.item() - c["loss_masked"]) < 1e-6
    np.testing.assert_allclose(xg2.grad.cpu().numpy(), g["g_masked"], rtol=2e-3, atol=1e-9)
    cv, ca, cm = jmt_b200.cccmetric.cccva(torch.stack([y, ym], 1), torch.stack([x, x * 0.5], 1))
    np.testing.assert_allclose([cv, ca, cm], c["cccva"], rtol=1e-5)
    with pytest.raises(ValueError):
        jmt_b200.cccmetric.ccc(x[:1], y[:1])
    m = jmt_b200.cccmetric.CCCMetric()
    m.update(x[:400], y[:400])
    m.update(x[400:], y[400:])
    assert abs(m.get() - c["metric"]) < 1e-6


def test_padseq_golden(golden_meta, golden_dir):
    import os
    m = golden_meta["padseq"]
    g = np.load(os.path.join(golden_dir, "padseq.npz"))
    gen = torch.Generator().manual_seed(m["seed"])
    specs = []
    for w in m["widths"]:
        torch.randn(16, 3, 2, 4, 4, generator=gen)
        specs.append((torch.randn(16, 1, 64, w, generator=gen) + 3.0).cuda())
        torch.randn(16, generator=gen), torch.randn(16, generator=gen)
    out = jmt_b200.padseq.pad_spectrograms(specs).cpu()
    assert np.array_equal((out == 0).numpy(), g["zero_mask"])
    np.testing.assert_allclose(out.double().sum().numpy(), g["audio_sum"], rtol=1e-12)


def _pad_restated(specs):
    """padSequence.py:9-24 zero-fill written directly in torch (right-aligned copy)."""
    mw = max(s.shape[3] for s in specs)
    out = torch.zeros(len(specs), 16, 1, 64, mw)
    for i, s in enumerate(specs):
        out[i, :, :, :, mw - s.shape[3]:] = s.cpu()
    return out


def test_val_and_test_padsequence_tuples():
    gen = torch.Generator().manual_seed(77)
    widths = [104, 97, 104, 88]
    samples = []
    for i, w in enumerate(widths):
        clip = torch.randn(16, 3, 2, 4, 4, generator=gen).cuda()
        spec = torch.randn(16, 1, 64, w, generator=gen).cuda()
        samples.append((clip, spec, list(range(i, i + 16)), f"vid{i}", 100 + i,
                        torch.randn(16, generator=gen).cuda(), torch.randn(16, generator=gen).cuda(), f"w{i}.wav"))
    want_audio = _pad_restated([s[1] for s in samples])
    vis, aud, fids, vids, vlens, lv, la, wav = jmt_b200.padseq.ValPadSequence()(samples)
    assert torch.equal(aud.cpu(), want_audio)
    assert torch.equal(vis, torch.stack([s[0] for s in samples]))
    assert fids == [s[2] for s in samples] and vids == [s[3] for s in samples] and vlens == [s[4] for s in samples]
    assert torch.equal(lv, torch.stack([s[5] for s in samples])) and torch.equal(la, torch.stack([s[6] for s in samples]))
    assert wav == [s[7] for s in samples]
    tsamples = [(s[0], s[1], s[2], s[3], s[4], s[7]) for s in samples]
    out = jmt_b200.padseq.TestPadSequence()(tsamples)
    assert len(out) == 6 and torch.equal(out[1].cpu(), want_audio) and out[5] == wav and out[4] == vlens


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("misalign", [0, 1])
def test_regressor_tail_bwd_kernel(dt, misalign):
    """Backward of the fused Linear(128, 1) tails (two_transformers.py:104-114): d(hidden) through the ReLU mask, dW, db for
    two groups sharing one hidden buffer (accumulate) and a third with its own; vector path and (misaligned rows) scalar path."""
    import ctypes as C
    torch.manual_seed(11)
    dev = torch.device("cuda")
    lib, st = L.lib(), E._stream()
    B, T, ld = 5, 37, 264
    M = B * T
    code = L.BF16 if dt == torch.bfloat16 else L.F32
    hbuf = [torch.relu(torch.randn(M * ld + 8)).to(dt).to(dev) for _ in range(2)]
    hs = [hbuf[0][misalign:misalign + M * ld].view(M, ld), hbuf[1][misalign:misalign + M * ld].view(M, ld)]
    groups = [0, 0, 1]                                   # groups 0 and 1 read the same hidden rows (columns 0..127)
    ws = [torch.randn(128, device=dev) for _ in groups]
    douts = [torch.randn(T, B, device=dev) for _ in groups]          # (T, B) output layout: row m = b*T + t -> [t, b]
    dhbuf = [torch.full((M * ld + 8,), 7.0, device=dev).to(dt) for _ in range(2)]
    dhs = [dhbuf[i][misalign:misalign + M * ld].view(M, ld) for i in range(2)]
    dws = [torch.zeros(128, device=dev) for _ in groups]
    dbs = [torch.zeros(1, device=dev) for _ in groups]
    scale = [1.0, 2.0, 1.25]
    arr = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    L.check(lib.jmt_regressor_tail_bwd(3, arr([hs[g] for g in groups]), ld, code, arr(ws), arr(douts),
                                       arr([dhs[g] for g in groups]), (C.c_int * 3)(0, 1, 0), (C.c_float * 3)(*scale),
                                       arr(dws), arr(dbs), M, T, 1, B, st), "tail_bwd")
    torch.cuda.synchronize()
    tol = 2e-2 if dt == torch.bfloat16 else 1e-5
    want_dh = [torch.zeros(M, 128, dtype=torch.float64), torch.zeros(M, 128, dtype=torch.float64)]
    for g, hg in enumerate(groups):
        h = hs[hg][:, :128].double().cpu()
        go = douts[g].t().reshape(M).double().cpu()              # row m = b*T + t
        want_dh[hg] += (h > 0) * go[:, None] * ws[g].double().cpu()[None, :] * scale[g]
        assert _close(dws[g].cpu(), (go[:, None] * h).sum(0), 1e-4)
        assert _close(dbs[g].cpu(), go.sum()[None], 1e-4)
    for hg in range(2):
        assert _close(dhs[hg][:, :128].cpu(), want_dh[hg], tol)
        assert torch.all(dhs[hg][:, 128:] == 7.0)                    # columns past the 128 hidden units are not touched


def _close(got, want, tol):
    got, want = got.double(), want.double()
    return float((got - want).abs().max()) <= tol * max(1.0, float(want.abs().max()))


def test_small_kernels():
    torch.manual_seed(5)
    dev = torch.device("cuda")
    lib = L.lib()
    st = E._stream()
    # transpose with cast
    x = torch.randn(3, 37, 70)
    o = torch.empty(3, 70, 37, device=dev, dtype=torch.bfloat16)
    x_d = x.to(dev)
    L.check(lib.jmt_transpose(E._ptr(x_d), L.F32, E._ptr(o), L.BF16, 3, 37, 70, st), "tr")
    assert torch.equal(o.cpu(), x.transpose(1, 2).to(torch.bfloat16))
    # bf16 -> bf16 fast path (8-byte loads, 16-byte stores): the TCN input geometry (1024 channels x 300 frames) with a padded
    # output batch stride, a ragged tile in both directions, and canaries around every output sequence
    for (nb, R, Cc, pad) in ((3, 1024, 300, 32), (2, 72, 44, 0), (1, 8, 4, 5)):
        xb = torch.randn(nb, R, Cc).to(torch.bfloat16)
        ob = torch.full((nb, pad + Cc, R), 7.0, device=dev, dtype=torch.bfloat16)
        xb_d = xb.to(dev)
        L.check(lib.jmt_transpose_strided(E._ptr(xb_d), L.BF16, R * Cc, E._ptr(ob[:, pad:]), L.BF16, (pad + Cc) * R, nb, R, Cc, st), "trb")
        torch.cuda.synchronize()
        assert torch.equal(ob[:, pad:].cpu(), xb.transpose(1, 2)), (nb, R, Cc)
        assert (ob[:, :pad].cpu() == 7.0).all()
    # colsum
    a = torch.randn(1000, 130).to(torch.bfloat16)
    out = torch.zeros(130, device=dev)
    a_d = a.to(dev)
    L.check(lib.jmt_colsum(E._ptr(a_d), L.BF16, 130, 1000, 130, E._ptr(out), st), "cs")
    assert (out.cpu() - a.float().sum(0)).abs().max() < 1e-3
    # vectorised path: 4x unrolled body + tail, on a strided column slice (ld > cols), accumulating into out
    for rows in (5003, 9, 76800):
        a = torch.randn(rows, 1536).to(torch.bfloat16)
        a_d = a.to(dev)
        out = torch.ones(512, device=dev)
        L.check(lib.jmt_colsum(E._ptr(a_d[:, 512:1024]), L.BF16, 1536, rows, 512, E._ptr(out), st), "cs")
        want = 1.0 + a[:, 512:1024].double().sum(0)
        assert (out.cpu().double() - want).abs().max() < 2e-3 * max(1.0, rows ** 0.5 / 30)
    # fused activation-gradient + channel-dropout replay + bias gradient
    for dt, tol in ((torch.float32, 1e-4), (torch.bfloat16, 2e-2)):
        nb, Lr, Cc = 3, 37, 136
        dyf = torch.randn(nb * Lr, Cc).to(dt)
        yf = torch.randn(nb * Lr, Cc).to(dt)
        mk = (torch.rand(nb, Cc) > 0.3).to(torch.uint8)
        want = dyf.float() * mk.float().repeat_interleave(Lr, 0) * 1.25 * torch.where(yf.float() > 0, 1.0, 0.01)
        dy_d, y_d, mk_d = dyf.to(dev), yf.to(dev), mk.to(dev)
        dx = torch.empty_like(dy_d)
        cs = torch.zeros(Cc, device=dev)
        L.check(lib.jmt_act_bwd_fused(E._ptr(dy_d), E._ptr(y_d), E._ptr(mk_d), E._ptr(dx), nb * Lr, Cc, Lr, 1.25, 0.01,
                                      E._ptr(cs), L.F32 if dt == torch.float32 else L.BF16, st), "abf")
        assert (dx.float().cpu() - want).abs().max() < tol
        assert (cs.cpu() - want.sum(0)).abs().max() < 2e-3       # column sums are taken before the bf16 rounding of dx
        dx2 = torch.empty_like(dy_d)
        L.check(lib.jmt_act_bwd_fused(E._ptr(dy_d), E._ptr(y_d), None, E._ptr(dx2), nb * Lr, Cc, 1, 1.0, 0.0, None,
                                      L.F32 if dt == torch.float32 else L.BF16, st), "abf")
        assert (dx2.float().cpu() - dyf.float() * (yf.float() > 0)).abs().max() < tol
    # weight norm fwd/bwd
    cout, cin, k = 24, 40, 5
    g = torch.rand(cout, 1, 1) + 0.5
    v = torch.randn(cout, cin, k)
    wf = torch.empty(cout, k * cin, device=dev)
    wd = torch.empty(cin, k * cout, device=dev)
    nrm = torch.empty(cout, device=dev)
    g_d, v_d = g.to(dev), v.to(dev)      # keep device operands alive across the async launches
    L.check(lib.jmt_weight_norm_fwd(E._ptr(g_d), E._ptr(v_d), E._ptr(wf), E._ptr(wd), L.F32, E._ptr(nrm), cout, cin, k, st), "wn")
    gg = g.double().requires_grad_(True)
    vv = v.double().requires_grad_(True)
    w = vv * (gg / vv.pow(2).sum((1, 2), keepdim=True).sqrt())
    assert (wf.cpu().double() - w.permute(0, 2, 1).reshape(cout, -1)).abs().max() < 1e-5
    assert (wd.cpu().double() - w.permute(1, 2, 0).reshape(cin, -1)).abs().max() < 1e-5
    dw = torch.randn(cout, cin, k)
    w.backward(dw.double())
    dwf = dw.permute(0, 2, 1).reshape(cout, -1).contiguous().to(dev)
    dg = torch.empty(cout, 1, 1, device=dev)
    dv = torch.empty(cout, cin, k, device=dev)
    L.check(lib.jmt_weight_norm_bwd(E._ptr(dwf), E._ptr(g_d), E._ptr(v_d), E._ptr(nrm), E._ptr(dg), E._ptr(dv), cout, cin, k, st), "wnb")
    assert (dg.cpu().double() - gg.grad).abs().max() < 1e-4
    assert (dv.cpu().double() - vv.grad).abs().max() < 1e-4
    # dropout mask statistics + determinism
    n = 1 << 20
    m1 = torch.empty(n, dtype=torch.uint8, device=dev)
    m2 = torch.empty(n, dtype=torch.uint8, device=dev)
    L.check(lib.jmt_dropout_mask(E._ptr(m1), n, 0.3, 1234, 0, None, st), "dm")
    L.check(lib.jmt_dropout_mask(E._ptr(m2), n, 0.3, 1234, 0, None, st), "dm")
    assert torch.equal(m1, m2) and abs(m1.float().mean().item() - 0.7) < 3e-3
    # device-resident RNG state: same host arguments, fresh mask after jmt_rng_advance (what a graph replay relies on)
    state = torch.zeros(2, dtype=torch.int64, device=dev)
    L.check(lib.jmt_dropout_mask(E._ptr(m2), n, 0.3, 1234, 0, E._ptr(state), st), "dm")
    assert torch.equal(m1, m2)
    L.check(lib.jmt_rng_advance(E._ptr(state), (n + 3) // 4, st), "adv")
    L.check(lib.jmt_dropout_mask(E._ptr(m2), n, 0.3, 1234, 0, E._ptr(state), st), "dm")
    assert not torch.equal(m1, m2) and abs(m2.float().mean().item() - 0.7) < 3e-3
    # tiny attention fwd/bwd vs torch
    Ls, N, Em, h = 6, 50, 64, 2
    qkv = torch.randn(Ls, N, 3 * Em, requires_grad=True, dtype=torch.float64)
    q, kk, vv2 = qkv[..., :Em], qkv[..., Em:2 * Em], qkv[..., 2 * Em:]
    dh = Em // h
    sh = lambda t: t.reshape(Ls, N, h, dh).permute(1, 2, 0, 3)  # noqa: E731
    att = torch.softmax(sh(q) @ sh(kk).transpose(-1, -2) / math.sqrt(dh), -1) @ sh(vv2)
    ref = att.permute(2, 0, 1, 3).reshape(Ls, N, Em)
    do = torch.randn(Ls, N, Em, dtype=torch.float64)
    ref.backward(do)
    qd = qkv.detach().float().to(dev)
    out = torch.empty(Ls, N, Em, device=dev)
    probs = torch.empty(N * h, Ls, Ls, device=dev)
    L.check(lib.jmt_attn_small_fwd(E._ptr(qd), E._ptr(out), E._ptr(probs), Ls, N, Em, h, 1 / math.sqrt(dh), L.F32, st), "as")
    assert (out.cpu().double() - ref.detach()).abs().max() < 1e-4
    dq = torch.empty(Ls, N, 3 * Em, device=dev)
    do_d = do.float().to(dev)
    L.check(lib.jmt_attn_small_bwd(E._ptr(qd), E._ptr(do_d), E._ptr(probs), E._ptr(dq), Ls, N, Em, h, 1 / math.sqrt(dh), L.F32, st), "asb")
    assert (dq.cpu().double() - qkv.grad).abs().max() < 1e-4
    torch.cuda.synchronize()
