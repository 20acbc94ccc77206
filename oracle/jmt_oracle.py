"""CPU oracle for the JMT fusion + TCN + CCC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and there only as the checker / reported baseline.

This is a restatement of the reference algorithm in explicit tensor algebra
(``matmul`` / ``softmax`` / ``mean`` ... on CPU torch tensors, fp32 or fp64) -- it
does NOT call ``nn.MultiheadAttention``, ``nn.LayerNorm``, ``nn.Conv1d`` or
``weight_norm``; every contraction is spelled out so the oracle documents the
arithmetic the CUDA kernels must reproduce.  Autograd on these plain ops gives the
gradient oracle.

Parity pin: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by ``tests/golden/make_golden.py``
(imports ``/root/reference`` modules, runs them under the installed torch) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every
function here against those fixtures.

All ``file:line`` citations are relative to the reference repository root.
Parameters are passed as a flat ``dict`` keyed with the reference ``state_dict``
names so that the same dict loads into the reference modules, the oracle and the
CUDA-backed modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

LN_EPS = 1e-5          # nn.LayerNorm default (mm_multi_transformers.py:57-58)
NORMALIZE_EPS = 1e-12  # F.normalize default (two_transformers.py:118-119)
LEAKY_SLOPE = 0.01     # nn.LeakyReLU() default (temporal_convolutional_model.py:28,35,42)
IGNORE_LABEL = -5.0    # label sentinel (losses/CCCLoss.py:8, dataset_new.py:255-256)


# --------------------------------------------------------------------------- #
# primitives
# --------------------------------------------------------------------------- #
def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """nn.Linear: y = x W^T + b."""
    y = torch.matmul(x, w.t())
    return y if b is None else y + b


def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = LN_EPS) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc / torch.sqrt(var + eps) * g + b


def l2_normalize(x: Tensor, eps: float = NORMALIZE_EPS) -> Tensor:
    """F.normalize(x, dim=-1) (two_transformers.py:118-119)."""
    n = torch.sqrt((x * x).sum(dim=-1, keepdim=True))
    return x / torch.clamp(n, min=eps)


def mha(query: Tensor, key: Tensor, value: Tensor, p: Params, prefix: str,
        num_heads: int) -> Tensor:
    """nn.MultiheadAttention(E, h) forward, batch_first=False, no masks, no dropout.

    Inputs are (L, N, E) / (S, N, E).  Follows torch's math path (the reference always
    runs it because need_weights defaults to True; call sites
    mm_multi_transformers.py:62,142-167,186): packed in-projection ordered q|k|v,
    q scaled by 1/sqrt(dh) BEFORE QK^T, softmax over keys, heads = contiguous split
    of E.  The averaged attention weights the reference discards are not computed.
    """
    L, N, E = query.shape
    S = key.shape[0]
    dh = E // num_heads
    w_in = p[prefix + "in_proj_weight"]
    b_in = p[prefix + "in_proj_bias"]
    q = linear(query, w_in[:E], b_in[:E])
    k = linear(key, w_in[E:2 * E], b_in[E:2 * E])
    v = linear(value, w_in[2 * E:], b_in[2 * E:])
    # (L, N, E) -> (N*h, L, dh)
    q = q.reshape(L, N * num_heads, dh).transpose(0, 1)
    k = k.reshape(S, N * num_heads, dh).transpose(0, 1)
    v = v.reshape(S, N * num_heads, dh).transpose(0, 1)
    q = q * math.sqrt(1.0 / float(dh))
    s = torch.matmul(q, k.transpose(1, 2))           # (N*h, L, S)
    s = s - s.max(dim=-1, keepdim=True).values
    e = torch.exp(s)
    pr = e / e.sum(dim=-1, keepdim=True)
    o = torch.matmul(pr, v)                          # (N*h, L, dh)
    o = o.transpose(0, 1).reshape(L, N, E)
    return linear(o, p[prefix + "out_proj.weight"], p[prefix + "out_proj.bias"])


def encoder_layer(x: Tensor, p: Params, prefix: str, num_heads: int) -> Tensor:
    """TransformerEncoderLayer.forward (mm_multi_transformers.py:60-70; duplicates at
    mm_transformers.py:74-84 and intra_modal_transformer_fusion.py:61-71):
    post-LN:  x = LN1(x + MHA(x,x,x));  x = LN2(x + W2 relu(W1 x + b1) + b2)."""
    a = mha(x, x, x, p, prefix + "attention.", num_heads)
    x = layer_norm(x + a, p[prefix + "layer_norm1.weight"], p[prefix + "layer_norm1.bias"])
    h = torch.relu(linear(x, p[prefix + "feed_forward.0.weight"], p[prefix + "feed_forward.0.bias"]))
    f = linear(h, p[prefix + "feed_forward.2.weight"], p[prefix + "feed_forward.2.bias"])
    return layer_norm(x + f, p[prefix + "layer_norm2.weight"], p[prefix + "layer_norm2.bias"])


def encoder_block(x: Tensor, p: Params, prefix: str, num_heads: int, num_layers: int) -> Tensor:
    """TransformerEncoderBlock (mm_multi_transformers.py:36-45)."""
    for i in range(num_layers):
        x = encoder_layer(x, p, f"{prefix}layers.{i}.", num_heads)
    return x


# --------------------------------------------------------------------------- #
# fusion variants
# --------------------------------------------------------------------------- #
def w_jr_forward(visual: Tensor, phys: Tensor, p: Params, prefix: str, num_heads: int,
                 num_layers: int, output_format: str) -> Tensor:
    """MultimodalTransformer_w_JR.forward (mm_multi_transformers.py:118-214).

    Inputs (B, T, 512).  Output: 'FC' -> (T, B, 1024) [the reference never permutes
    back, SURVEY Q1]; 'SELF_ATTEN' -> (B, T, 512)."""
    jr = linear(torch.cat((visual, phys), dim=2),
                p[prefix + "out_layer_pv.weight"], p[prefix + "out_layer_pv.bias"])   # :120-124
    v = visual.permute(1, 0, 2)                                                      # :127-129
    a = phys.permute(1, 0, 2)
    j = jr.permute(1, 0, 2)
    v = encoder_block(v, p, prefix + "visual_encoder.", num_heads, num_layers)        # :132-136
    a = encoder_block(a, p, prefix + "physiological_encoder.", num_heads, num_layers)
    j = encoder_block(j, p, prefix + "joint_representation_encoder.", num_heads, num_layers)
    cv, cp, cpv = prefix + "cross_attention_v.", prefix + "cross_attention_p.", prefix + "cross_attention_pv."
    outs = [
        mha(v, a, a, p, cv, num_heads),    # v <- p   :142-144
        mha(a, v, v, p, cp, num_heads),    # p <- v   :147-148
        mha(j, v, v, p, cpv, num_heads),   # jr <- v  :151-152
        mha(v, j, j, p, cv, num_heads),    # v <- jr  :155-157
        mha(j, a, a, p, cpv, num_heads),   # jr <- p  :160-162
        mha(a, j, j, p, cp, num_heads),    # p <- jr  :165-167
    ]
    if output_format == "FC":
        cat = torch.cat(outs, dim=2)                                                 # :203-208
        return linear(cat, p[prefix + "out_layer1.weight"], p[prefix + "out_layer1.bias"])
    if output_format == "SELF_ATTEN":
        st = torch.stack(outs, dim=2)            # (T, B, 6, 512)   :173-178
        st = st.permute(1, 0, 2, 3)              # (B, T, 6, 512)   :179
        b_size, seq_size = st.shape[0], st.shape[1]
        flat = st.flatten(0, 1).permute(1, 0, 2)  # (6, B*T, 512)    :180
        enc = encoder_block(flat, p, prefix + "final_visual_encoder.", num_heads, num_layers)
        fa = mha(enc, enc, enc, p, prefix + "final_self_attention.", num_heads)      # :186-188
        fa = fa.permute(1, 0, 2).unflatten(0, (b_size, seq_size))                    # :189-191
        return fa[:, :, -1, :]                                                       # :193
    raise NotImplementedError(output_format)


def wo_jr_forward(visual: Tensor, phys: Tensor, p: Params, prefix: str, num_heads: int,
                  num_layers: int) -> Tensor:
    """MultimodalTransformer_wo_JR.forward (mm_transformers.py:119-146).  The encoders are
    fed (B, T, E) un-permuted, so they attend ACROSS THE BATCH (L=B, N=T)."""
    v = encoder_block(visual, p, prefix + "visual_encoder.", num_heads, num_layers)           # :120
    a = encoder_block(phys, p, prefix + "physiological_encoder.", num_heads, num_layers)      # :121-122
    vt, at = v.permute(1, 0, 2), a.permute(1, 0, 2)
    cv = mha(vt, at, at, p, prefix + "cross_attention_v.", num_heads).permute(1, 0, 2)        # :125-129
    cp = mha(at, vt, vt, p, prefix + "cross_attention_p.", num_heads).permute(1, 0, 2)        # :132-135
    return linear(torch.cat((cv, cp), dim=2),
                  p[prefix + "final_layer.weight"], p[prefix + "final_layer.bias"])          # :140-144


def feature_concat_fc(visual: Tensor, audio: Tensor, p: Params, prefix: str) -> Tensor:
    """FeatureConcatFC.forward (mm_multi_transformers.py:222-225)."""
    return linear(torch.cat((visual, audio), dim=2), p[prefix + "fc.weight"], p[prefix + "fc.bias"])


def regressor(x: Tensor, p: Params, prefix: str) -> Tensor:
    """Linear(dim,128) -> ReLU -> Dropout(eval: identity) -> Linear(128,k)
    (two_transformers.py:104-114)."""
    h = torch.relu(linear(x, p[prefix + "0.weight"], p[prefix + "0.bias"]))
    return linear(h, p[prefix + "3.weight"], p[prefix + "3.bias"])


def two_transformers_forward(f1_audio: Tensor, f2_visual: Tensor, p: Params, num_heads: int,
                             num_layers: int, joint_modalities: str,
                             output_format: str = "FC") -> Tuple[Tensor, Tensor]:
    """Two_transformers.forward (two_transformers.py:116-128), eval mode (dropout off)."""
    video = l2_normalize(f2_visual)
    audio = l2_normalize(f1_audio)
    if "linear.weight" in p:                                                        # :120-121
        video = linear(video, p["linear.weight"], p["linear.bias"])
    if joint_modalities == "TRANSFORMER":
        feats = w_jr_forward(video, audio, p, "mm_transformer.", num_heads, num_layers, output_format)
    elif joint_modalities == "FC":
        feats = feature_concat_fc(video, audio, p, "mm_transformer.")
    elif joint_modalities == "NONE":
        feats = wo_jr_forward(video, audio, p, "mm_transformer.", num_heads, num_layers)
    else:
        raise NotImplementedError(joint_modalities)
    v = regressor(feats, p, "vregressor.").squeeze(2)
    a = regressor(feats, p, "aregressor.").squeeze(2)
    return v, a


def single_backbone_pretrainer_forward(x: Tensor, p: Params) -> Tuple[Tensor, Tensor]:
    """SingleBackbonePretrainer.forward (two_transformers.py:151-162)."""
    out = regressor(x, p, "regressor.")
    return out[:, :, 0], out[:, :, 1]


def intra_modal_forward(fa: Tensor, fb: Tensor, p: Params, num_heads: int, num_layers: int) -> Tensor:
    """Intra_modal_transformer_fusion.forward (intra_modal_transformer_fusion.py:84-111):
    a length-2 'sequence' per (b, t); keep the last token."""
    if fa.shape[-1] == 768:
        fa = linear(fa, p["fc.weight"], p["fc.bias"])
    if fb.shape[-1] == 768:
        fb = linear(fb, p["fc.weight"], p["fc.bias"])
    st = torch.stack((fa, fb), dim=2)                       # (B, T, 2, 512)
    b_size, seq_size = st.shape[0], st.shape[1]
    flat = st.flatten(0, 1).permute(1, 0, 2)                # (2, B*T, 512)
    enc = encoder_block(flat, p, "final_visual_encoder.", num_heads, num_layers)
    out = mha(enc, enc, enc, p, "final_self_attention.", num_heads)
    out = out.permute(1, 0, 2).unflatten(0, (b_size, seq_size))
    return out[:, :, -1, :]


def fc_layer_forward(x: Tensor, p: Params) -> Tensor:
    """FcLayer.forward (fc_layer.py:11-12)."""
    return linear(x, p["fc_layer.weight"], p["fc_layer.bias"])


# --------------------------------------------------------------------------- #
# TCN
# --------------------------------------------------------------------------- #
def weight_norm_weight(g: Tensor, v: Tensor) -> Tensor:
    """Legacy torch weight_norm, dim=0: w = g * v / ||v|| with the norm over (Cin, k) per
    output channel (temporal_convolutional_model.py:24-26,31-33)."""
    n = torch.sqrt((v * v).sum(dim=(1, 2), keepdim=True))
    return v * (g / n)


def causal_dilated_conv1d(x: Tensor, w: Tensor, b: Tensor, dilation: int) -> Tensor:
    """Conv1d(padding=(k-1)d, dilation=d) followed by Chomp1d((k-1)d)
    (temporal_convolutional_model.py:12-18,24-27):
    out[n,co,t] = b[co] + sum_j sum_ci w[co,ci,j] * x[n,ci,t-(k-1-j)d], zero for t<0."""
    n, cin, L = x.shape
    cout, _, k = w.shape
    out = b.view(1, cout, 1).expand(n, cout, L).clone()
    for j in range(k):
        shift = (k - 1 - j) * dilation
        if shift >= L:
            continue
        xs = torch.zeros_like(x)
        xs[:, :, shift:] = x[:, :, :L - shift]
        out = out + torch.einsum("oc,ncl->nol", w[:, :, j], xs)
    return out


def leaky_relu(x: Tensor, slope: float = LEAKY_SLOPE) -> Tensor:
    return torch.where(x >= 0, x, x * slope)


def tcn_forward(x: Tensor, p: Params, num_levels: int, prefix: str = "network.") -> Tensor:
    """TemporalConvNet.forward / TemporalBlock.forward (temporal_convolutional_model.py:54-57,
    81-82), eval mode (Dropout2d identity).  x: (N, C, L) -> (N, C_last, L)."""
    for i in range(num_levels):
        d = 2 ** i                                                               # :67
        pre = f"{prefix}{i}."
        w1 = weight_norm_weight(p[pre + "conv1.weight_g"], p[pre + "conv1.weight_v"])
        w2 = weight_norm_weight(p[pre + "conv2.weight_g"], p[pre + "conv2.weight_v"])
        h = leaky_relu(causal_dilated_conv1d(x, w1, p[pre + "conv1.bias"], d))
        h = leaky_relu(causal_dilated_conv1d(h, w2, p[pre + "conv2.bias"], d))
        if (pre + "downsample.weight") in p:                                     # :41
            res = torch.einsum("oc,ncl->nol", p[pre + "downsample.weight"][:, :, 0], x) \
                + p[pre + "downsample.bias"].view(1, -1, 1)
        else:
            res = x
        x = leaky_relu(h + res)                                                  # :57
    return x


# --------------------------------------------------------------------------- #
# CCC: the three formulas, each also as a closed form of six sums
# --------------------------------------------------------------------------- #
def ccc_loss_live(x: Tensor, y: Tensor, eps: float = 1e-8) -> Tensor:
    """losses/loss.py:18-32 with digitize_num == 1: unbiased std, eps only in rho's
    denominator, NO -5 masking."""
    x = x.reshape(-1)
    y = y.reshape(-1)
    vx = x - x.mean()
    vy = y - y.mean()
    rho = (vx * vy).sum() / (torch.sqrt((vx ** 2).sum()) * torch.sqrt((vy ** 2).sum()) + eps)
    x_m, y_m = x.mean(), y.mean()
    n = x.numel()
    x_s = torch.sqrt((vx ** 2).sum() / (n - 1))
    y_s = torch.sqrt((vy ** 2).sum() / (n - 1))
    ccc = 2 * rho * x_s * y_s / (x_s ** 2 + y_s ** 2 + (x_m - y_m) ** 2)
    return 1 - ccc


def label_mask(y_true: Tensor, ignore: float = IGNORE_LABEL) -> Tensor:
    """The reference's only 'padding mask': y_true != -5.0 (losses/CCCLoss.py:19)."""
    return y_true != ignore


def ccc_loss_masked(y_pred: Tensor, y_true: Tensor, ignore: float = IGNORE_LABEL) -> Tensor:
    """losses/CCCLoss.py:12-43: masks y_true != ignore, unbiased variances, divides by the
    PRE-mask length; returns 0 when <= 1 valid element."""
    batch_size = y_pred.shape[0]
    idx = label_mask(y_true, ignore)
    yt = y_true[idx]
    yp = y_pred[idx]
    if yt.shape[0] <= 1:
        return torch.zeros((), dtype=y_pred.dtype)
    x_m, y_m = yp.mean(), yt.mean()
    nv = yt.shape[0]
    var_t = ((yt - y_m) ** 2).sum() / (nv - 1)
    var_p = ((yp - x_m) ** 2).sum() / (nv - 1)
    s_xy = ((yp - x_m) * (yt - y_m)).sum()
    ccc = 2 * s_xy / ((var_t + var_p + (x_m - y_m) ** 2 + 1e-8) * batch_size)
    return 1 - ccc


def ccc_metric(x: np.ndarray, y: np.ndarray) -> float:
    """EvaluationMetrics/cccmetric.py:4-21 (numpy, population std). len<=1 -> the reference
    calls sys.exit(); the oracle raises ValueError instead."""
    x = np.asarray(x)
    y = np.asarray(y)
    if len(y) <= 1:
        raise ValueError("ccc needs more than one sample (reference: sys.exit())")
    vx = x - np.mean(x)
    vy = y - np.mean(y)
    rho = np.sum(vx * vy) / (np.sqrt(np.sum(vx ** 2)) * np.sqrt(np.sum(vy ** 2)))
    x_m, y_m = np.mean(x), np.mean(y)
    x_s, y_s = np.std(x), np.std(y)
    return float(2 * rho * x_s * y_s / (x_s ** 2 + y_s ** 2 + (x_m - y_m) ** 2))


def cccva(y_true: np.ndarray, y_pred: np.ndarray) -> Tuple[float, float, float]:
    """cccmetric.py:24-38."""
    cv = ccc_metric(y_true[:, 0], y_pred[:, 0])
    ca = ccc_metric(y_true[:, 1], y_pred[:, 1])
    return cv, ca, (cv + ca) / 2


def ccc_numpy(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """cccmetric.py:41-56: np.cov (N-1) over np.var (N), +1e-8."""
    y_true = np.asarray(y_true, dtype=np.float64)
    y_pred = np.asarray(y_pred, dtype=np.float64)
    n = y_true.shape[0]
    x_m, y_m = y_true.mean(), y_pred.mean()
    s_xy = ((y_true - x_m) * (y_pred - y_m)).sum() / (n - 1)
    return float(2.0 * s_xy / (y_true.var() + y_pred.var() + (x_m - y_m) ** 2 + 1e-8))


def six_sums(x: np.ndarray, y: np.ndarray, ignore: Optional[float] = None) -> np.ndarray:
    """(N, sum x, sum y, sum xy, sum x^2, sum y^2) in fp64 over elements with y != ignore --
    the statistic the CUDA reduction kernel produces."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    if ignore is not None:
        keep = np.asarray(y, dtype=np.float32) != np.float32(ignore)
        x, y = x[keep], y[keep]
    return np.array([x.size, x.sum(), y.sum(), (x * y).sum(), (x * x).sum(), (y * y).sum()])


def ccc_from_sums(s: np.ndarray, kind: str, n_all: Optional[int] = None, eps: float = 1e-8) -> float:
    """Closed forms of the three CCC variants from the six sums.
    kind: 'metric' (cccmetric.ccc), 'loss_live' (losses/loss.py), 'loss_masked'
    (losses/CCCLoss.py; needs n_all = pre-mask length), 'ccc_numpy'."""
    n, sx, sy, sxy, sxx, syy = [float(v) for v in s]
    if kind == "loss_masked" and n <= 1:
        return 0.0
    cxx = sxx - sx * sx / n
    cyy = syy - sy * sy / n
    cxy = sxy - sx * sy / n
    d2 = (sx / n - sy / n) ** 2
    if kind == "metric":
        return 2 * (cxy / n) / (cxx / n + cyy / n + d2)
    if kind == "loss_live":
        rho = cxy / (math.sqrt(cxx) * math.sqrt(cyy) + eps)
        xs, ys = math.sqrt(cxx / (n - 1)), math.sqrt(cyy / (n - 1))
        return 1 - 2 * rho * xs * ys / (xs * xs + ys * ys + d2)
    if kind == "loss_masked":
        return 1 - 2 * cxy / ((cxx / (n - 1) + cyy / (n - 1) + d2 + 1e-8) * n_all)
    if kind == "ccc_numpy":
        return 2 * (cxy / (n - 1)) / (cxx / n + cyy / n + d2 + 1e-8)
    raise ValueError(kind)


# --------------------------------------------------------------------------- #
# collate zero-fill semantics (P1)
# --------------------------------------------------------------------------- #
def pad_spectrograms(specs: Sequence[Tensor]) -> Tensor:
    """The audio half of Train/Val/TestPadSequence.__call__ (padSequence.py:9-21): zeros
    (B,16,1,64,maxW); each spectrogram is RIGHT-aligned when ``shape[2] < maxW`` -- the
    reference compares the mel-bin dim (64), not the width (bug kept) -- else
    copied whole (which only works when widths are equal)."""
    widths = [int(s.shape[3]) for s in specs]
    max_w = max(widths)
    out = torch.zeros(len(specs), 16, 1, 64, max_w, dtype=specs[0].dtype)
    for i, s in enumerate(specs):
        if s.shape[2] < max_w:
            out[i, :, :, :, max_w - s.shape[3]:] = s
        else:
            out[i, :, :, :, :] = s
    return out


# --------------------------------------------------------------------------- #
# parameter inventories (shapes under the reference state_dict names) and
# deterministic synthetic parameters -- shared by the golden generator and tests
# --------------------------------------------------------------------------- #
def _mha_shapes(prefix: str, e: int) -> List[Tuple[str, Tuple[int, ...]]]:
    return [(prefix + "in_proj_weight", (3 * e, e)), (prefix + "in_proj_bias", (3 * e,)),
            (prefix + "out_proj.weight", (e, e)), (prefix + "out_proj.bias", (e,))]


def _enc_shapes(prefix: str, e: int, hidden: int, num_layers: int):
    out = []
    for i in range(num_layers):
        pre = f"{prefix}layers.{i}."
        out += _mha_shapes(pre + "attention.", e)
        out += [(pre + "feed_forward.0.weight", (hidden, e)), (pre + "feed_forward.0.bias", (hidden,)),
                (pre + "feed_forward.2.weight", (e, hidden)), (pre + "feed_forward.2.bias", (e,)),
                (pre + "layer_norm1.weight", (e,)), (pre + "layer_norm1.bias", (e,)),
                (pre + "layer_norm2.weight", (e,)), (pre + "layer_norm2.bias", (e,))]
    return out


def _regressor_shapes(prefix: str, dim: int, k: int):
    return [(prefix + "0.weight", (128, dim)), (prefix + "0.bias", (128,)),
            (prefix + "3.weight", (k, 128)), (prefix + "3.bias", (k,))]


def two_transformers_shapes(num_layers: int, joint_modalities: str, output_format: str = "FC",
                            vision_in_ft: int = 512, include_dead: bool = True):
    """Ordered (name, shape) list == reference ``Two_transformers(...).state_dict()``
    (two_transformers.py:18-114; mm_multi_transformers.py:74-116; mm_transformers.py:88-117).
    ``include_dead=False`` drops final_encoder (40.9 M params never used)."""
    s: List[Tuple[str, Tuple[int, ...]]] = []
    if vision_in_ft != 512:
        s += [("linear.weight", (512, vision_in_ft)), ("linear.bias", (512,))]
    m = "mm_transformer."
    if joint_modalities == "TRANSFORMER":
        s += _enc_shapes(m + "visual_encoder.", 512, 512, num_layers)
        s += _enc_shapes(m + "physiological_encoder.", 512, 512, num_layers)
        s += _enc_shapes(m + "joint_representation_encoder.", 512, 512, num_layers)
        if include_dead:
            s += _enc_shapes(m + "final_encoder.", 3072, 512, num_layers)
        s += _mha_shapes(m + "cross_attention_v.", 512)
        s += _mha_shapes(m + "cross_attention_p.", 512)
        s += _mha_shapes(m + "cross_attention_pv.", 512)
        s += [(m + "out_layer_pv.weight", (512, 1024)), (m + "out_layer_pv.bias", (512,))]
        if output_format == "FC":
            s += [(m + "out_layer1.weight", (1024, 3072)), (m + "out_layer1.bias", (1024,))]
            dim = 1024
        else:
            s += _enc_shapes(m + "final_visual_encoder.", 512, 512, num_layers)
            s += _mha_shapes(m + "final_self_attention.", 512)
            dim = 512
    elif joint_modalities == "FC":
        s += [(m + "fc.weight", (512, 1024)), (m + "fc.bias", (512,))]
        dim = 512
    elif joint_modalities == "NONE":
        s += _enc_shapes(m + "visual_encoder.", 512, 512, num_layers)
        s += _enc_shapes(m + "physiological_encoder.", 512, 512, num_layers)
        s += _mha_shapes(m + "cross_attention_v.", 512)
        s += _mha_shapes(m + "cross_attention_p.", 512)
        s += [(m + "gated_attention.weight", (1, 1024)), (m + "gated_attention.bias", (1,))]
        s += [(m + "final_layer.weight", (512, 1024)), (m + "final_layer.bias", (512,))]
        dim = 512
    else:
        raise NotImplementedError(joint_modalities)
    s += _regressor_shapes("vregressor.", dim, 1)
    s += _regressor_shapes("aregressor.", dim, 1)
    return s


def intra_modal_shapes(num_layers: int, feat_dim: int = 512, hidden: int = 512):
    """Intra_modal_transformer_fusion state_dict (intra_modal_transformer_fusion.py:75-82)."""
    s = _enc_shapes("final_visual_encoder.", feat_dim, hidden, num_layers)
    s += _mha_shapes("final_self_attention.", 512)
    s += [("fc.weight", (512, 768)), ("fc.bias", (512,))]
    return s


def tcn_shapes(num_inputs: int, num_channels: Sequence[int], kernel_size: int):
    """TemporalConvNet parameters = ``named_parameters()`` order
    (temporal_convolutional_model.py:61-79; legacy weight_norm registers bias, weight_g,
    weight_v in that order).  The reference ``state_dict`` additionally lists every conv twice
    (``net.0`` aliases ``conv1``, ``net.4`` aliases ``conv2``; :38-39) -- see
    ``tcn_state_dict`` below."""
    s = []
    for i, cout in enumerate(num_channels):
        cin = num_inputs if i == 0 else num_channels[i - 1]
        pre = f"network.{i}."
        s += [(pre + "conv1.bias", (cout,)), (pre + "conv1.weight_g", (cout, 1, 1)),
              (pre + "conv1.weight_v", (cout, cin, kernel_size)),
              (pre + "conv2.bias", (cout,)), (pre + "conv2.weight_g", (cout, 1, 1)),
              (pre + "conv2.weight_v", (cout, cout, kernel_size))]
        if cin != cout:
            s += [(pre + "downsample.weight", (cout, cin, 1)), (pre + "downsample.bias", (cout,))]
    return s


def tcn_state_dict(params: Params) -> Params:
    """Expand canonical TCN params into the reference's full state_dict key set/order:
    per block conv1.*, conv2.*, net.0.* (= conv1), net.4.* (= conv2), downsample.*."""
    out: Params = {}
    blocks = sorted({int(k.split(".")[1]) for k in params})
    for i in blocks:
        pre = f"network.{i}."
        for c in ("conv1.", "conv2."):
            for f in ("bias", "weight_g", "weight_v"):
                out[pre + c + f] = params[pre + c + f]
        for alias, c in (("net.0.", "conv1."), ("net.4.", "conv2.")):
            for f in ("bias", "weight_g", "weight_v"):
                out[pre + alias + f] = params[pre + c + f]
        for f in ("downsample.weight", "downsample.bias"):
            if pre + f in params:
                out[pre + f] = params[pre + f]
    return out


def synth_params(shapes, seed: int, dtype=torch.float32) -> Params:
    """Deterministic synthetic parameters for a (name, shape) list: independent of any module
    constructor so the golden generator (reference modules) and the tests (oracle, CUDA
    modules) build bit-identical weights.  Matrices ~ N(0, 1/fan_in)*1.5 (attention logits and
    FFN activations of useful size), biases ~ N(0, 0.05^2), LayerNorm gains 1 + N(0, 0.1^2),
    weight_norm g ~ |N(1, 0.2^2)|."""
    gen = torch.Generator().manual_seed(seed)
    out: Params = {}
    for name, shape in shapes:
        if name.endswith("weight_g"):
            t = (1.0 + 0.2 * torch.randn(shape, generator=gen)).abs() + 0.1
        elif "layer_norm" in name and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            t = torch.randn(shape, generator=gen) * (1.5 / math.sqrt(fan_in))
        else:
            t = 0.05 * torch.randn(shape, generator=gen)
        out[name] = t.to(dtype)
    return out


def synth_features(b: int, t: int, dims: Sequence[int], seed: int, dtype=torch.float32) -> List[Tensor]:
    """Seeded N(0,1) feature sequences (B, T, d) for each d in dims."""
    gen = torch.Generator().manual_seed(seed)
    return [torch.randn(b, t, d, generator=gen).to(dtype) for d in dims]


def synth_labels(b: int, t: int, seed: int, frac_ignored: float = 0.05) -> Tuple[Tensor, Tensor]:
    """Labels U(-1,1) with ~5 % of positions set to exactly -5.0."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(2):
        y = torch.rand(b, t, generator=gen) * 2 - 1
        drop = torch.rand(b, t, generator=gen) < frac_ignored
        out.append(torch.where(drop, torch.full_like(y, IGNORE_LABEL), y))
    return out[0], out[1]
