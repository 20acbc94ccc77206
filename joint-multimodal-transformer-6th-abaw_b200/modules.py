"""Drop-in nn.Modules mirroring the reference's hot-path classes: same class names, constructor
signatures, forward signatures, state_dict keys/shapes and (seeded) initialisation, so that
main.py / train.py / val.py can swap them in.  ``forward`` never touches ATen math:
it runs the tape engine over libjmt_b200.so and is wired into torch.autograd with one Function.

Extra keyword (not in the reference): ``precision`` in {'bf16' (default), 'bf16x3', 'fp32'} (engine.py).
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from . import engine as E

__all__ = ["JMTPipeline", "Two_transformers", "SingleBackbonePretrainer", "MultimodalTransformer_w_JR",
           "MultimodalTransformer_wo_JR", "FeatureConcatFC", "Intra_modal_transformer_fusion", "FcLayer",
           "TemporalConvNet", "TemporalBlock", "TransformerEncoderBlock", "TransformerEncoderLayer"]


# ----------------------------------------------------------------------------- parameter holders
class _Linear(nn.Module):
    """Parameters of an nn.Linear (same names, shapes and default init); no forward."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1 / math.sqrt(in_features) if in_features > 0 else 0
        nn.init.uniform_(self.bias, -bound, bound)


class _Slot(nn.Module):
    """Parameter-free placeholder keeping the reference's nn.Sequential indices (ReLU, Dropout, Chomp1d)."""


class _LayerNorm(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class _MultiheadAttention(nn.Module):
    """Parameters of nn.MultiheadAttention(E, h): in_proj_weight (3E,E) q|k|v, in_proj_bias, out_proj.{weight,bias},
    initialised in torch's order (out_proj Linear init, then xavier in_proj, zero biases)."""

    def __init__(self, embed_dim: int, num_heads: int):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = _Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.0)
        nn.init.constant_(self.out_proj.bias, 0.0)


class TransformerEncoderLayer(nn.Module):
    """Parameters of the reference TransformerEncoderLayer (mm_multi_transformers.py:48-58)."""

    def __init__(self, input_dim, num_heads, hidden_dim):
        super().__init__()
        self.attention = _MultiheadAttention(input_dim, num_heads)
        self.feed_forward = nn.Sequential(_Linear(input_dim, hidden_dim), _Slot(), _Linear(hidden_dim, input_dim))
        self.layer_norm1 = _LayerNorm(input_dim)
        self.layer_norm2 = _LayerNorm(input_dim)


class TransformerEncoderBlock(nn.Module):
    """mm_multi_transformers.py:36-45."""

    def __init__(self, input_dim, num_heads, hidden_dim, num_layers):
        super().__init__()
        self.layers = nn.Sequential(*[TransformerEncoderLayer(input_dim, num_heads, hidden_dim)
                                      for _ in range(num_layers)])


# ----------------------------------------------------------------------------- autograd bridge
class _TapeFn(torch.autograd.Function):
    """Runs `runner(ctx, *inputs)` forward and replays the engine tape in backward.
    runner returns (outputs: list[Tensor], out_grad_setters: list[callable], in_grad_getters: list)."""

    @staticmethod
    def forward(actx, owner, runner, record, n_in, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        names = owner._live_names_fast()
        pd = {n: p for n, p in zip(names, params)}
        owner._calls = getattr(owner, "_calls", 0) + 1
        seed = (torch.initial_seed() * 1000003 + owner._calls) & ((1 << 62) - 1)
        dev_key = params[0].device.index if params else 0
        wc = owner.__dict__.setdefault("_wcache", {}).setdefault(dev_key, {})
        # dropout draws: host seed (varies per eager call) + a device-resident counter that every forward advances, so a
        # CUDA-graph replay (host arguments frozen at capture) still gets fresh masks
        rng = None
        if owner.training:
            rs = owner.__dict__.setdefault("_rng_state", {})
            rng = rs.get(dev_key)
            if rng is None:
                rng = rs[dev_key] = torch.zeros(2, dtype=torch.int64, device=params[0].device if params else inputs[0].device)
        ctx = E.Ctx(pd, owner.precision, record, owner.training, wc, seed, rng)
        outs, setters, getters = runner(ctx, *inputs)
        ctx.finish_forward()
        actx.jctx, actx.setters, actx.getters, actx.names, actx.owner = ctx, setters, getters, names, owner
        actx.n_in = n_in
        actx.in_shapes = [tuple(i.shape) for i in inputs]
        return tuple(outs)

    @staticmethod
    def backward(actx, *gouts):
        ctx = actx.jctx
        if ctx is None:
            raise RuntimeError("jmt_b200: backward through the same forward twice is not supported")
        ctx.prepare_param_grads(actx.names)
        hook = getattr(actx.owner, "_grad_sync", None)
        ctx.grad_sync = hook
        for s, g in zip(actx.setters, gouts):
            if g is not None:
                s(g)
        ctx.backward()
        if hook is not None:                      # NCCL all-reduce of the flat live-gradient bucket
            if ctx.synced and hasattr(hook, "start"):
                for lo, hi in ctx.unsynced_ranges():          # the other slices are already in flight (overlap)
                    hook.start(ctx.bucket[lo:hi])
                hook.finish()
            else:
                hook(ctx.bucket)
        gin = [g().view(sh) if g is not None else None for g, sh in zip(actx.getters, actx.in_shapes)]
        gp = [ctx.pgrads[n] for n in actx.names]
        actx.jctx = None
        return (None, None, None, None, *gin, *gp)


class _JmtModule(nn.Module):
    """Common plumbing: live-parameter inventory and the call into the tape engine."""
    precision = "bf16"

    def _dead_prefixes(self) -> Tuple[str, ...]:
        return ()

    def _live_names(self) -> List[str]:
        dead = self._dead_prefixes()
        return [n for n, _ in self.named_parameters() if not any(n.startswith(d) for d in dead)]

    def live_parameters(self):
        """Parameters that can receive gradients (excludes constructed-but-unused ones)."""
        dead = self._dead_prefixes()
        return [p for n, p in self.named_parameters() if not any(n.startswith(d) for d in dead)]

    def _live_slots(self):
        """(name, owner module, attribute) of every live parameter in `_live_names()` order.  The walk over the module tree
        (named_parameters over ~60 sub-modules, 1 ms per call) is cached; the module count invalidates the cache when sub-modules are
        added or removed, and because the slots are re-read on every use a re-assigned Parameter is picked up without a re-walk."""
        n_mod = sum(1 for _ in self.modules())
        cache = self.__dict__.get("_live_slots_cache")
        if cache is None or cache[0] != n_mod:
            dead = self._dead_prefixes()
            slots, seen = [], set()
            for mname, mod in self.named_modules():
                for pname, par in mod._parameters.items():
                    if par is None or id(par) in seen:
                        continue
                    seen.add(id(par))
                    full = (mname + "." if mname else "") + pname
                    if not any(full.startswith(d) for d in dead):
                        slots.append((full, mod, pname))
            order = {n: i for i, n in enumerate(self._live_names())}
            slots.sort(key=lambda s: order[s[0]])
            assert len(slots) == len(order), "live-parameter inventory mismatch"
            cache = (n_mod, slots)
            self.__dict__["_live_slots_cache"] = cache
        return cache[1]

    def _live_params_fast(self) -> List[nn.Parameter]:
        return [mod._parameters[pname] for _, mod, pname in self._live_slots()]

    def _live_names_fast(self) -> List[str]:
        return [n for n, _, _ in self._live_slots()]

    def set_grad_sync(self, fn):
        """fn(flat_fp32_bucket) is called at the end of backward (data-parallel all-reduce hook)."""
        self.__dict__["_grad_sync"] = fn

    def _run(self, runner, *inputs):
        inputs = tuple(i.float() if i.dtype == torch.float16 else i for i in inputs)
        for i in inputs:
            E.require_cuda(i)
        params = self._live_params_fast()
        E.require_cuda(*params)
        record = torch.is_grad_enabled() and any(t.requires_grad for t in (*inputs, *params))
        with torch.autocast("cuda", enabled=False):
            return _TapeFn.apply(self, runner, record, len(inputs), *inputs, *params)


# ----------------------------------------------------------------------------- fusion graphs
def _w_jr_graph(ctx, video, audio, prefix, heads, layers, output_format, B, T):
    """MultimodalTransformer_w_JR.forward (mm_multi_transformers.py:118-214) on (B*T, 512) row-major
    activations (row = b*T + t).  Returns (features Var, time_major flag)."""
    M = B * T
    g = E.AttnGeom(T, B, 1, T)
    # joint representation: Linear(1024->512)(cat(v, a)) as two K=512 GEMMs on the column halves of W
    jr = E.linear(ctx, video, prefix + "out_layer_pv.weight", prefix + "out_layer_pv.bias", w_cols=(0, 512))
    E.linear(ctx, audio, prefix + "out_layer_pv.weight", None, w_cols=(512, 1024), accumulate_into=jr)
    v = E.encoder_block(ctx, video, prefix + "visual_encoder.", heads, layers, g)
    a = E.encoder_block(ctx, audio, prefix + "physiological_encoder.", heads, layers, g)
    j = E.encoder_block(ctx, jr, prefix + "joint_representation_encoder.", heads, layers, g)
    cv, cp, cpv = prefix + "cross_attention_v.", prefix + "cross_attention_p.", prefix + "cross_attention_pv."
    pairs = [(v, a, cv), (a, v, cp), (j, v, cpv), (v, j, cv), (j, a, cpv), (a, j, cp)]     # :142-167
    if output_format == "FC":
        cat = E.Var(ctx.empty((M, 6 * 512)))
        for i, (q, kv, pre) in enumerate(pairs):
            E.mha_cross(ctx, q, kv, pre, heads, g, g, out=cat.data[:, i * 512:(i + 1) * 512],
                        grad_from=(cat, i * 512, (i + 1) * 512))
        out = E.linear(ctx, cat, prefix + "out_layer1.weight", prefix + "out_layer1.bias")
        return out, True                                   # (T, B) outputs: SURVEY Q1
    outs = [E.mha_cross(ctx, q, kv, pre, heads, g, g) for (q, kv, pre) in pairs]
    st = E.stack_rows(ctx, outs)                            # (6*M, 512), row = l*M + m
    enc = E.encoder_block(ctx, st, prefix + "final_visual_encoder.", heads, layers, None, small=(6, M))
    fa = E.mha_self(ctx, enc, prefix + "final_self_attention.", heads, None, small=(6, M))
    return E.take_rows(ctx, fa, 5 * M, 6 * M), False


def _wo_jr_graph(ctx, video, audio, prefix, heads, layers, B, T):
    """MultimodalTransformer_wo_JR.forward (mm_transformers.py:119-146): encoders attend across the
    BATCH (seq = b, batch = t), cross-attention over time."""
    M = B * T
    g_batch = E.AttnGeom(B, T, T, 1)
    g_time = E.AttnGeom(T, B, 1, T)
    v = E.encoder_block(ctx, video, prefix + "visual_encoder.", heads, layers, g_batch)
    a = E.encoder_block(ctx, audio, prefix + "physiological_encoder.", heads, layers, g_batch)
    cat = E.Var(ctx.empty((M, 1024)))
    E.mha_cross(ctx, v, a, prefix + "cross_attention_v.", heads, g_time, g_time, out=cat.data[:, 0:512],
                grad_from=(cat, 0, 512))
    E.mha_cross(ctx, a, v, prefix + "cross_attention_p.", heads, g_time, g_time, out=cat.data[:, 512:1024],
                grad_from=(cat, 512, 1024))
    return E.linear(ctx, cat, prefix + "final_layer.weight", prefix + "final_layer.bias")


def _concat_fc_graph(ctx, video, audio, prefix):
    """FeatureConcatFC.forward (mm_multi_transformers.py:222-225) without materialising the cat."""
    y = E.linear(ctx, video, prefix + "fc.weight", prefix + "fc.bias", w_cols=(0, 512))
    E.linear(ctx, audio, prefix + "fc.weight", None, w_cols=(512, 1024), accumulate_into=y)
    return y


class MultimodalTransformer_w_JR(_JmtModule):
    """mm_multi_transformers.py:73-214 (parameters incl. the never-used final_encoder)."""

    def __init__(self, visual_dim, audio_dim, num_heads, hidden_dim, num_layers, output_format: str,
                 precision: str = "bf16"):
        super().__init__()
        assert output_format in ['FC', 'SELF_ATTEN'], output_format
        assert visual_dim == 512 and audio_dim == 512 and hidden_dim == 512, "the reference hard-codes 512"
        self.output_format = output_format
        self.num_heads, self.num_layers, self.precision = num_heads, num_layers, precision
        self.visual_encoder = TransformerEncoderBlock(visual_dim, num_heads, hidden_dim, num_layers)
        self.physiological_encoder = TransformerEncoderBlock(audio_dim, num_heads, hidden_dim, num_layers)
        self.joint_representation_encoder = TransformerEncoderBlock(audio_dim, num_heads, hidden_dim, num_layers)
        self.final_encoder = TransformerEncoderBlock(3072, num_heads, hidden_dim, num_layers)
        self.cross_attention_v = _MultiheadAttention(visual_dim, num_heads)
        self.cross_attention_p = _MultiheadAttention(audio_dim, num_heads)
        self.cross_attention_pv = _MultiheadAttention(512, num_heads)
        self.out_layer_pv = _Linear(1024, 512)
        if output_format == 'FC':
            self.out_layer1 = _Linear(3072, 1024)
        else:
            self.final_visual_encoder = TransformerEncoderBlock(visual_dim, num_heads, hidden_dim, num_layers)
            self.final_self_attention = _MultiheadAttention(512, num_heads)

    def _dead_prefixes(self):
        return ("final_encoder.",)

    def forward(self, visual_features, physiological_features):
        B, T = visual_features.shape[0], visual_features.shape[1]

        def runner(ctx, vis, aud):
            v, gv = E.from_external(ctx, vis, vis.requires_grad)
            a, ga = E.from_external(ctx, aud, aud.requires_grad)
            feats, time_major = _w_jr_graph(ctx, v, a, "", self.num_heads, self.num_layers, self.output_format, B, T)
            out, setter = _features_out(ctx, feats, B, T, time_major)
            return [out], [setter], [gv, ga]
        return self._run(runner, visual_features, physiological_features)[0]


def _features_out(ctx, feats, B, T, time_major):
    """Feature Var (row = b*T + t) -> external fp32 (B,T,D), or (T,B,D) when time_major."""
    D = feats.data.shape[1]
    if not time_major:
        return E.to_external(ctx, feats, (B, T, D))
    out = ctx.empty((T, B, D), torch.float32)
    ld = feats.data.stride(0)
    # out[t, b, :] = feats[b*T + t, :]: one launch (row permutation + cast)
    L.check(ctx.lib.jmt_copy3d(E._ptr(feats.data), ctx.acode, T * ld, ld, E._ptr(out), L.F32, D, B * D, B, T, D, E._stream()), "jmt_copy3d")
    g = {"t": None}
    if ctx.record:
        def bwd():
            if g["t"] is None:
                return
            d = g["t"]
            if not d.is_contiguous() or d.dtype not in E._DT:
                raise RuntimeError("jmt_b200: the gradient of the (T, B, D) feature output must be a contiguous fp32 / bf16 tensor")
            gb = E.GradBuf(ctx.empty(feats.data.shape))
            L.check(ctx.lib.jmt_copy3d(E._ptr(d), E._DT[d.dtype], D, B * D, E._ptr(gb.t), ctx.acode, T * D, D, B, T, D, E._stream()),
                    "jmt_copy3d")
            ctx.add_grad(feats, gb)
            gb.refs -= 1
        ctx.tape.append(bwd)
    return out, (lambda t: g.__setitem__("t", t))


class MultimodalTransformer_wo_JR(_JmtModule):
    """mm_transformers.py:87-146 (gated_attention is constructed but unused)."""

    def __init__(self, visual_dim, audio_dim, num_heads, hidden_dim, num_layers, output_format: str,
                 precision: str = "bf16"):
        super().__init__()
        assert output_format in ['FC'], output_format
        assert visual_dim == 512 and audio_dim == 512 and hidden_dim == 512
        self.output_format = output_format
        self.num_heads, self.num_layers, self.precision = num_heads, num_layers, precision
        self.visual_encoder = TransformerEncoderBlock(visual_dim, num_heads, hidden_dim, num_layers)
        self.physiological_encoder = TransformerEncoderBlock(audio_dim, num_heads, hidden_dim, num_layers)
        self.cross_attention_v = _MultiheadAttention(visual_dim, num_heads)
        self.cross_attention_p = _MultiheadAttention(audio_dim, num_heads)
        self.gated_attention = _Linear(visual_dim + audio_dim, 1)
        self.final_layer = _Linear(1024, 512)

    def _dead_prefixes(self):
        return ("gated_attention.",)

    def forward(self, visual_features, physiological_features):
        B, T = visual_features.shape[0], visual_features.shape[1]

        def runner(ctx, vis, aud):
            v, gv = E.from_external(ctx, vis, vis.requires_grad)
            a, ga = E.from_external(ctx, aud, aud.requires_grad)
            feats = _wo_jr_graph(ctx, v, a, "", self.num_heads, self.num_layers, B, T)
            out, setter = _features_out(ctx, feats, B, T, False)
            return [out], [setter], [gv, ga]
        return self._run(runner, visual_features, physiological_features)[0]


class FeatureConcatFC(_JmtModule):
    """mm_multi_transformers.py:217-225."""

    def __init__(self, visual_dim, audio_dim, precision: str = "bf16"):
        super().__init__()
        assert visual_dim == 512 and audio_dim == 512
        self.precision = precision
        self.fc = _Linear(visual_dim + audio_dim, 512)

    def forward(self, visual_features, audio_features):
        B, T = visual_features.shape[0], visual_features.shape[1]

        def runner(ctx, vis, aud):
            v, gv = E.from_external(ctx, vis, vis.requires_grad)
            a, ga = E.from_external(ctx, aud, aud.requires_grad)
            out, setter = _features_out(ctx, _concat_fc_graph(ctx, v, a, ""), B, T, False)
            return [out], [setter], [gv, ga]
        return self._run(runner, visual_features, audio_features)[0]


class Two_transformers(_JmtModule):
    """two_transformers.py:17-128.  Constructor signature, asserts and error behaviour as the reference."""

    def __init__(self, v_dropout: float, a_dropout: float, num_heads: int, num_layers: int, joint_modalities: str,
                 output_format: str = 'FC', vision_in_ft: int = 512, precision: str = "bf16"):
        super().__init__()
        assert isinstance(v_dropout, float), type(v_dropout)
        assert 0.0 <= v_dropout < 1., v_dropout
        self.v_dropout = v_dropout
        assert isinstance(a_dropout, float), type(a_dropout)
        assert 0.0 <= a_dropout < 1., a_dropout
        self.a_dropout = a_dropout
        assert isinstance(num_heads, int), type(num_heads)
        assert num_heads > 0, num_heads
        self.num_heads = num_heads
        assert isinstance(num_layers, int), type(num_layers)
        assert num_layers > 0, num_layers
        self.num_layers = num_layers
        assert isinstance(joint_modalities, str), type(joint_modalities)
        assert joint_modalities in ['NONE', 'TRANSFORMER', 'FC'], joint_modalities
        self.joint_modalities = joint_modalities
        assert isinstance(vision_in_ft, int), type(vision_in_ft)
        assert vision_in_ft > 0, vision_in_ft
        self.vision_in_ft = vision_in_ft
        assert precision in E.PRECISIONS, precision
        self.precision = precision

        self.linear = None
        if vision_in_ft != 512:
            self.linear = _Linear(vision_in_ft, 512)
        assert output_format in ['FC', 'SELF_ATTEN'], output_format
        self.output_format = output_format

        if joint_modalities == 'TRANSFORMER':
            self.mm_transformer = MultimodalTransformer_w_JR(512, 512, num_heads, 512, num_layers, output_format,
                                                             precision)
            dim = 1024 if output_format == 'FC' else 512
        elif joint_modalities == 'FC':
            self.mm_transformer = FeatureConcatFC(512, 512, precision)
            dim = 512
        elif joint_modalities == 'NONE':
            assert output_format in ['FC'], output_format
            self.mm_transformer = MultimodalTransformer_wo_JR(512, 512, num_heads, 512, num_layers, output_format,
                                                              precision)
            dim = 512
        else:
            raise NotImplementedError(joint_modalities)
        self.vregressor = nn.Sequential(_Linear(dim, 128), _Slot(), _Slot(), _Linear(128, 1))
        self.aregressor = nn.Sequential(_Linear(dim, 128), _Slot(), _Slot(), _Linear(128, 1))

    def _dead_prefixes(self):
        return ("mm_transformer.final_encoder.", "mm_transformer.gated_attention.")

    def forward(self, f1_norm, f2_norm):
        """f1_norm: audio (B,T,512); f2_norm: visual (B,T,vision_in_ft) -> (vouts, aouts), each (B,T) --
        or (T,B) for TRANSFORMER+FC exactly as the reference."""
        assert f1_norm.dim() == 3 and f2_norm.dim() == 3
        B, T = f2_norm.shape[0], f2_norm.shape[1]
        assert f1_norm.shape[0] == B and f1_norm.shape[1] == T and f1_norm.shape[2] == 512
        assert f2_norm.shape[2] == self.vision_in_ft

        def runner(ctx, f1, f2):
            video, gv = E.l2norm(ctx, f2, f2.requires_grad)          # :118
            audio, ga = E.l2norm(ctx, f1, f1.requires_grad)          # :119
            outs, setters = _two_transformers_graph(ctx, self, "", video, audio, B, T)
            return outs, setters, [ga, gv]
        v, a = self._run(runner, f1_norm, f2_norm)
        return v, a


def _two_transformers_graph(ctx, mod, prefix, video, audio, B, T):
    """Two_transformers.forward after the L2 normalisation (two_transformers.py:120-128)."""
    if mod.linear is not None:
        video = E.linear(ctx, video, prefix + "linear.weight", prefix + "linear.bias")
    time_major = False
    m = prefix + "mm_transformer."
    if mod.joint_modalities == 'TRANSFORMER':
        feats, time_major = _w_jr_graph(ctx, video, audio, m, mod.num_heads, mod.num_layers, mod.output_format, B, T)
    elif mod.joint_modalities == 'FC':
        feats = _concat_fc_graph(ctx, video, audio, m)
    else:
        feats = _wo_jr_graph(ctx, video, audio, m, mod.num_heads, mod.num_layers, B, T)
    drop_on = ctx.training and (mod.v_dropout > 0.0 or mod.a_dropout > 0.0)
    if not drop_on and os.environ.get("JMT_REGRESSOR_PAIR", "1") != "0":
        # (the reference default, config_file.json:69-70: p = 0) both hidden layers as one N = 256 GEMM
        outs, set_gout = E.regressor_heads(ctx, feats, (prefix + "vregressor.", prefix + "aregressor."), B, T, time_major)
        return outs, [lambda t: set_gout(0, t), lambda t: set_gout(1, t)]
    hv = E.linear(ctx, feats, prefix + "vregressor.0.weight", prefix + "vregressor.0.bias", act=L.ACT_RELU)
    ha = E.linear(ctx, feats, prefix + "aregressor.0.weight", prefix + "aregressor.0.bias", act=L.ACT_RELU)
    hv, _ = E.dropout(ctx, hv, mod.v_dropout)
    ha, _ = E.dropout(ctx, ha, mod.a_dropout)
    outs, set_gout = E.regressor_tail(ctx, [hv, ha], [prefix + "vregressor.3.weight", prefix + "aregressor.3.weight"],
                                      [prefix + "vregressor.3.bias", prefix + "aregressor.3.bias"], [0, 0], B, T, time_major)
    return outs, [lambda t: set_gout(0, t), lambda t: set_gout(1, t)]


class SingleBackbonePretrainer(_JmtModule):
    """two_transformers.py:131-162."""

    def __init__(self, v_dropout: float, a_dropout: float, precision: str = "bf16"):
        super().__init__()
        assert isinstance(v_dropout, float), type(v_dropout)
        assert 0.0 <= v_dropout < 1., v_dropout
        self.v_dropout = v_dropout
        assert isinstance(a_dropout, float), type(a_dropout)
        assert 0.0 <= a_dropout < 1., a_dropout
        self.a_dropout = a_dropout
        self.precision = precision
        self.regressor = nn.Sequential(_Linear(512, 128), _Slot(), _Slot(), _Linear(128, 2))

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert x.ndim == 3, x.ndim
        B, T = x.shape[0], x.shape[1]

        def runner(ctx, xin):
            xv, gx = E.from_external(ctx, xin, xin.requires_grad)
            h = E.linear(ctx, xv, "regressor.0.weight", "regressor.0.bias", act=L.ACT_RELU)
            h, _ = E.dropout(ctx, h, self.a_dropout)
            outs, set_gout = E.regressor_tail(ctx, [h, h], ["regressor.3.weight"] * 2, ["regressor.3.bias"] * 2, [0, 1],
                                              B, T, False)
            return outs, [lambda t: set_gout(0, t), lambda t: set_gout(1, t)], [gx]
        v, a = self._run(runner, x)
        return v, a


class Intra_modal_transformer_fusion(_JmtModule):
    """intra_modal_transformer_fusion.py:74-111: length-2 'sequence' per (b,t), keep the last token."""

    def __init__(self, feat_dim, num_heads, hidden_dim, num_layers, reduce_dim_for_audio=False,
                 precision: str = "bf16"):
        super().__init__()
        assert feat_dim == 512, "final_self_attention is hard-wired to 512 in the reference"
        self.num_heads, self.num_layers, self.precision = num_heads, num_layers, precision
        self.final_visual_encoder = TransformerEncoderBlock(feat_dim, num_heads, hidden_dim, num_layers)
        self.final_self_attention = _MultiheadAttention(512, num_heads)
        self.fc = _Linear(768, 512)

    def forward(self, features_a, features_b):
        B, T = features_a.shape[0], features_a.shape[1]
        M = B * T

        def runner(ctx, fa, fb):
            va, ga = E.from_external(ctx, fa, fa.requires_grad)
            vb, gb = E.from_external(ctx, fb, fb.requires_grad)
            if fa.shape[-1] == 768:
                va = E.linear(ctx, va, "fc.weight", "fc.bias")
            if fb.shape[-1] == 768:
                vb = E.linear(ctx, vb, "fc.weight", "fc.bias")
            st = E.stack_rows(ctx, [va, vb])
            enc = E.encoder_block(ctx, st, "final_visual_encoder.", self.num_heads, self.num_layers, None, small=(2, M))
            fo = E.mha_self(ctx, enc, "final_self_attention.", self.num_heads, None, small=(2, M))
            out, setter = E.to_external(ctx, E.take_rows(ctx, fo, M, 2 * M), (B, T, 512))
            return [out], [setter], [ga, gb]
        return self._run(runner, features_a, features_b)[0]


class FcLayer(_JmtModule):
    """fc_layer.py:6-12."""

    def __init__(self, input_dim, output_dim, precision: str = "bf16"):
        super().__init__()
        self.precision = precision
        self.fc_layer = _Linear(input_dim, output_dim)

    def forward(self, x):
        shape = tuple(x.shape[:-1]) + (self.fc_layer.out_features,)

        def runner(ctx, xin):
            xv, gx = E.from_external(ctx, xin, xin.requires_grad)
            y = E.linear(ctx, xv, "fc_layer.weight", "fc_layer.bias")
            out, setter = E.to_external(ctx, y, shape)
            return [out], [setter], [gx]
        return self._run(runner, x)[0]


# ----------------------------------------------------------------------------- TCN
class _WNConv1d(nn.Module):
    """Parameters of weight_norm(nn.Conv1d(cin, cout, k)): bias, weight_g (cout,1,1), weight_v (cout,cin,k)."""

    def __init__(self, cin, cout, k):
        super().__init__()
        w = torch.empty(cout, cin, k)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.bias = nn.Parameter(torch.empty(cout))
        bound = 1 / math.sqrt(cin * k)
        nn.init.uniform_(self.bias, -bound, bound)
        self.weight_g = nn.Parameter(w.reshape(cout, -1).norm(dim=1).reshape(cout, 1, 1))
        self.weight_v = nn.Parameter(w)


class _Conv1x1(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 1))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        self.bias = nn.Parameter(torch.empty(cout))
        nn.init.uniform_(self.bias, -1 / math.sqrt(cin), 1 / math.sqrt(cin))


class TemporalBlock(nn.Module):
    """Parameters of temporal_convolutional_model.py:21-52 (incl. the `net` aliases in the state_dict)."""

    def __init__(self, n_inputs, n_outputs, kernel_size, stride, dilation, padding, dropout=0.2):
        super().__init__()
        assert stride == 1 and padding == (kernel_size - 1) * dilation
        self.n_inputs, self.n_outputs, self.kernel_size, self.dilation, self.p = n_inputs, n_outputs, kernel_size, dilation, dropout
        self.conv1 = _WNConv1d(n_inputs, n_outputs, kernel_size)
        self.conv2 = _WNConv1d(n_outputs, n_outputs, kernel_size)
        self.net = nn.Sequential(self.conv1, _Slot(), _Slot(), _Slot(), self.conv2, _Slot(), _Slot(), _Slot())
        self.downsample = _Conv1x1(n_inputs, n_outputs) if n_inputs != n_outputs else None
        # init_weights (:45-52): xavier on the derived conv weights has no effect on g/v
        # but consumes RNG; only the downsample init is effective.
        nn.init.xavier_uniform_(torch.empty(n_outputs, n_inputs, kernel_size), gain=math.sqrt(2))
        nn.init.xavier_uniform_(torch.empty(n_outputs, n_outputs, kernel_size), gain=math.sqrt(2))
        if self.downsample is not None:
            nn.init.xavier_uniform_(self.downsample.weight, gain=math.sqrt(2))


class TemporalConvNet(_JmtModule):
    """temporal_convolutional_model.py:60-82.  attention=1 (AttentionBlock) is unreachable in the
    reference (its only call site passes attention=0) and raises NotImplementedError here."""

    def __init__(self, num_inputs, num_channels, kernel_size=2, dropout=0.2, max_length=200, attention=0,
                 precision: str = "bf16"):
        super().__init__()
        if attention == 1:
            raise NotImplementedError("AttentionBlock (attention=1) is out of scope: unreachable in the reference")
        self.precision = precision
        layers = []
        for i in range(len(num_channels)):
            d = 2 ** i
            cin = num_inputs if i == 0 else num_channels[i - 1]
            layers.append(TemporalBlock(cin, num_channels[i], kernel_size, stride=1, dilation=d,
                                        padding=(kernel_size - 1) * d, dropout=dropout))
        self.network = nn.Sequential(*layers)

    def _live_names(self):
        # `net.0/net.4` alias conv1/conv2: named_parameters() already de-duplicates them
        return [n for n, _ in self.named_parameters()]

    def block_specs(self):
        return [(b.n_inputs, b.n_outputs, b.kernel_size, b.dilation, b.p, b.downsample is not None) for b in self.network]

    def forward(self, x):
        """x: (N, C, L) -> (N, C_last, L)."""
        N, C0, Ls = x.shape
        specs = self.block_specs()

        def runner(ctx, xin):
            pad = E.tcn_pad(specs)
            h, gx = E.transpose_in(ctx, xin, xin.requires_grad, pad)       # flat padded channels-last (N*(pad+L), C)
            h = _tcn_graph(ctx, h, "", specs, N, Ls, pad)
            out, setter = E.transpose_out(ctx, h, N, Ls, h.data.shape[1], pad)
            return [out], [setter], [gx]
        return self._run(runner, x)[0]

    def forward_sequence_features(self, x, max_over_time: bool = False):
        """The way the reference consumes the TCN: `temporal(features).transpose(1, 2)` ->
        (N, L, C_last) (I3D_WSDDA.forward, I3DWSDDA.py:44); with max_over_time the `torch.max(ft, 1)` of
        tsav.py:216 is fused in -> (N, C_last).  x: (N, C, L).  No (N, C_last, L) tensor is ever materialised."""
        N, C0, Ls = x.shape
        specs = self.block_specs()

        def runner(ctx, xin):
            pad = E.tcn_pad(specs)
            h, gx = E.transpose_in(ctx, xin, xin.requires_grad, pad)
            h = _tcn_graph(ctx, h, "", specs, N, Ls, pad)
            Cl = h.data.shape[1]
            if max_over_time:
                out, setter = E.to_external(ctx, E.time_max(ctx, h, N, Ls, pad), (N, Cl))
            else:
                out, setter = E.to_external(ctx, E.unpad_rows(ctx, h, N, Ls, pad), (N, Ls, Cl))
            return [out], [setter], [gx]
        return self._run(runner, x)[0]


def _tcn_graph(ctx, h, prefix, specs, N, Ls, pad):
    """TemporalConvNet.forward on the flat padded channels-last layout (row = n*(pad+L) + pad + t, engine.py "TCN ops"):
    per level two weight-normed dilated causal convs (+LeakyReLU, channel dropout), residual (1x1 conv when
    Cin != Cout), LeakyReLU (temporal_convolutional_model.py:54-57, 81-82).  Padding rows stay zero throughout."""
    ks = {k for (_ci, _co, k, _d, _p, _ds) in specs}
    wn = None
    if len(ks) == 1 and 2 * len(specs) <= 16:      # all weight_norm reparametrisations of the net in one launch per layout
        convs = [(f"{prefix}network.{i}.conv{c}.", cout, cin if c == 1 else cout)
                 for i, (cin, cout, _k, _d, _p, _ds) in enumerate(specs) for c in (1, 2)]
        wn = E.weight_norm_all(ctx, convs, next(iter(ks)))
    ps = {p for (_ci, _co, _k, _d, p, _ds) in specs}
    masks = None
    if len(ps) == 1:                               # every Dropout2d keep-mask of the net in one launch
        masks = E.channel_dropout_masks(ctx, N, [cout for (_ci, cout, _k, _d, _p, _ds) in specs for _ in (1, 2)], next(iter(ps)))
    for i, (cin, cout, k, d, p, has_ds) in enumerate(specs):
        pre = f"{prefix}network.{i}."
        ctx.sync_point(pre)          # reached in backward once this level's dW / weight-norm gradients are final
        w1 = w2 = None
        if wn is not None:
            E.weight_norm_bwd_group(ctx, wn, [(pre + "conv1.", cout, cin), (pre + "conv2.", cout, cout)], k)
            w1, w2 = wn[pre + "conv1."], wn[pre + "conv2."]
        m1, m2 = (masks[2 * i], masks[2 * i + 1]) if masks is not None else (None, None)
        y = E.causal_conv(ctx, h, pre + "conv1.", N, Ls, cin, cout, k, d, L.ACT_LEAKY, drop_p=p, pad=pad, weights=w1, fold_act=True,
                          keep_mask=m1)
        y = E.causal_conv(ctx, y, pre + "conv2.", N, Ls, cout, cout, k, d, L.ACT_LEAKY, drop_p=p, pad=pad, weights=w2, keep_mask=m2)
        res = _conv1x1(ctx, h, pre + "downsample.", (Ls + pad, pad)) if has_ds else h
        h = E.add_act(ctx, y, res, L.ACT_LEAKY, E.LEAKY_SLOPE, a_exclusive=True)      # y (conv2's output) has no other consumer
    return h


def _conv1x1(ctx, x, prefix, zero_rows=(0, 0)):
    """nn.Conv1d(cin, cout, 1) on channels-last rows = a Linear whose (cout, cin, 1) weight is viewed 2-D; the
    padding rows of the flat layout must not pick up the bias."""
    return E.linear(ctx, x, prefix + "weight", prefix + "bias", zero_rows=zero_rows)


class JMTPipeline(_JmtModule):
    """The BASELINE.json C2 pipeline as ONE tape (no fp32 round trips between modules):

        visual (B, 1024, T) --TemporalConvNet--> (B, T, 512)   [I3D_WSDDA.forward, I3DWSDDA.py:40-45]
        audio  (B, T, 768)  --FcLayer(768,512)--> (B, T, 512)  [main.py:360, train.py:265]
        Two_transformers(audio, visual) -> (vouts, aouts)      [train.py:287]

    Sub-modules keep their reference names/state_dicts (`fusion`, `fc_audio`, `tcn`); either front-end may
    be None (features are then fed directly).  Inputs may be fp32 or bf16."""

    def __init__(self, fusion: "Two_transformers", fc_audio: Optional["FcLayer"] = None,
                 tcn: Optional["TemporalConvNet"] = None):
        super().__init__()
        self.fusion, self.fc_audio, self.tcn = fusion, fc_audio, tcn
        self.precision = fusion.precision

    def _dead_prefixes(self):
        return ("fusion.mm_transformer.final_encoder.", "fusion.mm_transformer.gated_attention.")

    def forward(self, audio, visual):
        B = audio.shape[0]
        T = audio.shape[1]
        fusion = self.fusion

        def runner(ctx, aud, vis):
            if self.fc_audio is not None:
                a0, ga = E.from_external(ctx, aud, aud.requires_grad)
                a1 = E.linear(ctx, a0, "fc_audio.fc_layer.weight", "fc_audio.fc_layer.bias")
                # two_transformers.py:119 normalises the FcLayer output
                audio_n = _l2norm_var(ctx, a1)
            else:
                audio_n, ga = E.l2norm(ctx, aud, aud.requires_grad)
            if self.tcn is not None:
                N, C0, Ls = vis.shape
                assert N == B and Ls == T
                specs = self.tcn.block_specs()
                pad = E.tcn_pad(specs)
                h, gv = E.transpose_in(ctx, vis, vis.requires_grad, pad)
                h = _tcn_graph(ctx, h, "tcn.", specs, N, Ls, pad)
                # (B*T, 512) == transpose(1, 2), normalised: read straight from / differentiated straight into the padded layout
                video_n = _l2norm_var(ctx, h, seq=(N, Ls, pad)) if pad > 0 else _l2norm_var(ctx, h)
            else:
                video_n, gv = E.l2norm(ctx, vis, vis.requires_grad)
            # Backward runs the tape in reverse: everything recorded after this point (the fusion) is done when this entry
            # executes, so the fusion slice of the gradient bucket is all-reduced while the TCN / FcLayer backward still
            # runs; the TCN levels likewise start theirs as each level's weight gradients complete (_tcn_graph).
            ctx.sync_point("fusion.")
            outs, setters = _two_transformers_graph(ctx, fusion, "fusion.", video_n, audio_n, B, T)
            return outs, setters, [ga, gv]
        v, a = self._run(runner, audio, visual)
        return v, a


def _l2norm_var(ctx, x, seq=None):
    """F.normalize on an activation Var (rows, D) staying in the activation dtype.  seq = (N, L, pad): x is the TCN's flat padded
    layout (N * (pad + L) rows); the result is compact (N * L rows, row = n*L + t) -- `E.unpad_rows` fused in, forward and backward."""
    D = x.data.shape[1]
    if seq is not None:
        N, Ls, pad = seq
        rows = N * Ls
        assert x.data.shape[0] == N * (Ls + pad) and x.data.is_contiguous()
    else:
        rows = x.data.shape[0]
    out = ctx.empty((rows, D))
    inv = ctx.empty((rows,), torch.float32) if ctx.record else None
    if seq is not None:
        L.check(ctx.lib.jmt_l2norm_fwd_seq(E._ptr(x.data), ctx.acode, x.data.stride(0), E._ptr(out), ctx.acode, N, Ls, Ls + pad, pad, D,
                                           1e-12, E._ptr(inv), E._stream()), "jmt_l2norm_fwd_seq")
    else:
        L.check(ctx.lib.jmt_l2norm_fwd(E._ptr(x.data), ctx.acode, x.data.stride(0), E._ptr(out), ctx.acode, rows, D, 1e-12,
                                       E._ptr(inv), E._stream()), "jmt_l2norm_fwd")
    y = E.Var(out)
    if ctx.record:
        def bwd():
            if y.grad is None:
                return
            gb = E.GradBuf(ctx.empty(x.data.shape))      # straight into the activation dtype (no fp32 round trip)
            if seq is not None:
                L.check(ctx.lib.jmt_l2norm_bwd_seq(E._ptr(y.grad), E._ptr(out), ctx.acode, E._ptr(inv), 1e-12, E._ptr(gb.t), ctx.acode,
                                                   N, Ls, Ls + pad, pad, D, E._stream()), "jmt_l2norm_bwd_seq")
            else:
                L.check(ctx.lib.jmt_l2norm_bwd(E._ptr(y.grad), E._ptr(out), ctx.acode, E._ptr(inv), 1e-12, E._ptr(gb.t), ctx.acode, rows, D,
                                               E._stream()), "jmt_l2norm_bwd")
            ctx.add_grad(x, gb)
            gb.refs -= 1
            ctx.release(y)
        ctx.tape.append(bwd)
    return y
