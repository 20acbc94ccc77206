"""ctypes binding of libjmt_b200.so (the C-ABI in include/jmt_b200.h).

The library is built in-tree by ``build.py`` / ``__graft_entry__.build()``.  There is no fallback:
if the shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JMT_B200_LIB") or os.path.join(_HERE, "libjmt_b200.so")   # (override: A/B runs of two builds)

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
MAJOR_K, MAJOR_MN = 0, 1
STORE, ACCUMULATE, ATOMIC_ADD = 0, 1, 2
CCC_METRIC, CCC_LOSS_LIVE, CCC_LOSS_MASKED, CCC_NUMPY = 0, 1, 2, 3


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("d", C.c_void_p), ("bias", C.c_void_p),
        ("a_major", C.c_int32), ("b_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_rows", C.c_int32), ("b_rows", C.c_int32),
        ("a_ld", C.c_int64), ("a_bs0", C.c_int64), ("a_bs1", C.c_int64),
        ("b_ld", C.c_int64), ("b_bs0", C.c_int64), ("b_bs1", C.c_int64),
        ("d_ld", C.c_int64), ("d_bs0", C.c_int64), ("d_bs1", C.c_int64),
        ("nb0", C.c_int32), ("nb1", C.c_int32),
        ("d_dtype", C.c_int32), ("act", C.c_int32),
        ("alpha", C.c_float), ("slope", C.c_float),
        ("store_mode", C.c_int32), ("ntaps", C.c_int32),
        ("a_shift0", C.c_int32), ("a_shift_step", C.c_int32),
        ("b_shift0", C.c_int32), ("b_shift_step", C.c_int32),
        ("reduce_batch", C.c_int32), ("split_k", C.c_int32),
        ("colmask", C.c_void_p), ("colmask_scale", C.c_float),
        ("colmask_row_period", C.c_int32), ("zero_row_period", C.c_int32), ("zero_row_count", C.c_int32),
        ("epi_aux", C.c_void_p), ("aux_slope", C.c_float), ("d_colsum", C.c_void_p), ("colsum_bs0", C.c_int64),
    ]


class AttnDesc(C.Structure):
    """Mirror of jmt_attn_desc (include/jmt_b200.h)."""
    _fields_ = [
        ("a1", C.c_void_p), ("b1", C.c_void_p), ("b2", C.c_void_p), ("p_in", C.c_void_p), ("o_in", C.c_void_p), ("delta_in", C.c_void_p), ("x", C.c_void_p), ("d", C.c_void_p),
        ("mode", C.c_int32), ("Lq", C.c_int32), ("S", C.c_int32), ("dh", C.c_int32), ("heads", C.c_int32), ("NB", C.c_int32),
        ("a1_ld", C.c_int64), ("a1_hs", C.c_int64), ("a1_bs", C.c_int64),
        ("b1_ld", C.c_int64), ("b1_hs", C.c_int64), ("b1_bs", C.c_int64),
        ("b2_ld", C.c_int64), ("b2_hs", C.c_int64), ("b2_bs", C.c_int64),
        ("d_ld", C.c_int64), ("d_hs", C.c_int64), ("d_bs", C.c_int64),
        ("x_ld", C.c_int64), ("scale", C.c_float), ("store_mode", C.c_int32), ("lse_out", C.c_void_p),
    ]


class AttnBwdPart(C.Structure):
    """Mirror of jmt_attn_bwd_part (include/jmt_b200.h)."""
    _fields_ = [
        ("a", C.c_void_p), ("a_ld", C.c_int64), ("a_hs", C.c_int64), ("a_bs", C.c_int64),
        ("x", C.c_void_p), ("x_trans", C.c_int32),
        ("d", C.c_void_p), ("d_ld", C.c_int64), ("d_hs", C.c_int64), ("d_bs", C.c_int64),
        ("store_mode", C.c_int32), ("alpha", C.c_float), ("colsum", C.c_void_p),
    ]


class AttnBwdDesc(C.Structure):
    """Mirror of jmt_attn_bwd_desc (include/jmt_b200.h)."""
    _fields_ = [
        ("part", AttnBwdPart * 3),
        ("Lq", C.c_int32), ("S", C.c_int32), ("dh", C.c_int32), ("heads", C.c_int32), ("NB", C.c_int32),
        ("x_ld", C.c_int64),
    ]


_P, _I, _L, _F, _D, _U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64

# name -> argtypes (restype is always int unless listed in _RESTYPES)
SIGNATURES = {
    "jmt_abi_version": [],
    "jmt_last_error": [],
    "jmt_launch_count": [],
    "jmt_gemm_bf16": [C.POINTER(GemmDesc), _P],
    "jmt_gemm_f32": [C.POINTER(GemmDesc), _P],
    "jmt_gemm_bf16x3": [C.POINTER(GemmDesc), _P, _P, _P],
    "jmt_split_bf16x2": [_P, _P, _P, _L, _P],
    "jmt_gemm_set_profile_buffer": [_P],
    "jmt_gemm_set_bres_mode": [_I],
    "jmt_attn_chain_supported": [C.POINTER(AttnDesc)],
    "jmt_attn_chain_bf16": [C.POINTER(AttnDesc), _P],
    "jmt_attn_set_profile_buffer": [_P],
    "jmt_attn_bwd_dqkv_supported": [C.POINTER(AttnBwdDesc)],
    "jmt_attn_bwd_dqkv_bf16": [C.POINTER(AttnBwdDesc), _P],
    "jmt_attn_bwd_set_profile_buffer": [_P],
    "jmt_attn_merge": [_P, _L, _L, _L, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "jmt_rowdot_bf16": [_P, _P, _L, _L, _L, _I, _I, _I, _I, _P, _P],
    "jmt_l2norm_fwd": [_P, _I, _L, _P, _I, _L, _I, _F, _P, _P],
    "jmt_l2norm_bwd": [_P, _P, _I, _P, _F, _P, _I, _L, _I, _P],
    "jmt_l2norm_fwd_seq": [_P, _I, _L, _P, _I, _L, _I, _I, _I, _I, _F, _P, _P],
    "jmt_l2norm_bwd_seq": [_P, _P, _I, _P, _F, _P, _I, _L, _I, _I, _I, _I, _P],
    "jmt_add_layernorm_fwd": [_P, _P, _P, _P, _F, _P, _P, _P, _L, _I, _I, _P],
    "jmt_add_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _L, _I, _I, _P],
    "jmt_softmax_fwd": [_P, _L, _P, _I, _L, _L, _I, _P],
    "jmt_softmax_bwd": [_P, _I, _L, _P, _L, _P, _I, _L, _L, _I, _P],
    "jmt_attn_small_fwd": [_P, _P, _P, _I, _L, _I, _I, _F, _I, _P],
    "jmt_attn_small_bwd": [_P, _P, _P, _P, _I, _L, _I, _I, _F, _I, _P],
    "jmt_regressor_tail_fwd": [_I, _P, _L, _I, _P, _P, _P, _L, _L, _L, _L, _P],
    "jmt_regressor_tail_bwd": [_I, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _L, _P],
    "jmt_act_bwd": [_P, _P, _P, _L, _F, _I, _P],
    "jmt_colsum": [_P, _I, _L, _L, _I, _P, _P],
    "jmt_act_bwd_fused": [_P, _P, _P, _P, _L, _I, _I, _F, _F, _P, _I, _P],
    "jmt_add_act_bwd_fused": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _F, _F, _F, _P, _I, _P],
    "jmt_cast": [_P, _I, _P, _I, _L, _P],
    "jmt_axpy": [_P, _P, _F, _L, _I, _P],
    "jmt_cast_multi": [_I, _P, _P, _P, _P],
    "jmt_copy2d": [_P, _I, _L, _P, _I, _L, _L, _I, _P],
    "jmt_transpose": [_P, _I, _P, _I, _L, _I, _I, _P],
    "jmt_transpose_strided": [_P, _I, _L, _P, _I, _L, _L, _I, _I, _P],
    "jmt_copy_rows3d": [_P, _I, _L, _P, _I, _L, _L, _L, _I, _P],
    "jmt_copy3d": [_P, _I, _L, _L, _P, _I, _L, _L, _L, _L, _I, _P],
    "jmt_add_act": [_P, _P, _P, _L, _I, _F, _I, _P],
    "jmt_time_max_fwd": [_P, _L, _L, _I, _I, _P, _P, _I, _P],
    "jmt_time_max_bwd": [_P, _P, _L, _L, _I, _P, _I, _P],
    "jmt_apply_mask": [_P, _P, _P, _L, _I, _I, _I, _F, _I, _P],
    "jmt_dropout_mask": [_P, _L, _F, _U64, _U64, _P, _P],
    "jmt_rng_advance": [_P, _U64, _P],
    "jmt_weight_norm_fwd": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _P],
    "jmt_weight_norm_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "jmt_weight_norm_fwd_batched": [_I, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P],
    "jmt_weight_norm_bwd_batched": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "jmt_ccc_sums": [_P, _P, _L, _I, _L, _I, _F, _P, _P],
    "jmt_ccc_finalize": [_P, _I, _I, _D, _D, _P, _P, _P],
    "jmt_ccc_bwd": [_P, _P, _L, _I, _L, _P, _P, _I, _I, _F, _P, _P],
    "jmt_label_mask": [_P, _L, _F, _P, _P],
    "jmt_valpost_scatter": [_P, _P, _P, _P, _P, _P, _L, _P, _F, _U64, _P, _P, _P, _P, _P, _P],
    "jmt_valpost_finalize": [_P, _P, _P, _P, _P, _I, _L, _I, _I, _P, _P, _P, _P],
    "jmt_pad_right_align": [_P, _L, _I, _P, _I, _P],
}
_RESTYPES = {"jmt_last_error": C.c_char_p, "jmt_launch_count": C.c_int64}

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python build.py` (nvcc, sm_100a). "
                "jmt_b200 has no CPU / PyTorch fallback.")
        h = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(h, name)           # AttributeError if a declared symbol is not exported
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        if h.jmt_abi_version() != 9:
            raise RuntimeError("libjmt_b200.so ABI version mismatch")
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().jmt_last_error()
        raise RuntimeError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def launch_count():
    return int(lib().jmt_launch_count())
