"""Tape engine: forward/backward orchestration of the JMT hot path over the C-ABI kernels.

PyTorch supplies device memory (caching allocator), the current stream and the outer autograd
hook; every byte and FLOP of the path itself runs in libjmt_b200.so.  The engine keeps its own
reverse-mode tape so that (a) no ATen kernel runs on the path, (b) gradient accumulation is
fused into kernel epilogues (GEMM ``ACCUMULATE`` stores), (c) parameter gradients land in ONE
flat fp32 bucket (the NCCL all-reduce unit) and (d) the step is CUDA-graph
capturable (no host synchronisation anywhere).

Precision modes (SURVEY.md 7 "hard parts" #4):
  'bf16'   : bf16 operands / activations, fp32 accumulate (tcgen05 path), fp32 LN/softmax statistics
  'bf16x3' : fp32 activations; every GEMM operand is split into two bf16 parts (x = hi + lo) and the tcgen05 kernel issues
             three MMAs per k-step (hi*hi + hi*lo + lo*hi) into one fp32 TMEM accumulator -- the tensor-core mode that
             meets the 1e-3 prediction / 1e-4 CCC parity gate of north_star
  'fp32'   : fp32 operands / activations on the FFMA GEMM -- kept as an independent on-device cross-check only
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}
LEAKY_SLOPE = 0.01
PRECISIONS = ("bf16", "bf16x3", "fp32")
_PARAM_GENERATION = [0]


def bump_param_generation():
    """Invalidate every cached bf16 operand copy of a parameter: call after parameters were changed behind autograd's back
    (a CUDA-graph replay containing optimizer.step(): in-place `_version` counters do not move)."""
    _PARAM_GENERATION[0] += 1


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    # raw handle of the current stream of the current device (torch.cuda.current_stream() builds a Stream object through several
    # Python layers: ~15 us a call, 3 ms of an eagerly launched step)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("jmt_b200 runs on CUDA tensors only (there is no CPU fallback)")


# --------------------------------------------------------------------------- CUDA runtime memset
_cudart = None


def cuda_memset0(t: torch.Tensor):
    """Stream-ordered zero fill through cudaMemsetAsync (copy engine, not an ATen kernel)."""
    global _cudart
    if t.numel() == 0:
        return
    if _cudart is None:
        import glob
        import os
        cands = sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*"))) + \
            sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))
        lib = None
        for c in cands:
            try:
                lib = C.CDLL(c)
                break
            except OSError:
                continue
        if lib is None:
            raise RuntimeError("libcudart not found")
        lib.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
        lib.cudaMemsetAsync.restype = C.c_int
        _cudart = lib
    assert t.is_contiguous()
    rc = _cudart.cudaMemsetAsync(C.c_void_p(t.data_ptr()), 0, t.numel() * t.element_size(), _stream())
    if rc != 0:
        raise RuntimeError(f"cudaMemsetAsync failed ({rc})")


class GradBuf:
    """A gradient buffer.  ``written`` is None when every column holds a valid value; otherwise it lists the column
    ranges written so far of a buffer that was allocated WITHOUT zero-filling (projection gradients whose Q | K | V
    column slices are each produced by one attention-backward GEMM: the first write of a slice is a plain store, so the
    236 MB memset and the read-modify-write of an accumulate are skipped)."""
    __slots__ = ("t", "refs", "written")

    def __init__(self, t, written=None):
        self.t = t
        self.refs = 1
        self.written = written


class Var:
    """An activation on the tape: ``data`` is a 2-D (rows, cols) view with unit inner stride.

    Backward-epilogue bookkeeping (bf16 mode, jmt_gemm_desc.epi_aux / d_colsum):
      track_colsum : this Var is the output of a Linear / conv with a bias: GEMMs that write its gradient also add the column
                     sums of what they write into ``colsum_tmp`` (= the bias gradient) -- valid as long as EVERY writer did
                     (``colsum_ok``); otherwise the producer falls back to a column-sum pass over the finished gradient.
      colsum_direct: the graph GUARANTEES that every writer is such a GEMM (projections feeding the GEMM-based attention backward,
                     folded activations): the column sums go straight into the bias gradient, no accumulator, no extra launch.
      fold         : (slope, mask, mask_scale, mask_rows) when this Var is act(pre) [* channel-dropout mask] and its only
                     consumer is a GEMM: that GEMM's dgrad epilogue multiplies by act'(data) (and the mask), so the gradient
                     buffer holds d(pre) (``folded``) and no act-bwd pass runs."""
    __slots__ = ("data", "gbuf", "needs_grad", "track_colsum", "colsum_tmp", "colsum_ok", "fold", "folded", "colsum_direct",
                 "act_pending", "act_done")

    def __init__(self, data: torch.Tensor, needs_grad: bool = True):
        self.data = data
        self.gbuf: Optional[GradBuf] = None
        self.needs_grad = needs_grad
        self.track_colsum = False
        self.colsum_tmp: Optional[torch.Tensor] = None
        self.colsum_ok = True
        self.fold: Optional[tuple] = None
        self.folded = False
        self.colsum_direct: Optional[Callable[[], torch.Tensor]] = None   # getter of the bias-gradient slice the writers add into
        # (slope, mask, mask_rows, mask_scale, bias-gradient getter) of a conv output whose activation / dropout backward may be taken
        # over by its only consumer (add_act, a_exclusive): act_done is then set and the gradient buffer holds d(pre-activation)
        self.act_pending: Optional[tuple] = None
        self.act_done = False

    @property
    def grad(self) -> Optional[torch.Tensor]:
        if self.gbuf is None:
            return None
        w = self.gbuf.written
        if w is not None:                 # slice-wise buffer: it may only be consumed once every column has been written
            cols, pos = self.gbuf.t.shape[-1], 0
            for c0, c1 in sorted(w):
                if c0 > pos:
                    break
                pos = max(pos, c1)
            if pos < cols:
                raise RuntimeError(f"gradient buffer consumed with unwritten columns [{pos}, ...) of {cols}")
            self.gbuf.written = None
        return self.gbuf.t


class Ctx:
    """One forward(+backward) pass: precision, parameter access, tape, flat gradient bucket."""

    def __init__(self, params: Dict[str, torch.Tensor], precision: str, record: bool, training: bool,
                 wcache: Optional[dict] = None, seed: int = 0, rng_state: Optional[torch.Tensor] = None):
        assert precision in PRECISIONS, precision
        self.params = params
        self.precision = precision
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.x3 = precision == "bf16x3"
        self.x3_const: List[Tuple[int, int]] = []      # address ranges of parameters (their splits are reused within this pass)
        self.x3_cache: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor]] = {}
        self.acode = _DT[self.adt]
        self.record = record
        self.training = training
        self.tape: List[Callable[[], None]] = []
        self.lib = L.lib()
        self.wcache = wcache if wcache is not None else {}
        self._w: Dict[str, torch.Tensor] = {}
        self._bulk_done = False
        self.pgrads: Dict[str, torch.Tensor] = {}
        self.bucket: Optional[torch.Tensor] = None
        self.seed = seed
        self.rng_offset = 0
        self.rng_state = rng_state          # device int64[2]: graph-replay-safe RNG state (see jmt_dropout_mask)
        self.qcache: Dict[Tuple[str, int], Var] = {}
        self.dev = next(iter(params.values())).device if params else torch.device("cuda", torch.cuda.current_device())
        self.keep: List[object] = []          # host arrays that must outlive async launches
        self.bucket_ranges: Dict[str, Tuple[int, int]] = {}
        self.synced: List[Tuple[int, int]] = []   # bucket ranges whose all-reduce has already been started (overlap)
        self.grad_sync = None                 # the owner's GradSync hook (set by the autograd bridge)

    # ---------------------------------------------------------------- memory helpers
    def empty(self, shape, dtype=None) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype or self.adt, device=self.dev)

    def zeros(self, shape, dtype=None) -> torch.Tensor:
        t = torch.empty(shape, dtype=dtype or self.adt, device=self.dev)
        cuda_memset0(t)
        return t

    # ---------------------------------------------------------------- parameters
    def p(self, name: str) -> torch.Tensor:
        return self.params[name]

    def w(self, name: str) -> torch.Tensor:
        """Parameter as a GEMM operand in the activation dtype (cast once per parameter version)."""
        par = self.params[name]
        if self.adt == torch.float32:
            if self.x3 and name not in self._w:
                self._w[name] = par
                self.x3_const.append((par.data_ptr(), par.data_ptr() + par.numel() * 4))
            return par
        t = self._w.get(name)
        if t is None:
            self._cast_all_weights()
            t = self._w.get(name)
        if t is None:               # not a bulk-cast candidate (e.g. a 1-D parameter used as an operand): cast it alone
            out = torch.empty(par.shape, dtype=self.adt, device=par.device)
            L.check(self.lib.jmt_cast(_ptr(par), L.F32, _ptr(out), self.acode, par.numel(), _stream()), "jmt_cast")
            self._w[name] = t = out
        return t

    def _cast_all_weights(self):
        """bf16 operand copies of ALL weight matrices of this pass in one launch (jmt_cast_multi) into one buffer.  The cached
        buffer is valid for one (storages, in-place versions, replay generation): a CUDA-graph replay that contains
        optimizer.step() updates the parameters on the device WITHOUT bumping `_version`, so every replay
        (GraphedStep.replay) bumps the generation instead.  Under capture the cast is always PART of the graph (a replay runs
        after optimizer steps the version check at capture time cannot see) and stays private to it: a graph-pool tensor has no
        content before the first replay and must never be handed to a later eager forward."""
        if self._bulk_done:
            return
        self._bulk_done = True
        names = [n for n, p in self.params.items() if p.dim() >= 2 and not n.endswith("weight_v") and not n.endswith("weight_g")]
        if not names:
            return
        pars = [self.params[n] for n in names]
        key = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in pars) + (_PARAM_GENERATION[0],)
        capturing = torch.cuda.is_current_stream_capturing()
        ent = None if capturing else self.wcache.get("__bulk__")
        if ent is None or ent[0] != key:
            offs, total = [], 0
            for p in pars:
                offs.append(total)
                total += (p.numel() + 63) // 64 * 64
            buf = torch.empty((total,), dtype=self.adt, device=pars[0].device)
            views = [buf[o:o + p.numel()].view(p.shape) for o, p in zip(offs, pars)]
            n = len(pars)
            src = (C.c_void_p * n)(*[p.data_ptr() for p in pars])
            dst = (C.c_void_p * n)(*[v.data_ptr() for v in views])
            cnt = (C.c_int64 * n)(*[p.numel() for p in pars])
            L.check(self.lib.jmt_cast_multi(n, src, dst, cnt, _stream()), "jmt_cast_multi")
            ent = (key, dict(zip(names, views)), buf)
            if not capturing:
                self.wcache["__bulk__"] = ent
        self._w.update(ent[1])

    def cat_params(self, names: Sequence[str], as_operand: bool) -> torch.Tensor:
        """Row-wise concatenation of same-width parameters in one buffer (operand dtype when `as_operand`, else fp32): lets
        independent Linears that read the same input run as ONE GEMM (the valence | arousal regressor heads,
        two_transformers.py:104-114).  Cached like w() (per storage / version / replay generation, never across a capture)."""
        key_name = "|".join(names) + ("|op" if as_operand else "|f32")
        t = self._w.get(key_name)
        if t is not None:
            return t
        pars = [self.params[n] for n in names]
        dt = self.adt if as_operand else torch.float32
        key = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in pars) + (_PARAM_GENERATION[0],)
        capturing = torch.cuda.is_current_stream_capturing()
        ent = None if capturing else self.wcache.get(key_name)
        if ent is None or ent[0] != key:
            rows = sum(p.shape[0] for p in pars)
            out = torch.empty((rows,) + tuple(pars[0].shape[1:]), dtype=dt, device=pars[0].device)
            r0 = 0
            for p in pars:
                L.check(self.lib.jmt_cast(_ptr(p), L.F32, _ptr(out[r0:r0 + p.shape[0]]), _DT[dt], p.numel(), _stream()), "jmt_cast")
                r0 += p.shape[0]
            ent = (key, out)
            if not capturing:
                self.wcache[key_name] = ent
        t = ent[1]
        self._w[key_name] = t
        if self.x3 and as_operand:
            self.x3_const.append((t.data_ptr(), t.data_ptr() + t.numel() * 4))
        return t

    def prepare_param_grads(self, names: Sequence[str]):
        """Allocate the flat fp32 gradient bucket (zeroed) with one 256-byte aligned view per parameter."""
        total = 0
        offs = []
        for n in names:
            k = self.params[n].numel()
            offs.append((n, total, k))
            total += (k + 63) // 64 * 64
        self.bucket = self.zeros((max(total, 64),), torch.float32)
        self.bucket_ranges = {n: (o, (k + 63) // 64 * 64) for n, o, k in offs}
        for n, o, k in offs:
            self.pgrads[n] = self.bucket[o:o + k].view(self.params[n].shape)

    def pgrad(self, name: str) -> torch.Tensor:
        return self.pgrads[name]

    def sync_point(self, prefix: str):
        """Record a tape entry that -- when backward reaches it, i.e. after everything recorded AFTER this call has run its
        backward -- starts the asynchronous all-reduce of the bucket slice holding the parameters named `prefix*` (their
        gradients are final then; the slice is contiguous because parameters are laid out in registration order).  NCCL runs
        it on its own stream under the rest of the backward pass."""
        if not self.record:
            return

        def start():
            hook = self.grad_sync
            if hook is None or not hasattr(hook, "start") or os.environ.get("JMT_GRAD_OVERLAP", "1") == "0":
                return
            mine = [(o, o + k) for n, (o, k) in self.bucket_ranges.items() if n.startswith(prefix)]
            if not mine:
                return
            lo, hi = min(a for a, _ in mine), max(b for _, b in mine)
            other = [(o, o + k) for n, (o, k) in self.bucket_ranges.items() if not n.startswith(prefix)]
            if any(a < hi and lo < b for a, b in other) or any(a < hi and lo < b for a, b in self.synced):
                return                                    # not contiguous / already in flight: leave it to the final pass
            hook.start(self.bucket[lo:hi])
            self.synced.append((lo, hi))
        self.tape.append(start)

    def unsynced_ranges(self) -> List[Tuple[int, int]]:
        """Bucket ranges not covered by an all-reduce started at a sync_point."""
        total = self.bucket.numel() if self.bucket is not None else 0
        out, pos = [], 0
        for lo, hi in sorted(self.synced):
            if lo > pos:
                out.append((pos, lo))
            pos = max(pos, hi)
        if pos < total:
            out.append((pos, total))
        return out

    # ---------------------------------------------------------------- gradient plumbing
    def ext_on(self) -> bool:
        """backward-epilogue extensions (activation-gradient fold, bias-gradient column sums) are a bf16 tensor-core feature"""
        return self.precision == "bf16" and EPI_EXT

    def colsum_target(self, v: Var, c0: int = 0, c1: Optional[int] = None) -> Optional[torch.Tensor]:
        """fp32 accumulator slice for the column sums of a gradient contribution to v[:, c0:c1] written by a capable GEMM
        (None: not tracked).  Callers must have obtained the gradient buffer with capable=True."""
        if not (v.track_colsum and self.ext_on() and (EPI_COLSUM or v.colsum_direct is not None)):
            return None
        c1 = c1 if c1 is not None else v.data.shape[1]
        if v.colsum_direct is not None:
            return v.colsum_direct()[c0:c1]
        if v.colsum_tmp is None:
            v.colsum_tmp = self.zeros((v.data.shape[1],), torch.float32)
        return v.colsum_tmp[c0:c1]

    def _incapable_writer(self, v: Var):
        """a writer of d(v) that cannot fold the activation gradient / emit column sums"""
        v.colsum_ok = False
        if v.fold is not None or v.colsum_direct is not None:
            raise RuntimeError("jmt_b200 engine: an activation whose gradient must be written by GEMMs only (fold_act / "
                               "gemm_writers_only) got another kind of consumer")

    def grad_target(self, v: Var, capable: bool = False) -> Tuple[torch.Tensor, int]:
        """Tensor to write d(v) into and the store mode (STORE the first time, ACCUMULATE after).  capable: the caller is a
        GEMM that honours v.fold and adds its column sums to colsum_target(v)."""
        if not (capable and self.ext_on()):
            self._incapable_writer(v)
        if v.gbuf is None:
            v.gbuf = GradBuf(self.empty(v.data.shape, v.data.dtype))
            return v.gbuf.t, L.STORE
        if v.gbuf.refs > 1:                       # buffer aliased by another Var: copy-on-write
            old = v.gbuf
            new = self.empty(old.t.shape, old.t.dtype)
            copy2d(self, old.t, new)
            old.refs -= 1
            v.gbuf = GradBuf(new, None if old.written is None else list(old.written))
        return v.gbuf.t, L.ACCUMULATE

    def add_grad(self, v: Var, gb: GradBuf):
        """d(v) += gb.  A Var without gradient aliases the buffer (ref-counted) instead of copying."""
        if not v.needs_grad:
            return
        self._incapable_writer(v)
        if v.gbuf is None:
            gb.refs += 1
            v.gbuf = gb
            return
        t, _ = self.grad_target(v)
        g = gb.t
        assert g.is_contiguous() and t.is_contiguous()
        L.check(self.lib.jmt_axpy(_ptr(g), _ptr(t), 1.0, g.numel(), _DT[g.dtype], _stream()), "jmt_axpy")

    def release(self, v: Var):
        if v.gbuf is not None:
            v.gbuf.refs -= 1
            v.gbuf = None

    def finish_forward(self):
        """Advance the device RNG counter past everything this forward drew (stream-ordered after its last mask)."""
        if self.rng_state is not None and self.rng_offset > 0:
            L.check(self.lib.jmt_rng_advance(_ptr(self.rng_state), self.rng_offset, _stream()), "jmt_rng_advance")

    def backward(self):
        for fn in reversed(self.tape):
            fn()
        self.tape = []


# --------------------------------------------------------------------------- raw op wrappers
PROFILE: Optional[list] = None      # when a list: gemm() appends (kind, algorithmic flops, start event, end event)


def gemm_flops(M, N, K, nb, ntaps, a_shift, b_shift, a_rows, b_rows, a_major, b_major) -> float:
    """Algorithmic FLOPs (multiply-add = 2) of one launch; taps count only rows whose shifted index is in
    range (useful taps)."""
    if ntaps == 1 and a_shift == (0, 0) and b_shift == (0, 0):
        return 2.0 * M * N * K * nb
    total = 0.0
    for j in range(ntaps):
        ash = a_shift[0] + j * a_shift[1]
        bsh = b_shift[0] + j * b_shift[1]
        if a_major == L.MAJOR_K:      # shift moves the output row m
            valid_m = max(0, min(M, a_rows - ash) - max(0, -ash))
            valid_k = K
        else:                         # shift moves the reduction index k
            valid_m = M
            valid_k = max(0, min(K, a_rows - ash) - max(0, -ash))
        if b_major == L.MAJOR_MN and bsh != 0:
            valid_k = min(valid_k, max(0, min(K, b_rows - bsh) - max(0, -bsh)))
        total += 2.0 * valid_m * N * valid_k * nb
    return total


def gemm(ctx: Ctx, a: torch.Tensor, b: torch.Tensor, d: torch.Tensor, *, M: int, N: int, K: int,
         a_major=L.MAJOR_K, b_major=L.MAJOR_K, a_rows=None, b_rows=None, a_ld=None, b_ld=None, d_ld=None,
         nb0=1, nb1=1, a_bs=(0, 0), b_bs=(0, 0), d_bs=(0, 0), bias=None, act=L.ACT_NONE, slope=0.0, alpha=1.0,
         store=L.STORE, ntaps=1, a_shift=(0, 0), b_shift=(0, 0), reduce_batch=False, split_k=1,
         colmask=None, colmask_scale=1.0, colmask_row_period=0, zero_rows=(0, 0), alg_flops=None,
         epi_aux=None, aux_slope=0.0, d_colsum=None, colsum_bs0=0):
    """One GEMM of the family in include/jmt_b200.h.  a/b/d supply base pointers and dtypes (views
    allowed); strides are in elements."""
    require_cuda(a, b, d)
    g = L.GemmDesc()
    g.a, g.b, g.d = a.data_ptr(), b.data_ptr(), d.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.a_major, g.b_major = a_major, b_major
    g.M, g.N, g.K = M, N, K
    g.a_rows = a_rows if a_rows is not None else (M if a_major == L.MAJOR_K else K)
    g.b_rows = b_rows if b_rows is not None else (N if b_major == L.MAJOR_K else K)
    g.a_ld = a_ld if a_ld is not None else a.stride(-2)
    g.b_ld = b_ld if b_ld is not None else b.stride(-2)
    g.d_ld = d_ld if d_ld is not None else d.stride(-2)
    g.a_bs0, g.a_bs1 = a_bs
    g.b_bs0, g.b_bs1 = b_bs
    g.d_bs0, g.d_bs1 = d_bs
    g.nb0, g.nb1 = nb0, nb1
    g.d_dtype = _DT[d.dtype]
    g.act, g.alpha, g.slope = act, alpha, slope
    g.store_mode = store
    g.ntaps = ntaps
    g.a_shift0, g.a_shift_step = a_shift
    g.b_shift0, g.b_shift_step = b_shift
    g.reduce_batch = 1 if reduce_batch else 0
    g.split_k = split_k
    g.colmask = colmask.data_ptr() if colmask is not None else None
    g.colmask_scale = colmask_scale
    g.colmask_row_period = colmask_row_period
    g.zero_row_period, g.zero_row_count = zero_rows
    g.epi_aux = epi_aux.data_ptr() if epi_aux is not None else None
    g.aux_slope = aux_slope
    g.d_colsum = d_colsum.data_ptr() if d_colsum is not None else None
    g.colsum_bs0 = colsum_bs0
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16:
        L.check(ctx.lib.jmt_gemm_bf16(C.byref(g), _stream()), "jmt_gemm_bf16")
        kind = "gemm_tc_kernel"
    elif ctx.x3 and a.dtype == torch.float32 and b.dtype == torch.float32:
        # bf16x3: split both fp32 operands over the address span the descriptor can touch (same strides for hi / lo)
        a_inner = K if a_major == L.MAJOR_K else M
        b_inner = ntaps * K if b_major == L.MAJOR_K else N
        a_hi, a_lo = _split_x3(ctx, a, (nb1 - 1) * g.a_bs1 + (nb0 - 1) * g.a_bs0 + (g.a_rows - 1) * g.a_ld + a_inner)
        b_hi, b_lo = _split_x3(ctx, b, (nb1 - 1) * g.b_bs1 + (nb0 - 1) * g.b_bs0 + (g.b_rows - 1) * g.b_ld + b_inner)
        g.a, g.b = a_hi.data_ptr(), b_hi.data_ptr()
        L.check(ctx.lib.jmt_gemm_bf16x3(C.byref(g), _ptr(a_lo), _ptr(b_lo), _stream()), "jmt_gemm_bf16x3")
        kind = "gemm_tc_kernel"
    elif a.dtype == torch.float32 and b.dtype == torch.float32:
        L.check(ctx.lib.jmt_gemm_f32(C.byref(g), _stream()), "jmt_gemm_f32")
        kind = "gemm_simt_kernel"
    else:
        raise RuntimeError(f"gemm operand dtypes {a.dtype}/{b.dtype}")
    if prof is not None:
        e1.record()
        fl = alg_flops if alg_flops is not None else gemm_flops(M, N, K, nb0 * nb1, ntaps, tuple(a_shift), tuple(b_shift),
                                                                g.a_rows, g.b_rows, a_major, b_major)
        prof.append((kind, fl, e0, e1, (M, N, K, nb0 * nb1, ntaps)))


def _split_x3(ctx: Ctx, t: torch.Tensor, span: int):
    """(hi, lo) bf16 buffers of `span` elements with hi[i] + lo[i] ~= t.flat[i] for the elements from t's first one on
    (jmt_split_bf16x2); element offsets -- hence ld / batch strides -- are those of the fp32 operand.  Splits of parameters are
    reused for the rest of the pass (forward GEMM and dgrad read the same weight)."""
    ptr = t.data_ptr()
    const = any(lo <= ptr < hi for lo, hi in ctx.x3_const)
    if const:
        hit = ctx.x3_cache.get((ptr, span))
        if hit is not None:
            return hit
    hi = torch.empty((span,), dtype=torch.bfloat16, device=t.device)
    lo = torch.empty((span,), dtype=torch.bfloat16, device=t.device)
    L.check(ctx.lib.jmt_split_bf16x2(C.c_void_p(ptr), _ptr(hi), _ptr(lo), span, _stream()), "jmt_split_bf16x2")
    if const:
        ctx.x3_cache[(ptr, span)] = (hi, lo)
    return hi, lo


def copy2d(ctx: Ctx, src: torch.Tensor, dst: torch.Tensor):
    """dst[r, c] = cast(src[r, c]) for 2-D views with unit inner stride."""
    src = src.reshape(-1, src.shape[-1]) if src.dim() != 2 else src
    dst = dst.reshape(-1, dst.shape[-1]) if dst.dim() != 2 else dst
    rows, cols = src.shape
    assert src.stride(1) == 1 and dst.stride(1) == 1 and tuple(dst.shape) == (rows, cols)
    L.check(ctx.lib.jmt_copy2d(_ptr(src), _DT[src.dtype], src.stride(0), _ptr(dst), _DT[dst.dtype], dst.stride(0),
                               rows, cols, _stream()), "jmt_copy2d")


def split_k_waves(red: int, tiles: int, units: int = 74) -> int:
    """Split factor for a weight-gradient GEMM (`tiles` output tiles, `units` CTAs / CTA pairs on the chip): the split whose
    tiles*split fills whole waves best, lightly penalised per split (every split is one more fp32 reduce-add pass over the
    output), keeping at least 8 k-blocks per split."""
    kblocks = max(1, (red + 63) // 64)
    best, best_score = 1, -1.0
    for sk in range(1, max(1, min(64, kblocks // 8)) + 1):
        n = tiles * sk
        score = n / (units * ((n + units - 1) // units)) - 0.004 * sk
        if score > best_score:
            best, best_score = sk, score
    return best


def wgrad_split(red: int, m_out: int, n_out: int, batch: int = 1) -> int:
    """split_k for dW (m_out x n_out, `batch` independent outputs) = dy^T x over `red` rows, mirroring jmt_gemm_bf16's tiling
    (csrc/gemm_tc.cu): CTA pairs own 256 x 256 tiles, or 256 x 512 "wide" tiles when n_out is a multiple of 512; an odd number
    (< 9) of 128-row tiles runs unpaired 128 x 256 tiles on all 148 CTAs."""
    m_tiles = (m_out + 127) // 128
    if m_tiles >= 2 and (m_tiles % 2 == 0 or m_tiles >= 9):
        n_tiles = n_out // 512 if n_out % 512 == 0 else (n_out + 255) // 256
        return split_k_waves(red, ((m_tiles + 1) // 2) * n_tiles * batch, 74)
    return split_k_waves(red, m_tiles * ((n_out + 255) // 256) * batch, 148)


# --------------------------------------------------------------------------- differentiable ops
def from_external(ctx: Ctx, x: torch.Tensor, needs_grad: bool):
    """External (.., D) tensor -> activation-dtype Var (cast copy).  Returns (Var, grad getter)."""
    require_cuda(x)
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1 or x2.dtype not in _DT:
        raise RuntimeError("inputs must be fp32/bf16 with a contiguous feature dimension")
    if x2.dtype == ctx.adt and x2.is_contiguous() and not needs_grad:
        # already in the activation dtype (bf16 feature shards): read in place, no copy.  Like any tensor autograd saves for
        # backward, the caller must not overwrite it between forward and backward.
        return Var(x2, False), None
    out = ctx.empty(x2.shape)
    copy2d(ctx, x2, out)
    v = Var(out, needs_grad)
    holder: dict = {}
    if ctx.record and needs_grad:
        def bwd():
            dx = ctx.empty(x2.shape, torch.float32)
            if v.grad is None:
                cuda_memset0(dx)
            else:
                copy2d(ctx, v.grad, dx)
            ctx.release(v)
            holder["dx"] = dx
        ctx.tape.append(bwd)
        return v, (lambda: holder["dx"])
    return v, None


def l2norm(ctx: Ctx, x: torch.Tensor, needs_grad: bool):
    """F.normalize over the last dim of an external (.., D) tensor -> activation-dtype Var.
    Returns (Var, getter of d(input) in fp32 valid after backward, or None)."""
    require_cuda(x)
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1 or x2.dtype not in _DT:
        raise RuntimeError("inputs must be fp32/bf16 with a contiguous feature dimension")
    rows, D = x2.shape
    out = ctx.empty((rows, D))
    rec = ctx.record and needs_grad
    inv = ctx.empty((rows,), torch.float32) if rec else None
    L.check(ctx.lib.jmt_l2norm_fwd(_ptr(x2), _DT[x2.dtype], x2.stride(0), _ptr(out), ctx.acode, rows, D, 1e-12,
                                   _ptr(inv), _stream()), "jmt_l2norm_fwd")
    v = Var(out, needs_grad)
    holder: dict = {}
    if rec:
        def bwd():
            dx = ctx.empty((rows, D), torch.float32)
            if v.grad is None:
                cuda_memset0(dx)
            else:
                L.check(ctx.lib.jmt_l2norm_bwd(_ptr(v.grad), _ptr(out), ctx.acode, _ptr(inv), 1e-12, _ptr(dx), L.F32, rows, D,
                                               _stream()), "jmt_l2norm_bwd")
            ctx.release(v)
            holder["dx"] = dx
        ctx.tape.append(bwd)
        return v, (lambda: holder["dx"])
    return v, None


def linear(ctx: Ctx, x: Var, wname: str, bname: Optional[str], act=L.ACT_NONE, slope=0.0,
           out: Optional[torch.Tensor] = None, w_rows: Optional[Tuple[int, int]] = None,
           w_cols: Optional[Tuple[int, int]] = None, accumulate_into: Optional[Var] = None,
           grad_from: Optional[Tuple[Var, int, int]] = None, bias_grad_external: bool = False,
           zero_rows: Tuple[int, int] = (0, 0), fold_act: bool = False, gemm_writers_only: bool = False,
           writers_emit_colsum: bool = False) -> Var:
    """y = act(x W[r0:r1, c0:c1]^T + b[r0:r1])  (nn.Linear).  ``out`` may be a strided (rows, N) view
    (concat-free epilogue: a GEMM writes straight into its slice of a wider buffer);
    ``accumulate_into`` adds into an existing Var (Linear over a concatenation = sum of Linears over
    column blocks of W, FeatureConcatFC / out_layer_pv without materialising the cat).
    ``bias_grad_external``: the bias gradient (column sums of dy) is produced by the consumer's backward kernel
    (add_layernorm's dz column sums), so no separate pass over dy runs here.
    ``fold_act``: the caller guarantees that y's only consumer is another GEMM op (Linear / conv): that op's dgrad epilogue
    multiplies by act'(y) and emits the column sums, so neither the act-bwd nor the bias-gradient pass runs (bf16 mode).
    ``gemm_writers_only``: the caller guarantees that every writer of d(y) is a GEMM that emits column sums (the projections
    consumed by attention_core): the bias gradient is accumulated in place by those GEMMs.
    ``writers_emit_colsum`` (with gemm_writers_only): the writers are the attention-backward kernel (jmt_attn_bwd_dqkv_bf16), whose
    transposed epilogue has the column sums for free: they go straight into the bias gradient even without JMT_EPI_EXT=2."""
    W = ctx.w(wname)
    if W.dim() == 3:                       # 1x1 Conv1d weight (cout, cin, 1)
        W = W.view(W.shape[0], W.shape[1])
    r0, r1 = w_rows if w_rows else (0, W.shape[0])
    c0, c1 = w_cols if w_cols else (0, W.shape[1])
    Wv = W[r0:r1, c0:c1]
    N, K = r1 - r0, c1 - c0
    M = x.data.shape[0]
    assert x.data.shape[1] == K, (tuple(x.data.shape), K, wname)
    bias = ctx.p(bname)[r0:r1] if bname else None
    if accumulate_into is not None:
        assert act == L.ACT_NONE
        y = accumulate_into
        gemm(ctx, x.data, Wv, y.data, M=M, N=N, K=K, bias=bias, store=L.ACCUMULATE)
    else:
        if out is None:
            out = ctx.empty((M, N))
        y = Var(out)
        gemm(ctx, x.data, Wv, out, M=M, N=N, K=K, bias=bias, act=act, slope=slope, zero_rows=zero_rows)
        if ctx.record and ctx.ext_on() and grad_from is None and out.is_contiguous():
            # writers of d(y) that are GEMMs add their column sums (= this bias gradient) on the fly
            y.track_colsum = bool(bname) and not bias_grad_external and (act == L.ACT_NONE or fold_act)
            if fold_act and act != L.ACT_NONE and zero_rows == (0, 0):
                y.fold = (slope, None, 1.0, 0)
            if y.track_colsum and ((EPI_COLSUM and (y.fold is not None or (gemm_writers_only and act == L.ACT_NONE))) or
                                   (writers_emit_colsum and gemm_writers_only and act == L.ACT_NONE)):
                y.colsum_direct = lambda: ctx.pgrad(bname)[r0:r1]
    if ctx.record:
        def bwd():
            if grad_from is not None:          # `out` is a column slice of a wider buffer owned by grad_from[0]
                pg = grad_from[0].grad
                dy = None if pg is None else pg[:, grad_from[1]:grad_from[2]]
            else:
                dy = y.grad
            if dy is None:
                return
            need_colsum = bool(bname) and not bias_grad_external
            if need_colsum and y.colsum_direct is not None and (act == L.ACT_NONE or y.folded):
                need_colsum = False            # the GEMMs that wrote dy added their column sums to the bias gradient already
            if need_colsum and y.track_colsum and y.colsum_ok and y.colsum_tmp is not None and (act == L.ACT_NONE or y.folded):
                # every writer of dy was a GEMM that already summed its columns: add the (N,) accumulator to the bias gradient
                L.check(ctx.lib.jmt_axpy(_ptr(y.colsum_tmp), _ptr(ctx.pgrad(bname)[r0:r1]), 1.0, N, L.F32, _stream()), "jmt_axpy")
                need_colsum = False
            if act != L.ACT_NONE and not y.folded:          # activation gradient and bias gradient in one pass over dy
                dy = _act_bwd(ctx, y, dy, slope, colsum=ctx.pgrad(bname)[r0:r1] if need_colsum else None)
            elif need_colsum:
                L.check(ctx.lib.jmt_colsum(_ptr(dy), _DT[dy.dtype], dy.stride(0), M, N, _ptr(ctx.pgrad(bname)[r0:r1]),
                                           _stream()), "jmt_colsum")
            # dW[r0:r1, c0:c1] += dy^T x   (both operands MN-major, split-K over the M rows, fp32 atomics)
            dWf = ctx.pgrad(wname)
            if dWf.dim() == 3:
                dWf = dWf.view(dWf.shape[0], dWf.shape[1])
            dW = dWf[r0:r1, c0:c1]
            gemm(ctx, dy, x.data, dW, M=N, N=K, K=M, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN,
                 store=L.ATOMIC_ADD, split_k=wgrad_split(M, N, K))
            if x.needs_grad:
                dx, mode = ctx.grad_target(x, capable=True)
                _dgrad_gemm(ctx, x, dy, Wv, dx, mode, M=M, N=K, K=N, b_major=L.MAJOR_MN)
            if accumulate_into is None:
                ctx.release(y)
        ctx.tape.append(bwd)
    return y


def _dgrad_gemm(ctx: Ctx, x: Var, dy: torch.Tensor, w: torch.Tensor, dx: torch.Tensor, mode: int, **kw):
    """The GEMM that writes d(x): with the backward-epilogue extensions it multiplies by act'(x) (and the channel-dropout mask)
    when x is marked `fold`, and adds the column sums of what it writes to x's bias-gradient accumulator."""
    if ctx.ext_on() and dx.dtype == torch.bfloat16:
        if x.fold is not None:
            slope, mask, mscale, mrows = x.fold
            kw.update(epi_aux=x.data, aux_slope=slope)
            if mask is not None:
                kw.update(colmask=mask, colmask_scale=mscale, colmask_row_period=mrows)
            x.folded = True
        cs = ctx.colsum_target(x)
        if cs is not None:
            kw.update(d_colsum=cs)
    elif x.fold is not None or x.track_colsum:
        ctx._incapable_writer(x)
    gemm(ctx, dy, w, dx, store=mode, **kw)


def _act_bwd(ctx: Ctx, y: Var, dy: torch.Tensor, slope: float, colsum: Optional[torch.Tensor] = None,
             mask: Optional[torch.Tensor] = None, mask_rows: int = 1, mask_scale: float = 1.0) -> torch.Tensor:
    """dy * [channel-dropout mask * scale] * act'(y), optionally with the column sums (bias gradient) of the
    result accumulated into ``colsum`` -- one pass (jmt_act_bwd_fused).  In place when this Var is the only owner
    of its gradient buffer; otherwise (buffer aliased by a sibling, e.g. the two branches of a residual add)
    into a fresh buffer."""
    assert dy.is_contiguous() and y.data.is_contiguous()
    dst = dy if (y.gbuf is not None and y.gbuf.refs == 1 and y.gbuf.t is dy) else ctx.empty(dy.shape, dy.dtype)
    rows, cols = dy.shape
    if cols % 8 == 0:
        L.check(ctx.lib.jmt_act_bwd_fused(_ptr(dy), _ptr(y.data), _ptr(mask), _ptr(dst), rows, cols, mask_rows, mask_scale,
                                          slope, _ptr(colsum), _DT[dy.dtype], _stream()), "jmt_act_bwd_fused")
        return dst
    assert mask is None
    L.check(ctx.lib.jmt_act_bwd(_ptr(dy), _ptr(y.data), _ptr(dst), dy.numel(), slope, _DT[dy.dtype], _stream()),
            "jmt_act_bwd")
    if colsum is not None:
        L.check(ctx.lib.jmt_colsum(_ptr(dst), _DT[dst.dtype], dst.stride(0), rows, cols, _ptr(colsum), _stream()), "jmt_colsum")
    return dst


def add_layernorm(ctx: Ctx, x: Var, res: Optional[Var], gname: str, bname: str, res_bias: Optional[str] = None) -> Var:
    """LayerNorm(x + res) with affine params (post-LN residual, mm_multi_transformers.py:62-69).
    ``res_bias``: name of the bias of the Linear that produced ``res`` (consumed only here): its gradient = column
    sums of dz, accumulated by the LN backward kernel itself (that Linear is built with bias_grad_external)."""
    rows, D = x.data.shape
    assert x.data.is_contiguous() and (res is None or res.data.is_contiguous())
    y = ctx.empty((rows, D))
    mean = ctx.empty((rows,), torch.float32)
    rstd = ctx.empty((rows,), torch.float32)
    L.check(ctx.lib.jmt_add_layernorm_fwd(_ptr(x.data), _ptr(res.data) if res else None, _ptr(ctx.p(gname)),
                                          _ptr(ctx.p(bname)), 1e-5, _ptr(y), _ptr(mean), _ptr(rstd), rows, D, ctx.acode,
                                          _stream()), "jmt_add_layernorm_fwd")
    out = Var(y)
    if ctx.record:
        def bwd():
            dy = out.grad
            if dy is None:
                return
            assert dy.is_contiguous()
            dz = GradBuf(ctx.empty((rows, D)))
            L.check(ctx.lib.jmt_add_layernorm_bwd(_ptr(dy), _ptr(x.data), _ptr(res.data) if res else None,
                                                  _ptr(ctx.p(gname)), _ptr(mean), _ptr(rstd), _ptr(dz.t), 0,
                                                  _ptr(ctx.pgrad(gname)), _ptr(ctx.pgrad(bname)),
                                                  _ptr(ctx.pgrad(res_bias)) if res_bias else None, rows, D, ctx.acode,
                                                  _stream()), "jmt_add_layernorm_bwd")
            ctx.add_grad(x, dz)
            if res is not None:
                ctx.add_grad(res, dz)
            dz.refs -= 1            # drop the creator's reference
            ctx.release(out)
        ctx.tape.append(bwd)
    return out


class AttnGeom:
    """How (seq, batch) index the rows of a (rows, E) activation matrix.
    Standard (B, T, E): seq_stride=1, batch_stride=T.  Attention across the batch dimension
    (MultimodalTransformer_wo_JR encoders): seq = b (stride T), batch = t (stride 1)."""

    def __init__(self, seq: int, batch: int, seq_stride: int, batch_stride: int):
        self.seq, self.batch, self.seq_stride, self.batch_stride = seq, batch, seq_stride, batch_stride


def _proj_grad(ctx: Ctx, v: Var) -> torch.Tensor:
    """Gradient buffer of a Var that is written piecewise with ACCUMULATE semantics (zero-filled on first touch)."""
    ctx._incapable_writer(v)
    if v.gbuf is None:
        v.gbuf = GradBuf(ctx.zeros(v.data.shape, v.data.dtype))
    elif v.gbuf.refs > 1:
        ctx.grad_target(v)
    elif v.gbuf.written is not None:
        _ = v.grad                       # checks full coverage
    return v.gbuf.t


def _proj_grad_slice(ctx: Ctx, v: Var, c0: int, c1: int, capable: bool = False) -> Tuple[torch.Tensor, int]:
    """Column slice [c0, c1) of the gradient buffer of projection matrix `v` and the store mode for writing it: the first
    write of a slice STOREs into an un-initialised buffer, later writes of the same slice (a Q projection shared by two
    attention calls) ACCUMULATE."""
    if not (capable and ctx.ext_on()):
        ctx._incapable_writer(v)
    if v.gbuf is None:
        v.gbuf = GradBuf(ctx.empty(v.data.shape, v.data.dtype), written=[])
    elif v.gbuf.refs > 1:
        ctx.grad_target(v, capable=capable)
    gb = v.gbuf
    if gb.written is None:
        return gb.t[:, c0:c1], L.ACCUMULATE
    for (a, b) in gb.written:
        if (a, b) == (c0, c1):
            return gb.t[:, c0:c1], L.ACCUMULATE
        if a < c1 and c0 < b:
            raise RuntimeError("overlapping projection-gradient slices")
    gb.written.append((c0, c1))
    return gb.t[:, c0:c1], L.STORE


FUSED_ATTENTION = True      # bf16 mode: True = QK^T -> softmax -> PV in one kernel when the geometry is supported; "p" = QK^T +
                            # softmax fused, PV a plain GEMM (measured 1.4 % slower per step); False = GEMM + softmax kernels
# Backward GEMM epilogue extensions (jmt_gemm_desc.epi_aux / d_colsum), JMT_EPI_EXT = 0 | 1 (default) | 2:
#   1: the activation gradient of FFN-ReLU / TCN conv1-LeakyReLU(+channel mask) is folded into the consumer's dgrad epilogue
#      (the act-bwd pass, 3 passes over the gradient, shrinks to the bias-gradient column sum, 1 pass);
#   2: additionally the GEMMs that write a gradient emit its column sums themselves (no colsum kernels).  Measured on B200
#      (profiles/gemm_table_r2q_*.txt): the butterfly transpose-reduce makes the attention-backward GEMMs 15-19 % slower, which
#      costs what the 0.33 ms of colsum kernels cost -- kept as an option, not the default.
_EPI_MODE = int(os.environ.get("JMT_EPI_EXT", "1"))
EPI_EXT = _EPI_MODE >= 1
EPI_COLSUM = _EPI_MODE >= 2
ADD_ACT_FUSED_BWD = os.environ.get("JMT_ADD_ACT_FUSED", "1") != "0"     # TemporalBlock: residual-ReLU backward fused with conv2's activation backward
ATTN_DELTA_IN_KERNEL = os.environ.get("JMT_ATTN_DELTA", "1") != "0"   # softmax-backward delta = rowsum(P o dP) inside the dS kernel
LONG_S_CHUNKED = True       # forward-only attention beyond the fused kernel's key limit: chunked fused kernel + log-sum-exp merge
LONG_S_CHUNK_MIN_BYTES = 2 << 30   # ... once the fp32 score tensor of the composed path would exceed this (measured: NONE eval at
                                   # B = 1024, 1.26 GB of scores, 7.4 ms composed vs 9.1 ms chunked; B = 4096, 20 GB: chunked only)
ATTN_BWD_DQKV = os.environ.get("JMT_ATTN_DQKV", "1") != "0"   # dQ / dK / dV of the "ds" backward by ONE launch of jmt_attn_bwd_dqkv_bf16
                                                               # (transposed CTA-pair kernel, bias-gradient column sums included) instead of
                                                               # three batched jmt_gemm_bf16 launches, when the geometry is supported
FUSED_ATTENTION_BWD = "ds"   # "ds": dP GEMM + softmax backward fused (dS on chip, dQ/dK/dV plain GEMMs); "full": dP -> dS -> dQ in one
                             # kernel (correct, not faster than the composition yet); False: GEMM + softmax_bwd kernel


def _attn_chain(ctx: Ctx, mode: int, a1, a1_geo, b1, b1_geo, b2, b2_geo, p_in, x, d, d_geo, Lq, S, dh, heads, NB, x_ld,
                scale, store, probe_only=False, o_in=None, delta_in=None, lse_out=None):
    """One launch of the fused attention core (include/jmt_b200.h: jmt_attn_desc).  *_geo = (ld, head stride,
    batch stride) in elements."""
    g = L.AttnDesc()
    g.a1, g.b1, g.b2 = a1.data_ptr(), b1.data_ptr(), (b2.data_ptr() if b2 is not None else None)
    g.p_in = p_in.data_ptr() if p_in is not None else None
    g.o_in = o_in.data_ptr() if o_in is not None else None
    g.delta_in = delta_in.data_ptr() if delta_in is not None else None
    g.x, g.d = (x.data_ptr() if x is not None else None), (d.data_ptr() if d is not None else None)
    g.lse_out = lse_out.data_ptr() if lse_out is not None else None
    g.mode = mode
    g.Lq, g.S, g.dh, g.heads, g.NB = Lq, S, dh, heads, NB
    g.a1_ld, g.a1_hs, g.a1_bs = a1_geo
    g.b1_ld, g.b1_hs, g.b1_bs = b1_geo
    g.b2_ld, g.b2_hs, g.b2_bs = b2_geo if b2_geo is not None else (0, 0, 0)
    g.d_ld, g.d_hs, g.d_bs = d_geo if d_geo is not None else (0, 0, 0)
    g.x_ld = x_ld
    g.scale = scale
    g.store_mode = store
    if probe_only:
        return bool(ctx.lib.jmt_attn_chain_supported(C.byref(g)))
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(ctx.lib.jmt_attn_chain_bf16(C.byref(g), _stream()), "jmt_attn_chain_bf16")
    if prof is not None:
        e1.record()
        prof.append(("attn_chain_kernel", (2.0 if d is None else 4.0) * NB * heads * Lq * S * dh, e0, e1, (Lq, S, dh, NB * heads, mode)))
    return True


def dqkv_planned(ctx: Ctx, dh: int, Lq: int, S: int, heads: int) -> bool:
    """Will attention_core's backward run jmt_attn_bwd_dqkv_bf16 for this geometry?  (Decided before the projections are built:
    they then let that kernel add the bias-gradient column sums.  attention_core falls back to column-sum-emitting GEMMs should
    the fused forward turn out to be unsupported.)"""
    return (ctx.record and ctx.ext_on() and FUSED_ATTENTION is True and FUSED_ATTENTION_BWD == "ds" and dqkv_geometry_ok(dh, Lq, S, heads))


def dqkv_geometry_ok(dh: int, Lq: int, S: int, heads: int) -> bool:
    """Geometries jmt_attn_bwd_dqkv_bf16 supports (jmt_attn_bwd_dqkv_supported, csrc/attn_bwd_tc.cu) and the switch for it."""
    return ATTN_BWD_DQKV and dh in (256, 512) and Lq <= 320 and S <= 320 and heads * dh <= 1024


def _attn_bwd_dqkv(ctx: Ctx, parts, Lq, S, dh, heads, NB, x_ld):
    """One launch of jmt_attn_bwd_dqkv_bf16 (include/jmt_b200.h).  parts: up to three (a, a_geo, x, x_trans, d, d_geo, store, alpha,
    colsum) tuples; *_geo = (ld, head stride, batch stride) in elements."""
    g = L.AttnBwdDesc()
    for i, (a, a_geo, x, x_trans, d, d_geo, store, alpha, colsum) in enumerate(parts):
        pt = g.part[i]
        pt.a, pt.x, pt.d = a.data_ptr(), x.data_ptr(), d.data_ptr()
        pt.a_ld, pt.a_hs, pt.a_bs = a_geo
        pt.d_ld, pt.d_hs, pt.d_bs = d_geo
        pt.x_trans, pt.store_mode, pt.alpha = x_trans, store, alpha
        pt.colsum = colsum.data_ptr() if colsum is not None else None
    g.Lq, g.S, g.dh, g.heads, g.NB, g.x_ld = Lq, S, dh, heads, NB, x_ld
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(ctx.lib.jmt_attn_bwd_dqkv_bf16(C.byref(g), _stream()), "jmt_attn_bwd_dqkv_bf16")
    if prof is not None:
        e1.record()
        prof.append(("attn_bwd_dqkv_kernel", 2.0 * len(parts) * NB * heads * Lq * S * dh, e0, e1, (Lq, S, dh, NB * heads, len(parts))))


def attention_core(ctx: Ctx, q: Var, qcol: int, k: Var, kcol: int, v: Var, vcol: int, E: int, heads: int,
                   gq: AttnGeom, gk: AttnGeom) -> Var:
    """softmax((Q/sqrt(dh)) K^T) V per (batch, head) -- torch's MHA math path minus the
    head-averaged weights the reference discards.  q/k/v are column slices [col, col+E) of projection
    matrices; the output (rows_q, E) is laid out like q's rows.
    bf16 mode: one fused kernel forward (QK^T -> softmax -> PV, scores never leave the SM) and one backward
    (dP -> dS -> dQ) plus two plain GEMMs (dV = P^T dO, dK = dS^T Q); fp32 mode / unsupported geometry: GEMM and
    softmax kernels."""
    dh = E // heads
    Lq, S, NB = gq.seq, gk.seq, gq.batch
    assert gk.batch == NB and E % heads == 0
    s_ld = (S + 7) // 8 * 8
    scale = 1.0 / math.sqrt(dh)
    qd, kd, vd = q.data[:, qcol:qcol + E], k.data[:, kcol:kcol + E], v.data[:, vcol:vcol + E]
    qld, kld, vld = q.data.stride(0), k.data.stride(0), v.data.stride(0)
    sb = (Lq * s_ld, heads * Lq * s_ld)
    q_geo = (gq.seq_stride * qld, dh, gq.batch_stride * qld)
    k_geo = (gk.seq_stride * kld, dh, gk.batch_stride * kld)
    v_geo = (gk.seq_stride * vld, dh, gk.batch_stride * vld)
    o_geo = (gq.seq_stride * E, dh, gq.batch_stride * E)
    o = ctx.empty((q.data.shape[0], E))
    fused = False
    if FUSED_ATTENTION and ctx.adt == torch.bfloat16:
        fused = _attn_chain(ctx, 0, qd, q_geo, kd, k_geo, vd, v_geo, None, o, o, o_geo, Lq, S, dh, heads, NB, s_ld, scale,
                            L.STORE, probe_only=True)
    if (not fused and FUSED_ATTENTION and ctx.adt == torch.bfloat16 and not ctx.record and LONG_S_CHUNKED and
            4 * NB * heads * Lq * s_ld >= LONG_S_CHUNK_MIN_BYTES):
        # Long key sequences, forward only (evaluation; SURVEY 7.6a: the batch-dimension attention of the NONE variant at
        # large batch): the fused kernel over key chunks + running log-sum-exp merge.  Neither the (L, S) scores nor the
        # probabilities ever reach memory.  (A backward needs the probabilities: training keeps the composed path below.)
        chunk = next((c for c in (320, 256, 192, 128, 64) if c <= S and _attn_chain(
            ctx, 0, qd, q_geo, kd, k_geo, vd, v_geo, None, None, o, o_geo, Lq, c, dh, heads, NB, 0, scale, L.STORE, probe_only=True)), 0)
        if chunk:
            oc = ctx.empty((q.data.shape[0], E))
            acc = ctx.empty((q.data.shape[0], E), torch.float32)
            lse_c = ctx.empty((NB, heads, Lq), torch.float32)
            lse_acc = ctx.empty((NB, heads, Lq), torch.float32)
            for s0 in range(0, S, chunk):
                sc = min(chunk, S - s0)
                kc, vc = kd[s0 * gk.seq_stride:], vd[s0 * gk.seq_stride:]       # key s <-> row s * seq_stride (+ batch offset)
                _attn_chain(ctx, 0, qd, q_geo, kc, k_geo, vc, v_geo, None, None, oc, o_geo, Lq, sc, dh, heads, NB, 0, scale,
                            L.STORE, lse_out=lse_c)
                last = s0 + chunk >= S
                L.check(ctx.lib.jmt_attn_merge(_ptr(oc), o_geo[0], o_geo[1], o_geo[2], _ptr(lse_c), _ptr(acc), _ptr(lse_acc),
                                               _ptr(o) if last else None, 1 if s0 == 0 else 0, NB, heads, Lq, dh, _stream()),
                        "jmt_attn_merge")
            return Var(o)
    if fused and FUSED_ATTENTION == "p":
        # QK^T + softmax on chip (fp32 scores never leave the SM), PV as a plain GEMM
        probs = ctx.empty((NB, heads, Lq, s_ld))
        _attn_chain(ctx, 0, qd, q_geo, kd, k_geo, None, None, None, probs, None, None, Lq, S, dh, heads, NB, s_ld, scale, L.STORE)
        gemm(ctx, probs, vd, o, M=Lq, N=dh, K=S, a_rows=Lq, b_major=L.MAJOR_MN, b_rows=S,
             a_ld=s_ld, b_ld=gk.seq_stride * vld, d_ld=gq.seq_stride * E,
             nb0=heads, nb1=NB, a_bs=sb, b_bs=(dh, gk.batch_stride * vld), d_bs=(dh, gq.batch_stride * E))
    elif fused:
        probs = ctx.empty((NB, heads, Lq, s_ld))
        _attn_chain(ctx, 0, qd, q_geo, kd, k_geo, vd, v_geo, None, probs, o, o_geo, Lq, S, dh, heads, NB, s_ld, scale, L.STORE)
    else:
        scores = ctx.empty((NB, heads, Lq, s_ld), torch.float32)
        gemm(ctx, qd, kd, scores, M=Lq, N=S, K=dh, a_rows=Lq, b_rows=S,
             a_ld=gq.seq_stride * qld, b_ld=gk.seq_stride * kld, d_ld=s_ld,
             nb0=heads, nb1=NB, a_bs=(dh, gq.batch_stride * qld), b_bs=(dh, gk.batch_stride * kld), d_bs=sb, alpha=scale)
        probs = ctx.empty((NB, heads, Lq, s_ld))
        rows = NB * heads * Lq
        L.check(ctx.lib.jmt_softmax_fwd(_ptr(scores), s_ld, _ptr(probs), ctx.acode, s_ld, rows, S, _stream()), "jmt_softmax_fwd")
        del scores
        gemm(ctx, probs, vd, o, M=Lq, N=dh, K=S, a_rows=Lq, b_major=L.MAJOR_MN, b_rows=S,
             a_ld=s_ld, b_ld=gk.seq_stride * vld, d_ld=gq.seq_stride * E,
             nb0=heads, nb1=NB, a_bs=sb, b_bs=(dh, gk.batch_stride * vld), d_bs=(dh, gq.batch_stride * E))
    out = Var(o)
    if ctx.record:
        def bwd():
            do = out.grad
            if do is None:
                return
            assert do.is_contiguous()
            rows = NB * heads * Lq
            # the plain GEMMs that write dQ / dK / dV also add their column sums to the projection's bias-gradient accumulator
            gemm_dq = not (fused and FUSED_ATTENTION_BWD and FUSED_ATTENTION_BWD != "ds")
            gq_s, q_mode = _proj_grad_slice(ctx, q, qcol, qcol + E, capable=gemm_dq)
            cs_q = dict(d_colsum=ctx.colsum_target(q, qcol, qcol + E), colsum_bs0=dh) if gemm_dq and q.track_colsum and ctx.ext_on() else {}
            q_gld = gq_s.stride(0)
            dq_geo = (gq.seq_stride * q_gld, dh, gq.batch_stride * q_gld)
            ds = ctx.empty((NB, heads, Lq, s_ld))
            if fused and FUSED_ATTENTION_BWD == "ds":
                # delta = rowsum(dO o O); dP = dO V^T -> dS = scale * P o (dP - delta) in one kernel (fp32 dP never leaves
                # the SM); dQ = dS K, dK = dS^T Q, dV = P^T dO stay plain GEMMs (the scale is already in dS)
                delta = None
                if not ATTN_DELTA_IN_KERNEL:        # delta = rowsum(dO o O) by a separate pass (round-1 path)
                    delta = ctx.empty((NB, heads, Lq), torch.float32)
                    L.check(ctx.lib.jmt_rowdot_bf16(_ptr(do), _ptr(o), o_geo[0], o_geo[1], o_geo[2], NB, heads, Lq, dh, _ptr(delta),
                                                    _stream()), "jmt_rowdot_bf16")
                _attn_chain(ctx, 1, do, o_geo, vd, v_geo, None, None, probs, ds, None, None, Lq, S, dh, heads, NB, s_ld, scale,
                            L.STORE, delta_in=delta)
                if dqkv_geometry_ok(dh, Lq, S, heads):
                    # dQ = dS K, dK = dS^T Q, dV = P^T dO (+ the projections' bias-gradient column sums) in ONE launch
                    gv, v_mode = _proj_grad_slice(ctx, v, vcol, vcol + E, capable=True)
                    gk_, k_mode = _proj_grad_slice(ctx, k, kcol, kcol + E, capable=True)
                    cs = [ctx.colsum_target(t, c0, c0 + E) if t.track_colsum and ctx.ext_on() else None
                          for t, c0 in ((q, qcol), (k, kcol), (v, vcol))]
                    gk_ld, gv_ld = gk_.stride(0), gv.stride(0)
                    _attn_bwd_dqkv(ctx, [
                        (kd, k_geo, ds, 0, gq_s, dq_geo, q_mode, 1.0, cs[0]),
                        (qd, q_geo, ds, 1, gk_, (gk.seq_stride * gk_ld, dh, gk.batch_stride * gk_ld), k_mode, 1.0, cs[1]),
                        (do, o_geo, probs, 1, gv, (gk.seq_stride * gv_ld, dh, gk.batch_stride * gv_ld), v_mode, 1.0, cs[2]),
                    ], Lq, S, dh, heads, NB, s_ld)
                    ctx.release(out)
                    return
                gemm(ctx, ds, kd, gq_s, M=Lq, N=dh, K=S, a_rows=Lq, b_major=L.MAJOR_MN, b_rows=S,
                     a_ld=s_ld, b_ld=gk.seq_stride * kld, d_ld=gq.seq_stride * q_gld,
                     nb0=heads, nb1=NB, a_bs=sb, b_bs=(dh, gk.batch_stride * kld), d_bs=(dh, gq.batch_stride * q_gld),
                     store=q_mode, **cs_q)
                dk_alpha = 1.0
            elif fused and FUSED_ATTENTION_BWD:
                # dP = dO V^T -> dS = scale * P o (dP - rowsum(P o dP)) -> dQ += dS K, one kernel; dS is saved for dK
                _attn_chain(ctx, 1, do, o_geo, vd, v_geo, kd, k_geo, probs, ds, gq_s, dq_geo, Lq, S, dh,
                            heads, NB, s_ld, scale, q_mode, o_in=o)
                dk_alpha = 1.0
            else:
                dp = ctx.empty((NB, heads, Lq, s_ld), torch.float32)          # dP = dO V^T
                gemm(ctx, do, vd, dp, M=Lq, N=S, K=dh, a_rows=Lq, b_rows=S,
                     a_ld=gq.seq_stride * E, b_ld=gk.seq_stride * vld, d_ld=s_ld,
                     nb0=heads, nb1=NB, a_bs=(dh, gq.batch_stride * E), b_bs=(dh, gk.batch_stride * vld), d_bs=sb)
                L.check(ctx.lib.jmt_softmax_bwd(_ptr(probs), ctx.acode, s_ld, _ptr(dp), s_ld, _ptr(ds), ctx.acode, s_ld,
                                                rows, S, _stream()), "jmt_softmax_bwd")
                del dp
                gemm(ctx, ds, kd, gq_s, M=Lq, N=dh, K=S, a_rows=Lq, b_major=L.MAJOR_MN, b_rows=S,
                     a_ld=s_ld, b_ld=gk.seq_stride * kld, d_ld=gq.seq_stride * q_gld,
                     nb0=heads, nb1=NB, a_bs=sb, b_bs=(dh, gk.batch_stride * kld), d_bs=(dh, gq.batch_stride * q_gld),
                     alpha=scale, store=q_mode, **cs_q)                        # dQ = scale * dS K
                dk_alpha = scale
            gv, v_mode = _proj_grad_slice(ctx, v, vcol, vcol + E, capable=True)         # dV = P^T dO
            cs_v = dict(d_colsum=ctx.colsum_target(v, vcol, vcol + E), colsum_bs0=dh) if v.track_colsum and ctx.ext_on() else {}
            gemm(ctx, probs, do, gv, M=S, N=dh, K=Lq, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN,
                 a_rows=Lq, b_rows=Lq, a_ld=s_ld, b_ld=gq.seq_stride * E, d_ld=gk.seq_stride * gv.stride(0),
                 nb0=heads, nb1=NB, a_bs=sb, b_bs=(dh, gq.batch_stride * E), d_bs=(dh, gk.batch_stride * gv.stride(0)),
                 store=v_mode, **cs_v)
            gk_, k_mode = _proj_grad_slice(ctx, k, kcol, kcol + E, capable=True)        # dK = scale * dS^T Q
            cs_k = dict(d_colsum=ctx.colsum_target(k, kcol, kcol + E), colsum_bs0=dh) if k.track_colsum and ctx.ext_on() else {}
            gemm(ctx, ds, qd, gk_, M=S, N=dh, K=Lq, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN,
                 a_rows=Lq, b_rows=Lq, a_ld=s_ld, b_ld=gq.seq_stride * qld, d_ld=gk.seq_stride * gk_.stride(0),
                 nb0=heads, nb1=NB, a_bs=sb, b_bs=(dh, gq.batch_stride * qld), d_bs=(dh, gk.batch_stride * gk_.stride(0)),
                 alpha=dk_alpha, store=k_mode, **cs_k)
            ctx.release(out)
        ctx.tape.append(bwd)
    return out


def _attn_bwd_by_gemms() -> bool:
    """attention_core's backward writes dQ / dK / dV with plain GEMMs (always, except in the experimental "full" chain mode)"""
    return FUSED_ATTENTION_BWD in ("ds", False, None)


def attention_small(ctx: Ctx, qkv: Var, Lseq: int, N: int, E: int, heads: int) -> Var:
    """Self-attention over tiny sequences (L <= 8): qkv is (L*N, 3E) with row = l*N + n."""
    scale = 1.0 / math.sqrt(E // heads)
    assert qkv.data.is_contiguous()
    o = ctx.empty((Lseq * N, E))
    probs = ctx.empty((N * heads, Lseq, Lseq), torch.float32)
    L.check(ctx.lib.jmt_attn_small_fwd(_ptr(qkv.data), _ptr(o), _ptr(probs), Lseq, N, E, heads, scale, ctx.acode,
                                       _stream()), "jmt_attn_small_fwd")
    out = Var(o)
    if ctx.record:
        def bwd():
            do = out.grad
            if do is None:
                return
            assert do.is_contiguous()
            dq = GradBuf(ctx.empty(qkv.data.shape))
            L.check(ctx.lib.jmt_attn_small_bwd(_ptr(qkv.data), _ptr(do), _ptr(probs), _ptr(dq.t), Lseq, N, E, heads,
                                               scale, ctx.acode, _stream()), "jmt_attn_small_bwd")
            ctx.add_grad(qkv, dq)
            dq.refs -= 1
            ctx.release(out)
        ctx.tape.append(bwd)
    return out


def mha_self(ctx: Ctx, x: Var, prefix: str, heads: int, geom: Optional[AttnGeom], small: Optional[Tuple[int, int]] = None,
             out_bias_grad_external: bool = False) -> Var:
    """nn.MultiheadAttention(x, x, x): packed QKV projection (one N=3E GEMM), attention, out-proj."""
    E = x.data.shape[1]
    wec = small is None and geom is not None and dqkv_planned(ctx, E // heads, geom.seq, geom.seq, heads)
    qkv = linear(ctx, x, prefix + "in_proj_weight", prefix + "in_proj_bias", gemm_writers_only=small is None and _attn_bwd_by_gemms(),
                 writers_emit_colsum=wec)
    if small is not None:
        o = attention_small(ctx, qkv, small[0], small[1], E, heads)
    else:
        o = attention_core(ctx, qkv, 0, qkv, E, qkv, 2 * E, E, heads, geom, geom)
    return linear(ctx, o, prefix + "out_proj.weight", prefix + "out_proj.bias", bias_grad_external=out_bias_grad_external)


def mha_cross(ctx: Ctx, xq: Var, xkv: Var, prefix: str, heads: int, gq: AttnGeom, gk: AttnGeom,
              out: Optional[torch.Tensor] = None, grad_from=None) -> Var:
    """nn.MultiheadAttention(xq, xkv, xkv).  The Q projection of a module applied twice to the same
    query (cross_attention_{v,p,pv} in MultimodalTransformer_w_JR) is computed once."""
    E = xq.data.shape[1]
    wec = dqkv_planned(ctx, E // heads, gq.seq, gk.seq, heads)
    key = (prefix, id(xq))
    q = ctx.qcache.get(key)
    if q is None:
        q = linear(ctx, xq, prefix + "in_proj_weight", prefix + "in_proj_bias", w_rows=(0, E), gemm_writers_only=_attn_bwd_by_gemms(),
                   writers_emit_colsum=wec)
        ctx.qcache[key] = q
    kv = linear(ctx, xkv, prefix + "in_proj_weight", prefix + "in_proj_bias", w_rows=(E, 3 * E), gemm_writers_only=_attn_bwd_by_gemms(),
                writers_emit_colsum=wec)
    o = attention_core(ctx, q, 0, kv, 0, kv, E, E, heads, gq, gk)
    return linear(ctx, o, prefix + "out_proj.weight", prefix + "out_proj.bias", out=out, grad_from=grad_from)


def encoder_layer(ctx: Ctx, x: Var, prefix: str, heads: int, geom: Optional[AttnGeom], small=None) -> Var:
    """TransformerEncoderLayer.forward (mm_multi_transformers.py:60-70): post-LN MHA + ReLU FFN."""
    a = mha_self(ctx, x, prefix + "attention.", heads, geom, small, out_bias_grad_external=True)
    x1 = add_layernorm(ctx, x, a, prefix + "layer_norm1.weight", prefix + "layer_norm1.bias",
                       res_bias=prefix + "attention.out_proj.bias")
    h = linear(ctx, x1, prefix + "feed_forward.0.weight", prefix + "feed_forward.0.bias", act=L.ACT_RELU, fold_act=True)
    f = linear(ctx, h, prefix + "feed_forward.2.weight", prefix + "feed_forward.2.bias", bias_grad_external=True)
    return add_layernorm(ctx, x1, f, prefix + "layer_norm2.weight", prefix + "layer_norm2.bias",
                         res_bias=prefix + "feed_forward.2.bias")


def encoder_block(ctx: Ctx, x: Var, prefix: str, heads: int, layers: int, geom, small=None) -> Var:
    for i in range(layers):
        x = encoder_layer(ctx, x, f"{prefix}layers.{i}.", heads, geom, small)
    return x


def dropout(ctx: Ctx, x: Var, p: float) -> Tuple[Var, float]:
    """nn.Dropout(p) in training mode with a Philox keep-mask (replayed in backward).  Identity (and no
    kernel) when p == 0 or in eval mode.  Returns (Var, scale)."""
    if p <= 0.0 or not ctx.training:
        return x, 1.0
    n = x.data.numel()
    assert x.data.is_contiguous()
    mask = ctx.empty((n,), torch.uint8)
    L.check(ctx.lib.jmt_dropout_mask(_ptr(mask), n, p, ctx.seed, ctx.rng_offset, _ptr(ctx.rng_state), _stream()), "jmt_dropout_mask")
    ctx.rng_offset += (n + 3) // 4
    scale = 1.0 / (1.0 - p)
    y = ctx.empty(x.data.shape)
    L.check(ctx.lib.jmt_apply_mask(_ptr(x.data), _ptr(mask), _ptr(y), 1, 1, n, 0, scale, ctx.acode, _stream()), "jmt_apply_mask")
    out = Var(y)
    if ctx.record:
        def bwd():
            dy = out.grad
            if dy is None:
                return
            dx = GradBuf(ctx.empty(x.data.shape))
            L.check(ctx.lib.jmt_apply_mask(_ptr(dy), _ptr(mask), _ptr(dx.t), 1, 1, n, 0, scale, ctx.acode, _stream()), "jmt_apply_mask")
            ctx.add_grad(x, dx)
            dx.refs -= 1
            ctx.release(out)
        ctx.tape.append(bwd)
    return out, scale


def _ptr_array(ctx: Ctx, tensors):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    ctx.keep.append(arr)
    return arr


def contig_f32_2d(ctx: Ctx, d: torch.Tensor) -> torch.Tensor:
    """A 2-D gradient handed in by autograd as a contiguous fp32 tensor: `.sum().backward()` and friends deliver expanded
    (stride-0) or otherwise strided views; they are materialised by one strided-copy launch (jmt_copy3d), not by ATen."""
    require_cuda(d)
    if d.is_contiguous() and d.dtype == torch.float32:
        return d
    if d.dim() != 2 or d.dtype not in _DT:
        return d.contiguous().to(torch.float32)          # exotic layouts / dtypes only
    out = ctx.empty(tuple(d.shape), torch.float32)
    L.check(ctx.lib.jmt_copy3d(_ptr(d), _DT[d.dtype], d.stride(0), d.stride(1), _ptr(out), L.F32, d.shape[1], 1,
                               d.shape[0], d.shape[1], 1, _stream()), "jmt_copy3d")
    return out


def regressor_tail(ctx: Ctx, hidden: List[Var], wnames: List[str], bnames: List[str], w_row: List[int],
                   B: int, T: int, time_major: bool):
    """Final Linear(128, k) of the heads as one fused kernel over G groups (hidden[g] may repeat).
    Returns (list of fp32 output tensors, list of output 'grad setters')."""
    G = len(hidden)
    M = B * T
    shape = (T, B) if time_major else (B, T)
    sb, st = (1, B) if time_major else (T, 1)
    outs = [ctx.empty(shape, torch.float32) for _ in range(G)]
    ws = [ctx.p(wnames[g])[w_row[g]] for g in range(G)]
    bs = [ctx.p(bnames[g])[w_row[g]:w_row[g] + 1] for g in range(G)]
    hl = [h.data for h in hidden]
    h_ld = hl[0].stride(0)
    assert all(h.stride(0) == h_ld and h.shape[1] == 128 for h in hl)
    L.check(ctx.lib.jmt_regressor_tail_fwd(G, _ptr_array(ctx, hl), h_ld, ctx.acode, _ptr_array(ctx, ws), _ptr_array(ctx, bs),
                                           _ptr_array(ctx, outs), M, T, sb, st, _stream()), "jmt_regressor_tail_fwd")
    gouts: List[Optional[torch.Tensor]] = [None] * G
    if ctx.record:
        def bwd():
            dhs, acc, seen = [], [], {}
            for g in range(G):
                hv = hidden[g]
                if id(hv) in seen:
                    dhs.append(seen[id(hv)].t)
                    acc.append(1)
                else:
                    gb = GradBuf(ctx.empty(hv.data.shape))
                    seen[id(hv)] = gb
                    dhs.append(gb.t)
                    acc.append(0)
            douts = []
            for g in range(G):
                d = gouts[g]
                if d is None:
                    d = ctx.zeros(shape, torch.float32)
                douts.append(contig_f32_2d(ctx, d))
            dws = [ctx.pgrad(wnames[g])[w_row[g]] for g in range(G)]
            dbs = [ctx.pgrad(bnames[g])[w_row[g]:w_row[g] + 1] for g in range(G)]
            acc_arr = (C.c_int * G)(*acc)
            sc_arr = (C.c_float * G)(*([1.0] * G))
            ctx.keep += [acc_arr, sc_arr]
            L.check(ctx.lib.jmt_regressor_tail_bwd(G, _ptr_array(ctx, hl), h_ld, ctx.acode, _ptr_array(ctx, ws),
                                                   _ptr_array(ctx, douts), _ptr_array(ctx, dhs), acc_arr, sc_arr,
                                                   _ptr_array(ctx, dws), _ptr_array(ctx, dbs), M, T, sb, st, _stream()),
                    "jmt_regressor_tail_bwd")
            done = set()
            for g in range(G):
                hv = hidden[g]
                if id(hv) in done:
                    continue
                done.add(id(hv))
                gb = seen[id(hv)]
                ctx.add_grad(hv, gb)
                gb.refs -= 1
        ctx.tape.append(bwd)

    def set_gout(g, t):
        gouts[g] = t
    return outs, set_gout


def regressor_heads(ctx: Ctx, x: Var, pre: Sequence[str], B: int, T: int, time_major: bool):
    """The two regressor heads `Linear(dim,128) -> ReLU -> Dropout(p = 0) -> Linear(128,1)` (two_transformers.py:104-114) on the
    same features: the hidden layers run as ONE N = 256 GEMM on the stacked weights (valence rows | arousal rows), the
    Linear(128,1) pair as the fused regressor tail, the backward as one batched wgrad launch (the two CTAs of a pair take the
    two heads and share the x tile), one K = 256 dgrad and the tail kernel (which already applies the ReLU mask).
    `pre` = the two parameter prefixes, e.g. ("vregressor.", "aregressor.").  Returns (outputs, gradient setter)."""
    wn, bn = [p + "0.weight" for p in pre], [p + "0.bias" for p in pre]
    Wc = ctx.cat_params(wn, True)                       # (256, dim) operand dtype
    bc = ctx.cat_params(bn, False)                      # (256,) fp32
    M, K = x.data.shape
    H = ctx.empty((M, 256))
    gemm(ctx, x.data, Wc, H, M=M, N=256, K=K, bias=bc, act=L.ACT_RELU)
    shape = (T, B) if time_major else (B, T)
    sb, st = (1, B) if time_major else (T, 1)
    outs = [ctx.empty(shape, torch.float32) for _ in range(2)]
    hl = [H[:, 0:128], H[:, 128:256]]
    ws = [ctx.p(p + "3.weight")[0] for p in pre]
    bs = [ctx.p(p + "3.bias")[0:1] for p in pre]
    L.check(ctx.lib.jmt_regressor_tail_fwd(2, _ptr_array(ctx, hl), 256, ctx.acode, _ptr_array(ctx, ws), _ptr_array(ctx, bs),
                                           _ptr_array(ctx, outs), M, T, sb, st, _stream()), "jmt_regressor_tail_fwd")
    gouts: List[Optional[torch.Tensor]] = [None, None]
    if ctx.record:
        def bwd():
            douts = []
            for g in range(2):
                d = gouts[g]
                douts.append(ctx.zeros(shape, torch.float32) if d is None else contig_f32_2d(ctx, d))
            dH = ctx.empty((M, 256))
            dhl = [dH[:, 0:128], dH[:, 128:256]]
            dws = [ctx.pgrad(p + "3.weight")[0] for p in pre]
            dbs = [ctx.pgrad(p + "3.bias")[0:1] for p in pre]
            acc_arr, sc_arr = (C.c_int * 2)(0, 0), (C.c_float * 2)(1.0, 1.0)
            ctx.keep += [acc_arr, sc_arr]
            # dH = d(pre-activation hidden): the tail kernel applies the ReLU mask (h > 0) itself
            L.check(ctx.lib.jmt_regressor_tail_bwd(2, _ptr_array(ctx, hl), 256, ctx.acode, _ptr_array(ctx, ws), _ptr_array(ctx, douts),
                                                   _ptr_array(ctx, dhl), acc_arr, sc_arr, _ptr_array(ctx, dws), _ptr_array(ctx, dbs),
                                                   M, T, sb, st, _stream()), "jmt_regressor_tail_bwd")
            for g in range(2):
                L.check(ctx.lib.jmt_colsum(_ptr(dhl[g]), ctx.acode, 256, M, 128, _ptr(ctx.pgrad(bn[g])), _stream()), "jmt_colsum")
            # dW_v | dW_a = dH_v^T x | dH_a^T x: one launch, batch entry = head (A = column half of dH, B = x shared, D = the two
            # parameter gradients at their distance in the flat bucket)
            g0, g1 = ctx.pgrad(wn[0]), ctx.pgrad(wn[1])
            dist = (g1.data_ptr() - g0.data_ptr()) // 4
            if dist > 0 and dist % 4 == 0:
                gemm(ctx, dH, x.data, g0, M=128, N=K, K=M, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, a_rows=M, b_rows=M,
                     a_ld=256, b_ld=x.data.stride(0), d_ld=K, nb1=2, a_bs=(0, 128), b_bs=(0, 0), d_bs=(0, dist),
                     store=L.ATOMIC_ADD, split_k=wgrad_split(M, 128, K, 2))
            else:
                for g, gw in enumerate((g0, g1)):
                    gemm(ctx, dhl[g], x.data, gw, M=128, N=K, K=M, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, a_rows=M, b_rows=M,
                         a_ld=256, store=L.ATOMIC_ADD, split_k=wgrad_split(M, 128, K))
            if x.needs_grad:
                dx, mode = ctx.grad_target(x, capable=True)
                _dgrad_gemm(ctx, x, dH, Wc, dx, mode, M=M, N=K, K=256, b_major=L.MAJOR_MN)
        ctx.tape.append(bwd)

    def set_gout(g, t):
        gouts[g] = t
    return outs, set_gout


# --------------------------------------------------------------------------- TCN ops
# Flat padded layout: all N sequences live in ONE channels-last matrix of N * (pad + L) rows, row = n*(pad+L) + pad + t,
# with `pad` >= (k-1)*max_dilation ZERO rows in front of every sequence.  Those rows are the causal left padding of the
# forward convs (Chomp1d, temporal_convolutional_model.py:12-18) and -- being the rows right after the previous
# sequence -- the right padding of its anti-causal dgrad, so the implicit GEMM runs over flat 128-row tiles with plain
# row shifts instead of one M = L problem per sequence (L = 300 fills 2.34 tiles of 128 rows: 22 % of every conv GEMM was
# padding; L = 7 clips: 95 %).  Every GEMM that writes activations keeps the padding rows zero (zero_rows).
def tcn_pad(specs) -> int:
    """Padding rows per sequence for a TemporalConvNet: the largest causal reach (k-1)*dilation of its levels."""
    return max((k - 1) * d for (_cin, _cout, k, d, _p, _ds) in specs)


def transpose_in(ctx: Ctx, x: torch.Tensor, needs_grad: bool, pad: int = 0):
    """(N, C, L) external tensor -> channels-last activation Var of shape (N*(pad+L), C), padding rows zero."""
    require_cuda(x)
    xc = x.contiguous()
    if xc.dtype not in _DT:
        raise RuntimeError(f"unsupported input dtype {xc.dtype}")
    N, Cc, Ls = xc.shape
    Lp = Ls + pad
    out = ctx.zeros((N * Lp, Cc)) if pad else ctx.empty((N * Lp, Cc))
    L.check(ctx.lib.jmt_transpose_strided(_ptr(xc), _DT[xc.dtype], 0, _ptr(out[pad:]), ctx.acode, Lp * Cc, N, Cc, Ls, _stream()),
            "jmt_transpose")
    v = Var(out, needs_grad)
    holder: dict = {}
    if ctx.record and needs_grad:
        def bwd():
            dx = ctx.empty((N, Cc, Ls), torch.float32)
            if v.grad is None:
                cuda_memset0(dx)
            else:
                L.check(ctx.lib.jmt_transpose_strided(_ptr(v.grad[pad:]), ctx.acode, Lp * Cc, _ptr(dx), L.F32, 0, N, Ls, Cc,
                                                      _stream()), "jmt_transpose")
            ctx.release(v)
            holder["dx"] = dx
        ctx.tape.append(bwd)
        return v, (lambda: holder["dx"])
    return v, None


def weight_norm_conv_weights(ctx: Ctx, prefix: str, cout: int, cin: int, k: int):
    """w = g * v / ||v|| (legacy weight_norm), emitted in the two implicit-GEMM layouts:
    forward (Cout, k*Cin) tap-major and dgrad (Cin, k*Cout).  Records the weight_norm backward that turns
    the fp32 tap-major dW into d(weight_g), d(weight_v)."""
    g, v = ctx.p(prefix + "weight_g"), ctx.p(prefix + "weight_v")
    w_fwd = ctx.empty((cout, k * cin))
    w_dg = ctx.empty((cin, k * cout)) if ctx.record else None
    norm = ctx.empty((cout,), torch.float32)
    L.check(ctx.lib.jmt_weight_norm_fwd(_ptr(g), _ptr(v), _ptr(w_fwd), _ptr(w_dg), ctx.acode, _ptr(norm), cout, cin, k,
                                        _stream()), "jmt_weight_norm_fwd")
    dw = {"t": None}
    if ctx.record:
        def bwd():
            if dw["t"] is None:
                return
            L.check(ctx.lib.jmt_weight_norm_bwd(_ptr(dw["t"]), _ptr(g), _ptr(v), _ptr(norm),
                                                _ptr(ctx.pgrad(prefix + "weight_g")), _ptr(ctx.pgrad(prefix + "weight_v")),
                                                cout, cin, k, _stream()), "jmt_weight_norm_bwd")
        ctx.tape.append(bwd)
    return w_fwd, w_dg, dw


def channel_dropout_masks(ctx: Ctx, N: int, couts: Sequence[int], p: float) -> List[Optional[torch.Tensor]]:
    """Philox keep-masks (N, cout_i) of every Dropout2d of a TemporalConvNet in ONE launch (eight ~3 us launches before);
    None entries when dropout is off."""
    if p <= 0.0 or not ctx.training:
        return [None] * len(couts)
    sizes = [(N * c + 15) // 16 * 16 for c in couts]          # 16-byte aligned slices
    buf = ctx.empty((sum(sizes),), torch.uint8)
    L.check(ctx.lib.jmt_dropout_mask(_ptr(buf), buf.numel(), p, ctx.seed, ctx.rng_offset, _ptr(ctx.rng_state), _stream()), "jmt_dropout_mask")
    ctx.rng_offset += (buf.numel() + 3) // 4
    out, o = [], 0
    for c, sz in zip(couts, sizes):
        out.append(buf[o:o + N * c])
        o += sz
    return out


def weight_norm_all(ctx: Ctx, convs: Sequence[Tuple[str, int, int]], k: int):
    """weight_norm of several convolutions (prefix, cout, cin) -- every conv of a TemporalConvNet -- in ONE launch per layout
    (jmt_weight_norm_fwd_batched; the per-conv kernels are latency-bound, 8 x 2 x ~10 us per pass).  Returns
    {prefix: (w_fwd, w_dgrad, dw holder, norm)}; the backward is recorded per group by weight_norm_bwd_group."""
    out, g, v, wf, wd, nm, co, ci = {}, [], [], [], [], [], [], []
    for prefix, cout, cin in convs:
        w_fwd = ctx.empty((cout, k * cin))
        w_dg = ctx.empty((cin, k * cout)) if ctx.record else None
        norm = ctx.empty((cout,), torch.float32)
        out[prefix] = (w_fwd, w_dg, {"t": None}, norm)
        g.append(ctx.p(prefix + "weight_g")); v.append(ctx.p(prefix + "weight_v"))
        wf.append(w_fwd); wd.append(w_dg); nm.append(norm); co.append(cout); ci.append(cin)
    n = len(convs)
    arr = lambda ts: (C.c_void_p * n)(*[(t.data_ptr() if t is not None else None) for t in ts])     # noqa: E731
    ia = lambda xs: (C.c_int * n)(*xs)                                                               # noqa: E731
    L.check(ctx.lib.jmt_weight_norm_fwd_batched(n, arr(g), arr(v), arr(wf), arr(wd), ctx.acode, arr(nm), ia(co), ia(ci), k, _stream()),
            "jmt_weight_norm_fwd_batched")
    return out


def weight_norm_bwd_group(ctx: Ctx, weights: dict, convs: Sequence[Tuple[str, int, int]], k: int):
    """Record ONE weight_norm backward launch for a group of convolutions (a TCN level): record it BEFORE the convs so that it
    runs after their weight-gradient GEMMs in the reversed tape."""
    if not ctx.record:
        return

    def bwd():
        n = len(convs)
        dws = [weights[p][2]["t"] for p, _, _ in convs]
        if all(d is None for d in dws):
            return
        arr = lambda ts: (C.c_void_p * n)(*[(t.data_ptr() if t is not None else None) for t in ts])     # noqa: E731
        ia = lambda xs: (C.c_int * n)(*xs)                                                               # noqa: E731
        L.check(ctx.lib.jmt_weight_norm_bwd_batched(
            n, arr(dws), arr([ctx.p(p + "weight_g") for p, _, _ in convs]), arr([ctx.p(p + "weight_v") for p, _, _ in convs]),
            arr([weights[p][3] for p, _, _ in convs]), arr([ctx.pgrad(p + "weight_g") for p, _, _ in convs]),
            arr([ctx.pgrad(p + "weight_v") for p, _, _ in convs]), ia([c for _, c, _ in convs]), ia([c for _, _, c in convs]), k,
            _stream()), "jmt_weight_norm_bwd_batched")
    ctx.tape.append(bwd)


def causal_conv(ctx: Ctx, x: Var, prefix: str, N: int, Ls: int, cin: int, cout: int, k: int, dil: int, act: int,
                drop_p: float = 0.0, pad: int = 0, weights: Optional[tuple] = None, fold_act: bool = False,
                keep_mask: Optional[torch.Tensor] = None) -> Var:
    """weight-normed dilated causal Conv1d + Chomp1d + LeakyReLU + Dropout2d (temporal_convolutional_model.py:24-29)
    as an implicit GEMM on the flat padded channels-last layout (see above): taps are K blocks whose A rows are
    shifted by -(k-1-j)*dil; the zero padding rows in front of every sequence supply the causal zeros.
    Channel dropout (training, p > 0; SURVEY Q12: whole channels per sample) is a per-(sample, channel) scale in the
    GEMM epilogue; its backward, the activation gradient and the bias gradient are one fused pass."""
    assert pad >= (k - 1) * dil, "the padding rows must cover the causal reach of this conv"
    Lp = Ls + pad
    R = N * Lp
    assert x.data.shape[0] == R
    if weights is not None:            # computed (and differentiated) together with the other convs: weight_norm_all
        w_fwd, w_dg, dwh = weights[0], weights[1], weights[2]
    else:
        w_fwd, w_dg, dwh = weight_norm_conv_weights(ctx, prefix, cout, cin, k)   # recorded first => runs last in backward
    y = ctx.empty((R, cout))
    bias = ctx.p(prefix + "bias")
    mask, mscale = None, 1.0
    if drop_p > 0.0 and ctx.training:
        if keep_mask is not None:          # drawn together with the masks of the other convs (channel_dropout_masks)
            mask = keep_mask
        else:
            mask = ctx.empty((N * cout,), torch.uint8)
            L.check(ctx.lib.jmt_dropout_mask(_ptr(mask), N * cout, drop_p, ctx.seed, ctx.rng_offset, _ptr(ctx.rng_state), _stream()), "jmt_dropout_mask")
            ctx.rng_offset += (N * cout + 3) // 4
        mscale = 1.0 / (1.0 - drop_p)
    # algorithmic FLOPs = useful taps only: tap j touches L - (k-1-j)*dil positions of each sequence
    tap_flops = [2.0 * N * max(0, Ls - (k - 1 - j) * dil) * cout * cin for j in range(k)]
    # the GEMM epilogue stages the keep-flags of at most two samples per 128-row tile: sequences shorter than that (the
    # reference-faithful placement with L = 7 clips, tsav.py:214-216) get the mask from a separate pass over y
    fuse_mask = mask is not None and Lp >= 127
    gemm(ctx, x.data, w_fwd, y, M=R, N=cout, K=cin, a_rows=R, b_rows=cout, a_ld=cin, b_ld=k * cin, d_ld=cout,
         bias=bias, act=act, slope=LEAKY_SLOPE, ntaps=k, a_shift=(-(k - 1) * dil, dil),
         colmask=mask if fuse_mask else None, colmask_scale=mscale, colmask_row_period=Lp if fuse_mask else 0,
         zero_rows=(Lp, pad), alg_flops=sum(tap_flops))
    if mask is not None and not fuse_mask:
        L.check(ctx.lib.jmt_apply_mask(_ptr(y), _ptr(mask), _ptr(y), N, Lp, cout, 1, mscale, ctx.acode, _stream()), "jmt_apply_mask")
    out = Var(y)
    if ctx.record and ctx.ext_on() and fold_act and act != L.ACT_NONE and (mask is None or fuse_mask):
        # the only consumer is the next conv's GEMM: its dgrad epilogue applies the channel mask and LeakyReLU'(y) and sums the
        # columns (this conv's bias gradient), so no pass over d(y) runs here
        out.fold = (LEAKY_SLOPE, mask, mscale, Lp if mask is not None else 0)
        out.track_colsum = True
        if EPI_COLSUM:
            out.colsum_direct = lambda: ctx.pgrad(prefix + "bias")
    if ctx.record and out.fold is None and (act != L.ACT_NONE or mask is not None) and cout % 8 == 0 and (mask is None or fuse_mask):
        out.act_pending = (LEAKY_SLOPE if act != L.ACT_NONE else 1.0, mask, Lp, mscale, lambda: ctx.pgrad(prefix + "bias"))
    if ctx.record:
        def bwd():
            dy = out.grad
            if dy is None:
                return
            assert dy.is_contiguous()
            # where a channel was dropped y = 0 and the masked dy is 0, so act'(y) of the post-dropout y is exact;
            # padding rows of dy are zero (every producer keeps them so) and stay zero
            if out.act_done:
                pass        # dy is d(pre-activation) already and the bias gradient has been summed (add_act's fused backward)
            elif out.folded:
                # dy is d(pre-activation) already (consumer's dgrad epilogue); its column sums = the bias gradient
                if out.colsum_direct is None:
                    L.check(ctx.lib.jmt_colsum(_ptr(dy), _DT[dy.dtype], cout, R, cout, _ptr(ctx.pgrad(prefix + "bias")), _stream()), "jmt_colsum")
            elif act != L.ACT_NONE or mask is not None:
                dy = _act_bwd(ctx, out, dy, LEAKY_SLOPE if act != L.ACT_NONE else 1.0, colsum=ctx.pgrad(prefix + "bias"),
                              mask=mask, mask_rows=Lp, mask_scale=mscale)
            else:
                L.check(ctx.lib.jmt_colsum(_ptr(dy), _DT[dy.dtype], cout, R, cout, _ptr(ctx.pgrad(prefix + "bias")),
                                           _stream()), "jmt_colsum")
            # wgrad: dW_j (Cout, Cin) = dy^T shift_j(x) over all flat rows (split-K, fp32 reduce-add).  The k taps are ONE
            # launch: tap j is batch entry j, whose B operand is x moved down by j*dil rows (an overlapping batch stride) under a
            # common shift of -(k-1)*dil, whose A operand dy is shared (stride 0) and whose D is column block j of dW.  The
            # rows the common shift zero-fills but tap j would have read (x rows 0 .. j*dil-1) are padding rows of the first
            # sequence, i.e. zeros (pad >= (k-1)*dil is asserted above).
            dw = ctx.zeros((cout, k * cin), torch.float32)
            # (B's row extent stops (k-1)*dil short of R: with tap j's base moved down by j*dil rows nothing past row R-1 of
            # x is ever addressed, also not by the zero-filled tail of the last 64-row k-block)
            gemm(ctx, dy, x.data, dw, M=cout, N=cin, K=R, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, a_rows=R,
                 b_rows=R - (k - 1) * dil, a_ld=cout, b_ld=cin, d_ld=k * cin, nb1=k, a_bs=(0, 0), b_bs=(0, dil * cin), d_bs=(0, cin),
                 b_shift=(-(k - 1) * dil, 0), store=L.ATOMIC_ADD, split_k=wgrad_split(R, cout, cin, k), alg_flops=sum(tap_flops))
            dwh["t"] = dw
            if x.needs_grad:
                # dgrad: dx[r] = sum_j W_j^T dy[r + (k-1-j) dil]; rows past a sequence's end are the next one's zero padding
                dx, mode = ctx.grad_target(x, capable=True)
                _dgrad_gemm(ctx, x, dy, w_dg, dx, mode, M=R, N=cin, K=cout, a_rows=R, b_rows=cin, a_ld=cout, b_ld=k * cout, d_ld=cin,
                            ntaps=k, a_shift=((k - 1) * dil, -dil), zero_rows=(Lp, pad), alg_flops=sum(tap_flops))
            ctx.release(out)
        ctx.tape.append(bwd)
    return out


def add_act(ctx: Ctx, a: Var, b: Var, act: int, slope: float, a_exclusive: bool = False) -> Var:
    """act(a + b)  (TemporalBlock residual, temporal_convolutional_model.py:54-57).
    a_exclusive: the caller guarantees that this op is the ONLY consumer of `a`; when `a` is a conv output with a pending
    activation / channel-dropout backward (Var.act_pending), one fused pass then produces d(b) and d(pre-activation of a) together
    with that conv's bias gradient (jmt_add_act_bwd_fused) instead of jmt_act_bwd followed by jmt_act_bwd_fused."""
    assert a.data.is_contiguous() and b.data.is_contiguous()
    y = ctx.empty(a.data.shape)
    L.check(ctx.lib.jmt_add_act(_ptr(a.data), _ptr(b.data), _ptr(y), y.numel(), act, slope, ctx.acode, _stream()), "jmt_add_act")
    out = Var(y)
    if ctx.record:
        def bwd():
            dy = out.grad
            if dy is None:
                return
            dz = GradBuf(ctx.empty(y.shape))
            pend = a.act_pending if (a_exclusive and a.gbuf is None and a.needs_grad and a is not b and act != L.ACT_NONE and
                                     dy.is_contiguous() and ADD_ACT_FUSED_BWD) else None
            if pend is not None:
                slope2, mask, mrows, mscale, bias_grad = pend
                dz2 = GradBuf(ctx.empty(y.shape))
                rows, cols = y.shape
                L.check(ctx.lib.jmt_add_act_bwd_fused(_ptr(dy), _ptr(y), _ptr(a.data), _ptr(mask), _ptr(dz.t), _ptr(dz2.t), rows, cols,
                                                      mrows, mscale, slope, slope2, _ptr(bias_grad()), ctx.acode, _stream()),
                        "jmt_add_act_bwd_fused")
                a.gbuf = dz2                       # d(pre-activation) of the conv that produced a: its backward skips the activation pass
                a.act_done = True
            else:
                L.check(ctx.lib.jmt_act_bwd(_ptr(dy), _ptr(y), _ptr(dz.t), y.numel(), slope, ctx.acode, _stream()), "jmt_act_bwd")
                ctx.add_grad(a, dz)
            ctx.add_grad(b, dz)
            dz.refs -= 1
            ctx.release(out)
        ctx.tape.append(bwd)
    return out


def transpose_out(ctx: Ctx, x: Var, N: int, Ls: int, Cc: int, pad: int = 0):
    """flat padded channels-last (N*(pad+L), C) Var -> external fp32 (N, C, L) tensor; returns (tensor, grad setter)."""
    Lp = Ls + pad
    out = ctx.empty((N, Cc, Ls), torch.float32)
    L.check(ctx.lib.jmt_transpose_strided(_ptr(x.data[pad:]), ctx.acode, Lp * Cc, _ptr(out), L.F32, 0, N, Ls, Cc, _stream()),
            "jmt_transpose")
    g = {"t": None}
    if ctx.record:
        def bwd():
            if g["t"] is None:
                return
            d = g["t"].contiguous()
            gb = GradBuf(ctx.zeros((N * Lp, Cc)) if pad else ctx.empty((N * Lp, Cc)))
            L.check(ctx.lib.jmt_transpose_strided(_ptr(d), _DT[d.dtype], 0, _ptr(gb.t[pad:]), ctx.acode, Lp * Cc, N, Cc, Ls,
                                                  _stream()), "jmt_transpose")
            ctx.add_grad(x, gb)
            gb.refs -= 1
        ctx.tape.append(bwd)
    return out, (lambda t: g.__setitem__("t", t))


def unpad_rows(ctx: Ctx, x: Var, N: int, Ls: int, pad: int) -> Var:
    """flat padded (N*(pad+L), C) Var -> compact (N*L, C) Var (row = n*L + t), e.g. the (B, T, 512) view
    I3D_WSDDA.forward returns (I3DWSDDA.py:44) that the fusion consumes."""
    if pad == 0:
        return x
    Cc = x.data.shape[1]
    Lp = Ls + pad
    y = ctx.empty((N * Ls, Cc))
    L.check(ctx.lib.jmt_copy_rows3d(_ptr(x.data[pad:]), ctx.acode, Lp * Cc, _ptr(y), ctx.acode, Ls * Cc, N, Ls, Cc, _stream()),
            "jmt_copy_rows3d")
    out = Var(y)
    if ctx.record:
        def bwd():
            if out.grad is None:
                return
            gb = GradBuf(ctx.zeros((N * Lp, Cc)))
            L.check(ctx.lib.jmt_copy_rows3d(_ptr(out.grad), ctx.acode, Ls * Cc, _ptr(gb.t[pad:]), ctx.acode, Lp * Cc, N, Ls, Cc,
                                            _stream()), "jmt_copy_rows3d")
            ctx.add_grad(x, gb)
            gb.refs -= 1
            ctx.release(out)
        ctx.tape.append(bwd)
    return out


def time_max(ctx: Ctx, x: Var, N: int, Ls: int, pad: int) -> Var:
    """max over the L positions of every sequence of a flat padded channels-last Var -> (N, C) Var
    (`torch.max(temporal(...).transpose(1,2), 1)`: I3DWSDDA.py:44 + tsav.py:216)."""
    Cc = x.data.shape[1]
    Lp = Ls + pad
    y = ctx.empty((N, Cc))
    arg = ctx.empty((N, Cc), torch.int32)
    L.check(ctx.lib.jmt_time_max_fwd(_ptr(x.data[pad:]), Lp * Cc, N, Ls, Cc, _ptr(y), _ptr(arg), ctx.acode, _stream()), "jmt_time_max_fwd")
    out = Var(y)
    if ctx.record:
        def bwd():
            if out.grad is None:
                return
            gb = GradBuf(ctx.zeros((N * Lp, Cc)))
            L.check(ctx.lib.jmt_time_max_bwd(_ptr(out.grad), _ptr(arg), Lp * Cc, N, Cc, _ptr(gb.t[pad:]), ctx.acode, _stream()),
                    "jmt_time_max_bwd")
            ctx.add_grad(x, gb)
            gb.refs -= 1
            ctx.release(out)
        ctx.tape.append(bwd)
    return out


def to_external(ctx: Ctx, x: Var, shape):
    """Activation Var -> external fp32 tensor of `shape` (row-major compatible); returns (tensor, setter)."""
    out = ctx.empty(shape, torch.float32)
    copy2d(ctx, x.data, out.view(-1, shape[-1]))
    g = {"t": None}
    if ctx.record:
        def bwd():
            if g["t"] is None:
                return
            d = g["t"].contiguous()
            gb = GradBuf(ctx.empty(x.data.shape))
            copy2d(ctx, d.view(-1, shape[-1]), gb.t)
            ctx.add_grad(x, gb)
            gb.refs -= 1
        ctx.tape.append(bwd)
    return out, (lambda t: g.__setitem__("t", t))


def stack_rows(ctx: Ctx, parts: List[Var]) -> Var:
    """torch.stack(parts, dim=2).flatten(0,1).permute(1,0,2) of the reference (length-L 'sequences' per
    (b, t); intra_modal_transformer_fusion.py:93-99, mm_multi_transformers.py:173-184): rows l*M + m."""
    M, E = parts[0].data.shape
    buf = ctx.empty((len(parts) * M, E))
    for l, pv in enumerate(parts):
        copy2d(ctx, pv.data, buf[l * M:(l + 1) * M])
    out = Var(buf)
    if ctx.record:
        def bwd():
            g = out.grad
            if g is None:
                return
            for l, pv in enumerate(parts):
                if not pv.needs_grad:
                    continue
                gb = GradBuf(ctx.empty((M, E)))
                copy2d(ctx, g[l * M:(l + 1) * M], gb.t)
                ctx.add_grad(pv, gb)
                gb.refs -= 1
            ctx.release(out)
        ctx.tape.append(bwd)
    return out


def take_rows(ctx: Ctx, x: Var, r0: int, r1: int) -> Var:
    """x[r0:r1] as a new Var (the reference's `[:, :, -1, :]` last-token select)."""
    out = Var(x.data[r0:r1])
    if ctx.record:
        def bwd():
            g = out.grad
            if g is None:
                return
            full = _proj_grad(ctx, x)
            t = full[r0:r1]
            L.check(ctx.lib.jmt_axpy(_ptr(g), _ptr(t), 1.0, g.numel(), _DT[g.dtype], _stream()), "jmt_axpy")
            ctx.release(out)
        ctx.tape.append(bwd)
    return out
