import os, sys, torch, ctypes as C, math
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
dev = 'cuda'
lib = L.lib()
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
names = ["tma_wait_empty", "tma_total", "mma_wait_full1", "mma_wait_full2", "mma_wait_x", "mma_wait_tempty", "mma_total", "row_wait_t1", "row_op", "row_wait_t2", "row_epi", "row_total"]
NB, T, Eh, h = 256, 300, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = E.Ctx({}, 'bf16', True, False)
qkv = E.Var((torch.randn(NB * T, 3 * Eh, device=dev) * 0.5).bfloat16())
g = E.AttnGeom(T, NB, 1, T)
def run(tag, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    lib.jmt_attn_set_profile_buffer(C.c_void_p(buf.data_ptr()))
    buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    lib.jmt_attn_set_profile_buffer(None)
    b = buf.view(148, 16).cpu().double()
    s = f"{tag}: {e0.elapsed_time(e1)*1e3:.1f}us |"
    for i, n in enumerate(names):
        s += f" {n}={b[:, i].mean().item():.0f}"
    print(s, flush=True)
dh = Eh // h
s_ld = (T + 7) // 8 * 8
probs = torch.empty(NB, h, T, s_ld, device=dev, dtype=torch.bfloat16)
ds = torch.empty_like(probs)
o = torch.empty(NB * T, Eh, device=dev, dtype=torch.bfloat16)
do = (torch.randn(NB * T, Eh, device=dev) * 0.5).bfloat16()
dq = torch.zeros(NB * T, 3 * Eh, device=dev, dtype=torch.bfloat16)
qd, kd, vd = qkv.data[:, :Eh], qkv.data[:, Eh:2*Eh], qkv.data[:, 2*Eh:]
geo = (3 * Eh, dh, T * 3 * Eh); ogeo = (Eh, dh, T * Eh)
sc = 1 / math.sqrt(dh)
run("fwd", lambda: E._attn_chain(ctx, 0, qd, geo, kd, geo, vd, geo, None, probs, o, ogeo, T, T, dh, h, NB, s_ld, sc, L.STORE))
run("bwd", lambda: E._attn_chain(ctx, 1, do, ogeo, vd, geo, kd, geo, probs, ds, dq[:, :Eh], geo, T, T, dh, h, NB, s_ld, sc, L.ACCUMULATE, o_in=o))
run("bwd dS-only (the mode of the step: delta = rowsum(P o dP) in kernel, no GEMM2)",
    lambda: E._attn_chain(ctx, 1, do, ogeo, vd, geo, None, None, probs, ds, None, None, T, T, dh, h, NB, s_ld, sc, L.STORE))
# E1: separate contiguous Q/K/V (row pitch 1024 B instead of 3072 B)
qc, kc, vc = qd.contiguous(), kd.contiguous(), vd.contiguous()
geo2 = (Eh, dh, T * Eh)
run("fwd contiguous qkv", lambda: E._attn_chain(ctx, 0, qc, geo2, kc, geo2, vc, geo2, None, probs, o, ogeo, T, T, dh, h, NB, s_ld, sc, L.STORE))
