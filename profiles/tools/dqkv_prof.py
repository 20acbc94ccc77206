"""Isolated timing + per-role cycle counters (jmt_attn_bwd_set_profile_buffer) of jmt_attn_bwd_dqkv_bf16 at the C2 attention geometry
(256 windows, Lq = S = 300, dh = 512), next to the three batched jmt_gemm_bf16 launches it replaces.   python profiles/tools/dqkv_prof.py"""
import sys, torch, ctypes as C
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
dev = 'cuda'
lib = L.lib()
NB, T, Eh, h = 256, 300, 512, 1
dh = Eh // h
ctx = E.Ctx({}, 'bf16', False, False)
s_ld = (T + 7) // 8 * 8
qkv = (torch.randn(NB * T, 3 * Eh, device=dev) * 0.5).bfloat16()
do = (torch.randn(NB * T, Eh, device=dev) * 0.5).bfloat16()
ds = (torch.randn(NB, h, T, s_ld, device=dev) * 0.1).bfloat16()
pr = torch.rand(NB, h, T, s_ld, device=dev).bfloat16()
g = torch.zeros(NB * T, 3 * Eh, device=dev, dtype=torch.bfloat16)
geo = (3 * Eh, dh, T * 3 * Eh); ogeo = (Eh, dh, T * Eh)
qd, kd, vd = qkv[:, :Eh], qkv[:, Eh:2 * Eh], qkv[:, 2 * Eh:]
sb = (T * s_ld, h * T * s_ld)
flops = 3 * 2.0 * NB * h * T * T * dh
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def new(nparts=3):
    parts = [(kd, geo, ds, 0, g[:, :Eh], geo, L.STORE, 1.0, None),
             (qd, geo, ds, 1, g[:, Eh:2 * Eh], geo, L.STORE, 1.0, None),
             (do, ogeo, pr, 1, g[:, 2 * Eh:], geo, L.STORE, 1.0, None)][:nparts]
    E._attn_bwd_dqkv(ctx, parts, T, T, dh, h, NB, s_ld)


def old():
    E.gemm(ctx, ds, kd, g[:, :Eh], M=T, N=dh, K=T, a_rows=T, b_major=L.MAJOR_MN, b_rows=T, a_ld=s_ld, b_ld=3 * Eh, d_ld=3 * Eh,
           nb0=h, nb1=NB, a_bs=sb, b_bs=(dh, T * 3 * Eh), d_bs=(dh, T * 3 * Eh))
    E.gemm(ctx, pr, do, g[:, 2 * Eh:], M=T, N=dh, K=T, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, a_rows=T, b_rows=T, a_ld=s_ld, b_ld=Eh,
           d_ld=3 * Eh, nb0=h, nb1=NB, a_bs=sb, b_bs=(dh, T * Eh), d_bs=(dh, T * 3 * Eh))
    E.gemm(ctx, ds, qd, g[:, Eh:2 * Eh], M=T, N=dh, K=T, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, a_rows=T, b_rows=T, a_ld=s_ld, b_ld=3 * Eh,
           d_ld=3 * Eh, nb0=h, nb1=NB, a_bs=sb, b_bs=(dh, T * 3 * Eh), d_bs=(dh, T * 3 * Eh))


us_old = timed(old)
print(f"three jmt_gemm_bf16 launches: {us_old:8.1f} us  {flops / us_old / 1e6:7.1f} TFLOP/s")
for npart in (3, 1):
    us = timed(lambda: new(npart))
    fl = flops * npart / 3
    buf.zero_()
    lib.jmt_attn_bwd_set_profile_buffer(C.c_void_p(buf.data_ptr()))
    new(npart)
    torch.cuda.synchronize()
    lib.jmt_attn_bwd_set_profile_buffer(None)
    b = buf.view(148, 16).cpu().double()
    lead = b[:, 3] > 0
    m = lambda c, sel: float(b[sel][:, c].mean())       # noqa: E731
    allc = b[:, 5] > 0
    print(f"jmt_attn_bwd_dqkv_bf16 ({npart} parts): {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s | MMA issuer total {m(3, lead):9.0f} wait-A {m(0, lead):8.0f} "
          f"wait-X {m(1, lead):8.0f} wait-TMEM {m(2, lead):8.0f} | epilogue total {m(5, allc):9.0f} wait-acc {m(4, allc):8.0f}")
