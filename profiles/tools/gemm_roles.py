"""Per-role cycle counters (jmt_gemm_set_profile_buffer) + isolated timing of selected GEMM shapes of the C2 step:
where does a CTA's time go -- MMA issuer waiting for operands (full) or for the epilogue (tmem-empty), TMA producer waiting
for free slots, epilogue waiting for the accumulator?   python profiles/tools/gemm_roles.py"""
import sys
import torch
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
dev = 'cuda'
ctx = E.Ctx({}, 'bf16', False, False)
lib = L.lib()
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)


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


def roles(name, fn, flops):
    us = timed(fn)
    prof.zero_()
    lib.jmt_gemm_set_profile_buffer(E._ptr(prof))
    fn()
    torch.cuda.synchronize()
    lib.jmt_gemm_set_profile_buffer(None)
    p = prof.view(148, 16).double().cpu()
    act = p[:, 4] > 0                      # CTAs whose producer ran
    lead = p[:, 2] > 0                     # CTAs that issued MMAs (leaders of a pair / every CTA of a 1-CTA launch)
    m = lambda c, sel: float(p[sel][:, c].mean()) if sel.any() else 0.0   # noqa: E731
    print(f"{name:44s} {us:8.1f} us {flops / us / 1e6:7.1f} TF/s | MMA total {m(2, lead):9.0f} wait-full {m(0, lead):8.0f} wait-tmem {m(1, lead):8.0f}"
          f" | TMA total {m(4, act):9.0f} wait-empty {m(3, act):8.0f} | EPI total {m(8, act):9.0f} wait-tfull {m(5, act):8.0f} bias-bar {m(6, act):7.0f}",
          flush=True)


M = 76800
for (m_, n_, k_) in [(M, 512, 512), (M, 1536, 512), (M, 512, 1024), (M, 1024, 3072)]:
    a = torch.randn(m_, k_, device=dev).bfloat16()
    b = torch.randn(n_, k_, device=dev).bfloat16()
    d = torch.empty(m_, n_, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(n_, device=dev)
    roles(f"linear {m_}x{n_}x{k_}", lambda: E.gemm(ctx, a, b, d, M=m_, N=n_, K=k_, bias=bias), 2.0 * m_ * n_ * k_)
    bt = b.t().contiguous()
    roles(f"dgrad  {m_}x{n_}x{k_} (B MN-major)", lambda: E.gemm(ctx, a, bt, d, M=m_, N=n_, K=k_, b_major=L.MAJOR_MN), 2.0 * m_ * n_ * k_)
# attention backward GEMMs of one (B=256, T=300, E=512, h=1) attention
B, T, Eh, s_ld = 256, 300, 512, 304
ds = torch.randn(B, 1, T, s_ld, device=dev).bfloat16()
kv = torch.randn(B * T, 2 * Eh, device=dev).bfloat16()
q = torch.randn(B * T, Eh, device=dev).bfloat16()
do = torch.randn(B * T, Eh, device=dev).bfloat16()
dq = torch.empty(B * T, Eh, device=dev, dtype=torch.bfloat16)
dkv = torch.empty(B * T, 2 * Eh, device=dev, dtype=torch.bfloat16)
sb = (T * s_ld, T * s_ld)
fl = 2.0 * B * T * T * Eh
roles("attn dQ = dS K   300x512x300 nb256", lambda: E.gemm(ctx, ds, kv[:, :Eh], dq, M=T, N=Eh, K=T, a_rows=T, b_major=L.MAJOR_MN, b_rows=T,
      a_ld=s_ld, b_ld=2 * Eh, d_ld=Eh, nb0=1, nb1=B, a_bs=sb, b_bs=(Eh, T * 2 * Eh), d_bs=(Eh, T * Eh)), fl)
roles("attn dV = P^T dO 300x512x300 nb256", lambda: E.gemm(ctx, ds, do, dkv[:, Eh:], M=T, N=Eh, K=T, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN,
      a_rows=T, b_rows=T, a_ld=s_ld, b_ld=Eh, d_ld=2 * Eh, nb0=1, nb1=B, a_bs=sb, b_bs=(Eh, T * Eh), d_bs=(Eh, T * 2 * Eh)), fl)
