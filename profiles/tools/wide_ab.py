"""A/B of the wide (256 x 512) pair tile: run once per JMT_GEMM_WIDE value (the switch is read once per process)."""
import os, sys, torch
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
dev = 'cuda'
ctx = E.Ctx({}, 'bf16', False, False)
def bench(name, fn, flops, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"WIDE={os.environ.get('JMT_GEMM_WIDE', 'default'):8s} {name:36s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s", flush=True)
M = 76800
for (m, n, k) in [(M, 512, 512), (M, 512, 1024), (M, 512, 1536), (M, 1024, 3072), (M, 3072, 1024)]:
    a = torch.randn(m, k, device=dev).bfloat16(); b = torch.randn(n, k, device=dev).bfloat16()
    d = torch.empty(m, n, device=dev, dtype=torch.bfloat16); bias = torch.randn(n, device=dev)
    bench(f"linear {m}x{n}x{k}", lambda: E.gemm(ctx, a, b, d, M=m, N=n, K=k, bias=bias), 2.0 * m * n * k)
# linear wgrad
m, n, k = 512, 512, M
dy = torch.randn(k, m, device=dev).bfloat16(); x = torch.randn(k, n, device=dev).bfloat16()
dw = torch.zeros(m, n, device=dev)
for sk in (18, 37):
    bench(f"wgrad 512x512x{k} split{sk}", lambda: E.gemm(ctx, dy, x, dw, M=m, N=n, K=k, a_major=1, b_major=1, store=2, split_k=sk), 2.0 * m * n * k)
# conv dgrad on the flat padded layout + batched conv wgrad
Nn, Ls, pad, cin, cout, kk, dil = 256, 300, 32, 512, 512, 5, 2
Lp = Ls + pad; R = Nn * Lp
dyc = torch.randn(R, cout, device=dev).bfloat16(); wdg = torch.randn(cin, kk * cout, device=dev).bfloat16()
dx = torch.empty(R, cin, device=dev, dtype=torch.bfloat16)
bench("conv dgrad 84992x512x512 taps5", lambda: E.gemm(ctx, dyc, wdg, dx, M=R, N=cin, K=cout, a_rows=R, b_rows=cin, a_ld=cout, b_ld=kk * cout, d_ld=cin,
      ntaps=kk, a_shift=((kk - 1) * dil, -dil), zero_rows=(Lp, pad)), 2.0 * R * cin * cout * kk)
xc = torch.randn(R, cin, device=dev).bfloat16(); dwc = torch.zeros(cout, kk * cin, device=dev)
for sk in (11, 7):
    bench(f"conv wgrad batched split{sk}", lambda: E.gemm(ctx, dyc, xc, dwc, M=cout, N=cin, K=R, a_major=1, b_major=1, a_rows=R, b_rows=R - (kk - 1) * dil,
          a_ld=cout, b_ld=cin, d_ld=kk * cin, nb1=kk, a_bs=(0, 0), b_bs=(0, dil * cin), d_bs=(0, cin), b_shift=(-(kk - 1) * dil, 0), store=2, split_k=sk),
          2.0 * R * cin * cout * kk)
