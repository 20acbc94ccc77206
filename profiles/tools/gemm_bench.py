import sys, torch, time
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
dev = 'cuda'
ctx = E.Ctx({}, 'bf16', False, False)
def bench(name, fn, flops, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:40s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s", flush=True)
M = 76800
shapes = [(M, 512, 512), (M, 1536, 512), (M, 1024, 3072), (M, 512, 1024), (8192, 8192, 8192)]
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
for (m, n, k) in shapes:
    a = torch.randn(m, k, device=dev).bfloat16(); b = torch.randn(n, k, device=dev).bfloat16()
    d = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(n, device=dev)
    bench(f"linear {m}x{n}x{k} bf16out", lambda: E.gemm(ctx, a, b, d, M=m, N=n, K=k, bias=bias), 2.0 * m * n * k)
    if which == 'all':
        bench(f"  torch.matmul (cuBLAS) same", lambda: torch.matmul(a, b.t()), 2.0 * m * n * k)
# wgrad
m, n, k = 512, 512, M
dy = torch.randn(k, m, device=dev).bfloat16(); x = torch.randn(k, n, device=dev).bfloat16()
dw = torch.zeros(m, n, device=dev)
for sk in (9, 18, 37):
    bench(f"wgrad 512x512x{k} split{sk}", lambda: E.gemm(ctx, dy, x, dw, M=m, N=n, K=k, a_major=1, b_major=1, store=2, split_k=sk), 2.0 * m * n * k)
# attention S and PV
B, T, Eh = 256, 300, 512
q = torch.randn(B * T, 3 * Eh, device=dev).bfloat16()
s = torch.empty(B, 1, T, 304, device=dev)
bench("S=QK^T 300x300x512 nb256 f32out", lambda: E.gemm(ctx, q[:, :Eh], q[:, Eh:2*Eh], s, M=T, N=T, K=Eh, a_rows=T, b_rows=T, a_ld=3*Eh, b_ld=3*Eh, d_ld=304, nb0=1, nb1=B, a_bs=(Eh, T*3*Eh), b_bs=(Eh, T*3*Eh), d_bs=(T*304, T*304)), 2.0 * B * T * T * Eh)
p = torch.randn(B, 1, T, 304, device=dev).bfloat16()
o = torch.empty(B * T, Eh, device=dev, dtype=torch.bfloat16)
bench("O=PV 300x512x300 nb256", lambda: E.gemm(ctx, p, q[:, 2*Eh:], o, M=T, N=Eh, K=T, a_rows=T, b_major=1, b_rows=T, a_ld=304, b_ld=3*Eh, d_ld=Eh, nb0=1, nb1=B, a_bs=(T*304, T*304), b_bs=(Eh, T*3*Eh), d_bs=(Eh, T*Eh)), 2.0 * B * T * T * Eh)
