import os, sys, torch
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
dev = 'cuda'
ctx = E.Ctx({}, 'bf16', False, False)
def bench1(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
def bench(fn, iters=20):
    return min(bench1(fn, iters) for _ in range(3))
M = 76800
only = sys.argv[1] if len(sys.argv) > 1 else None
shapes = [(M, 512, 512), (M, 1536, 512), (M, 1024, 3072), (M, 512, 1024), (8192, 8192, 8192)]
for (m, n, k) in shapes:
    a = torch.randn(m, k, device=dev).bfloat16(); b = torch.randn(n, k, device=dev).bfloat16()
    d = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(n, device=dev)
    fl = 2.0 * m * n * k
    res = []
    for cl in ("0", "1"):
        os.environ["JMT_GEMM_CLUSTER"] = cl
        us = bench(lambda: E.gemm(ctx, a, b, d, M=m, N=n, K=k, bias=bias))
        res.append(f"cl{cl} {us:8.1f}us {fl/us/1e6:7.1f}TF")
    if only != "nocublas":
        us = bench(lambda: torch.matmul(a, b.t()))
        res.append(f"cublas {us:8.1f}us {fl/us/1e6:7.1f}TF")
    print(f"linear {m}x{n}x{k}: " + " | ".join(res), flush=True)
# conv fwd as in the TCN: (N*L, 512) x (512, 5*512)
N_, Ls, cin, cout, k = 256, 300, 512, 512, 5
x = torch.randn(N_ * Ls, cin, device=dev).bfloat16(); w = torch.randn(cout, k * cin, device=dev).bfloat16()
y = torch.empty(N_ * Ls, cout, device=dev, dtype=torch.bfloat16); bias = torch.randn(cout, device=dev)
fl = 2.0 * N_ * Ls * cout * cin * k
for cl in ("0", "1"):
    os.environ["JMT_GEMM_CLUSTER"] = cl
    us = bench(lambda: E.gemm(ctx, x, w, y, M=Ls, N=cout, K=cin, a_rows=Ls, b_rows=cout, a_ld=cin, b_ld=k * cin, d_ld=cout,
         nb0=1, nb1=N_, a_bs=(0, Ls * cin), d_bs=(0, Ls * cout), bias=bias, act=2, slope=0.01, ntaps=k, a_shift=(-(k - 1) * 2, 2)))
    print(f"conv fwd 256x300x512x512x5 cl{cl}: {us:8.1f}us {fl/us/1e6:7.1f}TF (nominal taps)", flush=True)
# wgrad
m, n, kk = 512, 512, M
dy = torch.randn(kk, m, device=dev).bfloat16(); xx = torch.randn(kk, n, device=dev).bfloat16()
dw = torch.zeros(m, n, device=dev)
for cl in ("0", "1"):
    os.environ["JMT_GEMM_CLUSTER"] = cl
    for sk in (18, 37):
        us = bench(lambda: E.gemm(ctx, dy, xx, dw, M=m, N=n, K=kk, a_major=1, b_major=1, store=2, split_k=sk))
        print(f"wgrad 512x512x{kk} split{sk} cl{cl}: {us:8.1f}us {2.0*m*n*kk/us/1e6:7.1f}TF", flush=True)
