import os, sys, torch
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
dev = 'cuda'
ctx = E.Ctx({}, 'bf16', False, False)
m, n, k = [int(v) for v in sys.argv[1:4]]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
a = torch.randn(m, k, device=dev).bfloat16(); b = torch.randn(n, k, device=dev).bfloat16()
d = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
bias = torch.randn(n, device=dev)
for _ in range(iters):
    E.gemm(ctx, a, b, d, M=m, N=n, K=k, bias=bias)
torch.cuda.synchronize()
