"""One linear-layer GEMM shape, a few launches (for `ncu -k regex:gemm_tc`):  python profiles/tools/gemm_one.py M N K iters"""
import sys
import torch
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E
M, N, K, iters = (int(a) for a in sys.argv[1:5])
ctx = E.Ctx({}, 'bf16', False, False)
a = torch.randn(M, K, device='cuda').bfloat16()
b = torch.randn(N, K, device='cuda').bfloat16()
d = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
bias = torch.randn(N, device='cuda')
for _ in range(iters):
    E.gemm(ctx, a, b, d, M=M, N=N, K=K, bias=bias)
torch.cuda.synchronize()
