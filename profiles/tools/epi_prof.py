"""Per-phase cycles of one epilogue warp (needs the diagnostic build: profiles/tools/epi_prof.sh, JMT_B200_LIB=profiles/tools/ab/lib_prof.so)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
lib = L.lib()
ctx = E.Ctx({}, 'bf16', False, False)
names = ["ld1", "math1+ld2issue", "wait_read", "sts1+ld2wait", "release+math2+sts2", "fence+tma"]
def run(tag, fn, tiles):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8)()
    lib.jmt_gemm_epi_prof_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.jmt_gemm_epi_prof_read(buf, 1)
    tot = sum(buf[:6])
    print(f"{tag}: {e0.elapsed_time(e1)*1e3:.1f}us epilogue-warp cycles {tot} | " + " ".join(f"{n}={buf[i]}" for i, n in enumerate(names)), flush=True)
M = 76800
for (m, n, k) in [(M, 512, 512), (M, 1536, 512), (M, 1024, 3072)]:
    a = torch.randn(m, k, device='cuda').bfloat16(); b = torch.randn(n, k, device='cuda').bfloat16()
    d = torch.empty(m, n, device='cuda', dtype=torch.bfloat16); bias = torch.randn(n, device='cuda')
    run(f"lin {m}x{n}x{k}", lambda: E.gemm(ctx, a, b, d, M=m, N=n, K=k, bias=bias), 0)
