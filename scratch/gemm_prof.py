import os, sys, torch, ctypes as C
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
dev = 'cuda'
ctx = E.Ctx({}, 'bf16', False, False)
lib = L.lib()
buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
names = ["mma_wait_full", "mma_wait_tempty", "mma_total", "tma_wait_empty", "tma_total", "epi_wait_tfull", "epi_bias_bar", "epi_wait_rd", "epi_total"]
def run(tag, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    lib.jmt_gemm_set_profile_buffer(C.c_void_p(buf.data_ptr()))
    buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    lib.jmt_gemm_set_profile_buffer(None)
    b = buf.view(148, 16).cpu().double()
    act = b[:, 2] > 0     # leader CTAs (MMA issuers)
    s = f"{tag}: {e0.elapsed_time(e1)*1e3:.1f}us |"
    for i, n in enumerate(names):
        col = b[:, i]
        sel = col[act] if i < 3 else col[col > 0]
        if sel.numel():
            s += f" {n}={sel.mean().item():.0f}"
    print(s, flush=True)
M = 76800
for (m, n, k) in [(M, 512, 512), (M, 1024, 3072), (M, 512, 1024)]:
    a = torch.randn(m, k, device=dev).bfloat16(); b = torch.randn(n, k, device=dev).bfloat16()
    d = torch.empty(m, n, device=dev, dtype=torch.bfloat16); bias = torch.randn(n, device=dev)
    for cl in ("0", "1"):
        os.environ["JMT_GEMM_CLUSTER"] = cl
        run(f"lin {m}x{n}x{k} cl{cl}", lambda: E.gemm(ctx, a, b, d, M=m, N=n, K=k, bias=bias))
N_, Ls, cin, cout, k = 256, 300, 512, 512, 5
x = torch.randn(N_ * Ls, cin, device=dev).bfloat16(); w = torch.randn(cout, k * cin, device=dev).bfloat16()
y = torch.empty(N_ * Ls, cout, device=dev, dtype=torch.bfloat16); bias = torch.randn(cout, device=dev)
for cl in ("0", "1"):
    os.environ["JMT_GEMM_CLUSTER"] = cl
    run(f"conv cl{cl}", lambda: E.gemm(ctx, x, w, y, M=Ls, N=cout, K=cin, a_rows=Ls, b_rows=cout, a_ld=cin, b_ld=k * cin, d_ld=cout,
         nb0=1, nb1=N_, a_bs=(0, Ls * cin), d_bs=(0, Ls * cout), bias=bias, act=2, slope=0.01, ntaps=k, a_shift=(-(k - 1) * 2, 2)))
