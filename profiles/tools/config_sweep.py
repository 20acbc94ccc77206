"""Throughput of the other BASELINE.json configs (parity-test cases, not bench lines): C4 intra-modal fusion with
T = 1024 windows, C5 Two_transformers eval sweep over the batch size followed by the CCC reduction.  Eager launches,
CUDA events, bf16.  Writes one JSON object per line."""
import json, sys, time
import torch
sys.path.insert(0, '.')
import jmt_b200
dev = torch.device('cuda')

def timed(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

out = []
torch.manual_seed(0)
# C4: Intra_modal_transformer_fusion, long windows (projection-GEMM bound)
for B in (32, 256):
    T = 1024
    m = jmt_b200.Intra_modal_transformer_fusion(512, 1, 512, 1, precision='bf16').to(dev).train()
    a = torch.randn(B, T, 512, device=dev).bfloat16(); b = torch.randn(B, T, 768, device=dev).bfloat16()
    w = torch.randn(B, T, 512, device=dev)
    def step():
        for p in m.parameters(): p.grad = None
        (m(a, b) * w).sum().backward()
    ms = timed(step)
    out.append({"config": "C4 intra-modal fwd+bwd", "B": B, "T": T, "ms": ms, "windows_per_s": B / ms * 1e3,
                "model_tflops": 3 * 11.56 * B / ms})
    del m, a, b, w
# C5: Two_transformers eval sweep + CCC reduction
for joint in ("TRANSFORMER", "NONE"):
    m = jmt_b200.Two_transformers(0.0, 0.0, 1, 1, joint, "FC", 512, precision='bf16').to(dev).eval()
    for B in (1, 4, 16, 64, 256, 1024, 4096):
        # NONE attends across the batch (L = S = B): beyond the fused kernel's key limit the no-grad forward runs the
        # fused kernel over key chunks + log-sum-exp merge (jmt_attn_merge), so B = 4096 needs no (T, B, B) score tensor (20 GB)
        T = 300
        aud = torch.randn(B, T, 512, device=dev).bfloat16(); vis = torch.randn(B, T, 512, device=dev).bfloat16()
        lab = torch.rand(B * T, device=dev) * 2 - 1
        def fwd():
            with torch.no_grad():
                v, a = m(aud, vis)
            return jmt_b200.losses.six_sums(v.reshape(1, -1), lab.reshape(1, -1))
        ms = timed(fwd, iters=5 if B >= 256 else 20)
        # the same forward + CCC sums replayed as one CUDA graph (small batches are launch-bound when launched eagerly)
        g = jmt_b200.GraphedStep(lambda *_: fwd(), [(aud, vis)], warmup=2)
        ms_g = timed(lambda: g.replay(0), iters=5 if B >= 256 else 20)
        out.append({"config": f"C5 Two_transformers({joint},FC) eval + CCC sums", "B": B, "T": T, "ms_eager": ms, "ms_graph": ms_g,
                    "windows_per_s_graph": B / ms_g * 1e3})
        del g
        del aud, vis, lab
    del m
for o in out:
    print(json.dumps(o), flush=True)
