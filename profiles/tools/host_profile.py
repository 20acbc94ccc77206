"""cProfile of the Python side of eagerly launched C2 training steps (what `host_enqueue_ms_per_eager_step` is made of).
   python profiles/tools/host_profile.py [B]"""
import cProfile, pstats, sys, time, io
import torch
sys.path.insert(0, '.')
import jmt_b200
dev = torch.device('cuda')
B, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 32), 300
torch.manual_seed(0)
fusion = jmt_b200.Two_transformers(0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512, precision='bf16')
fc = jmt_b200.FcLayer(768, 512, precision='bf16')
tcn = jmt_b200.TemporalConvNet(1024, [512] * 4, kernel_size=5, attention=0, dropout=0.1, precision='bf16')
pipe = jmt_b200.JMTPipeline(fusion, fc, tcn).to(dev).train()
opt = torch.optim.SGD(pipe.live_parameters(), lr=1e-3)
crit = jmt_b200.CCCLoss(digitize_num=1)
vis = torch.randn(B, 1024, T, device=dev).bfloat16(); aud = torch.randn(B, T, 768, device=dev).bfloat16()
lv = torch.rand(B, T, device=dev) * 2 - 1; la = torch.rand(B, T, device=dev) * 2 - 1
n = B * T


def step():
    v, a = pipe(aud, vis)
    loss = crit.forward_va(v.view(-1, n), lv.view(-1, n), a.view(-1, n), la.view(-1, n))
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue per eager step: {(t1 - t0) / 5 * 1e3:.2f} ms (B={B})")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28)
print(s.getvalue()[:6000])
