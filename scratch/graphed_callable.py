import sys, time, torch
sys.path.insert(0, '.')
import jmt_b200
dev = torch.device('cuda')
torch.manual_seed(0)
B, T = 16, 300
m = jmt_b200.Two_transformers(0.0, 0.0, 1, 1, "TRANSFORMER", "FC", 512, precision="bf16").to(dev).train()
aud = torch.randn(B, T, 512, device=dev, requires_grad=False)
vis = torch.randn(B, T, 512, device=dev, requires_grad=False)
crit = jmt_b200.CCCLoss(1)
lab = torch.rand(T, B, device=dev) * 2 - 1
opt = torch.optim.SGD(m.live_parameters(), lr=1e-3)
def step(mod):
    v, a = mod(aud, vis)
    loss = crit(v.reshape(1, -1), lab.reshape(1, -1)) + crit(a.reshape(1, -1), lab.reshape(1, -1))
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): l = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, float(l)
print("eager ms/step", timeit(lambda: step(m)))
try:
    gm = torch.cuda.make_graphed_callables(m, (aud, vis), allow_unused_input=True)
    print("graphed-callable ms/step", timeit(lambda: step(gm)))
except Exception as e:
    import traceback; traceback.print_exc()
    print("make_graphed_callables failed:", type(e).__name__, str(e)[:300])
