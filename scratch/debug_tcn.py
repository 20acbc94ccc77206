import sys, torch, numpy as np
sys.path.insert(0, '.')
import jmt_b200
from jmt_b200 import engine as E, _lib as L
from oracle import jmt_oracle as O
dev = 'cuda'
torch.manual_seed(0)
N, Ls, cin, cout, k = 2, 40, 64, 48, 5
for precision in ['fp32', 'bf16']:
  for dil in [1, 2, 4, 8]:
    g = torch.rand(cout, 1, 1) + 0.5; v = torch.randn(cout, cin, k) * 0.2; b = torch.randn(cout) * 0.1
    x = torch.randn(N, cin, Ls)
    if precision == 'bf16': x = x.bfloat16().float()
    params = {'c.weight_g': g.to(dev), 'c.weight_v': v.to(dev), 'c.bias': b.to(dev)}
    ctx = E.Ctx(params, precision, True, False)
    ctx.prepare_param_grads(list(params))
    xv, gx = E.transpose_in(ctx, x.to(dev), True)
    y = E.causal_conv(ctx, xv, 'c.', N, Ls, cin, cout, k, dil, L.ACT_LEAKY)
    out, setter = E.transpose_out(ctx, y, N, Ls, cout)
    wgt = torch.linspace(-1, 1, out.numel()).reshape(out.shape)
    setter(wgt.to(dev))
    ctx.backward()
    torch.cuda.synchronize()
    xo = x.double().requires_grad_(True); go = g.double().requires_grad_(True); vo = v.double().requires_grad_(True); bo = b.double().requires_grad_(True)
    w = O.weight_norm_weight(go, vo)
    yo = O.leaky_relu(O.causal_dilated_conv1d(xo, w, bo, dil))
    (yo * wgt.double()).sum().backward()
    rel = lambda a, r: float((a.double().cpu() - r).abs().max() / r.abs().max())
    dx = gx()
    pert = (dx.double().cpu() - xo.grad).abs().amax(dim=(0, 1)) / xo.grad.abs().max()
    print(precision, 'dil', dil, 'fwd', rel(out, yo.detach()), 'dx', rel(dx, xo.grad), 'dv', rel(ctx.pgrads['c.weight_v'], vo.grad),
          'dg', rel(ctx.pgrads['c.weight_g'], go.grad), 'db', rel(ctx.pgrads['c.bias'], bo.grad))
    print('   dx err per t:', ' '.join(f'{e:.0e}' for e in pert.tolist()))
