import sys, json, torch, numpy as np
sys.path.insert(0, '.')
import jmt_b200
from oracle import jmt_oracle as O
dev = 'cuda'
for nlev in [1, 2, 3, 4]:
  for seed in [51, 7]:
    cin, chans, k, N, Ls = 1024, [512] * nlev, 5, 2, 40
    params = O.synth_params(O.tcn_shapes(cin, chans, k), seed)
    model = jmt_b200.TemporalConvNet(cin, chans, kernel_size=k, attention=0, dropout=0.1, precision='fp32')
    model.load_state_dict(O.tcn_state_dict(params), strict=True)
    model = model.to(dev).eval()
    gen = torch.Generator().manual_seed(52)
    x = torch.randn(N, cin, Ls, generator=gen)
    xd = x.to(dev).requires_grad_(True)
    out = model(xd)
    w = torch.linspace(-1, 1, out.numel()).reshape(out.shape)
    (out * w.to(dev)).sum().backward()
    po = {kk: v.double().requires_grad_(True) for kk, v in params.items()}
    xo = x.double().requires_grad_(True)
    oo = O.tcn_forward(xo, po, nlev)
    (oo * w.double()).sum().backward()
    e = (xd.grad.double().cpu() - xo.grad).abs()
    pert = e.amax(dim=(0, 1)) / xo.grad.abs().max()
    print('levels', nlev, 'seed', seed, 'fwd', float((out.detach().double().cpu() - oo.detach()).abs().max() / oo.abs().max()), 'dx', float(e.max() / xo.grad.abs().max()))
    print('   per t:', ' '.join(f'{v:.0e}' for v in pert.tolist()))
    # near-zero pre-activations in the oracle (flip candidates)
