import sys, json, os, numpy as np, torch
sys.path.insert(0, '.')
import jmt_b200
from oracle import jmt_oracle as O
meta = json.load(open('tests/golden/golden_meta.json'))
def rl2(x, y): return float((x - y).norm() / (y.norm() + 1e-30))
for name in ['tt_transformer_sa_h2_l1', 'tt_transformer_fc_h4_l2', 'tt_transformer_fc_h1_l1']:
    m = meta[name]
    for scale in (1.0, 0.5, 0.25):
        params = O.synth_params(O.two_transformers_shapes(m["layers"], m["joint"], m["fmt"], m["vin"]), m["param_seed"])
        params = {k: (v * scale if (k.endswith('weight') and 'layer_norm' not in k) else v) for k, v in params.items()}
        aud, vis = O.synth_features(m["B"], m["T"], [512, m["vin"]], m["feat_seed"])
        res = {}
        torch.manual_seed(0)
        cot = None
        for prec in ("fp32", "bf16"):
            model = jmt_b200.Two_transformers(0.0, 0.0, m["heads"], m["layers"], m["joint"], m["fmt"], m["vin"], precision=prec)
            sd = model.state_dict(); model.load_state_dict({k: params[k] for k in sd}); model = model.cuda().eval()
            a = aud.cuda().requires_grad_(True); v = vis.cuda().requires_grad_(True)
            vo, ao = model(a, v)
            if cot is None: cot = (torch.randn_like(vo), torch.randn_like(ao))
            torch.autograd.backward([vo, ao], list(cot))
            res[prec] = dict(vo=vo.detach().cpu(), da=a.grad.cpu(), dv=v.grad.cpu())
        print(name, "wscale", scale, "vo", round(rl2(res["bf16"]["vo"], res["fp32"]["vo"]), 4), "da", round(rl2(res["bf16"]["da"], res["fp32"]["da"]), 4), "dv", round(rl2(res["bf16"]["dv"], res["fp32"]["dv"]), 4), flush=True)
