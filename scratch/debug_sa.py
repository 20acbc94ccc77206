import sys, json, os, numpy as np, torch
sys.path.insert(0, '.')
import jmt_b200
from oracle import jmt_oracle as O
meta = json.load(open('tests/golden/golden_meta.json'))
m = meta['tt_transformer_sa_h2_l1']
params = O.synth_params(O.two_transformers_shapes(m["layers"], m["joint"], m["fmt"], m["vin"]), m["param_seed"])
aud, vis = O.synth_features(m["B"], m["T"], [512, m["vin"]], m["feat_seed"])
lv, la = O.synth_labels(m["B"], m["T"], m["label_seed"])
res = {}
for prec in ("fp32", "bf16"):
    model = jmt_b200.Two_transformers(0.0, 0.0, m["heads"], m["layers"], m["joint"], m["fmt"], m["vin"], precision=prec)
    sd = model.state_dict(); model.load_state_dict({k: params[k] for k in sd}); model = model.cuda().eval()
    a = aud.cuda().requires_grad_(True); v = vis.cuda().requires_grad_(True)
    vo, ao = model(a, v)
    crit = jmt_b200.CCCLoss(digitize_num=1); n = vo.numel()
    loss = crit(vo.view(-1, n), lv.cuda().view(-1, n)) + crit(ao.view(-1, n), la.cuda().view(-1, n))
    loss.backward()
    res[prec] = dict(vo=vo.detach().cpu(), da=a.grad.cpu(), dv=v.grad.cpu(), loss=loss.item(),
                     g={k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None})
def rl2(x, y): return float((x - y).norm() / (y.norm() + 1e-30))
print("loss", res["fp32"]["loss"], res["bf16"]["loss"])
print("vo", rl2(res["bf16"]["vo"], res["fp32"]["vo"]), "da", rl2(res["bf16"]["da"], res["fp32"]["da"]), "dv", rl2(res["bf16"]["dv"], res["fp32"]["dv"]))
rows = sorted(((rl2(res["bf16"]["g"][k], res["fp32"]["g"][k]), k) for k in res["fp32"]["g"]), reverse=True)
for r, k in rows[:50]: print(f"{r:.3f} {k}")
