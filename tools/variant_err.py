"""Per-tensor gradient error of the ctc_variants cases (debugging aid): python tools/variant_err.py [name] [precision]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import llm_bci_b200 as lb
from test_oracle_golden import load, sub, CTC_KW, VARIANTS, variant_case
name = sys.argv[1] if len(sys.argv) > 1 else "gelu_factors"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
g = load("ctc_variants.npz")
cfg, params, batch = variant_case(g, name)
model = lb.NDT1(cfg, **CTC_KW, precision=prec)
model.load_state_dict({k: v.clone() for k, v in params.items()})
model = model.cuda().train()
out = model(**{k: v.cuda() for k, v in batch.items()})
out.loss.backward()
print("loss", float(out.loss), float(g[f"{name}/out/loss"]))
ref = sub(g, f"{name}/grad")
nscale = max(float(np.linalg.norm(v.astype(np.float64))) for v in ref.values())
for n, p in model.named_parameters():
    r = ref[n].astype(np.float64); q = p.grad.detach().cpu().double().numpy()
    print(f"{n:55s} rel-L2 {np.linalg.norm(q - r) / max(np.linalg.norm(r), 1e-3 * nscale):.3e}  |ref| {np.linalg.norm(r):.3e}")
