"""Per-tensor error of the fp32 CUDA mode against the float64 reference gradients (golden ctc_full_b4)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import llm_bci_b200 as lb
from oracle import ndt1_oracle as O
g = dict(np.load(os.path.join(ROOT, "tests/golden/ctc_full_b4.npz")))
tr = lb.default_trainer_config()
cfg = lb.update_config(tr.model, {"encoder": {"embedder": {"dropout": 0.0}, "transformer": {"dropout": 0.0}, "smooth_and_noise": {"noise": False}}})
for prec in sys.argv[1:] or ["fp32"]:
    torch.manual_seed(1)
    model = lb.NDT1(cfg, **tr.method.model_kwargs, precision=prec).to("cuda").train()
    batch = {k: v.to("cuda") for k, v in O.synthetic_ctc_batch(B=4, T=1000, N=256, seed=1).items()}
    out = model(**batch); out.loss.backward()
    print(prec, "loss", float(out.loss), "ref64", float(g["out64/loss"]))
    names = list(g["names"])
    got = {n: p.grad.detach().cpu().double().numpy() for n, p in model.named_parameters()}
    gn = np.array([np.linalg.norm(got[n]) for n in names])
    rel = np.abs(gn - g["grad_norm64"]) / g["grad_norm64"]
    relref = np.abs(g["grad_norm"] - g["grad_norm64"]) / g["grad_norm64"]
    for n, a, b in zip(names, rel, relref):
        if a > 3e-5 or b > 3e-5:
            print(f"  norm err ours {a:.2e}  ref-fp32 {b:.2e}  {n}")
    for k in [k for k in g if k.startswith("grad64/")]:
        n = k[len("grad64/"):]
        r64 = g[k].astype(np.float64); r32 = g["grad/" + n].astype(np.float64)
        print(f"  relL2 ours {np.linalg.norm(got[n]-r64)/np.linalg.norm(r64):.2e}  ref-fp32 {np.linalg.norm(r32-r64)/np.linalg.norm(r64):.2e}  {n}")
