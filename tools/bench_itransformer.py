#!/usr/bin/env python
"""SURVEY.md 8 f4 / BASELINE.json configs[3]: one training step (masker + forward + backward + AdamW) of the iTransformer SSL
configuration -- 16 trials x 100 bins x 669 neurons, 768 hidden, 8 heads of 96, 5 post-LN layers, dropout on -- on one B200
through this package, beside the UNMODIFIED reference (baseline/_ref) under torch-CUDA eager on the same GPU and on the host
CPU.  Prints one JSON line (kept under profiles/ per round); not the driver's bench (that is bench.py on configs[1]).

    python tools/bench_itransformer.py [--steps 20] [--warmup 5] [--no-reference]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def batch_of(torch, B, T, N, seed, dev):
    g = torch.Generator().manual_seed(seed)
    sp = torch.poisson(torch.full((B, T, N), 0.1), generator=g)
    return dict(spikes=sp.to(dev), spikes_mask=torch.ones(B, T, dtype=torch.int64, device=dev),
                spikes_timestamp=torch.arange(T)[None].expand(B, T).contiguous().to(dev))


def time_steps(torch, step, steps, warmup, cuda=True):
    for _ in range(warmup):
        step()
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) * 1e3 / steps


def reference_model(torch):
    """The unmodified reference module (baseline/_ref, verified against the manifest) with the two masker keys its yaml lacks."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import install_reference as ir
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not (os.path.isdir(ref) and ir.verify(ref)):
        return None
    cwd = os.getcwd()
    os.chdir(ref)
    sys.path.insert(0, ref)
    try:
        from utils.config_utils import update_config
        from models.itransformer import iTransformer as RefModel
        mk = {"active": True, "regions": None}
        cfg = update_config("configs/itransformer.yaml", {"masker": {"main": mk}, "encoder": {"embed_region": False}})
        torch.manual_seed(1)
        return RefModel(cfg, method_name="mlm", loss="poisson_nll", log_input=True)
    finally:
        os.chdir(cwd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--profile", action="store_true", help="also print the per-kernel launch profile of one step")
    args = ap.parse_args()
    import torch
    import llm_bci_b200 as lb
    from llm_bci_b200 import _C
    dev = "cuda"
    B, T, N = 16, 100, 669
    over = {"encoder": {"embed_region": False}}
    torch.manual_seed(1)
    model = lb.iTransformer(over, precision="bf16", method_name="mlm", loss="poisson_nll", log_input=True).to(dev).train()
    params = [p for p in model.parameters()]
    state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in params]
    batch = batch_of(torch, B, T, N, 1, dev)
    L = _C.lib()
    it = {"n": 0, "loss": None}

    def step():
        it["n"] += 1
        for p in params:
            p.grad = None
        out = model(**batch)
        (out.loss / out.n_examples).backward()
        for p, (m, v) in zip(params, state):                     # AdamW on this library's kernel (trainer.py:229 of the reference)
            _C.check(L.ndt1_adamw_step(p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-4, 0.9, 0.999, 1e-8, 0.01,
                                       it["n"], 1.0, _C.stream_ptr()), "ndt1_adamw_step")
        it["loss"] = out.loss

    l0 = L.ndt1_launch_counter()
    ms = time_steps(torch, step, args.steps, args.warmup)
    launches = (L.ndt1_launch_counter() - l0) / (args.steps + args.warmup)
    line = {"metric": "itransformer_ssl_train_trials_per_sec", "value": B / (ms * 1e-3), "unit": "trials/s", "n_gpus": 1, "ms_per_step": ms,
            "steps": args.steps, "warmup": args.warmup, "dtype": "bf16 GEMMs, fp32 LayerNorm / attention / loss", "data": "synthetic",
            "config": {"workload": "iTransformer SSL train step (masker + fwd + bwd + AdamW), BASELINE configs[3]: 16 trials x 100 bins x 669 neurons, "
                                   "768 hidden, 8 heads of 96, 5 post-LN layers, dropout 0.2 / 0.4 on", "global_batch": B},
            "gpu_launches_per_step": launches, "loss": float(it["loss"])}
    if args.profile:
        _C.profile_begin()
        step()
        torch.cuda.synchronize()
        prof = sorted(_C.profile_end(), key=lambda e: -e["ms"])
        line["kernels_alone_us"] = [{"kernel": e["name"][:70], "launches": e["launches"], "us": round(e["ms"] * 1e3, 1)} for e in prof[:12]]
        line["kernels_alone_ms_sum"] = sum(e["ms"] for e in prof)
    if not args.no_reference:
        ref = reference_model(torch)
        if ref is None:
            line["reference"] = {"unavailable": "baseline/_ref absent or modified"}
        else:
            rb = batch_of(torch, B, T, N, 1, "cpu")
            opt = torch.optim.AdamW(ref.parameters(), lr=1e-4, weight_decay=0.01)

            def ref_step(model=ref, b=rb, opt=opt):
                opt.zero_grad()
                out = model(**{k: v.clone() for k, v in b.items()})
                (out.loss / out.n_examples).backward()
                opt.step()

            ref.train()
            cpu_ms = time_steps(torch, ref_step, 3, 1, cuda=False)
            ref = ref.to(dev)
            gb = batch_of(torch, B, T, N, 1, dev)
            opt = torch.optim.AdamW(ref.parameters(), lr=1e-4, weight_decay=0.01)
            g32 = time_steps(torch, lambda: ref_step(ref, gb, opt), 10, 3)

            def ref_step_ac():
                opt.zero_grad()
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    out = ref(**{k: v.clone() for k, v in gb.items()})
                (out.loss.float() / out.n_examples).backward()
                opt.step()

            g16 = time_steps(torch, ref_step_ac, 10, 3)
            line["reference"] = {"cpu": {"value": B / (cpu_ms * 1e-3), "unit": "trials/s", "ms_per_step": cpu_ms, "cores": torch.get_num_threads()},
                                 "gpu_eager_fp32": {"value": B / (g32 * 1e-3), "ms_per_step": g32},
                                 "gpu_eager_bf16_autocast": {"value": B / (g16 * 1e-3), "ms_per_step": g16}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
