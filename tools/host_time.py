#!/usr/bin/env python
"""Host (enqueue) time of one DataParallelTrainer.train_step on the benchmark shape: 5 steps are enqueued behind a long-running
dummy kernel queue so that the measured wall time is pure CPU work (debugging aid)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llm_bci_b200 as lb
from oracle import ndt1_oracle as O

dev = "cuda"
tr = lb.default_trainer_config()
torch.manual_seed(1)
model = lb.NDT1(tr.model, **tr.method.model_kwargs, precision="bf16").to(dev)
trainer = lb.DataParallelTrainer(model, lr=1e-3)
batch = {k: v.to(dev) for k, v in O.synthetic_ctc_batch(B=32, T=1000, N=256, seed=1).items()}
for _ in range(5):
    trainer.train_step(batch)
torch.cuda.synchronize()
for n in (1, 3, 5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        trainer.train_step(batch)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{n} steps: host enqueue {1e3 * (t1 - t0) / n:.2f} ms/step, until the GPU is done {1e3 * (t2 - t0) / n:.2f} ms/step")
