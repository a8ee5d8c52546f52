#!/usr/bin/env python
"""Stand-alone timing of the tensor-core attention kernels (benchmark shape) and, for the pipelined key-side backward,
a per-CTA phase timeline (mean duration of every phase, and how back-to-back the CTAs of one SM run)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_bci_b200 import _C

B, L, H, NH = 32, 243, 1024, 8
dev = "cuda"
Lb = _C.lib()
torch.manual_seed(0)
qkv = (torch.randn(B, L, 3 * H, device=dev) * 0.5).bfloat16()
dout = (torch.randn(B, L, H, device=dev) * 0.1).bfloat16()
out, outd, dqkv = torch.empty(B, L, H, device=dev, dtype=torch.bfloat16), torch.empty(B, L, H, device=dev, dtype=torch.bfloat16), torch.empty_like(qkv)
lse = torch.empty(B, NH, L, device=dev)
kv = torch.ones(B, L, dtype=torch.int64, device=dev)
ws = torch.empty(Lb.ndt1_attention_workspace_bytes(B, L, NH), dtype=torch.uint8, device=dev)

def run(bwd=True):
    _C.check(Lb.ndt1_attention_bf16(qkv.data_ptr(), out.data_ptr(), outd.data_ptr(), lse.data_ptr(), kv.data_ptr(), B, L, H, NH, -2, -2, 0.4, 0.4, 7, 1, 2,
                                    dout.data_ptr() if bwd else None, dqkv.data_ptr() if bwd else None, ws.data_ptr(), 1, _C.stream_ptr()))

for _ in range(3):
    run()
torch.cuda.synchronize()
for name, bwd in (("fwd", False), ("fwd+bwd", True)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run(bwd)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per layer")

ncta = 2 * NH * B
buf = torch.zeros(ncta * 32, dtype=torch.int64, device=dev)
Lb.ndt1_debug_attention_timeline(buf.data_ptr())
run()
torch.cuda.synchronize()
Lb.ndt1_debug_attention_timeline(None)
t = buf.cpu().numpy().reshape(ncta, 32).astype(np.int64)
names = {0: "start", 1: "tmem alloc'd", 2: "setup sync done", 3: "K,V landed", 4: "S(0) issued", 5: "S(0) ready", 6: "S(1) ready", 7: "S(2) ready",
         8: "S(3) ready", 9: "last P written", 10: "accumulators done", 11: "stores issued", 12: "dV/dK(0) issued", 13: "dV/dK(1) issued",
         14: "dV/dK(2) issued", 15: "dV/dK(3) issued", 16: "exit", 17: "S(1) issued", 18: "S(2) issued", 19: "S(3) issued", 20: "(no S(4))",
         21: "P(0) seen by MMA", 22: "P(1) seen by MMA", 23: "P(2) seen by MMA", 24: "P(3) seen by MMA"}
base = t[:, 0:1]
rel = (t - base) / 1e3
print("key-side backward, mean time since CTA start (us):")
for k in sorted(names):
    v = rel[:, k][t[:, k] > 0]
    if len(v):
        print(f"  {names[k]:22s} {v.mean():7.2f}  (min {v.min():6.2f} max {v.max():6.2f})")
sm = t[:, 31]
gaps = []
for s_ in np.unique(sm):
    idx = np.where(sm == s_)[0]
    order = idx[np.argsort(t[idx, 0])]
    for a_, b_ in zip(order[:-1], order[1:]):
        gaps.append((t[b_, 0] - t[a_, 16]) / 1e3)
print(f"gap between a CTA's exit and the next CTA's start on the same SM: mean {np.mean(gaps):.2f} us, max {np.max(gaps):.2f} us; "
      f"kernel span {(t[:, 16].max() - t[:, 0].min()) / 1e3:.1f} us, CTA mean {rel[:, 16].mean():.2f} us")
