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
buf = torch.zeros(65536 + 148 * 64, dtype=torch.int64, device=dev)
Lb.ndt1_debug_attention_timeline(buf.data_ptr())
run(True)
torch.cuda.synchronize()
Lb.ndt1_debug_attention_timeline(None)
raw = buf.cpu().numpy().astype(np.int64)

# ---- forward kernel: 32 slots per CTA
t = raw[:ncta * 32].reshape(ncta, 32)
names = {0: "start", 1: "set-up done", 2: "Q, K landed", 3: "S ready", 4: "row maxima done", 5: "P written", 6: "V landed", 7: "O ready",
         8: "output staged", 9: "stores read smem", 16: "exit"}
rel = (t - t[:, 0:1]) / 1e3
print("forward, mean time since CTA start (us):")
for k in sorted(names):
    v = rel[:, k][t[:, k] > 0]
    if len(v):
        print(f"  {names[k]:22s} {v.mean():7.2f}  (min {v.min():6.2f} max {v.max():6.2f})")
sm = t[:, 31]
print(f"  kernel span {(t[:, 16].max() - t[:, 0].min()) / 1e3:.1f} us, CTA mean {rel[:, 16].mean():.2f} us, CTAs per SM {ncta / len(np.unique(sm)):.2f}")
starts = np.sort(t[:, 0] - t[:, 0].min()) / 1e3
print("  CTA start times (us), deciles:", " ".join(f"{starts[int(q * (ncta - 1))]:.1f}" for q in np.linspace(0, 1, 11)))

# ---- persistent key-side backward: 64 slots per CTA, 12 per item for the first four items
k3 = raw[65536:65536 + 148 * 64].reshape(148, 64)
live = k3[:, 0] > 0
k3 = k3[live]
t0 = k3[:, 0:1]
inames = {0: "item start", 11: "K/V requested", 2: "S(0) seen", 3: "S(1) seen", 4: "S(2) seen", 5: "S(3) seen", 1: "last P written",
          10: "last products queued", 6: "accumulators done", 7: "staging written", 8: "stores read smem"}
print(f"key-side backward (persistent), {live.sum()} CTAs, time since the CTA's first item started (us):")
for it in range(4):
    print(f" item {it}:")
    for k in (0, 11, 2, 3, 4, 5, 1, 10, 6, 7, 8):
        col = k3[:, it * 12 + k]
        v = ((col - t0[:, 0]) / 1e3)[col > 0]
        if len(v):
            print(f"  {inames[k]:22s} {v.mean():7.2f}  (min {v.min():6.2f} max {v.max():6.2f})")
ex = k3[:, 63]
print(f"  exit {((ex - t0[:, 0]) / 1e3)[ex > 0].mean():.2f} us; kernel span {(ex.max() - k3[:, 0].min()) / 1e3:.1f} us")

