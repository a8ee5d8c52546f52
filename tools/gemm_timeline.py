#!/usr/bin/env python
"""Per-CTA phase timeline of the tensor-core GEMM on the K = N = 1024 shapes of the benchmark step (debugging aid).
Runs y = x W^T (+ bias) through ndt1_linear_fwd in the bf16 mode and prints, per phase, the mean / min / max time since
the first CTA of the launch started, plus the launch-to-launch period of back-to-back launches."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_bci_b200 import _C

M, N, K = 7776, int(os.environ.get("GT_N", 1024)), int(os.environ.get("GT_K", 1024))
dev = "cuda"
Lb = _C.lib()
torch.manual_seed(0)
x, w, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev) / K ** 0.5, torch.randn(N, device=dev)
y = torch.empty(M, N, device=dev)
ws = torch.empty(4 * (M + N) * (K + 8) + 4096, dtype=torch.uint8, device=dev)

def run():
    _C.check(Lb.ndt1_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, 0, 1, ws.data_ptr(), ws.numel(), _C.stream_ptr()))

for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record(); torch.cuda.synchronize()
print(f"linear (2 casts + GEMM {M}x{N}x{K}): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")

buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
Lb.ndt1_debug_gemm_timeline(buf.data_ptr())
run()
torch.cuda.synchronize()
Lb.ndt1_debug_gemm_timeline(None)
t = buf.cpu().numpy().reshape(148, 16).astype(np.int64)
t = t[t[:, 0] > 0]
names = {0: "CTA start", 1: "barriers + TMEM ready", 2: "griddepcontrol.wait passed", 3: "first TMA issued", 5: "first operands landed (MMA starts)",
         4: "all loads issued", 6: "last MMA issued", 7: "first accumulator ready", 8: "last accumulator ready", 9: "epilogue warp 0/4 done",
         10: "epilogue warp 1/5 done", 11: "epilogue warp 2/6 done", 12: "epilogue warp 3/7 done", 14: "TMEM freed (exit)"}
t0 = t[:, 0].min()
print(f"{len(t)} CTAs; times in us since the first CTA started")
for k in (0, 1, 2, 3, 5, 7, 4, 6, 8, 9, 10, 11, 12, 14):
    v = (t[:, k][t[:, k] > 0] - t0) / 1e3
    if len(v):
        print(f"  {names[k]:36s} mean {v.mean():6.2f}  min {v.min():6.2f}  max {v.max():6.2f}  (n={len(v)})")
