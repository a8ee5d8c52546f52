#!/usr/bin/env python
"""Stand-alone timing of log-softmax + CTC (benchmark shape: 32 trials x 243 frames x 41 phonemes, 20-60 labels) with and
without the gradient pass (debugging aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_bci_b200 import _C

B, L, V, S = 32, 243, 41, 60
dev = "cuda"
Lb = _C.lib()
torch.manual_seed(0)
lg = torch.randn(B, L, V, device=dev)
tl = torch.randint(20, 61, (B,), device=dev)
tg = torch.randint(1, V, (B, S), device=dev) * (torch.arange(S, device=dev)[None] < tl[:, None])
il = torch.randint(150, L + 1, (B,), device=dev); il[0] = L
logp, nll, loss, dl = torch.empty_like(lg), torch.empty(B, device=dev), torch.zeros((), device=dev), torch.empty_like(lg)
ws = torch.empty(Lb.ndt1_ctc_workspace_bytes(B, L, S), dtype=torch.uint8, device=dev)

def run(grad):
    _C.check(Lb.ndt1_ctc_loss(lg.data_ptr(), logp.data_ptr(), tg.data_ptr(), il.data_ptr(), tl.data_ptr(), B, L, V, S, 0, 1, ws.data_ptr(),
                              nll.data_ptr(), loss.data_ptr(), dl.data_ptr() if grad else None, None, _C.stream_ptr()))

for grad in (False, True):
    for _ in range(3):
        run(grad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        run(grad)
    e1.record(); torch.cuda.synchronize()
    print(f"log-softmax + ctc, gradient {'on' if grad else 'off'}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
