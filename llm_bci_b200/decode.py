"""Greedy CTC decoding: argmax + the reference's collapse rule.

``format_ctc`` is the host function of utils/eval_bci.py:41-48 (emit when the id differs
from the last EMITTED id and is not blank; ``last`` only moves on emission, so
``A, blank, A`` yields one ``A``).  ``greedy_ctc_decode`` does argmax (main.py:69) and the
collapse on the device for a whole batch, over all L rows like the reference."""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import _C


def format_ctc(pred: Sequence[int], vocab: Sequence, blank_id: int) -> List:
    phonemes, last = [], -1
    for idx in pred:
        idx = int(idx)
        if idx != last and idx != blank_id:
            phonemes.append(vocab[idx])
            last = idx
    return phonemes


def greedy_ctc_decode(log_probs: torch.Tensor, blank_id: int = 0):
    """log_probs (B, L, V) on the GPU -> (ids (B, L) padded with -1, lengths (B))."""
    if not log_probs.is_cuda:
        raise RuntimeError("greedy_ctc_decode runs on the GPU only (use format_ctc on the host)")
    lp = log_probs.detach().contiguous().float()
    B, L, V = lp.shape
    ids = torch.empty((B, L), dtype=torch.int64, device=lp.device)
    lens = torch.empty((B,), dtype=torch.int64, device=lp.device)
    _C.check(_C.lib().ndt1_ctc_greedy_decode(lp.data_ptr(), B, L, V, int(blank_id), ids.data_ptr(), lens.data_ptr(), _C.stream_ptr()),
             "ndt1_ctc_greedy_decode")
    return ids, lens
