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


def ctc_error_counts(log_probs: torch.Tensor, targets: torch.Tensor, targets_lengths: torch.Tensor, blank_id: int = 0):
    """The CER metric of main.py:67-73 without leaving the device: greedy decode, then the edit distance of every trial to its
    target ids (word_error_count, utils/eval_bci.py:19-36, on ids instead of space-joined phoneme strings).  Returns
    (errors (B), words (B)) as int64 device tensors; CER = errors.sum() / words.sum().  An empty target counts as one word,
    as `"".split(" ")` does in the reference."""
    ids, lens = greedy_ctc_decode(log_probs, blank_id)
    tg = targets.detach().to(device=ids.device, dtype=torch.int64).contiguous()
    tl = targets_lengths.detach().to(device=ids.device, dtype=torch.int64).contiguous()
    B, L = ids.shape
    errors = torch.empty((B,), dtype=torch.int64, device=ids.device)
    _C.check(_C.lib().ndt1_edit_distance(ids.data_ptr(), lens.data_ptr(), L, tg.data_ptr(), tl.data_ptr(), int(tg.shape[1]), B,
                                         errors.data_ptr(), _C.stream_ptr()), "ndt1_edit_distance")
    return errors, tl.clamp(min=1)


def phoneme_error_rate(log_probs: torch.Tensor, targets: torch.Tensor, targets_lengths: torch.Tensor, blank_id: int = 0) -> torch.Tensor:
    """errors / n_phonemes over the batch (the `cer` metric function of main.py:67-73), a 0-dim device tensor."""
    errors, words = ctc_error_counts(log_probs, targets, targets_lengths, blank_id)
    return errors.sum().float() / words.sum().float()
