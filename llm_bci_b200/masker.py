"""Masker as a device kernel (reference models/masker.py:28-110).

Same constructor config, same ``forward(spikes, brain_regions=None) ->
(spikes, mask int64)`` contract, same in-place mutation of ``spikes``.  The
random draws are made in the reference's order:

* ``rng="reference"`` (default): the three Bernoulli tensors come from the torch
  CPU generator and the uniform tensor from the device generator, exactly like
  the reference, so a seeded reference run and a seeded run of this module
  produce bit-identical spikes and masks;
* ``rng="device"``: the tensors are drawn on the GPU from the library's own
  Philox streams (no host draws, no H2D copies); only the two scalar draws of
  the temporal mode stay on the host.

Masking itself (mode broadcast, temporal dilation, zeroing, the global max
taken after zeroing and the random replacement) is the kernel pair behind
``ndt1_masker_apply``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _C
from .config import DictConfig


class Masker(nn.Module):

    def __init__(self, config: DictConfig, rng: str = "reference"):
        super().__init__()
        self.force_active = config.force_active if "force_active" in config else False
        self.active = config.active
        self.mode = config.mode
        self.ratio = config.ratio
        self.zero_ratio = config.zero_ratio
        self.random_ratio = config.random_ratio
        self.expand_prob = config.expand_prob
        self.max_timespan = config.max_timespan
        self.regions = config.regions
        self.channels = config.channels
        self.rng = config.rng if "rng" in config else rng
        self._step = 0

    def is_active(self) -> bool:
        return bool(self.active) and (self.training or bool(self.force_active))

    def _mask_probs(self, spikes: torch.Tensor, brain_regions) -> Tuple[torch.Tensor, int]:
        """Probabilities of the first Bernoulli draw and the temporal dilation, masker.py:53-80."""
        B, T, N = spikes.shape
        timespan = 1
        if self.mode == "temporal":
            if torch.bernoulli(torch.tensor(self.expand_prob).float()):
                timespan = int(torch.randint(1, self.max_timespan + 1, (1,)).item())
            probs = torch.full((B, T), self.ratio / timespan)
        elif self.mode == "neuron":
            probs = torch.full((B, N), self.ratio)
        elif self.mode == "random":
            probs = torch.full((B, T, N), self.ratio)
        elif self.mode == "region":
            assert brain_regions is not None, "Can't mask region without brain region information"
            assert self.regions is not None, "No regions to mask"
            probs = torch.zeros(B, N)
            for region in self.regions:
                probs[torch.from_numpy(np.asarray(brain_regions == region))] = 1
        elif self.mode == "co-smooth":
            assert self.channels is not None, "No channels to mask"
            probs = torch.zeros(N)
            for c in self.channels:
                probs[c] = 1
        else:
            raise Exception(f"Masking mode {self.mode} not implemented")
        return probs, timespan

    def forward(self, spikes: torch.FloatTensor, brain_regions: Optional[np.ndarray] = None,
                targets_mask: Optional[torch.Tensor] = None, draws: Optional[dict] = None) -> Tuple[torch.FloatTensor, torch.LongTensor]:
        """``draws`` (test hook) injects {"mask","zero","random","rand","timespan"} instead of drawing."""
        if not self.is_active():
            return spikes, torch.zeros_like(spikes, dtype=torch.int64)
        if not spikes.is_cuda:
            raise RuntimeError("llm_bci_b200.Masker runs on the GPU only (no CPU fallback)")
        assert spikes.dtype == torch.float32 and spikes.is_contiguous(), "spikes must be contiguous float32"
        B, T, N = spikes.shape
        dev = spikes.device
        L = _C.lib()
        st = _C.stream_ptr()
        if draws is not None:
            u8 = lambda a: torch.as_tensor(np.asarray(a)).to(torch.uint8).to(dev).contiguous()
            mask_draw, zero_draw, random_draw = u8(draws["mask"]), u8(draws["zero"]), u8(draws["random"])
            rand = torch.as_tensor(np.asarray(draws["rand"])).float().to(dev).contiguous()
            timespan = int(draws.get("timespan", 1))
        else:
            probs, timespan = self._mask_probs(spikes, brain_regions)
        if draws is not None:
            pass
        elif self.rng == "reference":
            mask_draw = torch.bernoulli(probs).to(torch.uint8).to(dev, non_blocking=True)
            zero_draw = torch.bernoulli(torch.full((B, T, N), self.zero_ratio)).to(torch.uint8).to(dev, non_blocking=True)
            random_draw = torch.bernoulli(torch.full((B, T, N), self.random_ratio)).to(torch.uint8).to(dev, non_blocking=True)
            rand = torch.rand((B, T, N), device=dev)
        else:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            if self.mode in ("region", "co-smooth"):   # deterministic 0/1 table
                mask_draw = (probs > 0).to(torch.uint8).to(dev)
            else:
                mask_draw = torch.empty(probs.shape, dtype=torch.uint8, device=dev)
                _C.check(L.ndt1_bernoulli_u8(mask_draw.data_ptr(), mask_draw.numel(), float(probs.flatten()[0]), seed, 1, st))
            zero_draw = torch.empty((B, T, N), dtype=torch.uint8, device=dev)
            random_draw = torch.empty((B, T, N), dtype=torch.uint8, device=dev)
            rand = torch.empty((B, T, N), dtype=torch.float32, device=dev)
            _C.check(L.ndt1_bernoulli_u8(zero_draw.data_ptr(), zero_draw.numel(), float(self.zero_ratio), seed, 2, st))
            _C.check(L.ndt1_bernoulli_u8(random_draw.data_ptr(), random_draw.numel(), float(self.random_ratio), seed, 3, st))
            _C.check(L.ndt1_uniform_f32(rand.data_ptr(), rand.numel(), seed, 4, st))
        mask = torch.empty((B, T, N), dtype=torch.int64, device=dev)
        scratch = torch.empty(1, dtype=torch.int32, device=dev)
        _C.check(L.ndt1_masker_apply(spikes.data_ptr(), B, T, N, _C.MASK_MODE[self.mode], timespan, mask_draw.data_ptr(),
                                     zero_draw.data_ptr(), random_draw.data_ptr(), rand.data_ptr(), mask.data_ptr(),
                                     _C.ptr(targets_mask), scratch.data_ptr(), st), "ndt1_masker_apply")
        return spikes, mask

    @staticmethod
    def expand_timesteps(mask: torch.Tensor, width: int = 1) -> torch.Tensor:
        """masker.py:106-110 as index arithmetic: out[t] = OR_k m[t - (width-1)//2 + k]."""
        T = mask.shape[1]
        left = (width - 1) // 2
        out = torch.zeros_like(mask, dtype=torch.bool)
        m = mask.bool()
        for k in range(width):
            off = k - left
            lo, hi = max(0, -off), min(T, T - off)
            if lo < hi:
                out[:, lo:hi] |= m[:, lo + off:hi + off]
        return out
