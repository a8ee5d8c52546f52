"""Padded-batch collate with the reference's ``pad_dict`` semantics
(data_utils/datasets.py:191-272), on the host or -- the product path -- on the device.

``pad_collate_fn(batch, model_inputs, pad_dict)`` keeps the reference signature and
return value ``(padded_batch, unused_inputs)`` and runs on the host (numpy), for
callers that want CPU tensors.  ``DevicePadCollate`` is the B200 path: every padded
key is shipped as ONE ragged buffer (sum of lengths, not B x max length) plus an
offsets vector and scattered into its padded tensor by ``ndt1_pad_pack``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _C


def _pad_geometry(lengths: List[int], truncate: Optional[int], min_length: Optional[int]) -> Tuple[int, int]:
    """(length after padding, length after truncation), datasets.py:199-206."""
    longest = max(lengths)
    truncate = longest if truncate is None else truncate
    min_length = 0 if min_length is None else min_length
    assert min_length <= truncate, "Can't truncate below the minimum length"
    full = max(longest, min_length)
    return full, min(truncate, full)


def padded_array(arrays: List[np.ndarray], dim: int = 0, side: str = "right", value=0, truncate: Optional[int] = None,
                 min_length: Optional[int] = None) -> np.ndarray:
    """Stack ``arrays`` after padding axis ``dim`` with ``value`` on ``side`` to
    min(truncate, max(longest, min_length)); the FIRST entries survive truncation."""
    if side not in ("left", "right"):
        raise Exception(f' "side" can only take values "right" or "left", got {side}')
    full, keep = _pad_geometry([a.shape[dim] for a in arrays], truncate, min_length)
    out = []
    for a in arrays:
        # pad to `keep` (the reference pads to pad_size = min(truncate, full)), then keep the first `truncate` entries
        pad_shape = list(a.shape)
        pad_shape[dim] = max(0, keep - a.shape[dim])
        filler = np.full(pad_shape, value, dtype=a.dtype)
        joined = np.concatenate((filler, a) if side == "left" else (a, filler), axis=dim)
        limit = joined.shape[dim] if truncate is None else min(truncate, joined.shape[dim])
        out.append(np.take(joined, np.arange(limit), axis=dim))
    return np.stack(out, axis=0)


def _classify(batch):
    keys = list(batch[0].keys())
    arrays = [k for k in keys if isinstance(batch[0][k], np.ndarray) and batch[0][k].dtype.type != np.str_]
    strings = [k for k in keys if isinstance(batch[0][k], np.ndarray) and batch[0][k].dtype.type == np.str_]
    return keys, arrays, strings


def pad_collate_fn(batch, model_inputs: List[str], pad_dict: Dict[str, Dict[str, Any]]):
    """Host collate: (padded_batch[model inputs], unused_inputs).  datasets.py:236-272."""
    if isinstance(batch[0], list):      # the dataset already batched: flatten
        batch = [row for sub in batch for row in sub]
    keys, arrays, strings = _classify(batch)
    assert set(pad_dict.keys()).issubset(arrays), f"Can't pad keys which are not arrays: {set(pad_dict.keys()) - set(arrays)} "
    padded_batch, unused = {}, {}
    for key in keys:
        col = [row[key] for row in batch]
        if key in arrays:
            if key in pad_dict:
                value = torch.from_numpy(padded_array(col, **pad_dict[key])).clone()
            elif len({a.shape for a in col}) == 1:
                value = torch.from_numpy(np.stack(col, axis=0))
            else:
                value = [torch.from_numpy(a) for a in col]
        elif key in strings:
            value = np.stack(col, axis=0)
        else:
            value = col
        (padded_batch if key in model_inputs else unused)[key] = value
    return padded_batch, unused


class DevicePadCollate:
    """collate_fn whose padded model inputs are built on the GPU.

    Usable as ``DataLoader(collate_fn=DevicePadCollate(model_inputs, pad_dict, device))`` with
    ``num_workers=0`` (the reference's setting, models/trainer.py:216-222).  Keys padded along an
    axis other than 0, or of a dtype other than float32/int64, take the host path and are copied."""

    def __init__(self, model_inputs: List[str], pad_dict: Dict[str, Dict[str, Any]], device="cuda"):
        self.model_inputs = list(model_inputs)
        self.pad_dict = pad_dict
        self.device = torch.device(device)

    def _pack(self, col: List[np.ndarray], spec: Dict[str, Any]) -> torch.Tensor:
        side, value = spec.get("side", "right"), spec.get("value", 0)
        if side not in ("left", "right"):
            raise Exception(f' "side" can only take values "right" or "left", got {side}')
        lengths = [a.shape[0] for a in col]
        full, keep = _pad_geometry(lengths, spec.get("truncate"), spec.get("min_length"))
        inner_shape = col[0].shape[1:]
        inner = int(np.prod(inner_shape)) if inner_shape else 1
        dtype = col[0].dtype
        ragged = torch.from_numpy(np.ascontiguousarray(np.concatenate([a.reshape(a.shape[0], inner) for a in col], axis=0)))
        offsets = torch.from_numpy(np.concatenate(([0], np.cumsum(lengths))).astype(np.int64))
        ragged_d = ragged.pin_memory().to(self.device, non_blocking=True)
        offsets_d = offsets.pin_memory().to(self.device, non_blocking=True)
        tdtype = torch.float32 if dtype == np.float32 else torch.int64
        out = torch.empty((len(col), keep) + tuple(inner_shape), dtype=tdtype, device=self.device)
        _C.check(_C.lib().ndt1_pad_pack(ragged_d.data_ptr(), offsets_d.data_ptr(), out.data_ptr(), len(col), keep, inner,
                                        4 if dtype == np.float32 else 8, int(side == "left"), keep, float(value), _C.stream_ptr()),
                 "ndt1_pad_pack")
        return out

    def __call__(self, batch):
        if isinstance(batch[0], list):
            batch = [row for sub in batch for row in sub]
        keys, arrays, strings = _classify(batch)
        assert set(self.pad_dict.keys()).issubset(arrays), f"Can't pad keys which are not arrays: {set(self.pad_dict.keys()) - set(arrays)} "
        padded_batch, unused = {}, {}
        for key in keys:
            col = [row[key] for row in batch]
            is_input = key in self.model_inputs
            if key in arrays:
                spec = self.pad_dict.get(key)
                on_device = is_input and spec is not None and spec.get("dim", 0) == 0 and col[0].dtype in (np.float32, np.int64) and col[0].ndim >= 1
                if on_device:
                    value = self._pack(col, spec)
                elif spec is not None:
                    value = torch.from_numpy(padded_array(col, **spec)).clone()
                    value = value.to(self.device) if is_input else value
                elif len({a.shape for a in col}) == 1:
                    value = torch.from_numpy(np.stack(col, axis=0))
                    value = value.to(self.device, non_blocking=True) if is_input else value
                else:
                    value = [torch.from_numpy(a) for a in col]
            elif key in strings:
                value = np.stack(col, axis=0)
            else:
                value = col
            (padded_batch if is_input else unused)[key] = value
        return padded_batch, unused
