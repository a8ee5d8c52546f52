"""CPU oracle for the BCI coupler.  TEST INFRASTRUCTURE ONLY -- see oracle/ndt1_oracle.py for the rules.

Restates ``BCI.prepare_embeds`` (models/bci.py:107-168) on torch-CPU on top of ``ndt1_oracle.encoder_forward``: encoder ->
zero-pad to a multiple of ``projector.stacking`` -> view -> projector MLP (:88-96) -> stacked validity mask -> splice between
the two halves of the prompt (:143-166).  Pinned against outputs of the unmodified reference ``BCI`` run here with its debug
LLaMA (tests/golden/make_golden.py::bci_case -> tests/golden/bci_coupler.npz; ``peft`` is absent and stubbed for the import).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import ndt1_oracle as O

_ACTS = {"relu": F.relu, "gelu": F.gelu, "softsign": F.softsign, "identity": lambda x: x}


def prepare_embeds(params, cfg, text_embeds, attention_mask, input_split, spikes, spikes_mask, spikes_timestamp,
                   block_idx=None, day_idx=None, targets=None, training=False):
    """models/bci.py:107-168.  `params`: state_dict of the coupler with the keys ``ndt1.encoder.*`` and ``projector.*``;
    `text_embeds` = llm.get_input_embeddings()(input_ids) (the language model's own lookup, an input here)."""
    pc = cfg["projector"]
    enc_params = {k[len("ndt1."):]: v for k, v in params.items() if k.startswith("ndt1.")}
    # (:125 -- block_idx / day_idx land in the spikes_lengths / block_idx slots of NeuralEncoder.forward)
    feats, smask, _ = O.encoder_forward(enc_params, cfg["ndt1"]["encoder"], spikes, spikes_mask, spikes_timestamp, day_idx, None, training)
    B, T, H = feats.shape
    s = pc["stacking"]
    if T % s != 0:                                                              # :130-134
        new_T = math.ceil(T / s) * s
        feats = torch.cat((feats, torch.zeros(B, new_T - T, H).to(feats)), 1)
        smask = torch.cat((smask, torch.zeros(B, new_T - T).to(smask)), 1)
        T = new_T
    x = feats.view(B, T // s, H * s)                                            # :137
    if pc["inter_size"] is not None:                                            # :88-96
        x = F.linear(x, params["projector.0.weight"], params.get("projector.0.bias"))
        x = _ACTS[pc["act"]](x)
        x = F.linear(x, params["projector.2.weight"], params.get("projector.2.bias"))
    else:
        x = F.linear(x, params["projector.weight"], params.get("projector.bias"))
    smask = (smask.view(B, T // s, s).sum(-1) == s).to(attention_mask)          # :139-141
    embeds = torch.stack([torch.cat((t[:d], e, t[d:]), 0) for t, e, d in zip(text_embeds, x, input_split)], 0)          # :143-150
    amask = torch.stack([torch.cat((a[:d], m, a[d:]), 0) for a, m, d in zip(attention_mask, smask, input_split)], 0)    # :152-159
    if targets is not None:                                                     # :161-166
        targets = torch.stack([torch.cat((t[:d], torch.ones_like(m).to(t) * (-100), t[d:]), 0)
                               for t, m, d in zip(targets, smask, input_split)], 0)
    return embeds, amask, targets
