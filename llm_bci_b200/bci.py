"""BCI coupler: the NDT1 encoder feeding a causal language model (BASELINE.json configs[4]).

Plugin surface of the reference's ``BCI`` (models/bci.py:31-265): ``BCI(config, llm_path, lora, freeze_llm, **kwargs)``,
``forward(input_ids, attention_mask, input_split, spikes, spikes_mask, spikes_timestamp, spikes_lengths, block_idx, day_idx,
targets) -> BCIOutput``, ``prepare_embeds`` (:107-168), ``generate``, ``save_checkpoint`` / ``load_checkpoint`` with the files
``encoder.bin`` / ``encoder_config.pth`` / ``decoder.bin`` / ``projector.bin`` / ``projector_config.pth``.

What runs in this library's sm_100a kernels (through the C ABI, no fallback): the NDT1 encoder forward and backward
(``model.ndt1.encoder(...)``, trainable through ``ndt1_engine_backward_features``), the projector MLP forward and backward
(``ndt1_linear_fwd`` / ``ndt1_linear_bwd``: bias and activation fused, tcgen05 GEMMs in bf16 mode), the stacked validity mask
(``ndt1_stack_valid``) and the splice of the spike features into the prompt embeddings, its mask and its -100 targets
(``ndt1_splice_rows`` / ``ndt1_unsplice_rows``).  The language model itself is the caller's (any Hugging Face causal LM: the
reference loads LLaMA-7B + LoRA, which is not available here; ``debug=True`` builds its 2-layer stand-in, models/bci.py:51-53):
it is a consumer of this path, not part of it.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _C
from .config import DictConfig, update_config
from .model_output import ModelOutput
from .ndt1 import NDT1, _ActName

DEFAULT_CONFIG = "configs/bci.yaml"


@dataclass
class BCIOutput(ModelOutput):
    loss: Optional[torch.FloatTensor] = None
    n_examples: Optional[torch.LongTensor] = None
    mask: Optional[torch.LongTensor] = None
    preds: Optional[torch.FloatTensor] = None
    targets: Optional[torch.FloatTensor] = None


def _workspace(M: int, N: int, K: int, dev) -> torch.Tensor:
    n = max(int(_C.lib().ndt1_linear_workspace_bytes(M, N, K)), M * N * 4 + 256)
    return torch.empty(n, dtype=torch.uint8, device=dev)


class _LinearAct(torch.autograd.Function):
    """y = act(x W^T + b) with the library's GEMM kernels, forward and backward (one projector layer, models/bci.py:88-96)."""

    @staticmethod
    def forward(ctx, x, w, b, act: str, precision: str):
        if not x.is_cuda:
            raise RuntimeError("llm_bci_b200 runs on the GPU only (no CPU fallback)")
        x, w = x.contiguous().float(), w.contiguous().float()
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        pre = torch.empty_like(y) if act == "gelu" else None
        ws = _workspace(M, N, K, x.device)
        _C.check(_C.lib().ndt1_linear_fwd(x.data_ptr(), w.data_ptr(), _C.ptr(b), y.data_ptr(), _C.ptr(pre), M, N, K, _C.ACT[act],
                                          _C.PRECISION[precision], ws.data_ptr(), ws.numel(), _C.stream_ptr()), "ndt1_linear_fwd")
        ctx.save_for_backward(x, w, pre if pre is not None else y)
        ctx.act, ctx.precision, ctx.has_bias = act, precision, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, saved = ctx.saved_tensors
        dy = dy.contiguous().float()
        M, K = x.shape
        N = w.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(w) if ctx.needs_input_grad[1] else None
        db = torch.zeros(N, dtype=torch.float32, device=x.device) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        ws = _workspace(M, N, K, x.device)
        _C.check(_C.lib().ndt1_linear_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), saved.data_ptr(), _C.ptr(dx), _C.ptr(dw), _C.ptr(db),
                                          M, N, K, _C.ACT[ctx.act], _C.PRECISION[ctx.precision], ws.data_ptr(), ws.numel(), _C.stream_ptr()),
                 "ndt1_linear_bwd")
        return dx, dw, db, None, None


class _Splice(torch.autograd.Function):
    """out[b] = a[b, :split[b]] | ins[b] | a[b, split[b]:]  (models/bci.py:143-150); the backward routes the gradient back."""

    @staticmethod
    def forward(ctx, a, ins, split):
        a, ins = a.contiguous().float(), ins.contiguous().float()
        B, La, W = a.shape
        Ls = ins.shape[1]
        split = split.to(device=a.device, dtype=torch.int64).contiguous()
        out = torch.empty(B, La + Ls, W, dtype=torch.float32, device=a.device)
        _C.check(_C.lib().ndt1_splice_rows(a.data_ptr(), ins.data_ptr(), split.data_ptr(), out.data_ptr(), B, La, Ls, W, 4, 0, 0.0, _C.stream_ptr()),
                 "ndt1_splice_rows")
        ctx.save_for_backward(split)
        ctx.dims = (B, La, Ls, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        (split,) = ctx.saved_tensors
        B, La, Ls, W = ctx.dims
        dout = dout.contiguous().float()
        da = torch.empty(B, La, W, dtype=torch.float32, device=dout.device) if ctx.needs_input_grad[0] else None
        dins = torch.empty(B, Ls, W, dtype=torch.float32, device=dout.device) if ctx.needs_input_grad[1] else None
        _C.check(_C.lib().ndt1_unsplice_rows(dout.data_ptr(), split.data_ptr(), _C.ptr(da), _C.ptr(dins), B, La, Ls, W, _C.stream_ptr()),
                 "ndt1_unsplice_rows")
        return da, dins, None


def splice_int64(a: torch.Tensor, ins: Optional[torch.Tensor], split: torch.Tensor, Ls: int, fill: Optional[int] = None) -> torch.Tensor:
    """The integer twin of _Splice for the attention mask (ins = stacked validity) and the targets (constant -100)."""
    a = a.to(torch.int64).contiguous()
    B, La = a.shape
    split = split.to(device=a.device, dtype=torch.int64).contiguous()
    out = torch.empty(B, La + Ls, dtype=torch.int64, device=a.device)
    insp = None if ins is None else ins.to(torch.int64).contiguous()
    _C.check(_C.lib().ndt1_splice_rows(a.data_ptr(), _C.ptr(insp), split.data_ptr(), out.data_ptr(), B, La, Ls, 1, 8, int(fill is not None),
                                       float(fill if fill is not None else 0), _C.stream_ptr()), "ndt1_splice_rows")
    return out


class BCI(nn.Module):

    def __init__(self, config: DictConfig, llm_path: Optional[str] = None, lora: Optional[Dict] = None, freeze_llm: Optional[bool] = False,
                 **kwargs):
        super().__init__()
        config = update_config(DEFAULT_CONFIG, config)
        pt_path = dict(config).pop("from_pt", None)

        if "llm" in kwargs:
            llm = kwargs.pop("llm")
        else:
            from transformers import AutoModelForCausalLM, LlamaConfig
            if kwargs.get("debug"):     # models/bci.py:51-53: the reference's own tiny stand-in decoder
                llm = AutoModelForCausalLM.from_config(LlamaConfig(num_hidden_layers=2, hidden_size=32, intermediate_size=32, num_attention_heads=4))
            else:
                llm = AutoModelForCausalLM.from_pretrained(pt_path or llm_path)
            if lora is not None and pt_path is None:
                try:
                    from peft import LoraConfig, get_peft_model
                except ImportError as e:                       # the decoder is the caller's: LoRA needs the caller's peft install
                    raise RuntimeError("BCI(lora=...) needs the `peft` package (models/bci.py:57-63); pass a prepared model as llm=") from e
                lora = DictConfig(lora)
                llm = get_peft_model(llm, LoraConfig(inference_mode=False, r=lora.r, lora_alpha=lora.alpha, lora_dropout=lora.dropout,
                                                     target_modules=lora.target_modules, modules_to_save=lora.modules_to_save))
            if freeze_llm:
                for param in llm.parameters():
                    param.requires_grad = False
        llm.to(torch.float16)
        self.llm = llm
        self.llm_config = llm.config

        ndt1_pt_path = pt_path or kwargs.pop("load_ndt1_from_pt", None)
        if ndt1_pt_path is not None:
            config["ndt1"]["encoder"]["from_pt"] = ndt1_pt_path
        kwargs.pop("debug", None)
        self.precision = kwargs.get("precision", os.environ.get("NDT1_PRECISION", "bf16"))
        self.ndt1 = NDT1(config.ndt1, **kwargs)

        if pt_path is not None:
            projector_config = torch.load(os.path.join(pt_path, "projector_config.pth"), weights_only=False)
            config["projector"] = update_config(config.projector, projector_config)
        self.stacking = config.projector.stacking
        hidden = config.ndt1.encoder.transformer.hidden_size
        if config.projector.inter_size is not None:       # same Sequential indices as the reference: state_dict keys 0.* and 2.*
            self.projector = nn.Sequential(nn.Linear(hidden * self.stacking, config.projector.inter_size, bias=config.projector.bias),
                                           _ActName(config.projector.act),
                                           nn.Linear(config.projector.inter_size, llm.config.hidden_size, bias=config.projector.bias))
        else:
            self.projector = nn.Linear(hidden * self.stacking, llm.config.hidden_size, bias=config.projector.bias)
        if pt_path is not None:
            self.projector.load_state_dict(torch.load(os.path.join(pt_path, "projector.bin")))
        self.loss_fn = nn.CrossEntropyLoss(reduction="sum")
        self.config = config

    # ------------------------------------------------------------------
    def project(self, feats: torch.Tensor) -> torch.Tensor:
        """(B, T', H) encoder features -> zero-pad T' to a multiple of `stacking`, view (B, T'/s, s H), projector (models/bci.py:127-138)."""
        B, T, H = feats.shape
        s = self.stacking
        Ts = math.ceil(T / s)
        if T % s != 0:
            feats = torch.nn.functional.pad(feats, (0, 0, 0, Ts * s - T))
        x = feats.reshape(B * Ts, H * s)
        if isinstance(self.projector, nn.Sequential):
            l0, act, l2 = self.projector[0], self.projector[1].name, self.projector[2]
            h = _LinearAct.apply(x, l0.weight, l0.bias, act, self.precision)
            y = _LinearAct.apply(h, l2.weight, l2.bias, "identity", self.precision)
        else:
            y = _LinearAct.apply(x, self.projector.weight, self.projector.bias, "identity", self.precision)
        return y.view(B, Ts, -1)

    def prepare_embeds(self, input_ids, attention_mask, input_split, spikes, spikes_mask, spikes_timestamp, spikes_lengths,
                       block_idx=None, day_idx=None, targets=None):
        """models/bci.py:107-168.  Returns (inputs_embeds (B, L + T'/s, llm hidden) fp32, attention_mask, targets)."""
        text_embeds = (self.llm.get_input_embeddings())(input_ids)
        # (the reference passes block_idx / day_idx positionally into the spikes_lengths / block_idx slots, models/bci.py:125;
        #  harmless while tokens and adapt are off -- kept, so that the two behave alike when they are on)
        feats, smask, _ = self.ndt1.encoder(spikes, spikes_mask, spikes_timestamp, block_idx, day_idx)
        B, T, _ = feats.shape
        Ts = math.ceil(T / self.stacking)
        spikes_embeds = self.project(feats)
        smask = smask.to(torch.int64).contiguous()
        stacked = torch.empty(B, Ts, dtype=torch.int64, device=smask.device)
        _C.check(_C.lib().ndt1_stack_valid(smask.data_ptr(), stacked.data_ptr(), B, T, self.stacking, _C.stream_ptr()), "ndt1_stack_valid")
        inputs_embeds = _Splice.apply(text_embeds, spikes_embeds, input_split)
        attention_mask = splice_int64(attention_mask, stacked, input_split, Ts).to(attention_mask.dtype)
        if targets is not None:
            targets = splice_int64(targets, None, input_split, Ts, fill=-100).to(targets.dtype)
        return inputs_embeds, attention_mask, targets

    def forward(self, input_ids, attention_mask, input_split, spikes, spikes_mask, spikes_timestamp, spikes_lengths, block_idx=None,
                day_idx=None, targets=None) -> BCIOutput:
        inputs_embeds, attention_mask, targets = self.prepare_embeds(input_ids, attention_mask, input_split, spikes, spikes_mask,
                                                                     spikes_timestamp, spikes_lengths, block_idx, day_idx, targets)
        inputs_embeds = inputs_embeds.to(self.llm.dtype)
        logits = self.llm(inputs_embeds=inputs_embeds, attention_mask=attention_mask, return_dict=True).logits
        loss = n_examples = None
        if targets is not None:      # models/bci.py:199-212: tokens < n predict n, summed cross-entropy, -100 ignored
            shift_logits = logits[..., :-1, :].contiguous().view(-1, self.llm_config.vocab_size)
            shift_targets = targets[..., 1:].contiguous().view(-1).to(shift_logits.device)
            loss = self.loss_fn(shift_logits, shift_targets)
            n_examples = (shift_targets != -100).sum()
        return BCIOutput(loss=loss, n_examples=n_examples, preds=logits, targets=targets)

    def generate(self, input_ids, attention_mask, input_split, spikes, spikes_mask, spikes_timestamp, spikes_lengths, block_idx=None,
                 day_idx=None, inputs_embeds=None, **gen_config) -> List[torch.LongTensor]:
        if inputs_embeds is None:
            inputs_embeds, attention_mask, _ = self.prepare_embeds(input_ids, attention_mask, input_split, spikes, spikes_mask,
                                                                   spikes_timestamp, spikes_lengths, block_idx, day_idx, targets=None)
        return self.llm.generate(inputs_embeds=inputs_embeds.to(self.llm.dtype), attention_mask=attention_mask, **gen_config)

    def save_checkpoint(self, save_dir):
        """models/bci.py:253-260."""
        self.llm.save_pretrained(save_dir)
        self.ndt1.save_checkpoint(save_dir)
        torch.save(self.projector.state_dict(), os.path.join(save_dir, "projector.bin"))
        torch.save(dict(self.config.projector), os.path.join(save_dir, "projector_config.pth"))

    def load_checkpoint(self, load_dir):
        """models/bci.py:262-265."""
        from transformers import AutoModelForCausalLM
        self.llm = AutoModelForCausalLM.from_pretrained(load_dir).to(self.llm.device)
        self.ndt1.load_checkpoint(load_dir)
        self.projector.load_state_dict(torch.load(os.path.join(load_dir, "projector.bin")))
