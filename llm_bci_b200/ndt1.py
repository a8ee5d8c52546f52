"""NDT1 with a B200-native device path.

Plugin surface of the reference (models/ndt1.py:455-692): a class registered
under ``NAME2MODEL["NDT1"]``, built as ``NDT1(trainer_config.model,
**method.model_kwargs)``, called as ``model(**model_inputs)`` with the
parameter names the trainer introspects (models/trainer.py:161-171), returning
an ``NDT1Output`` (a ``ModelOutput``), with ``save_checkpoint`` /
``load_checkpoint`` writing ``encoder.bin`` / ``encoder_config.pth`` /
``decoder.bin``.  Parameter names, shapes and construction order are the
reference's, so ``torch.manual_seed(s)`` reproduces its
initialisation and state_dicts interchange.

The torch modules below only *hold* parameters.  All device math -- smoothing
and noise, masking, embedding, stack projection, attention, LayerNorm, MLP,
head, CTC / Poisson loss and the whole backward -- runs in the hand-written
sm_100a kernels of ``csrc/`` through the C ABI (``_C.py``).  There is no CPU or
PyTorch fallback: calling the model off-GPU raises.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _C
from .config import DictConfig, update_config
from .masker import Masker
from .model_output import ModelOutput

DEFAULT_CONFIG = "configs/ndt1.yaml"


@dataclass
class NDT1Output(ModelOutput):
    loss: Optional[torch.FloatTensor] = None
    n_examples: Optional[torch.LongTensor] = None
    mask: Optional[torch.LongTensor] = None
    preds: Optional[torch.FloatTensor] = None
    targets: Optional[torch.FloatTensor] = None


def create_context_mask(context_forward: int, context_backward: int, max_F: int) -> torch.LongTensor:
    """Band of allowed (query, key) pairs as a matrix (models/ndt1.py:30-41).  The engine never
    materialises it (it evaluates the predicate in the attention kernels); kept for API parity."""
    if context_forward == -2 and context_backward == -2:
        return torch.ones(max_F, max_F, dtype=torch.int64)
    cf = context_forward if context_forward >= -1 else max_F
    cb = context_backward if context_backward >= -1 else max_F
    i = torch.arange(max_F)[:, None]
    j = torch.arange(max_F)[None, :]
    return ((j <= i + cf) & (j >= i - cb)).to(torch.int64)


def gaussian_taps(smooth_sd: int) -> np.ndarray:
    """Normalised gaussian window of 1+6*sd points (models/ndt1.py:87-88), float64."""
    m = 1 + smooth_sd * 6
    n = np.arange(m, dtype=np.float64) - (m - 1) / 2.0
    w = np.exp(-0.5 * (n / float(smooth_sd)) ** 2)
    return w / w.sum()


def _seed_from_torch() -> int:
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class SmoothAndNoise(nn.Module):
    """models/ndt1.py:78-107 as one prologue kernel (``ndt1_smooth_noise``).

    ``rng="device"`` (default) draws the two Gaussian noises inside the kernel
    (Philox); ``rng="reference"`` draws them with ``torch.randn`` on the device in
    the reference's order and injects them."""

    def __init__(self, config: DictConfig, rng: str = "device"):
        super().__init__()
        self.noise = config.noise
        self.white_noise_sd = config.white_noise_sd
        self.constant_offset_sd = config.constant_offset_sd
        self.smooth = config.smooth_sd is not None
        self.rng = rng
        self.seed_tensor = None      # optional int64 CUDA tensor: element 0 holds the Philox key of the noise (read on the device when
                                     # the kernel runs, so a captured step is replayed with fresh noise; DataParallelTrainer sets it)
        if self.smooth:
            kernel = torch.from_numpy(gaussian_taps(config.smooth_sd))
            self.register_buffer("kernel", kernel, persistent=False)
            self._taps_host = kernel.numpy().astype(np.float32).copy()    # host copy for the launch (the taps are a by-value kernel argument):
                                                                          # reading the device buffer back every forward would synchronise the stream

    def forward(self, spikes: torch.Tensor, noise: Optional[Dict[str, torch.Tensor]] = None, bf16_out: bool = False) -> torch.Tensor:
        """``bf16_out``: write the result directly as the bfloat16 operand of the channel-embedding GEMM (what the engine would
        otherwise cast it to) instead of float32 -- only the engine consumes it then (``ndt1_batch.spikes_bf16``)."""
        if not spikes.is_cuda:
            raise RuntimeError("llm_bci_b200 runs on the GPU only (no CPU fallback)")
        B, T, N = spikes.shape
        spikes = spikes.contiguous().float()
        add_noise = bool(self.noise) and self.training
        if not self.smooth and not add_noise:
            return spikes
        bf16_out = bool(bf16_out) and N % 8 == 0
        out = torch.empty(spikes.shape, dtype=torch.bfloat16 if bf16_out else torch.float32, device=spikes.device)
        taps = self._taps_host if self.smooth else np.zeros(0, dtype=np.float32)
        taps_c = taps.ctypes.data_as(_C.C.POINTER(_C.C.c_float))
        white = offset = None
        wsd = osd = 0.0
        use_philox, seed, seed_ptr = 0, 0, None
        if add_noise:
            wsd = float(self.white_noise_sd) if self.white_noise_sd is not None else 0.0
            osd = float(self.constant_offset_sd) if self.constant_offset_sd is not None else 0.0
            if noise is not None:
                white, offset = noise.get("white"), noise.get("offset")
            elif self.rng == "reference":
                if self.white_noise_sd is not None:
                    white = torch.randn(B, T, N, dtype=spikes.dtype, device=spikes.device)
                if self.constant_offset_sd is not None:
                    offset = torch.randn(B, 1, N, dtype=spikes.dtype, device=spikes.device)
            else:
                use_philox = 1
                if self.seed_tensor is not None:
                    seed_ptr = self.seed_tensor.data_ptr()
                else:
                    seed = _seed_from_torch()
        _C.check(_C.lib().ndt1_smooth_noise(spikes.data_ptr(), None if bf16_out else out.data_ptr(), B, T, N, taps_c, len(taps), wsd, osd,
                                            _C.ptr(white), _C.ptr(offset), use_philox, seed, seed_ptr, out.data_ptr() if bf16_out else None, N,
                                            _C.stream_ptr()), "ndt1_smooth_noise")
        return out


class NeuralEmbeddingLayer(nn.Module):
    """Parameter container of models/ndt1.py:111-208; the arithmetic is in the engine."""

    def __init__(self, hidden_size: int, config: DictConfig):
        super().__init__()
        self.adapt = config.adapt
        self.pos = config.pos
        self.block_token = config.block_token
        self.day_token = config.day_token
        self.bias = config.bias
        self.input_dim = config.input_dim
        if self.adapt:
            self.embed_spikes = nn.ModuleList([nn.Linear(config.n_channels, self.input_dim, bias=config.bias) for _ in range(config.n_days)])
        else:
            self.embed_spikes = nn.Linear(config.n_channels, self.input_dim, bias=config.bias)
        self.stack = config.stack.active
        if self.stack:
            self.stack_size = config.stack.size
            self.stack_stride = config.stack.stride
            self.stack_projection = nn.Linear(self.input_dim * config.stack.size, hidden_size)
        else:
            self.projection = nn.Linear(self.input_dim, hidden_size)
        self.act_name = config.act
        if self.pos:
            self.embed_pos = nn.Embedding(config.max_F, hidden_size)
        if self.block_token:
            self.block_embedding = nn.Embedding(config.n_blocks, hidden_size)
        if self.day_token:
            self.day_embedding = nn.Embedding(config.n_days, hidden_size)
        self.dropout = nn.Dropout(config.dropout)

    def get_stacked_lens(self, lens: torch.Tensor) -> torch.Tensor:
        """models/ndt1.py:207-208."""
        return lens if not self.stack else (1 + (lens - self.stack_size) / self.stack_stride).to(lens.dtype)


class NeuralMLP(nn.Module):
    def __init__(self, hidden_size, inter_size, act, use_bias, dropout):
        super().__init__()
        self.up_proj = nn.Linear(hidden_size, inter_size, bias=use_bias)
        self.act_name = act
        self.down_proj = nn.Linear(inter_size, hidden_size, bias=use_bias)
        self.dropout = nn.Dropout(dropout)


class NeuralAttention(nn.Module):
    def __init__(self, idx, hidden_size, n_heads, use_bias, dropout, use_rope=False, base=10000., max_F=1024):
        super().__init__()
        self.idx = idx
        self.hidden_size = hidden_size
        self.n_heads = n_heads
        assert self.hidden_size % self.n_heads == 0, "Hidden dim is not multiple of head size"
        self.head_size = self.hidden_size // self.n_heads
        self.use_rope = use_rope
        self.query = nn.Linear(hidden_size, hidden_size, bias=use_bias)
        self.key = nn.Linear(hidden_size, hidden_size, bias=use_bias)
        self.value = nn.Linear(hidden_size, hidden_size, bias=use_bias)
        self.attn_dropout = dropout
        self.dropout = nn.Dropout(dropout)
        self.out_proj = nn.Linear(hidden_size, hidden_size, bias=use_bias)
        if use_rope:   # models/ndt1.py:44-53, 262-266: same op sequence, so the engine gets the reference's table values bit for bit
            inv_freq = 1.0 / (base ** (torch.arange(0, self.head_size, 2).float() / self.head_size))
            freqs = torch.einsum("i,j->ij", torch.arange(max_F, dtype=inv_freq.dtype), inv_freq)
            emb = torch.cat((freqs, freqs), dim=-1)
            self.register_buffer("cos", emb.cos().to(self.query.weight.dtype), persistent=False)
            self.register_buffer("sin", emb.sin().to(self.query.weight.dtype), persistent=False)


class NeuralEncoderLayer(nn.Module):
    def __init__(self, idx, max_F, config: DictConfig):
        super().__init__()
        self.idx = idx
        self.use_rope = config.use_rope
        self.ln1 = nn.LayerNorm(config.hidden_size)
        self.attn = NeuralAttention(idx, config.hidden_size, config.n_heads, config.attention_bias, config.dropout, config.use_rope,
                                    config.rope_theta, max_F)
        self.ln2 = nn.LayerNorm(config.hidden_size)
        self.mlp = NeuralMLP(config.hidden_size, config.inter_size, config.act, config.mlp_bias, config.dropout)
        if config.fixup_init:
            self.fixup_initialization(config.n_layers)

    @torch.no_grad()
    def fixup_initialization(self, n_layers: int) -> None:
        """models/ndt1.py:332-344: *_proj.weight scaled by 0.67*L^-1/4, value.weight additionally by sqrt(2)."""
        scale = 0.67 * (n_layers) ** (-1. / 4.)
        for name, param in self.named_parameters():
            if name.endswith("_proj.weight"):
                param.copy_(scale * param)
            elif name.endswith("value.weight"):
                param.copy_(scale * (param * (2 ** 0.5)))


class NeuralFactorsProjection(nn.Module):
    def __init__(self, hidden_size, config):
        super().__init__()
        self.out_size = config.size if config.active else hidden_size
        self.dropout = nn.Dropout(config.dropout)
        self.active = config.active
        if config.active:
            self.proj = nn.Sequential(nn.Linear(hidden_size, config.size, config.bias), _ActName(config.act))
            if config.fixup_init:
                self.proj[0].weight.data.uniform_(-config.init_range, config.init_range)
                if config.bias:
                    self.proj[0].bias.data.zero_()
        else:
            self.proj = nn.Identity()


class _ActName(nn.Module):
    """Parameter-free placeholder keeping Sequential indices (and state_dict keys) aligned."""

    def __init__(self, name: str):
        super().__init__()
        self.name = name


class NeuralEncoder(nn.Module):
    """models/ndt1.py:376-450.  ``forward`` is the sub-API used by the BCI coupler
    (models/bci.py:125): returns (features, stacked padding mask, targets_mask)."""

    def __init__(self, config: DictConfig, owner: "NDT1" = None):
        super().__init__()
        self.hidden_size = config.transformer.hidden_size
        self.n_layers = config.transformer.n_layers
        self.masker = nn.ModuleList([Masker(DictConfig(m)) for m in config.masker.values()])
        self.context_forward = config.context.forward
        self.context_backward = config.context.backward
        self.smooth_and_noise = SmoothAndNoise(config.smooth_and_noise)
        self.embedder = NeuralEmbeddingLayer(self.hidden_size, config.embedder)
        self.layers = nn.ModuleList([NeuralEncoderLayer(i, config.embedder.max_F, config.transformer) for i in range(self.n_layers)])
        self.out_norm = nn.LayerNorm(self.hidden_size)
        self.out_proj = NeuralFactorsProjection(self.hidden_size, config.factors)
        object.__setattr__(self, "_owner", owner)

    def prologue(self, spikes: torch.Tensor, noise=None, want_mask: bool = False, masker_draws=None, bf16_ok: bool = False):
        """Smoothing/noise then the maskers (models/ndt1.py:421-427).  Returns (spikes', targets_mask or None).
        ``bf16_ok``: nothing but the bf16 engine reads spikes' -> the prologue kernel may write it as bfloat16 directly."""
        active = [m for m in self.masker if m.is_active()]
        x = self.smooth_and_noise(spikes, noise, bf16_out=bf16_ok and not active)
        if active and x.data_ptr() == spikes.data_ptr():
            x = x.clone()   # the reference mutates its input here; this implementation never does
        targets_mask = None
        if want_mask or active:
            targets_mask = torch.zeros(x.shape, dtype=torch.int64, device=x.device)
        for i, m in enumerate(active):
            x, _ = m(x, targets_mask=targets_mask, draws=None if masker_draws is None else masker_draws[i])
        return x, targets_mask

    def forward(self, spikes, spikes_mask, spikes_timestamp, spikes_lengths=None, block_idx=None, day_idx=None):
        owner = self._owner
        if owner is None:
            raise RuntimeError("NeuralEncoder must be owned by an NDT1 instance")
        return owner._encode(spikes, spikes_mask, spikes_timestamp, spikes_lengths, block_idx, day_idx)


def _flat_param_table(model: "NDT1") -> List[Tuple[str, Optional[torch.nn.Parameter]]]:
    """(C-ABI slot name, parameter) in the order of ndt1_tensors."""
    enc, emb = model.encoder, model.encoder.embedder
    t: List[Tuple[str, Optional[torch.nn.Parameter]]] = []
    es = emb.embed_spikes
    if isinstance(es, nn.ModuleList):      # embedder.adapt: one Linear per day (models/ndt1.py:118-127), routed by day_idx in the engine
        if len(es) > _C.MAX_DAYS:
            raise RuntimeError(f"embedder.adapt supports at most {_C.MAX_DAYS} days (n_days = {len(es)})")
        t += [("embed_w", None), ("embed_b", None)]
        for d, lin in enumerate(es):
            t += [(f"embed_w_day.{d}", lin.weight), (f"embed_b_day.{d}", lin.bias)]
    else:
        t += [("embed_w", es.weight), ("embed_b", es.bias)]
    pr = emb.stack_projection if emb.stack else emb.projection
    t += [("proj_w", pr.weight), ("proj_b", pr.bias)]
    t += [("pos_w", emb.embed_pos.weight if emb.pos else None)]
    t += [("block_emb", emb.block_embedding.weight if emb.block_token else None)]
    t += [("day_emb", emb.day_embedding.weight if emb.day_token else None)]
    for i, layer in enumerate(enc.layers):
        a, m = layer.attn, layer.mlp
        for slot, p in (("ln1_w", layer.ln1.weight), ("ln1_b", layer.ln1.bias), ("q_w", a.query.weight), ("q_b", a.query.bias),
                        ("k_w", a.key.weight), ("k_b", a.key.bias), ("v_w", a.value.weight), ("v_b", a.value.bias),
                        ("o_w", a.out_proj.weight), ("o_b", a.out_proj.bias), ("ln2_w", layer.ln2.weight), ("ln2_b", layer.ln2.bias),
                        ("up_w", m.up_proj.weight), ("up_b", m.up_proj.bias), ("down_w", m.down_proj.weight), ("down_b", m.down_proj.bias)):
            t.append((f"layer.{i}.{slot}", p))
    t += [("out_norm_w", enc.out_norm.weight), ("out_norm_b", enc.out_norm.bias)]
    if enc.out_proj.active:
        t += [("factors_w", enc.out_proj.proj[0].weight), ("factors_b", enc.out_proj.proj[0].bias)]
    else:
        t += [("factors_w", None), ("factors_b", None)]
    t += [("dec_w", model.decoder[0].weight), ("dec_b", model.decoder[0].bias)]
    return t


def _fill_tensors(struct: "_C.Tensors", table, ptr_of) -> None:
    for slot, p in table:
        v = None if p is None else ptr_of(p)
        if slot.startswith("layer."):
            _, i, name = slot.split(".")
            setattr(struct.layer[int(i)], name, v)
        elif slot.startswith("embed_w_day.") or slot.startswith("embed_b_day."):
            name, d = slot.split(".")
            getattr(struct, name)[int(d)] = v
        else:
            setattr(struct, slot, v)


class _EngineFunction(torch.autograd.Function):
    """One autograd node for the whole model: forward = ndt1_engine_forward,
    backward = ndt1_engine_backward (gradients for every parameter at once)."""

    @staticmethod
    def forward(ctx, model: "NDT1", call: dict, *params):
        out = model._engine_forward(call, need_backward=True)
        ctx.model = model
        ctx.n_params = len(params)
        ctx.generation = model._generation
        ctx.mark_non_differentiable(out["preds"])
        return out["loss"], out["preds"]

    @staticmethod
    def backward(ctx, dloss, _dpreds):
        ctx.model._check_generation(ctx.generation)
        grads = ctx.model._engine_backward(dloss)
        return (None, None) + tuple(grads)


class _EncoderFunction(torch.autograd.Function):
    """NeuralEncoder.forward as a differentiable sub-module (models/bci.py:125 trains through it):
    forward = ndt1_engine_forward(encoder_only), backward = ndt1_engine_backward_features."""

    @staticmethod
    def forward(ctx, model: "NDT1", call: dict, *params):
        out = model._engine_forward(call, need_backward=True)
        ctx.model = model
        ctx.generation = model._generation
        ctx.mark_non_differentiable(out["out_mask"])
        return out["features"], out["out_mask"]

    @staticmethod
    def backward(ctx, dfeatures, _dmask):
        ctx.model._check_generation(ctx.generation)
        grads = ctx.model._engine_backward(None, dfeatures=dfeatures)
        return (None, None) + tuple(grads)


class NDT1(nn.Module):

    def __init__(self, config: DictConfig, **kwargs):
        super().__init__()
        self.precision = kwargs.pop("precision", os.environ.get("NDT1_PRECISION", "bf16"))
        self._max_batch = kwargs.pop("max_batch", None)
        self._max_T = kwargs.pop("max_T", None)
        kwargs.pop("device", None)
        kwargs.pop("engine", None)
        config = update_config(DEFAULT_CONFIG, config)
        self.method = kwargs["method_name"]

        encoder_pt_path = config["encoder"].pop("from_pt", None)
        if encoder_pt_path is not None:
            encoder_config = torch.load(os.path.join(encoder_pt_path, "encoder_config.pth"), weights_only=False)
            config["encoder"] = update_config(config.encoder, encoder_config)
        self.encoder = NeuralEncoder(config.encoder, owner=self)
        if encoder_pt_path is not None:
            self.encoder.load_state_dict(torch.load(os.path.join(encoder_pt_path, "encoder.bin")))

        if self.method == "mlm":
            assert config.encoder.masker.active, "Can't pretrain with inactive masking"
            assert not config.encoder.embedder.stack.active, "Can't pretrain with stacked inputs"
            n_outputs = config.encoder.embedder.n_channels
        elif self.method == "autoregressive":
            assert config.encoder.context.forward == 0, "Autoregressive training requires context.forward == 0"
            assert not config.encoder.embedder.stack.active, "Can't train autoregressive with stacked inputs"
            n_outputs = config.encoder.embedder.n_channels
        elif self.method in ["ctc", "endtoend"]:
            n_outputs = kwargs["vocab_size"]
        else:
            raise Exception(f"Method {self.method} not implemented yet for NDT1")

        decoder_layers: List[nn.Module] = [nn.Linear(self.encoder.out_proj.out_size, n_outputs)]
        self._decoder_relu = False
        if self.method in ["mlm", "autoregressive"] and (kwargs["loss"] == "mse" or not kwargs["log_input"]):
            decoder_layers.append(nn.ReLU())
            self._decoder_relu = True
        elif self.method in ["ctc", "endtoend"]:
            decoder_layers.append(nn.LogSoftmax(dim=-1))
        self.decoder = nn.Sequential(*decoder_layers)
        if encoder_pt_path is not None:
            self.decoder.load_state_dict(torch.load(os.path.join(encoder_pt_path, "decoder.bin")))

        if self.method in ["mlm", "autoregressive"]:
            self.loss_name = kwargs["loss"]
            self.log_input = kwargs["log_input"]
            if self.loss_name == "poisson_nll":
                self._loss_kind = _C.LOSS_POISSON_LOG if self.log_input else _C.LOSS_POISSON_RATE
            elif self.loss_name == "mse":
                self._loss_kind = _C.LOSS_MSE
            else:
                raise Exception(f"Loss {kwargs['loss']} not implemented yet for mlm")
            self._blank, self._zero_inf = 0, 1
        else:
            self._loss_kind = _C.LOSS_CTC
            self._blank = int(kwargs["blank_id"])
            self._zero_inf = int(bool(kwargs["zero_infinity"]))
        self._n_outputs = int(n_outputs)
        self.config = config
        self._engine = None
        self._engine_cap = (0, 0, 0)
        self._engine_epoch = 0        # bumped whenever the engine (and with it every arena address) is re-created
        self._ptable = None
        self._pstruct = None
        self._goffs = None            # cached _grad_offsets() of the current parameter table
        self._gstruct = {}            # gradient pointer tables by base address of the flat buffer they point into
        self._last = None
        self._last_out = None
        self._generation = 0          # forwards run so far: the engine keeps the activations of the LAST one only
        self._weight_shadow = None
        self._param_stream = None     # a side stream that may still be updating the parameters (DataParallelTrainer's optimizer)
        self._param_event = None      # ... or the C event it records after its last update (usable from inside a captured step)
        self._gather_hook = None      # sharded optimizer: all-gathers the fp32 master weights before they are read outside the engine
        self._seed_tensor = None      # int64[2] CUDA tensor holding (noise key, dropout key) when the step's keys live in device memory

    # ------------------------------------------------------------------ engine plumbing
    def _apply(self, fn, *a, **k):
        self._ptable = None   # parameters may have moved
        self._goffs, self._gstruct = None, {}
        return super()._apply(fn, *a, **k)

    def invalidate_param_cache(self) -> None:
        self._ptable = None
        self._goffs, self._gstruct = None, {}

    def _params(self):
        if self._ptable is None:
            self._ptable = _flat_param_table(self)
            self._pstruct = _C.Tensors()
            for slot, p in self._ptable:
                if p is not None:
                    if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                        raise RuntimeError(f"parameter {slot} must be a contiguous float32 CUDA tensor (got {p.dtype} on {p.device})")
            _fill_tensors(self._pstruct, self._ptable, lambda p: p.data_ptr())
        return self._ptable, self._pstruct

    def _engine_config(self, max_batch: int, max_T: int, max_S: int) -> "_C.Config":
        e, t, f = self.config.encoder.embedder, self.config.encoder.transformer, self.config.encoder.factors
        c = _C.Config()
        c.abi_version = _C.ABI_VERSION
        c.precision = _C.PRECISION[self.precision]
        c.n_channels, c.input_dim, c.max_F = e.n_channels, e.input_dim, e.max_F
        c.embed_bias, c.embed_act, c.pos = int(bool(e.bias)), _C.ACT[e.act], int(bool(e.pos))
        c.stack_active = int(bool(e.stack.active))
        c.stack_size, c.stack_stride = (e.stack.size, e.stack.stride) if e.stack.active else (0, 1)
        c.block_token, c.day_token, c.n_blocks, c.n_days, c.adapt = int(bool(e.block_token)), int(bool(e.day_token)), e.n_blocks, e.n_days, int(bool(e.adapt))
        c.n_layers, c.hidden, c.n_heads, c.inter = t.n_layers, t.hidden_size, t.n_heads, t.inter_size
        c.attention_bias, c.mlp_bias, c.mlp_act = int(bool(t.attention_bias)), int(bool(t.mlp_bias)), _C.ACT[t.act]
        c.use_rope, c.rope_theta = int(bool(t.use_rope)), float(t.rope_theta)
        c.context_forward, c.context_backward = self.config.encoder.context.forward, self.config.encoder.context.backward
        c.factors_active, c.factors_size, c.factors_act, c.factors_bias = int(bool(f.active)), f.size, _C.ACT[f.act], int(bool(f.bias))
        c.method = _C.METHOD[self.method]
        c.loss_kind, c.n_outputs, c.blank_id, c.zero_infinity = self._loss_kind, self._n_outputs, self._blank, self._zero_inf
        c.decoder_relu = int(self._decoder_relu)
        c.p_embed, c.p_transformer, c.p_factors = float(e.dropout), float(t.dropout), float(f.dropout)
        c.max_batch, c.max_T, c.max_targets = max_batch, max_T, max_S
        return c

    def _get_engine(self, B: int, T: int, S: int):
        cb, ct, cs = self._engine_cap
        if self._engine is None or B > cb or T > ct or S > cs:
            L = _C.lib()
            if self._engine is not None:
                torch.cuda.synchronize()
                L.ndt1_engine_destroy(self._engine)
                self._engine = None
            nb, nt, ns = max(B, cb, self._max_batch or 0), max(T, ct, self._max_T or 0), max(S, cs, 64)
            cfg = self._engine_config(nb, nt, ns)
            h = _C._p()
            _C.check(L.ndt1_engine_create(_C.C.byref(cfg), _C.C.byref(h)), "ndt1_engine_create")
            self._engine, self._engine_cap = h, (nb, nt, ns)
            self._engine_epoch += 1
            self._apply_weight_shadow()
            if self.config.encoder.transformer.use_rope:
                a = self.encoder.layers[0].attn
                dev = next(self.parameters()).device
                cs, sn = a.cos.to(dev, torch.float32).contiguous(), a.sin.to(dev, torch.float32).contiguous()
                _C.check(L.ndt1_engine_set_rope_tables(h, cs.data_ptr(), sn.data_ptr(), cs.shape[0]), "ndt1_engine_set_rope_tables")
        return self._engine

    def set_weight_shadow(self, flat_param: Optional[torch.Tensor], shadow: Optional[torch.Tensor]) -> None:
        """bf16 mode: `shadow` is a bfloat16 copy (same offsets) of the flat fp32 arena `flat_param` that the parameters
        are views of; whoever updates the parameters keeps it current (DataParallelTrainer's fused AdamW does), and the
        engine then reads the weights from it instead of casting them every forward.  None switches it off."""
        self._weight_shadow = None if shadow is None else (flat_param, shadow)
        self._apply_weight_shadow()

    def _apply_weight_shadow(self) -> None:
        if self._engine is None:
            return
        ws = getattr(self, "_weight_shadow", None)
        if ws is None or self.precision != "bf16":
            _C.check(_C.lib().ndt1_engine_set_weight_shadow(self._engine, None, None, 0), "ndt1_engine_set_weight_shadow")
        else:
            _C.check(_C.lib().ndt1_engine_set_weight_shadow(self._engine, ws[0].data_ptr(), ws[1].data_ptr(), ws[0].numel()),
                     "ndt1_engine_set_weight_shadow")

    def __del__(self):
        try:
            if getattr(self, "_engine", None) is not None and _C._lib is not None:
                _C._lib.ndt1_engine_destroy(self._engine)
        except Exception:
            pass

    def engine_launch_count(self) -> int:
        return int(_C.lib().ndt1_engine_launch_count(self._engine)) if self._engine is not None else 0

    def engine_arena_bytes(self) -> int:
        return int(_C.lib().ndt1_engine_arena_bytes(self._engine)) if self._engine is not None else 0

    def wait_for_parameters(self) -> None:
        """Order the current stream after a pending optimizer update on a side stream (see DataParallelTrainer.train_step)."""
        if self._param_event is not None:      # (an EXTERNAL wait node when the current stream is being captured into a graph)
            _C.check(_C.lib().ndt1_stream_wait_event(_C.stream_ptr(), self._param_event), "ndt1_stream_wait_event")
        elif self._param_stream is not None:
            torch.cuda.current_stream().wait_stream(self._param_stream)

    def _check_generation(self, generation: int) -> None:
        """The engine holds the activations, dropout seed and batch geometry of its most recent forward only (one arena,
        nothing allocated per step).  A backward through an OLDER forward -- ``(model(a).loss + model(b).loss).backward()``,
        or an encoder sub-API call followed by a full forward -- would silently use the newer activations: refuse it."""
        if generation != self._generation:
            raise RuntimeError("llm_bci_b200.NDT1 keeps one live autograd graph per model: this backward belongs to forward "
                               f"#{generation}, but forward #{self._generation} has since overwritten the engine's activations. "
                               "Call backward() before the next forward (accumulate gradients across backward calls instead of "
                               "summing losses), or use a second model instance.")

    def _engine_forward(self, call: dict, need_backward: bool) -> dict:
        L = _C.lib()
        self._generation += 1
        self.wait_for_parameters()     # (after the prologue: smoothing / noise / masking do not read a parameter)
        x = call["spikes"]
        B, T, N = x.shape
        dev = x.device
        x_bf16 = x if x.dtype == torch.bfloat16 else None      # the prologue already wrote the GEMM operand (SmoothAndNoise bf16_out)
        tg = call.get("targets")
        S = int(tg.shape[1]) if (tg is not None and self.method in ("ctc", "endtoend")) else 0
        eng = self._get_engine(B, T, S)
        _, pstruct = self._params()
        Tp = int(L.ndt1_engine_out_len(eng, T))
        n_prefix = int(bool(self.config.encoder.embedder.block_token)) + int(bool(self.config.encoder.embedder.day_token))
        enc_only = bool(call.get("encoder_only", False))
        loss = torch.empty((), dtype=torch.float32, device=dev)
        n_examples = torch.empty((), dtype=torch.int64, device=dev)
        out_mask = torch.empty((B, n_prefix + Tp), dtype=torch.int64, device=dev)
        hout = self.encoder.out_proj.out_size
        features = torch.empty((B, Tp, hout), dtype=torch.float32, device=dev) if enc_only else None
        preds = loss_mask = None
        if not enc_only:
            preds = torch.empty((B, Tp, self._n_outputs), dtype=torch.float32, device=dev)
            if self.method == "mlm":
                loss_mask = torch.empty((B, T, N), dtype=torch.int64, device=dev)
        i64 = lambda t: None if t is None else t.to(device=dev, dtype=torch.int64).contiguous()
        keep = dict(mask=i64(call["spikes_mask"]), ts=i64(call["spikes_timestamp"]), lens=i64(call.get("spikes_lengths")),
                    blk=i64(call.get("block_idx")), day=i64(call.get("day_idx")), tg=None, tl=None)
        if keep["lens"] is None:
            keep["lens"] = keep["mask"].sum(1)
        b = _C.Batch()
        b.spikes, b.spikes_mask, b.spikes_timestamp, b.spikes_lengths = x.data_ptr(), keep["mask"].data_ptr(), keep["ts"].data_ptr(), keep["lens"].data_ptr()
        if x_bf16 is not None:
            if self.precision != "bf16" or N % 8 != 0:
                raise RuntimeError("a bfloat16 input needs the bf16 engine and n_channels % 8 == 0")
            b.spikes, b.spikes_bf16 = None, x_bf16.data_ptr()
        b.block_idx, b.day_idx = _C.ptr(keep["blk"]), _C.ptr(keep["day"])
        if self.method in ("ctc", "endtoend") and not enc_only:
            keep["tg"], keep["tl"] = i64(tg), i64(call["targets_lengths"])
            b.targets, b.targets_lengths = keep["tg"].data_ptr(), keep["tl"].data_ptr()
        elif not enc_only:
            keep["rt"] = call["recon_targets"].contiguous().float()
            b.recon_targets = keep["rt"].data_ptr()
            b.targets_mask = _C.ptr(call.get("targets_mask"))
            keep["tm"] = call.get("targets_mask")
        b.B, b.T, b.S = B, T, S
        b.training, b.need_backward, b.encoder_only = int(self.training), int(need_backward), int(enc_only)
        if self._seed_tensor is not None and self.training:      # the key lives in device memory (element 1 of the trainer's seed pair)
            b.seed, b.seed_ptr = 0, self._seed_tensor.data_ptr() + 8
        else:
            b.seed, b.seed_ptr = (_seed_from_torch() if self.training else 0), None
        o = _C.Outputs()
        o.loss, o.n_examples, o.preds = loss.data_ptr(), n_examples.data_ptr(), _C.ptr(preds)
        o.out_mask, o.loss_mask, o.features = out_mask.data_ptr(), _C.ptr(loss_mask), _C.ptr(features)
        _C.check(L.ndt1_engine_forward(eng, _C.C.byref(pstruct), _C.C.byref(b), _C.C.byref(o), _C.stream_ptr()), "ndt1_engine_forward")
        keep["x"] = x   # inputs must outlive the backward (the engine reads them again for the wgrads)
        self._last = keep
        self._last_out = dict(loss=loss, n_examples=n_examples, preds=preds if preds is not None else features, out_mask=out_mask,
                              loss_mask=loss_mask, features=features)
        return self._last_out

    def _engine_backward(self, dloss: Optional[torch.Tensor], into: Optional[torch.Tensor] = None, dfeatures: Optional[torch.Tensor] = None):
        """Returns one gradient tensor per parameter (views of one flat fp32 buffer).  ``dfeatures`` continues an
        encoder-only forward from the gradient w.r.t. its features instead of from the loss."""
        table, pstruct = self._params()
        offs, total = self._grad_offsets()
        dev = next(p for _, p in table if p is not None).device
        flat = into if into is not None else torch.zeros(total, dtype=torch.float32, device=dev)
        base = flat.data_ptr()
        g = self._gstruct.get(base) if into is not None else None     # (a trainer hands in the same arena every step: ~0.4 ms of host work)
        if g is None:
            g = _C.Tensors()
            _fill_tensors(g, table, lambda p: base + 4 * offs[id(p)])
            if into is not None:
                self._gstruct = {base: g}
        if dfeatures is not None:
            df = dfeatures.detach().to(device=dev, dtype=torch.float32).contiguous()
            _C.check(_C.lib().ndt1_engine_backward_features(self._engine, _C.C.byref(pstruct), _C.C.byref(g), df.data_ptr(), _C.stream_ptr()),
                     "ndt1_engine_backward_features")
        else:
            dl = dloss.detach().to(device=dev, dtype=torch.float32).contiguous()
            _C.check(_C.lib().ndt1_engine_backward(self._engine, _C.C.byref(pstruct), _C.C.byref(g), dl.data_ptr(), _C.stream_ptr()),
                     "ndt1_engine_backward")
        if into is not None:
            return None                # the caller owns the arena (DataParallelTrainer: its parameters' .grad are views of it)
        return [flat[offs[id(p)]:offs[id(p)] + p.numel()].view_as(p) for p in self._autograd_params()]

    def _grad_offsets(self):
        """Offsets (in floats, 256-byte aligned) of every parameter inside one flat gradient / parameter arena, and its layout.

        Arena order: gradient STAGES in backward-completion order of the engine (ndt1_engine_wait_stage) -- embedder | layer 0 ..
        layer L-1 | head -- and inside every stage first the BIG tensors (GEMM weights, which the bf16 engine reads through the
        weight shadow) then the SMALL ones (biases, LayerNorm affines, the position / token tables: read in fp32).  Each layer's
        q|k|v weights (and biases) sit next to each other so the engine runs one (3H x H) weight-gradient GEMM, one bias reduction
        and one weight cast for the three.  The big region of a stage starts and ends on a multiple of 1024 floats, so it splits
        evenly over up to 16 ranks (DataParallelTrainer shards the optimizer over it).  ``self._arena_layout`` = one dict per
        stage {stage, big: (lo, hi), small: (lo, hi)} with the trainer's stage numbering (0 = head, 1.. = layers L-1..0,
        L+1 = embedder)."""
        table, _ = self._params()
        if self._goffs is not None:
            return self._goffs
        n_layers = self.config.encoder.transformer.n_layers
        big_slots = {"embed_w", "proj_w", "q_w", "k_w", "v_w", "o_w", "up_w", "down_w", "factors_w"}
        rank = {"q_w": 0, "k_w": 1, "v_w": 2, "q_b": 0, "k_b": 1, "v_b": 2}

        def stage_of(slot):
            if slot.startswith("layer."):
                return n_layers - int(slot.split(".")[1])
            if slot in ("out_norm_w", "out_norm_b", "factors_w", "factors_b", "dec_w", "dec_b"):
                return 0
            return n_layers + 1

        groups = {}
        for i, (slot, p_) in enumerate(table):
            if p_ is None:
                continue
            name = slot.split(".")[2] if slot.startswith("layer.") else slot.split(".")[0]
            big = name in big_slots and p_.numel() >= 65536
            groups.setdefault(stage_of(slot), {True: [], False: []})[big].append((rank.get(name, 3), i, p_))
        offs, off, layout = {}, 0, []
        for stage in sorted(groups, reverse=True):           # embedder first ... head last (any fixed order works; spans are per stage)
            ent = {"stage": stage}
            for big in (True, False):
                if big:
                    off = (off + 1023) // 1024 * 1024
                lo = off
                for _, _, p_ in sorted(groups[stage][big], key=lambda t: (t[0], t[1])):
                    offs[id(p_)] = off
                    off += (p_.numel() + 63) // 64 * 64
                if big:
                    off = (off + 1023) // 1024 * 1024
                ent["big" if big else "small"] = (lo, off)
            layout.append(ent)
        self._arena_layout = layout
        self._goffs = (offs, off)
        return self._goffs

    def _autograd_params(self) -> List[torch.nn.Parameter]:
        table, _ = self._params()
        return [p for _, p in table if p is not None]

    # ------------------------------------------------------------------ public API
    def _encode(self, spikes, spikes_mask, spikes_timestamp, spikes_lengths=None, block_idx=None, day_idx=None):
        """NeuralEncoder.forward (models/ndt1.py:408-450): (features, stacked mask, targets_mask).  With gradients enabled
        the features carry an autograd node, so a model stacked on the encoder (models/bci.py:125) trains through it."""
        x, targets_mask = self.encoder.prologue(spikes, want_mask=True, bf16_ok=self.precision == "bf16")
        call = dict(spikes=x, spikes_mask=spikes_mask, spikes_timestamp=spikes_timestamp, spikes_lengths=spikes_lengths,
                    block_idx=block_idx, day_idx=day_idx, encoder_only=True)
        params = [p for p in self._autograd_params() if not any(p is q for q in self.decoder.parameters())]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            feats, out_mask = _EncoderFunction.apply(self, call, *self._autograd_params())
            return feats, out_mask, targets_mask
        out = self._engine_forward(call, need_backward=False)
        return out["features"], out["out_mask"], targets_mask

    def forward(
        self,
        spikes: torch.FloatTensor,             # (bs, seq_len, n_channels)
        spikes_mask: torch.LongTensor,         # (bs, seq_len)
        spikes_timestamp: torch.LongTensor,    # (bs, seq_len)
        spikes_lengths: torch.LongTensor,      # (bs)
        targets: Optional[torch.FloatTensor] = None,      # (bs, tar_len)
        targets_lengths: Optional[torch.LongTensor] = None,  # (bs)
        block_idx: Optional[torch.LongTensor] = None,     # (bs)
        day_idx: Optional[torch.LongTensor] = None,       # (bs)
        noise: Optional[Dict[str, torch.Tensor]] = None,  # test hook: injected N(0,1) draws {"white","offset"}
        masker_draws: Optional[list] = None,              # test hook: injected masker draws, one dict per active masker
    ) -> NDT1Output:
        if not spikes.is_cuda:
            raise RuntimeError("llm_bci_b200.NDT1 runs on the GPU only (there is no CPU fallback); move the model and the batch to cuda")
        call = dict(spikes_mask=spikes_mask, spikes_timestamp=spikes_timestamp, spikes_lengths=spikes_lengths, block_idx=block_idx,
                    day_idx=day_idx)
        if self.method in ["mlm", "autoregressive"]:
            assert targets is None, "No targets needed for ssl"
            targets = spikes   # the reference clones because it mutates spikes; nothing here does
            call["recon_targets"] = targets
        else:
            call["targets"], call["targets_lengths"] = targets, targets_lengths
        x, targets_mask = self.encoder.prologue(spikes, noise, want_mask=self.method == "mlm", masker_draws=masker_draws,
                                                bf16_ok=self.precision == "bf16")
        call["spikes"], call["targets_mask"] = x, targets_mask
        params = self._autograd_params()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            loss, preds = _EngineFunction.apply(self, call, *params)
            out = self._last_out
        else:
            out = self._engine_forward(call, need_backward=False)
            loss, preds = out["loss"], out["preds"]
        if self.method == "mlm":
            return NDT1Output(loss=loss, n_examples=out["n_examples"], preds=preds, targets=targets, mask=out["loss_mask"])
        if self.method == "autoregressive":
            return NDT1Output(loss=loss, n_examples=out["n_examples"], preds=preds, targets=targets, mask=spikes_mask)
        return NDT1Output(loss=loss, n_examples=out["n_examples"], preds=preds, targets=targets)

    def forward_backward(self, batch: Dict[str, torch.Tensor], grad_buffer: torch.Tensor, dloss: Optional[torch.Tensor] = None) -> NDT1Output:
        """Training fast path used by DataParallelTrainer: forward, then the backward straight into
        ``grad_buffer`` (the flat fp32 gradient arena, accumulated into), without an autograd graph."""
        with torch.no_grad():
            spikes = batch["spikes"]
            call = {k: batch.get(k) for k in ("spikes_mask", "spikes_timestamp", "spikes_lengths", "block_idx", "day_idx")}
            targets = batch.get("targets")
            if self.method in ["mlm", "autoregressive"]:
                targets = spikes
                call["recon_targets"] = spikes
            else:
                call["targets"], call["targets_lengths"] = targets, batch.get("targets_lengths")
            x, targets_mask = self.encoder.prologue(spikes, batch.get("noise"), want_mask=self.method == "mlm", bf16_ok=self.precision == "bf16")
            call["spikes"], call["targets_mask"] = x, targets_mask
            out = self._engine_forward(call, need_backward=True)
            if dloss is None:
                dloss = torch.ones((), dtype=torch.float32, device=spikes.device)
            self._engine_backward(dloss, into=grad_buffer)
        mask = out["loss_mask"] if self.method == "mlm" else (batch.get("spikes_mask") if self.method == "autoregressive" else None)
        return NDT1Output(loss=out["loss"], n_examples=out["n_examples"], preds=out["preds"], targets=targets, mask=mask)

    # generation helpers (models/ndt1.py:592-682); host loops over forward
    def generate(self, spikes=None, spikes_mask=None, spikes_timestamp=None, spikes_lengths=None, block_idx=None, day_idx=None,
                 max_new_bins: int = 16):
        if self.method == "mlm":
            return self._generate(spikes, spikes_mask, spikes_timestamp, spikes_lengths, max_new_bins, append_blank=True)
        if self.method == "autoregressive":
            return self._generate(spikes, spikes_mask, spikes_timestamp, spikes_lengths, max_new_bins, append_blank=False)

    @torch.no_grad()
    def _generate(self, spikes, spikes_mask, spikes_timestamp, spikes_lengths, max_new_bins, append_blank):
        dev = next(self.parameters()).device
        N = self.config.encoder.embedder.n_channels
        inputs = spikes if spikes is not None else (None if append_blank else torch.ones(1, 1, N, device=dev))
        mask = spikes_mask if spikes_mask is not None else (None if append_blank else torch.ones(1, 1, device=dev, dtype=torch.int64))
        ts = spikes_timestamp if spikes_timestamp is not None else (None if append_blank else torch.zeros(1, 1, device=dev, dtype=torch.int64))
        bins, preds = [], []
        for _ in range(max_new_bins):
            if append_blank:   # mlm: append an empty bin, predict it, write the sample back (models/ndt1.py:662-680)
                inputs = torch.cat((inputs, torch.zeros_like(inputs)[:, :1, :]), 1) if inputs is not None else torch.ones(1, 1, N, device=dev)
                mask = torch.cat((mask, torch.ones_like(mask[:, -1:])), 1) if mask is not None else torch.ones(1, 1, device=dev, dtype=torch.int64)
                ts = torch.cat((ts, ts[:, -1:] + 1), 1) if ts is not None else torch.zeros(1, 1, device=dev, dtype=torch.int64)
            out = self(spikes=inputs, spikes_mask=mask, spikes_timestamp=ts, spikes_lengths=spikes_lengths)
            new_preds = new_bins = out.preds[:, -1:, :]
            if self.loss_name == "poisson_nll":
                if self.log_input:
                    new_preds, new_bins = new_preds.exp(), new_bins.exp()
                new_bins = torch.poisson(new_bins)
            if append_blank:
                inputs = inputs.clone()
                inputs[:, -1:, :] = new_bins
                bins.append(new_bins)
                preds.append(new_preds)
            else:              # autoregressive: append the sample (models/ndt1.py:625-640)
                inputs = torch.cat((inputs, new_bins), 1)
                mask = torch.cat((mask, torch.ones_like(mask[:, -1:])), 1)
                ts = torch.cat((ts, ts[:, -1:] + 1), 1)
                bins.append(new_bins[:, 0, :])
                preds.append(new_preds[:, 0, :])
        if append_blank:
            return torch.cat(preds, 1), torch.cat(bins, 1)
        return torch.stack(preds, 1), torch.stack(bins, 1)

    def save_checkpoint(self, save_dir):
        """models/ndt1.py:685-688: same three files, same keys.  (With DataParallelTrainer's sharded optimizer this is a collective:
        every rank calls it, like the reference's trainer does, models/trainer.py:405-409.)"""
        if self._gather_hook is not None:
            self._gather_hook()
        self.wait_for_parameters()
        torch.save(self.encoder.state_dict(), os.path.join(save_dir, "encoder.bin"))
        torch.save(self.config.encoder.get_dict(), os.path.join(save_dir, "encoder_config.pth"))
        torch.save(self.decoder.state_dict(), os.path.join(save_dir, "decoder.bin"))

    def load_checkpoint(self, load_dir):
        """models/ndt1.py:690-692."""
        self.wait_for_parameters()     # a pending side-stream optimizer bucket must not overwrite what is loaded
        self.encoder.load_state_dict(torch.load(os.path.join(load_dir, "encoder.bin")))
        self.decoder.load_state_dict(torch.load(os.path.join(load_dir, "decoder.bin")))
        self._refresh_weight_shadow()

    def load_state_dict(self, state_dict, *args, **kwargs):
        self.wait_for_parameters()
        r = super().load_state_dict(state_dict, *args, **kwargs)
        self._refresh_weight_shadow()
        return r

    def _refresh_weight_shadow(self) -> None:
        """Re-cast the bf16 weight shadow from the flat fp32 arena it mirrors (parameters changed behind the optimizer's back)."""
        ws = getattr(self, "_weight_shadow", None)
        if ws is not None:
            ws[1].copy_(ws[0])
