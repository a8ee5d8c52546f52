"""iTransformer: neurons as tokens (BASELINE.json configs[3]; reference models/itransformer.py).

Plugin surface of the reference's ``iTransformer`` (models/itransformer.py:212-412): ``iTransformer(config, **method_kwargs)``,
``forward(spikes, spikes_mask, spikes_timestamp, spikes_spacestamp, spikes_lengths, targets, targets_lengths, neuron_regions,
neuron_depths) -> iTransformerOutput``, ``save_checkpoint`` / ``load_checkpoint`` (``encoder.bin``, ``encoder_config.pth``,
``decoder.bin``, ``decoder_config.pth``) with the reference's ``state_dict`` keys: the parameters live in the same torch
containers the reference builds (``nn.Sequential`` MLP with the layer indices of torchvision's ``MLP``, ``nn.Embedding``,
``nn.TransformerEncoder``), constructed in the same order, so a seeded construction draws the same initial values and
checkpoints interchange in both directions.  Those containers are never CALLED: they only hold parameters.

What computes (this library's sm_100a kernels through the C ABI; GPU only, no fallback):
  * every Linear, forward and backward, with bias / activation / activation derivative fused (``ndt1_linear_fwd`` / ``_bwd``:
    tcgen05 GEMMs in the bf16 mode, CUDA-core fp32 in the strict mode),
  * every LayerNorm, forward and backward (``ndt1_layernorm_fwd`` / ``ndt1_layernorm_bwd``) -- the layers are POST-LN,
  * the unmasked multi-head attention over the [cls +] neuron tokens (670 of them, heads of 96, at the shipped size), forward and
    backward, with dropout on the probabilities: batched tcgen05 GEMMs with the probabilities in bf16 in the bf16 mode
    (``ndt1_attention_mm_fwd`` / ``_bwd``), fused CUDA-core fp32 kernels in the strict mode (``ndt1_attention_f32``),
  * every nn.Dropout: inside the epilogue of the Linear it follows (``ndt1_linear_drop_fwd`` / ``_bwd``), or in place
    (``ndt1_dropout_inplace``) after the cls concatenation -- the library's Philox streams, keyed per forward,
  * the maskers (``llm_bci_b200.Masker``), the masked Poisson-NLL / MSE loss and its gradient (``ndt1_recon_loss``).
Left to torch tensor ops (data movement, no arithmetic kernels of this library exist for them): the (B, T, N) -> (B, N, T)
transposes, the embedding row lookups, the cls concatenation and the residual additions.

Scope: the ``mlp`` embedder (the shipped configuration) and the methods ``mlm``, ``dyn_behaviour`` and ``stat_behaviour`` (the
three shipped trainer configs: ssl, wheel, choice).  The ``transformer`` embedder mode and the ``ctc`` method raise
``NotImplementedError`` (the reference's ctc path log-softmaxes the FLAT (bins x vocabulary) vector, :270, not per frame).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _C
from .bci import _LinearAct, _workspace
from .config import DictConfig, update_config
from .masker import Masker
from .model_output import ModelOutput

DEFAULT_CONFIG = "configs/itransformer.yaml"
_ACT_MODULES = {"relu": nn.ReLU, "gelu": nn.GELU, "softsign": nn.Softsign}


@dataclass
class iTransformerOutput(ModelOutput):
    loss: Optional[torch.FloatTensor] = None
    n_examples: Optional[torch.LongTensor] = None
    mask: Optional[torch.LongTensor] = None
    preds: Optional[torch.FloatTensor] = None
    targets: Optional[torch.FloatTensor] = None


def _need_cuda(x: torch.Tensor) -> None:
    if not x.is_cuda:
        raise RuntimeError("llm_bci_b200 runs on the GPU only (no CPU fallback)")


class _LinearActDrop(torch.autograd.Function):
    """y = dropout(act(x W^T + b)): one Linear -> activation -> nn.Dropout run in one GEMM (``ndt1_linear_drop_fwd`` / ``_bwd``:
    bias, activation and the keep mask in the epilogue; the backward masks dy and applies the activation derivative while it
    casts the operand of the two gradient GEMMs).  Same mask as ``ndt1_dropout_inplace`` with the same (seed, site)."""

    @staticmethod
    def forward(ctx, x, w, b, act: str, precision: str, p: float, seed: int, site: int):
        _need_cuda(x)
        x, w = x.contiguous().float(), w.contiguous().float()
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        pre = torch.empty_like(y) if act == "gelu" else None
        ws = _workspace(M, N, K, x.device)
        _C.check(_C.lib().ndt1_linear_drop_fwd(x.data_ptr(), w.data_ptr(), _C.ptr(b), y.data_ptr(), _C.ptr(pre), M, N, K, _C.ACT[act],
                                               _C.PRECISION[precision], ws.data_ptr(), ws.numel(), float(p), seed, site, _C.stream_ptr()),
                 "ndt1_linear_drop_fwd")
        ctx.save_for_backward(x, w, pre if pre is not None else y)
        ctx.cfg = (act, precision, b is not None, float(p), seed, site)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, saved = ctx.saved_tensors
        act, precision, has_bias, p, seed, site = ctx.cfg
        dy = dy.contiguous().float()
        M, K = x.shape
        N = w.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(w) if ctx.needs_input_grad[1] else None
        db = torch.zeros(N, dtype=torch.float32, device=x.device) if (has_bias and ctx.needs_input_grad[2]) else None
        ws = _workspace(M, N, K, x.device)
        _C.check(_C.lib().ndt1_linear_drop_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), saved.data_ptr(), _C.ptr(dx), _C.ptr(dw), _C.ptr(db),
                                               M, N, K, _C.ACT[act], _C.PRECISION[precision], ws.data_ptr(), ws.numel(), p, seed, site,
                                               _C.stream_ptr()), "ndt1_linear_drop_bwd")
        return dx, dw, db, None, None, None, None, None


class _LayerNorm(torch.autograd.Function):
    """nn.LayerNorm(H, eps=1e-5) on (rows, H) fp32."""

    @staticmethod
    def forward(ctx, x, gamma, beta):
        _need_cuda(x)
        x = x.contiguous().float()
        rows, H = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        _C.check(_C.lib().ndt1_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                             rows, H, 1e-5, _C.stream_ptr()), "ndt1_layernorm_fwd")
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous().float()
        rows, H = x.shape
        dx = torch.zeros_like(x)
        dg = torch.zeros_like(gamma)
        db = torch.zeros_like(gamma)
        _C.check(_C.lib().ndt1_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                                             dg.data_ptr(), db.data_ptr(), rows, H, _C.stream_ptr()), "ndt1_layernorm_bwd")
        return dx, dg, db


class _Dropout(torch.autograd.Function):
    """nn.Dropout(p) with the library's Philox stream (seed, site); the backward applies the same keep mask to the gradient."""

    @staticmethod
    def forward(ctx, x, p: float, seed: int, site: int):
        _need_cuda(x)
        y = x.contiguous().float().clone()
        _C.check(_C.lib().ndt1_dropout_inplace(y.data_ptr(), y.numel(), float(p), seed, site, _C.stream_ptr()), "ndt1_dropout_inplace")
        ctx.p, ctx.seed, ctx.site = float(p), seed, site
        return y

    @staticmethod
    def backward(ctx, dy):
        g = dy.contiguous().float().clone()
        _C.check(_C.lib().ndt1_dropout_inplace(g.data_ptr(), g.numel(), ctx.p, ctx.seed, ctx.site, _C.stream_ptr()), "ndt1_dropout_inplace")
        return g, None, None, None


class _Attention(torch.autograd.Function):
    """softmax(q k^T / sqrt(hd)) v over all L tokens of a trial, no mask (nn.MultiheadAttention inside nn.TransformerEncoderLayer,
    models/itransformer.py:157-165), dropout p on the probabilities.  qkv: (B*L, 3H) packed q | k | v -> (B*L, H)."""

    @staticmethod
    def forward(ctx, qkv, B: int, L: int, n_heads: int, p: float, seed: int, site: int):
        _need_cuda(qkv)
        qkv = qkv.contiguous().float()
        H = qkv.shape[1] // 3
        out = torch.empty(B * L, H, dtype=torch.float32, device=qkv.device)
        lse = torch.empty(B, n_heads, L, dtype=torch.float32, device=qkv.device)
        valid = torch.ones(B, L, dtype=torch.int64, device=qkv.device)
        _C.check(_C.lib().ndt1_attention_f32(qkv.data_ptr(), out.data_ptr(), None, lse.data_ptr(), valid.data_ptr(), B, L, H, n_heads, -2, -2,
                                             float(p), 0.0, seed, site, 0, None, None, None, _C.stream_ptr()), "ndt1_attention_f32")
        ctx.save_for_backward(qkv, out, lse, valid)
        ctx.dims = (B, L, H, n_heads, float(p), seed, site)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse, valid = ctx.saved_tensors
        B, L, H, n_heads, p, seed, site = ctx.dims
        dout = dout.contiguous().float()
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(B * n_heads * L + 16, dtype=torch.float32, device=qkv.device)
        _C.check(_C.lib().ndt1_attention_f32(qkv.data_ptr(), out.data_ptr(), None, lse.data_ptr(), valid.data_ptr(), B, L, H, n_heads, -2, -2,
                                             p, 0.0, seed, site, 0, dout.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), _C.stream_ptr()),
                 "ndt1_attention_f32 (backward)")
        return dqkv, None, None, None, None, None, None


class _AttentionMM(torch.autograd.Function):
    """The same operator on the tcgen05 GEMM (``ndt1_attention_mm_fwd`` / ``_bwd``): batched Q K^T, row softmax + dropout, P V
    and the five products of the backward, probabilities kept in bf16.  The bf16 mode's attention: 670 tokens with heads of 96
    are beyond the fused tensor-core kernels (256 tokens, head 128) and cost 12 ms per layer on the CUDA-core ones."""

    @staticmethod
    def forward(ctx, qkv, B: int, L: int, n_heads: int, p: float, seed: int, site: int):
        _need_cuda(qkv)
        qkv = qkv.contiguous().float()
        H = qkv.shape[1] // 3
        lib = _C.lib()
        out = torch.empty(B * L, H, dtype=torch.float32, device=qkv.device)
        saved = torch.empty(int(lib.ndt1_attention_mm_saved_bytes(B, L, H, n_heads, float(p))), dtype=torch.uint8, device=qkv.device)
        ws = _mm_workspace(int(lib.ndt1_attention_mm_workspace_bytes(B, L, H, n_heads)), qkv.device)
        _C.check(lib.ndt1_attention_mm_fwd(qkv.data_ptr(), out.data_ptr(), saved.data_ptr(), ws.data_ptr(), B, L, H, n_heads, float(p), seed, site,
                                           _C.stream_ptr()), "ndt1_attention_mm_fwd")
        ctx.save_for_backward(saved)
        ctx.dims = (B, L, H, n_heads, float(p), seed, site)
        return out

    @staticmethod
    def backward(ctx, dout):
        (saved,) = ctx.saved_tensors
        B, L, H, n_heads, p, seed, site = ctx.dims
        dout = dout.contiguous().float()
        lib = _C.lib()
        dqkv = torch.empty(B * L, 3 * H, dtype=torch.float32, device=dout.device)
        ws = _mm_workspace(int(lib.ndt1_attention_mm_workspace_bytes(B, L, H, n_heads)), dout.device)
        _C.check(lib.ndt1_attention_mm_bwd(dout.data_ptr(), saved.data_ptr(), ws.data_ptr(), dqkv.data_ptr(), B, L, H, n_heads, p, seed, site,
                                           _C.stream_ptr()), "ndt1_attention_mm_bwd")
        return dqkv, None, None, None, None, None, None


_MM_WS = {}


def _mm_workspace(nbytes: int, dev) -> torch.Tensor:
    """Scratch of one attention call, shared by all layers (calls on one stream run in order); grown on demand."""
    key = str(dev)
    t = _MM_WS.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _MM_WS[key] = t
    return t


class _ReconLoss(torch.autograd.Function):
    """sum over (b, t, n) of w * l(pred, target), w = tmask & pmask (models/itransformer.py:343-349, 357-362); also the count."""

    @staticmethod
    def forward(ctx, pred, target, tmask, pmask, kind: int):
        _need_cuda(pred)
        pred, target = pred.contiguous().float(), target.contiguous().float()
        B, T, N = pred.shape
        dpred = torch.empty_like(pred)
        loss = torch.zeros((), dtype=torch.float32, device=pred.device)
        count = torch.zeros((), dtype=torch.int64, device=pred.device)
        _C.check(_C.lib().ndt1_recon_loss(pred.data_ptr(), target.data_ptr(), dpred.data_ptr(), tmask.data_ptr(), pmask.data_ptr(), B, T, N,
                                          kind, 0, 0, loss.data_ptr(), count.data_ptr(), None, _C.stream_ptr()), "ndt1_recon_loss")
        ctx.save_for_backward(dpred)
        ctx.mark_non_differentiable(count)
        return loss, count

    @staticmethod
    def backward(ctx, dloss, _dcount):
        (dpred,) = ctx.saved_tensors
        return dpred * dloss, None, None, None, None


class _XentLoss(torch.autograd.Function):
    """nn.CrossEntropyLoss(reduction="none")(logits, labels).sum() (models/itransformer.py:300-301, 375)."""

    @staticmethod
    def forward(ctx, logits, labels):
        _need_cuda(logits)
        logits = logits.contiguous().float()
        B, V = logits.shape
        dlogits = torch.empty_like(logits)
        loss = torch.zeros((), dtype=torch.float32, device=logits.device)
        _C.check(_C.lib().ndt1_xent_loss(logits.data_ptr(), labels.contiguous().data_ptr(), dlogits.data_ptr(), loss.data_ptr(), B, V, None,
                                         _C.stream_ptr()), "ndt1_xent_loss")
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dlogits,) = ctx.saved_tensors
        return dlogits * dloss, None


class AverageTokens(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim


class iTransformerEncoder(nn.Module):
    """Parameter containers of models/itransformer.py:98-173 (same classes, same order: same init draws, same state_dict keys)
    and the forward of :175-209 on this library's kernels."""

    def __init__(self, config: DictConfig, use_cls: bool, precision: str):
        super().__init__()
        self.precision = precision
        self.mode = config.embedder.mode
        if self.mode != "mlp":
            raise NotImplementedError("iTransformer: only the `mlp` embedder is built (the shipped configuration); "
                                      f"got embedder.mode = {self.mode}")
        H = config.hidden_size
        self.act = config.activation
        act_cls = _ACT_MODULES[config.activation]
        # torchvision.ops.MLP(in, [H, H], activation_layer, bias, dropout) = Linear, act, Dropout, Linear, Dropout  (:108-116)
        self.embed = nn.Sequential(
            nn.Sequential(nn.Linear(config.embedder.max_n_bins, H, bias=config.bias), act_cls(), nn.Dropout(config.embedder.dropout),
                          nn.Linear(H, H, bias=config.bias), nn.Dropout(config.embedder.dropout)),
            nn.LayerNorm(H),
        )
        self.embed_channel = (config.max_n_channels != 0)
        if self.embed_channel:
            self.channel_embeddings = nn.Sequential(nn.Embedding(config.max_n_channels, H), nn.LayerNorm(H))
        self.embed_region = config.embed_region
        if self.embed_region:
            self.regions = config.regions
            self.region_to_indx = {r: i for i, r in enumerate(self.regions)}
            self.indx_to_region = {v: k for k, v in self.region_to_indx.items()}
            self.region_embeddings = nn.Sequential(nn.Embedding(len(self.region_to_indx), H), nn.LayerNorm(H))
        self.embed_depth = config.embed_depth
        if self.embed_depth:
            self.depth_embeddings = nn.Sequential(nn.Linear(1, H), act_cls(), nn.Linear(H, H), nn.LayerNorm(H))
        self.use_cls = use_cls
        if self.use_cls:
            self.cls_embed = nn.Embedding(1, H)
        self.p_embed = float(config.embedder.dropout)
        self.p = float(config.dropout)
        self.n_heads = config.n_heads
        self.hidden_size = H
        layer = nn.TransformerEncoderLayer(d_model=H, nhead=config.n_heads, dim_feedforward=4 * H, activation=act_cls(),
                                           dropout=config.dropout, batch_first=True)
        self.transformer = nn.TransformerEncoder(encoder_layer=layer, num_layers=config.n_layers, norm=nn.LayerNorm(H),
                                                 enable_nested_tensor=False)
        self._seed = 0
        self._site = 0

    # ---- small helpers over the autograd functions
    def _lin(self, x, lin: nn.Linear, act: str = "identity"):
        return _LinearAct.apply(x, lin.weight, lin.bias, act, self.precision)

    def _ln(self, x, ln: nn.LayerNorm):
        return _LayerNorm.apply(x, ln.weight, ln.bias)

    def _lin_drop(self, x, lin: nn.Linear, act: str, p: float):
        """Linear -> activation -> Dropout(p) in one GEMM (the site counter advances exactly as with a separate dropout)."""
        self._site += 1
        if not self.training or p <= 0.0:
            return _LinearAct.apply(x, lin.weight, lin.bias, act, self.precision)
        return _LinearActDrop.apply(x, lin.weight, lin.bias, act, self.precision, p, self._seed, self._site)

    def _drop(self, x, p: float):
        self._site += 1
        if not self.training or p <= 0.0:
            return x
        return _Dropout.apply(x, p, self._seed, self._site)

    def forward(self, spikes, spikes_timestamp=None, spikes_spacestamp=None, neuron_regions=None, neuron_depths=None):
        _need_cuda(spikes)
        B, T, N = spikes.shape
        H = self.hidden_size
        dev = spikes.device
        self._seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if self.training else 0      # one Philox key per forward
        self._site = 100
        mlp, ln0 = self.embed[0], self.embed[1]
        x = spikes.transpose(1, 2).reshape(B * N, T)                                          # (:183) one token per neuron
        x = self._lin_drop(x, mlp[0], self.act, self.p_embed)
        x = self._lin_drop(x, mlp[3], "identity", self.p_embed)
        tokens = self._ln(x, ln0)                                                             # (B*N, H)
        if self.embed_channel:                                                                # (:187-191)
            if spikes_spacestamp is None:
                ce = self._ln(self.channel_embeddings[0].weight[:N], self.channel_embeddings[1])
                tokens = (tokens.view(B, N, H) + ce.unsqueeze(0)).view(B * N, H)
            else:
                rows = self.channel_embeddings[0].weight[spikes_spacestamp.reshape(-1)]
                ce = self._ln(rows, self.channel_embeddings[1])
                tokens = tokens + (ce if spikes_spacestamp.dim() == 2 else ce.unsqueeze(0).expand(B, N, H).reshape(B * N, H))
        if self.embed_region:                                                                 # (:193-196)
            idx = torch.tensor([[self.region_to_indx[r] for r in row] for row in neuron_regions], dtype=torch.int64, device=dev)
            tokens = tokens + self._ln(self.region_embeddings[0].weight[idx.reshape(-1)], self.region_embeddings[1])
        if self.embed_depth:                                                                  # (:198-200)
            de = self.depth_embeddings
            d = self._lin(neuron_depths.reshape(B * N, 1).float(), de[0], self.act)
            tokens = tokens + self._ln(self._lin(d, de[2]), de[3])
        L = N
        tokens = tokens.view(B, N, H)
        if self.use_cls:                                                                      # (:203-205)
            tokens = torch.cat((self.cls_embed.weight[0].expand(B, 1, H), tokens), dim=1)
            L = N + 1
        x = self._drop(tokens.reshape(B * L, H), self.p_embed)
        for layer in self.transformer.layers:                                                 # post-LN encoder layers (:157-173)
            qkv = _LinearAct.apply(x, layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias, "identity", self.precision)
            self._site += 1
            attn = _AttentionMM if self.precision == "bf16" else _Attention
            att = attn.apply(qkv, B, L, self.n_heads, self.p if self.training else 0.0, self._seed, self._site)
            x = self._ln(x + self._lin_drop(att, layer.self_attn.out_proj, "identity", self.p), layer.norm1)
            h = self._lin_drop(x, layer.linear1, self.act, self.p)
            x = self._ln(x + self._lin_drop(h, layer.linear2, "identity", self.p), layer.norm2)
        x = self._ln(x, self.transformer.norm)
        return x.view(B, L, H)


class iTransformer(nn.Module):

    def __init__(self, config: DictConfig, precision: Optional[str] = None, **kwargs):
        super().__init__()
        self.method = kwargs["method_name"]
        self.precision = precision or os.environ.get("NDT1_PRECISION", "bf16")
        config = update_config(DEFAULT_CONFIG, config)
        encoder_pt_path = config["encoder"].pop("from_pt", None)
        if encoder_pt_path is not None:
            config["encoder"] = update_config(config.encoder, torch.load(os.path.join(encoder_pt_path, "encoder_config.pth")))
        decoder_pt_path = config["decoder"].pop("from_pt", None)
        if decoder_pt_path is not None:
            config["decoder"] = update_config(config.decoder, torch.load(os.path.join(decoder_pt_path, "decoder_config.pth")))

        self.masker = nn.ModuleDict({k: Masker(DictConfig(m)) for k, m in config.masker.items()})
        self.encoder = iTransformerEncoder(config.encoder, config.decoder.use_cls, self.precision)
        if encoder_pt_path is not None:
            self.encoder.load_state_dict(torch.load(os.path.join(encoder_pt_path, "encoder.bin")))

        H = config.encoder.hidden_size
        if self.method == "mlm":
            n_outputs = config.encoder.embedder.max_n_bins
        elif self.method == "dyn_behaviour":
            n_outputs = config.encoder.embedder.max_n_bins
        elif self.method == "stat_behaviour":
            if kwargs["loss"] == "xent":
                n_outputs = kwargs["n_labels"]
            elif kwargs["loss"] == "mse":
                n_outputs = 1
            else:
                raise Exception(f"Loss {kwargs['loss']} not implemented yet for stat_behaviour")
        elif self.method == "ctc":
            raise NotImplementedError("iTransformer: method ctc is not built (mlm, dyn_behaviour and stat_behaviour are)")
        else:
            raise Exception(f"Method {self.method} not implemented")
        layers = []
        self.use_cls = config.decoder.use_cls
        if self.method in ["ctc", "dyn_behaviour", "stat_behaviour"] and not self.use_cls:
            layers.append(AverageTokens(dim=1))
        self.mlp_decoder = bool(config.decoder.mlp_decoder)
        self.dec_act = config.decoder.activation
        if self.mlp_decoder:
            layers.append(nn.Linear(H, H))
            layers.append(_ACT_MODULES[config.decoder.activation]())
        layers.append(nn.Linear(H, n_outputs))
        self.log_input = bool(kwargs.get("log_input", True))
        if self.method == "mlm" and not self.log_input:
            layers.append(nn.ReLU())
        self.decoder = nn.Sequential(*layers)
        if decoder_pt_path is not None:
            self.decoder.load_state_dict(torch.load(os.path.join(decoder_pt_path, "decoder.bin")))
        if self.method == "mlm":
            self.loss_name = kwargs["loss"]
            if self.loss_name not in ("poisson_nll", "mse"):
                raise Exception(f"Loss {kwargs['loss']} not implemented yet for mlm")
        elif self.method == "stat_behaviour":
            self.loss_name = kwargs["loss"]
        self.config = config

    def _decode(self, x2d):
        lins = [m for m in self.decoder if isinstance(m, nn.Linear)]
        if self.mlp_decoder:
            x2d = _LinearAct.apply(x2d, lins[0].weight, lins[0].bias, self.dec_act, self.precision)
        relu_out = self.method == "mlm" and not self.log_input
        return _LinearAct.apply(x2d, lins[-1].weight, lins[-1].bias, "relu" if relu_out else "identity", self.precision)

    def forward(self, spikes, spikes_mask, spikes_timestamp, spikes_spacestamp=None, spikes_lengths=None, targets=None,
                targets_lengths=None, neuron_regions=None, neuron_depths=None) -> iTransformerOutput:
        _need_cuda(spikes)
        if self.method == "mlm":
            targets = spikes.clone()
        spikes = spikes.clone()                                     # (the reference's maskers write into the caller's tensor)
        targets_mask = torch.zeros_like(spikes, dtype=torch.int64)
        for masker in self.masker.values():
            spikes, new_mask = masker(spikes, neuron_regions)
            targets_mask = targets_mask | new_mask
        x = self.encoder(spikes, spikes_timestamp, spikes_spacestamp, neuron_regions=neuron_regions, neuron_depths=neuron_depths)
        B, L, H = x.shape
        if self.use_cls:
            x = x[:, 1:, :] if self.method == "mlm" else x[:, 0, :]
        elif self.method != "mlm":
            x = x.sum(dim=1)                                        # AverageTokens (:29-37)
        lead = x.shape[:-1]
        preds = self._decode(x.reshape(-1, H)).view(*lead, -1)
        if self.method == "mlm":
            preds = preds.transpose(1, 2).contiguous()              # (B, T, N)
            targets_mask = targets_mask & spikes_mask.unsqueeze(2)
            kind = (_C.LOSS_POISSON_LOG if self.log_input else _C.LOSS_POISSON_RATE) if self.loss_name == "poisson_nll" else _C.LOSS_MSE
            loss, n = _ReconLoss.apply(preds, targets, targets_mask.contiguous(), spikes_mask.contiguous(), kind)
            return iTransformerOutput(loss=loss, n_examples=n, preds=preds, targets=targets, mask=targets_mask)
        if self.method == "stat_behaviour":                         # (:371-385) one label / value per trial from the cls token (or token sum)
            targets_mask = targets_mask & spikes_mask.unsqueeze(2)
            if self.loss_name == "xent":
                loss = _XentLoss.apply(preds, targets.long().squeeze(1))
            else:
                ones = torch.ones(preds.shape[0], 1, 1, dtype=torch.int64, device=preds.device)
                loss, _ = _ReconLoss.apply(preds.reshape(-1, 1, 1), targets.float().reshape(-1, 1, 1), ones, ones[:, :, 0].contiguous(), _C.LOSS_MSE)
            n = torch.tensor(len(targets), device=loss.device, dtype=torch.long)
            return iTransformerOutput(loss=loss, n_examples=n, preds=preds, targets=targets, mask=targets_mask)
        # dyn_behaviour (:357-369): one value per bin from the cls token (or the token sum), MSE over the bins that are not padding
        ones = torch.ones(preds.shape[0], preds.shape[1], 1, dtype=torch.int64, device=preds.device)
        loss, n = _ReconLoss.apply(preds.unsqueeze(2), targets.float().unsqueeze(2), ones, spikes_mask.contiguous(), _C.LOSS_MSE)
        return iTransformerOutput(loss=loss, n_examples=n, preds=preds, targets=targets, mask=spikes_mask)

    def save_checkpoint(self, save_dir):
        torch.save(self.encoder.state_dict(), os.path.join(save_dir, "encoder.bin"))
        torch.save(dict(self.config.encoder), os.path.join(save_dir, "encoder_config.pth"))
        torch.save(self.decoder.state_dict(), os.path.join(save_dir, "decoder.bin"))
        torch.save(dict(self.config.decoder), os.path.join(save_dir, "decoder_config.pth"))

    def load_checkpoint(self, load_dir):
        self.encoder.load_state_dict(torch.load(os.path.join(load_dir, "encoder.bin")))
        self.decoder.load_state_dict(torch.load(os.path.join(load_dir, "decoder.bin")))
