"""CPU oracle for the iTransformer path.  TEST INFRASTRUCTURE ONLY -- see oracle/ndt1_oracle.py for the rules:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product path never does.

Functional restatement on torch-CPU of ``iTransformer.forward`` (models/itransformer.py:312-385) and everything under it:
  * the ``mlp`` embedder -- torchvision ``MLP(max_n_bins -> H -> H)`` = Linear, act, Dropout, Linear, Dropout, then LayerNorm
    (:108-118) -- over the spikes transposed to (batch, channels, bins) (:182-183),
  * channel / region / depth embeddings, each through its own LayerNorm (:126-150, 187-201), cls token (:203-205), dropout,
  * ``nn.TransformerEncoder`` of POST-LN layers (``norm_first`` = False): x = LN1(x + drop(MHA(x))), x = LN2(x + drop(W2 drop(act(W1 x)))),
    packed ``in_proj`` = q | k | v, no attention mask, dropout on the attention probabilities, final LayerNorm (:157-173, 207),
  * the decoder (:249-272) and the four losses (:330-385); ``mlm``, ``dyn_behaviour`` and ``stat_behaviour`` are restated here.
Pinned against outputs of the unmodified reference run in the build container (tests/golden/make_golden.py::itransformer_cases ->
tests/golden/itransformer_small.npz, itransformer_config3.npz; the shipped yaml needs the masker keys `active` / `regions`,
SURVEY.md component 8 -- a config-only fix).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

_ACTS = {"relu": F.relu, "gelu": F.gelu, "softsign": F.softsign}


def _drop(x, site, drop_scales):
    """Dropout with an injected keep-scale tensor (0 or 1/(1-p)); absent -> identity (eval, or p = 0)."""
    if drop_scales is None or site not in drop_scales:
        return x
    return x * drop_scales[site].to(x.dtype).reshape(x.shape)


def encoder_layer(p: Dict[str, torch.Tensor], pre: str, x: torch.Tensor, n_heads: int, act, drop_scales=None, li: int = 0):
    """``nn.TransformerEncoderLayer.forward`` with norm_first=False, batch_first=True, no masks (torch/nn/modules/transformer.py;
    built at models/itransformer.py:157-165)."""
    B, L, H = x.shape
    hd = H // n_heads
    qkv = F.linear(x, p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.in_proj_bias"])
    q, k, v = [t.view(B, L, n_heads, hd).transpose(1, 2) for t in qkv.split(H, dim=-1)]
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
    a = _drop(a, f"attn{li}", drop_scales)
    o = (a @ v).transpose(1, 2).reshape(B, L, H)
    o = F.linear(o, p[pre + "self_attn.out_proj.weight"], p[pre + "self_attn.out_proj.bias"])
    x = F.layer_norm(x + _drop(o, f"res1_{li}", drop_scales), (H,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
    h = act(F.linear(x, p[pre + "linear1.weight"], p[pre + "linear1.bias"]))
    h = F.linear(_drop(h, f"ffn{li}", drop_scales), p[pre + "linear2.weight"], p[pre + "linear2.bias"])
    return F.layer_norm(x + _drop(h, f"res2_{li}", drop_scales), (H,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)


def encoder_forward(p: Dict[str, torch.Tensor], cfg: dict, use_cls: bool, spikes: torch.Tensor, spikes_spacestamp=None,
                    region_indx=None, neuron_depths=None, drop_scales=None) -> torch.Tensor:
    """``iTransformerEncoder.forward`` (models/itransformer.py:175-209), ``mlp`` embedder.  `region_indx`: (B, N) int64, the
    host-side mapping of the reference's region names (:193-194) already applied."""
    assert cfg["embedder"]["mode"] == "mlp", "oracle restates the mlp embedder (the shipped configuration)"
    act = _ACTS[cfg["activation"]]
    H = cfg["hidden_size"]
    x = spikes.transpose(1, 2)                                                             # (B, N, T)  :183
    x = act(F.linear(x, p["encoder.embed.0.0.weight"], p.get("encoder.embed.0.0.bias")))
    x = _drop(x, "embed_mlp0", drop_scales)
    x = _drop(F.linear(x, p["encoder.embed.0.3.weight"], p.get("encoder.embed.0.3.bias")), "embed_mlp1", drop_scales)
    tokens = F.layer_norm(x, (H,), p["encoder.embed.1.weight"], p["encoder.embed.1.bias"], 1e-5)
    if cfg["max_n_channels"] != 0:                                                         # :187-191
        if spikes_spacestamp is None:
            spikes_spacestamp = torch.arange(tokens.shape[1])
        ce = F.layer_norm(p["encoder.channel_embeddings.0.weight"][spikes_spacestamp], (H,), p["encoder.channel_embeddings.1.weight"],
                          p["encoder.channel_embeddings.1.bias"], 1e-5)
        tokens = tokens + ce
    if cfg["embed_region"]:                                                                # :193-196
        re = F.layer_norm(p["encoder.region_embeddings.0.weight"][region_indx], (H,), p["encoder.region_embeddings.1.weight"],
                          p["encoder.region_embeddings.1.bias"], 1e-5)
        tokens = tokens + re
    if cfg["embed_depth"]:                                                                 # :198-200
        d = act(F.linear(neuron_depths.unsqueeze(2), p["encoder.depth_embeddings.0.weight"], p["encoder.depth_embeddings.0.bias"]))
        d = F.linear(d, p["encoder.depth_embeddings.2.weight"], p["encoder.depth_embeddings.2.bias"])
        tokens = tokens + F.layer_norm(d, (H,), p["encoder.depth_embeddings.3.weight"], p["encoder.depth_embeddings.3.bias"], 1e-5)
    if use_cls:                                                                            # :203-205
        tokens = torch.cat((p["encoder.cls_embed.weight"][0].expand(tokens.shape[0], 1, H), tokens), dim=1)
    x = _drop(tokens, "embed", drop_scales)
    for li in range(cfg["n_layers"]):
        x = encoder_layer(p, f"encoder.transformer.layers.{li}.", x, cfg["n_heads"], act, drop_scales, li)
    return F.layer_norm(x, (H,), p["encoder.transformer.norm.weight"], p["encoder.transformer.norm.bias"], 1e-5)


def decoder_forward(p: Dict[str, torch.Tensor], cfg: dict, method: str, x: torch.Tensor, log_input: bool = True) -> torch.Tensor:
    """models/itransformer.py:249-272: [AverageTokens] -> [Linear + act] -> Linear -> [ReLU for mlm on rates]."""
    dc = cfg["decoder"]
    if method in ("ctc", "dyn_behaviour", "stat_behaviour") and not dc["use_cls"]:
        x = x.sum(dim=1)
    last = "decoder.0."
    keys = sorted({k.split(".")[1] for k in p if k.startswith("decoder.")}, key=int)
    if dc["mlp_decoder"]:
        x = _ACTS[dc["activation"]](F.linear(x, p[f"decoder.{keys[0]}.weight"], p[f"decoder.{keys[0]}.bias"]))
    last = f"decoder.{keys[-1]}."
    x = F.linear(x, p[last + "weight"], p[last + "bias"])
    if method == "mlm" and not log_input:
        x = F.relu(x)
    return x


def itransformer_forward(p: Dict[str, torch.Tensor], cfg: dict, method_kwargs: dict, batch: dict, masked: Optional[dict] = None,
                         drop_scales=None):
    """``iTransformer.forward`` (:312-385) for ``mlm`` and ``dyn_behaviour``.  `masked` = {"spikes", "mask"}: the maskers' result
    (models/masker.py; restated in ndt1_oracle.masker_apply), applied by the caller so that the draws are explicit.
    Returns (loss, n_examples, preds, mask)."""
    method = method_kwargs["method_name"]
    spikes, smask = batch["spikes"], batch["spikes_mask"]
    targets = spikes.clone() if method == "mlm" else batch.get("targets")
    tmask = torch.zeros_like(spikes, dtype=torch.int64)
    if masked is not None:
        spikes, tmask = masked["spikes"], masked["mask"]
    use_cls = cfg["decoder"]["use_cls"]
    x = encoder_forward(p, cfg["encoder"], use_cls, spikes, batch.get("spikes_spacestamp"), batch.get("region_indx"),
                        batch.get("neuron_depths"), drop_scales)
    if use_cls:
        x = x[:, 1:, :] if method == "mlm" else x[:, 0, :]                                 # :334-338
    preds = decoder_forward(p, cfg, method, x, method_kwargs.get("log_input", True))
    if method == "mlm":
        preds = preds.transpose(1, 2)                                                      # (B, T, N)  :344
        tmask = tmask & smask.unsqueeze(2)
        if method_kwargs.get("loss", "poisson_nll") == "poisson_nll":
            el = F.poisson_nll_loss(preds, targets, log_input=method_kwargs.get("log_input", True), reduction="none")
        else:
            el = F.mse_loss(preds, targets, reduction="none")
        return (el * tmask).sum(), tmask.sum(), preds, tmask
    if method == "dyn_behaviour":                                                          # :357-369
        el = F.mse_loss(preds, targets, reduction="none")
        return (el * smask).sum(), smask.sum(), preds, smask
    if method == "stat_behaviour":                                                         # :371-385
        tmask = tmask & smask.unsqueeze(2)
        if method_kwargs["loss"] == "xent":
            loss = F.cross_entropy(preds, targets.long().squeeze(1), reduction="none").sum()
        else:
            loss = F.mse_loss(preds.squeeze(1), targets.squeeze(1), reduction="none").sum()
        return loss, torch.tensor(len(targets)), preds, tmask
    raise NotImplementedError(f"oracle: method {method} not restated")
