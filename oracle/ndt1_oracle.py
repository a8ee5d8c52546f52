"""CPU oracle for the NDT1 hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU, the algorithm of the reference's NDT1
encoder forward/backward (colehurwitz/llm_bci).  It exists so that the CUDA
path in ``llm_bci_b200`` can be checked without the reference being present
(``/root/reference`` does not exist on the GPU box).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; the product never does.

Pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the *unmodified reference module run
in the build container*: ``tests/golden/make_golden.py`` imports
``/root/reference/models/ndt1.py`` and ``models/masker.py`` and
``data_utils/datasets.py``, runs them under fixed seeds and commits inputs and
outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays
them through this file.

Integer / byte / index work (masker, collate, context band, greedy CTC
collapse, stacked lengths) is numpy and bit-exact.  Floating-point work is
plain torch fp32/fp64 on the CPU with autograd for the backward; an
independent numpy float64 alpha/beta CTC (``ctc_loss_np``) cross-checks the
loss and its gradient.

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# integer / index pieces (numpy, bit exact)
# --------------------------------------------------------------------------


def context_band(context_forward: int, context_backward: int, max_f: int) -> np.ndarray:
    """Band matrix of allowed (query i, key j) pairs.  models/ndt1.py:30-41.

    -2 on both sides -> all ones; a single -2 -> unbounded on that side;
    -1 excludes the diagonal on that side.
    """
    if context_forward == -2 and context_backward == -2:
        return np.ones((max_f, max_f), dtype=np.int64)
    cf = context_forward if context_forward >= -1 else max_f
    cb = context_backward if context_backward >= -1 else max_f
    i = np.arange(max_f)[:, None]
    j = np.arange(max_f)[None, :]
    return ((j <= i + cf) & (j >= i - cb)).astype(np.int64)


def attention_allowed(band: np.ndarray, key_valid: np.ndarray) -> np.ndarray:
    """allowed[b,i,j] = (i==j) | (band[i,j] & key_valid[b,j]).  models/ndt1.py:435-437
    (Python precedence: ``&`` binds before ``|``)."""
    L = key_valid.shape[1]
    eye = np.eye(L, dtype=np.int64)[None]
    return eye | (band[None, :L, :L] & key_valid[:, None, :].astype(np.int64))


def stacked_lengths(lens: np.ndarray, stack: bool, size: int, stride: int) -> np.ndarray:
    """models/ndt1.py:207-208: ``(1 + (lens - size) / stride).to(int)`` (true division,
    truncation toward zero on the cast)."""
    lens = np.asarray(lens)
    if not stack:
        return lens
    return np.trunc(1 + (lens.astype(np.float64) - size) / stride).astype(lens.dtype)


def stacked_mask(mask: np.ndarray, size: int, stride: int) -> np.ndarray:
    """Window AND of the padding mask, models/ndt1.py:182-183."""
    B, T = mask.shape
    Tp = (T - size) // stride + 1
    out = np.ones((B, Tp), dtype=mask.dtype)
    for r in range(Tp):
        out[:, r] = mask[:, r * stride:r * stride + size].prod(axis=1)
    return out


def expand_timesteps(mask: np.ndarray, width: int) -> np.ndarray:
    """Temporal mask dilation, models/masker.py:106-110: a ones-kernel 'same' conv,
    i.e. out[t] = OR_{k<width} m[t - (width-1)//2 + k]."""
    B, T = mask.shape
    left = (width - 1) // 2
    out = np.zeros((B, T), dtype=bool)
    m = mask.astype(bool)
    for k in range(width):
        off = k - left
        lo, hi = max(0, -off), min(T, T - off)
        if lo < hi:
            out[:, lo:hi] |= m[:, lo + off:hi + off]
    return out


def masker_apply(
    spikes: np.ndarray,        # (B,T,N) float32
    mode: str,                 # temporal | neuron | random | region | co-smooth
    mask_draw: np.ndarray,     # Bernoulli draw in the mode's own shape (0/1)
    zero_draw: np.ndarray,     # (B,T,N) Bernoulli(zero_ratio) draw
    random_draw: np.ndarray,   # (B,T,N) Bernoulli(random_ratio) draw
    rand: np.ndarray,          # (B,T,N) uniform [0,1) float32
    timespan: int = 1,
) -> Tuple[np.ndarray, np.ndarray]:
    """models/masker.py:44-104 given the random draws.  Returns (spikes', mask int64).

    The replacement value is ``max(spikes after zeroing) * rand`` with the max
    over the WHOLE batch tensor (masker.py:100-102); arithmetic is float32.
    """
    spikes = np.array(spikes, dtype=np.float32, copy=True)
    B, T, N = spikes.shape
    if mode == "temporal":
        m = mask_draw.astype(bool)
        if timespan > 1:
            m = expand_timesteps(m, timespan)
        mask = np.broadcast_to(m[:, :, None], (B, T, N))
    elif mode in ("neuron", "region"):
        mask = np.broadcast_to(mask_draw.astype(bool)[:, None, :], (B, T, N))
    elif mode == "co-smooth":
        mask = np.broadcast_to(mask_draw.astype(bool)[None, None, :], (B, T, N))
    elif mode == "random":
        mask = mask_draw.astype(bool)
    else:
        raise ValueError(f"masking mode {mode} not implemented")
    zero_idx = zero_draw.astype(bool) & mask
    spikes[zero_idx] = 0
    random_idx = random_draw.astype(bool) & mask & ~zero_idx
    mx = np.float32(spikes.max())
    repl = (mx * rand.astype(np.float32)).astype(np.float32)
    spikes[random_idx] = repl[random_idx]
    return spikes, mask.astype(np.int64)


def format_ctc(pred_ids: Sequence[int], blank_id: int) -> List[int]:
    """Greedy collapse of utils/eval_bci.py:41-48 on ids: emit when the id differs
    from the last EMITTED id and is not blank (``last`` only moves on emission)."""
    out: List[int] = []
    last = -1
    for idx in pred_ids:
        idx = int(idx)
        if idx != last and idx != blank_id:
            out.append(idx)
            last = idx
    return out


def edit_distance(source: Sequence, target: Sequence) -> int:
    """editdistance.eval (third-party `editdistance`, unpinned in the reference; utils/eval_bci.py:11-14 calls it on word
    lists): the classic Levenshtein distance, unit costs for insert / delete / substitute."""
    prev = list(range(len(target) + 1))
    for i, a in enumerate(source, 1):
        cur = [i]
        for j, b in enumerate(target, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a != b)))
        prev = cur
    return prev[-1]


def word_error_count(preds: Sequence[str], targets: Sequence[str]):
    """utils/eval_bci.py:19-36: summed word edit distance and summed target word count of space-joined strings."""
    errors = words = 0
    for p, t in zip(preds, targets):
        ps, ts = p.split(" "), t.split(" ")
        errors += edit_distance(ps, ts)
        words += len(ts)
    return errors, words


def padded_array(arrays, dim=0, side="right", value=0, truncate=None, min_length=None) -> np.ndarray:
    """data_utils/datasets.py:191-221."""
    max_size = max(a.shape[dim] for a in arrays)
    if truncate is None:
        truncate = max_size
    if min_length is None:
        min_length = 0
    assert min_length <= truncate, "Can't truncate below the minimum length"
    pad_size = min(truncate, max(max_size, min_length))
    if side not in ("left", "right"):
        raise Exception(f' "side" can only take values "right" or "left", got {side}')
    out = []
    for a in arrays:
        n = max(0, pad_size - a.shape[dim])
        widths = [(0, 0)] * a.ndim
        widths[dim] = (n, 0) if side == "left" else (0, n)
        p = np.pad(a, widths, mode="constant", constant_values=value)
        sl = [slice(None)] * a.ndim
        sl[dim] = slice(0, truncate)
        out.append(p[tuple(sl)])
    return np.stack(out, axis=0)


def pad_collate(batch, model_inputs, pad_dict):
    """data_utils/datasets.py:236-272 with numpy outputs (tensors in the reference)."""
    if isinstance(batch[0], list):
        batch = [row for sub in batch for row in sub]
    keys = list(batch[0].keys())
    array_keys = [k for k in keys if isinstance(batch[0][k], np.ndarray) and batch[0][k].dtype.type != np.str_]
    string_keys = [k for k in keys if isinstance(batch[0][k], np.ndarray) and batch[0][k].dtype.type == np.str_]
    assert set(pad_dict.keys()).issubset(array_keys), "Can't pad keys which are not arrays"
    padded, unused = {}, {}
    for k in keys:
        if k in array_keys:
            if k in pad_dict:
                v = padded_array([r[k] for r in batch], **pad_dict[k])
            elif len(set(r[k].shape for r in batch)) == 1:
                v = np.stack([r[k] for r in batch], axis=0)
            else:
                v = [r[k] for r in batch]
        elif k in string_keys:
            v = np.stack([r[k] for r in batch], axis=0)
        else:
            v = [r[k] for r in batch]
        (padded if k in model_inputs else unused)[k] = v
    return padded, unused


# --------------------------------------------------------------------------
# CTC in numpy float64 (independent of torch's kernel)
# --------------------------------------------------------------------------


def _lse(a, b):
    if a == -np.inf:
        return b
    if b == -np.inf:
        return a
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def ctc_loss_np(log_probs: np.ndarray, targets: np.ndarray, input_length: int, target_length: int, blank: int = 0,
                zero_infinity: bool = True) -> Tuple[float, np.ndarray]:
    """One trial.  log_probs (T,V) are log-softmax outputs.  Returns (nll, d nll / d logits)
    with the conventions of torch's CTCLoss(reduction='none') followed by the
    log-softmax backward: grad = softmax - posterior for t < input_length,
    0 afterwards; infeasible -> (0, 0) when zero_infinity.  models/ndt1.py:517,581.
    """
    T, V = log_probs.shape
    lp = log_probs.astype(np.float64)
    S = int(target_length)
    L = 2 * S + 1
    ext = np.full(L, blank, dtype=np.int64)
    ext[1::2] = targets[:S]
    Tn = int(input_length)
    grad = np.zeros((T, V), dtype=np.float64)
    if Tn == 0:
        nll = 0.0 if S == 0 else np.inf
        if nll == np.inf and zero_infinity:
            nll = 0.0
        return nll, grad
    NEG = -np.inf
    alpha = np.full((Tn, L), NEG)
    alpha[0, 0] = lp[0, ext[0]]
    if L > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, Tn):
        for s in range(L):
            a = alpha[t - 1, s]
            if s >= 1:
                a = _lse(a, alpha[t - 1, s - 1])
            if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                a = _lse(a, alpha[t - 1, s - 2])
            alpha[t, s] = a + lp[t, ext[s]] if a != NEG else NEG
    ll = alpha[Tn - 1, L - 1]
    if L > 1:
        ll = _lse(ll, alpha[Tn - 1, L - 2])
    nll = -ll
    if not np.isfinite(nll):
        return (0.0 if zero_infinity else np.inf), grad
    beta = np.full((Tn, L), NEG)
    beta[Tn - 1, L - 1] = lp[Tn - 1, ext[L - 1]]
    if L > 1:
        beta[Tn - 1, L - 2] = lp[Tn - 1, ext[L - 2]]
    for t in range(Tn - 2, -1, -1):
        for s in range(L):
            b = beta[t + 1, s]
            if s + 1 < L:
                b = _lse(b, beta[t + 1, s + 1])
            if s + 2 < L and ext[s + 2] != blank and ext[s + 2] != ext[s]:
                b = _lse(b, beta[t + 1, s + 2])
            beta[t, s] = b + lp[t, ext[s]] if b != NEG else NEG
    for t in range(Tn):
        post = np.zeros(V)
        for s in range(L):
            ab = alpha[t, s] + beta[t, s]
            if ab != NEG:
                post[ext[s]] += math.exp(ab - lp[t, ext[s]] - ll)
        grad[t] = np.exp(lp[t]) - post
    return float(nll), grad


# --------------------------------------------------------------------------
# floating-point model (torch CPU, autograd backward)
# --------------------------------------------------------------------------


def gaussian_kernel(smooth_sd: int) -> np.ndarray:
    """models/ndt1.py:87-88: scipy.signal gaussian window of 1+6*sd points, std sd,
    built in float64 and normalised to unit sum."""
    m = 1 + smooth_sd * 6
    n = np.arange(m, dtype=np.float64) - (m - 1) / 2.0
    w = np.exp(-0.5 * (n / float(smooth_sd)) ** 2)
    return w / w.sum()


def smooth_and_noise(spikes: torch.Tensor, cfg: dict, training: bool, noise: Optional[dict]) -> torch.Tensor:
    """models/ndt1.py:92-107.  ``noise`` carries the injected draws
    {"white": (B,T,N), "offset": (B,1,N)} standing in for torch.randn."""
    B, T, N = spikes.shape
    if cfg["smooth_sd"] is not None:
        k = torch.from_numpy(gaussian_kernel(cfg["smooth_sd"])).to(spikes.dtype)
        w = k.view(1, 1, -1).expand(N, 1, k.numel())
        spikes = F.conv1d(spikes.transpose(1, 2), w, padding="same", groups=N).transpose(1, 2)
    if cfg["noise"] and training and noise is not None:
        if cfg["white_noise_sd"] is not None:
            spikes = spikes + cfg["white_noise_sd"] * noise["white"].to(spikes.dtype)
        if cfg["constant_offset_sd"] is not None:
            spikes = spikes + cfg["constant_offset_sd"] * noise["offset"].to(spikes.dtype)
    return spikes


_ACTS: Dict[str, Callable[[torch.Tensor], torch.Tensor]] = {
    "softsign": lambda x: x / (1 + x.abs()),
    "gelu": lambda x: F.gelu(x),  # erf form (HF GELUActivation -> F.gelu)
    "relu": lambda x: F.relu(x),
    "identity": lambda x: x,
}


def _drop(x: torch.Tensor, site: str, drop_scales: Optional[dict]) -> torch.Tensor:
    """Dropout with an injected keep-scale tensor (0 or 1/(1-p)); absent -> identity."""
    if drop_scales is None:
        return x
    if "torch_dropout" in drop_scales:     # CPU-baseline mode: the reference's own nn.Dropout / SDPA dropout cost
        td = drop_scales["torch_dropout"]
        p = td.get("factors", 0.0) if site == "factors" else td["embed" if site == "embed" else "transformer"]
        return F.dropout(x, p, True) if p > 0 else x
    if site not in drop_scales:
        return x
    return x * drop_scales[site].to(x.dtype).reshape(x.shape)


def encoder_forward(
    params: Dict[str, torch.Tensor],
    enc_cfg: dict,
    spikes: torch.Tensor,
    spikes_mask: torch.Tensor,
    spikes_timestamp: torch.Tensor,
    block_idx: Optional[torch.Tensor] = None,
    day_idx: Optional[torch.Tensor] = None,
    training: bool = False,
    noise: Optional[dict] = None,
    masker_draws: Optional[List[dict]] = None,
    drop_scales: Optional[dict] = None,
    prefix: str = "encoder.",
):
    """NeuralEncoder.forward, models/ndt1.py:408-450, composed of
    SmoothAndNoise (:92-107), Masker (masker.py:44-104), NeuralEmbeddingLayer (:160-203),
    mask assembly (:435-437), NeuralEncoderLayer x L (:317-330) with NeuralAttention
    (:266-292) and NeuralMLP (:224-227), out_norm (:442) and NeuralFactorsProjection (:372-373).
    ``params`` uses the reference state_dict keys.
    """
    P = lambda k: params[prefix + k]
    emb, tr = enc_cfg["embedder"], enc_cfg["transformer"]
    B, T, N = spikes.shape
    x = smooth_and_noise(spikes, enc_cfg["smooth_and_noise"], training, noise)

    targets_mask = torch.zeros(B, T, N, dtype=torch.int64)
    mk_cfgs = list(enc_cfg["masker"].values())
    for i, mcfg in enumerate(mk_cfgs):
        active = mcfg["active"] and (training or mcfg.get("force_active", False))
        if not active:
            continue
        d = masker_draws[i]
        xs, m = masker_apply(x.detach().numpy().astype(np.float32), mcfg["mode"], d["mask"], d["zero"], d["random"], d["rand"],
                             d.get("timespan", 1))
        # masking is a data-dependent overwrite: no gradient flows to replaced bins
        keep = torch.from_numpy((xs == x.detach().numpy().astype(np.float32))).to(x.dtype)
        x = x * keep + torch.from_numpy(xs).to(x.dtype) * (1 - keep)
        targets_mask = targets_mask | torch.from_numpy(m)

    # --- embedding (models/ndt1.py:160-203)
    if emb["adapt"]:
        h = torch.stack([F.linear(f, P(f"embedder.embed_spikes.{int(day_idx[i])}.weight"),
                                  P(f"embedder.embed_spikes.{int(day_idx[i])}.bias") if emb["bias"] else None)
                         for i, f in enumerate(x)], 0)
    else:
        h = F.linear(x, P("embedder.embed_spikes.weight"), P("embedder.embed_spikes.bias") if emb["bias"] else None)
    h = _ACTS[emb["act"]](h)
    mask = spikes_mask
    ts = spikes_timestamp
    if emb["stack"]["active"]:
        S, st = emb["stack"]["size"], emb["stack"]["stride"]
        Tp = (T - S) // st + 1
        D = h.shape[-1]
        win = torch.stack([h[:, r * st:r * st + S, :].reshape(B, S * D) for r in range(Tp)], 1)  # time-major flatten
        h = F.linear(win, P("embedder.stack_projection.weight"), P("embedder.stack_projection.bias"))
        ts = ts[:, :Tp]
        mask = torch.from_numpy(stacked_mask(mask.numpy(), S, st))
    else:
        h = F.linear(h, P("embedder.projection.weight"), P("embedder.projection.bias"))
    if emb["pos"]:
        h = h + P("embedder.embed_pos.weight")[ts]
    n_prefix = 0
    if emb["block_token"]:
        h = torch.cat((P("embedder.block_embedding.weight")[block_idx].unsqueeze(1), h), 1)
        mask = torch.cat((torch.ones_like(mask[:, :1]), mask), 1)
        n_prefix += 1
    if emb["day_token"]:
        h = torch.cat((P("embedder.day_embedding.weight")[day_idx].unsqueeze(1), h), 1)
        mask = torch.cat((torch.ones_like(mask[:, :1]), mask), 1)
        n_prefix += 1
    h = _drop(h, "embed", drop_scales)

    # --- attention mask (models/ndt1.py:30-41, 435-437)
    L = h.shape[1]
    band = context_band(enc_cfg["context"]["forward"], enc_cfg["context"]["backward"], emb["max_F"])
    allowed = torch.from_numpy(attention_allowed(band, mask.numpy())).bool()  # (B,L,L)

    H = tr["hidden_size"]
    nh = tr["n_heads"]
    hd = H // nh
    for li in range(tr["n_layers"]):
        lp = f"layers.{li}."
        a_in = F.layer_norm(h, (H,), P(lp + "ln1.weight"), P(lp + "ln1.bias"), 1e-5)
        ab = tr["attention_bias"]
        q = F.linear(a_in, P(lp + "attn.query.weight"), P(lp + "attn.query.bias") if ab else None).view(B, L, nh, hd).transpose(1, 2)
        k = F.linear(a_in, P(lp + "attn.key.weight"), P(lp + "attn.key.bias") if ab else None).view(B, L, nh, hd).transpose(1, 2)
        v = F.linear(a_in, P(lp + "attn.value.weight"), P(lp + "attn.value.bias") if ab else None).view(B, L, nh, hd).transpose(1, 2)
        if tr["use_rope"]:
            q, k = _rope(q, k, ts if n_prefix == 0 else None, hd, tr["rope_theta"], emb["max_F"])
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        s = s.masked_fill(~allowed[:, None], float("-inf"))
        p = torch.softmax(s, dim=-1)
        p = _drop(p, f"attn_p.{li}", drop_scales)
        o = (p @ v).transpose(1, 2).reshape(B, L, H)
        o = _drop(o, f"attn_o.{li}", drop_scales)
        h = h + F.linear(o, P(lp + "attn.out_proj.weight"), P(lp + "attn.out_proj.bias") if ab else None)
        m_in = F.layer_norm(h, (H,), P(lp + "ln2.weight"), P(lp + "ln2.bias"), 1e-5)
        mb = tr["mlp_bias"]
        u = _ACTS[tr["act"]](F.linear(m_in, P(lp + "mlp.up_proj.weight"), P(lp + "mlp.up_proj.bias") if mb else None))
        d = F.linear(u, P(lp + "mlp.down_proj.weight"), P(lp + "mlp.down_proj.bias") if mb else None)
        h = h + _drop(d, f"mlp.{li}", drop_scales)
    h = F.layer_norm(h, (H,), P("out_norm.weight"), P("out_norm.bias"), 1e-5)
    if n_prefix:
        h = h[:, n_prefix:, :]
    fc = enc_cfg["factors"]
    h = _drop(h, "factors", drop_scales)       # models/ndt1.py:372: self.proj(self.dropout(x)), proj = Identity when inactive
    if fc["active"]:
        h = _ACTS[fc["act"]](F.linear(h, P("out_proj.proj.0.weight"), P("out_proj.proj.0.bias") if fc["bias"] else None))
    return h, mask, targets_mask


def _rope(q, k, pos_ids, dim, base, max_f):
    """models/ndt1.py:46-71."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
    t = torch.arange(max_f, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(q.dtype)[pos_ids].unsqueeze(1), emb.sin().to(q.dtype)[pos_ids].unsqueeze(1)

    def rot(x):
        x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
        return torch.cat((-x2, x1), -1)

    return q * cos + rot(q) * sin, k * cos + rot(k) * sin


def ndt1_forward(
    params: Dict[str, torch.Tensor],
    cfg: dict,                 # merged model config (keys: encoder, decoder)
    method_kwargs: dict,       # method_name, vocab_size/blank_id/zero_infinity or loss/log_input
    spikes, spikes_mask, spikes_timestamp, spikes_lengths,
    targets=None, targets_lengths=None, block_idx=None, day_idx=None,
    training: bool = False, noise=None, masker_draws=None, drop_scales=None,
):
    """NDT1.forward, models/ndt1.py:523-589.  Returns a dict with the fields of
    NDT1Output (:20-26): loss (sum, not mean), n_examples, preds, targets, mask."""
    method = method_kwargs["method_name"]
    enc = cfg["encoder"]
    if method in ("mlm", "autoregressive"):
        assert targets is None, "No targets needed for ssl"
        targets = spikes.clone()
    x, mask, targets_mask = encoder_forward(params, enc, spikes, spikes_mask, spikes_timestamp, block_idx, day_idx,
                                            training, noise, masker_draws, drop_scales)
    emb = enc["embedder"]
    lens = torch.from_numpy(stacked_lengths(spikes_lengths.numpy(), emb["stack"]["active"], emb["stack"]["size"],
                                            emb["stack"]["stride"]))
    preds = F.linear(x, params["decoder.0.weight"], params["decoder.0.bias"])
    if method in ("mlm", "autoregressive"):
        loss_name, log_input = method_kwargs["loss"], method_kwargs["log_input"]
        if loss_name == "mse" or not log_input:
            preds = F.relu(preds)
        if loss_name == "poisson_nll":
            lf = lambda p, t: F.poisson_nll_loss(p, t, log_input=log_input, reduction="none")
        elif loss_name == "mse":
            lf = lambda p, t: F.mse_loss(p, t, reduction="none")
        else:
            raise Exception(f"Loss {loss_name} not implemented")
        if method == "mlm":
            tm = targets_mask & mask.unsqueeze(2)
            loss = (lf(preds, targets) * tm).sum()
            return dict(loss=loss, n_examples=tm.sum(), preds=preds, targets=targets, mask=tm)
        sm = mask[:, :-1]
        loss = (lf(preds[:, :-1, :], targets[:, 1:, :]) * sm.unsqueeze(2)).sum()
        return dict(loss=loss, n_examples=sm.sum() * targets.size(2), preds=preds, targets=targets, mask=mask)
    if method in ("ctc", "endtoend"):
        preds = F.log_softmax(preds, dim=-1)
        loss = F.ctc_loss(preds.transpose(0, 1), targets, lens, targets_lengths, blank=method_kwargs["blank_id"],
                          reduction="none", zero_infinity=method_kwargs["zero_infinity"]).sum()
        return dict(loss=loss, n_examples=torch.tensor(spikes.size(0), dtype=torch.int64), preds=preds, targets=targets,
                    mask=None)
    raise Exception(f"Method {method} not implemented yet for NDT1")


def ndt1_loss_and_grads(params, cfg, method_kwargs, batch: dict, dtype=torch.float32, **kw):
    """Forward + autograd backward.  Returns (outputs, {name: grad})."""
    p = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in params.items()}
    b = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
    out = ndt1_forward(p, cfg, method_kwargs, **b, **kw)
    out["loss"].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    return out, grads


# --------------------------------------------------------------------------
# synthetic workload (BASELINE.md section 3 / SURVEY.md 8d) and the CPU train step
# --------------------------------------------------------------------------


def synthetic_ctc_batch(B=32, T=1000, N=256, seed=1, fixed_length=False, stack=(32, 4)):
    """Speech-BCI shaped batch: z-scored float features, lengths U{0.6T..T} with one
    full-length trial, right padding with 0, phoneme ids U{1..40}, target lengths
    U{20..60} clipped so that CTC stays feasible."""
    g = torch.Generator().manual_seed(seed)
    spikes = torch.randn(B, T, N, generator=g)
    if fixed_length:
        lens = torch.full((B,), T, dtype=torch.int64)
    else:
        lens = torch.randint(int(0.6 * T), T + 1, (B,), generator=g)
        lens[0] = T
    t = torch.arange(T)[None, :]
    mask = (t < lens[:, None]).to(torch.int64)
    spikes = spikes * mask[:, :, None]
    ts = t.expand(B, T) * mask
    tl = torch.randint(20, 61, (B,), generator=g)
    out_len = 1 + (lens - stack[0]) // stack[1] if stack else lens
    tl = torch.minimum(tl, torch.clamp(out_len // 2, min=1))
    S = int(tl.max())
    tg = torch.randint(1, 41, (B, S), generator=g)
    tg = tg * (torch.arange(S)[None, :] < tl[:, None])
    return dict(spikes=spikes, spikes_mask=mask, spikes_timestamp=ts, spikes_lengths=lens, targets=tg,
                targets_lengths=tl)


def synthetic_ssl_batch(B=16, T=100, N=668, seed=1):
    """BASELINE.json configs[0] inputs (SURVEY.md 8d, configs/ndt1.yaml shape): Poisson(0.1) spike counts, full-length trials
    (the same generator calls as tests/golden/make_golden.py::ssl_batch)."""
    g = torch.Generator().manual_seed(seed)
    sp = torch.poisson(torch.full((B, T, N), 0.1), generator=g)
    msk = torch.ones(B, T, dtype=torch.int64)
    return dict(spikes=sp, spikes_mask=msk, spikes_timestamp=torch.arange(T)[None].expand(B, T).contiguous(),
                spikes_lengths=torch.full((B,), T, dtype=torch.int64))
