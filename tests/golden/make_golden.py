"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (colehurwitz/llm_bci, mounted read-only at /root/reference) on the CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own, so
these files are what pins the oracle (oracle/ndt1_oracle.py) and, through it,
the CUDA path.  Nothing here is imported by the product.

Work-arounds applied to the environment, not to the reference sources:
  * scipy>=1.13 removed ``scipy.signal.gaussian`` (used at models/ndt1.py:87):
    alias it to ``scipy.signal.windows.gaussian``.
  * ``editdistance`` is absent: a stub module lets utils/eval_bci.py import so
    that ``format_ctc`` (eval_bci.py:41-48) can be called.
  * mlm needs the masker entry to be *named* ``active`` (ndt1.py:481 reads
    ``config.encoder.masker.active`` although ``masker`` is a dict of maskers).
"""
import os
import sys
import types
import copy

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    os.chdir(REF)
    sys.path.insert(0, REF)
    import scipy.signal
    import scipy.signal.windows
    scipy.signal.gaussian = scipy.signal.windows.gaussian
    sys.modules.setdefault("editdistance", types.ModuleType("editdistance"))
    from utils.config_utils import update_config, DictConfig
    from models.ndt1 import NDT1, create_context_mask
    from models.masker import Masker
    from data_utils.datasets import pad_collate_fn, padded_array
    from utils.eval_bci import format_ctc
    return dict(update_config=update_config, DictConfig=DictConfig, NDT1=NDT1, Masker=Masker,
                create_context_mask=create_context_mask, pad_collate_fn=pad_collate_fn, padded_array=padded_array,
                format_ctc=format_ctc)


def small_ctc_overrides():
    return {"encoder": {
        "embedder": {"n_channels": 16, "input_dim": 16, "max_F": 64, "dropout": 0.0,
                     "stack": {"active": True, "size": 32, "stride": 4}},
        "transformer": {"n_layers": 2, "hidden_size": 64, "n_heads": 4, "inter_size": 64, "dropout": 0.0},
        "smooth_and_noise": {"noise": False},
    }}


def make_ctc_batch(B, T, N, seed, S_lo=3, S_hi=8):
    g = torch.Generator().manual_seed(seed)
    spikes = torch.randn(B, T, N, generator=g)
    lens = torch.randint(int(0.6 * T), T + 1, (B,), generator=g)
    lens[0] = T
    t = torch.arange(T)[None, :]
    mask = (t < lens[:, None]).to(torch.int64)
    spikes = spikes * mask[:, :, None]
    ts = t.expand(B, T) * mask
    tl = torch.randint(S_lo, S_hi + 1, (B,), generator=g)
    S = int(tl.max())
    tg = torch.randint(1, 41, (B, S), generator=g)
    tg = tg * (torch.arange(S)[None, :] < tl[:, None])
    return dict(spikes=spikes, spikes_mask=mask, spikes_timestamp=ts, spikes_lengths=lens, targets=tg,
                targets_lengths=tl)


def flat(prefix, d):
    return {f"{prefix}/{k}": (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def run_ref(model, batch, train=False, seed=None):
    model.train(train)
    model.zero_grad()
    if seed is not None:
        torch.manual_seed(seed)
    b = {k: v.clone() for k, v in batch.items()}
    out = model(**b)
    out.loss.backward()
    grads = {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}
    return out, grads


def main():
    R = import_reference()
    update_config, NDT1 = R["update_config"], R["NDT1"]
    torch.set_num_threads(8)
    only_full = "--only-full" in sys.argv
    if "--only-variants" in sys.argv:
        variant_cases(R)
        return
    if "--bci" in sys.argv:
        bci_case()
        return
    if "--itransformer" in sys.argv:
        itransformer_cases(R)
        return
    if "--round2" in sys.argv:            # fixtures added in round 2 (the earlier files regenerate bit-identically and are left alone)
        autocast_error_cases(R)
        ssl_full_case(R)
        full_b32_case(R)
        return
    if not only_full:
        small_cases(R)
        variant_cases(R)
    full_case(R)


def small_cases(R):
    update_config, NDT1 = R["update_config"], R["NDT1"]

    # ------------------------------------------------------------------ 1. small CTC, eval-like numerics
    trainer = update_config("configs/trainer_ctc_ndt1.yaml", None)
    cfg = update_config(copy.deepcopy(dict(trainer.model)), small_ctc_overrides())
    torch.manual_seed(11)
    model = NDT1(cfg, **trainer.method.model_kwargs)
    batch = make_ctc_batch(3, 120, 16, seed=5)
    out, grads = run_ref(model, batch, train=True)   # train mode; dropout 0, noise off, masker inactive
    d = {}
    d.update(flat("param", dict(model.state_dict())))
    d.update(flat("batch", batch))
    d.update(flat("grad", grads))
    d["out/loss"] = out.loss.detach().numpy()
    d["out/preds"] = out.preds.detach().numpy()
    d["out/n_examples"] = out.n_examples.numpy()
    d["out/decoded"] = np.array([len(R["format_ctc"](p.argmax(-1).tolist(), list(range(41)), 0)) for p in out.preds.detach()])
    dec = [R["format_ctc"](p.argmax(-1).tolist(), list(range(41)), 0) for p in out.preds.detach()]
    d["out/decoded_flat"] = np.array([x for s in dec for x in s] + [-1], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "ctc_small.npz"), **d)
    print("ctc_small loss", float(out.loss))

    # ------------------------------------------------------------------ 2. small CTC with injected noise (train mode)
    cfg_n = update_config(copy.deepcopy(dict(cfg)), {"encoder": {"smooth_and_noise": {"noise": True}}})
    cfg_n["encoder"]["from_pt"] = None
    torch.manual_seed(11)
    model_n = NDT1(cfg_n, **trainer.method.model_kwargs)
    out_n, grads_n = run_ref(model_n, batch, train=True, seed=123)
    torch.manual_seed(123)
    white = torch.randn(3, 120, 16)
    offset = torch.randn(3, 1, 16)
    d = {"noise/white": white.numpy(), "noise/offset": offset.numpy(), "out/loss": out_n.loss.detach().numpy(),
         "out/preds": out_n.preds.detach().numpy()}
    d.update(flat("grad", {k: v for k, v in grads_n.items() if k.endswith("stack_projection.weight") or k.endswith("decoder.0.bias")
                           or k.endswith("embed_spikes.weight")}))
    np.savez_compressed(os.path.join(HERE, "ctc_small_noise.npz"), **d)
    print("ctc_small_noise loss", float(out_n.loss))

    # ------------------------------------------------------------------ 3. small masked-LM (SSL) with a temporal masker
    mk = {"active": True, "mode": "temporal", "ratio": 0.3, "zero_ratio": 0.8, "random_ratio": 0.5, "expand_prob": 1.0,
          "max_timespan": 3, "regions": None, "channels": None}
    ssl_over = {"encoder": {
        "masker": {"active": mk},
        "embedder": {"n_channels": 24, "input_dim": 16, "max_F": 64, "dropout": 0.0, "stack": {"active": False}},
        "transformer": {"n_layers": 2, "hidden_size": 64, "n_heads": 4, "inter_size": 64, "dropout": 0.0},
        "smooth_and_noise": {"noise": False},
        "context": {"forward": 2, "backward": 5},
    }}
    cfg_s = update_config("configs/ndt1.yaml", ssl_over)
    del cfg_s["encoder"]["masker"]["neuron"]
    torch.manual_seed(12)
    model_s = NDT1(cfg_s, method_name="mlm", loss="poisson_nll", log_input=True)
    g = torch.Generator().manual_seed(9)
    B, T, N = 4, 40, 24
    sp = torch.poisson(torch.full((B, T, N), 0.6), generator=g)
    lens = torch.tensor([40, 33, 40, 25])
    msk = (torch.arange(T)[None] < lens[:, None]).to(torch.int64)
    sp = sp * msk[:, :, None]
    sbatch = dict(spikes=sp, spikes_mask=msk, spikes_timestamp=torch.arange(T)[None].expand(B, T) * msk, spikes_lengths=lens)
    out_s, grads_s = run_ref(model_s, sbatch, train=True, seed=321)
    # replay the draws: CPU generator order
    torch.manual_seed(321)
    expand = bool(torch.bernoulli(torch.tensor(mk["expand_prob"]).float()))
    timespan = int(torch.randint(1, mk["max_timespan"] + 1, (1,)).item()) if expand else 1
    ratio = mk["ratio"] / timespan
    m_draw = torch.bernoulli(torch.full((B, T), ratio))
    z_draw = torch.bernoulli(torch.full((B, T, N), mk["zero_ratio"]))
    r_draw = torch.bernoulli(torch.full((B, T, N), mk["random_ratio"]))
    rnd = torch.rand((B, T, N))
    d = {}
    d.update(flat("param", dict(model_s.state_dict())))
    d.update(flat("batch", sbatch))
    d.update(flat("grad", grads_s))
    d.update({"draw/mask": m_draw.numpy(), "draw/zero": z_draw.numpy(), "draw/random": r_draw.numpy(), "draw/rand": rnd.numpy(),
              "draw/timespan": np.array(timespan), "out/loss": out_s.loss.detach().numpy(), "out/preds": out_s.preds.detach().numpy(),
              "out/n_examples": out_s.n_examples.numpy(), "out/mask": out_s.mask.numpy()})
    np.savez_compressed(os.path.join(HERE, "mlm_small.npz"), **d)
    print("mlm_small loss", float(out_s.loss), "n", int(out_s.n_examples), "timespan", timespan)

    # ------------------------------------------------------------------ 4. masker modes, bit exact
    Masker, DictConfig = R["Masker"], R["DictConfig"]
    d = {}
    B, T, N = 3, 17, 10
    base = torch.randn(B, T, N, generator=torch.Generator().manual_seed(3)) * 2 + 1
    modes = {
        "temporal1": dict(mode="temporal", ratio=0.4, zero_ratio=0.7, random_ratio=0.6, expand_prob=0.0, max_timespan=1),
        "temporal4": dict(mode="temporal", ratio=0.5, zero_ratio=0.5, random_ratio=1.0, expand_prob=1.0, max_timespan=4),
        "neuron": dict(mode="neuron", ratio=0.3, zero_ratio=1.0, random_ratio=1.0, expand_prob=0.0, max_timespan=1),
        "random": dict(mode="random", ratio=0.25, zero_ratio=0.3, random_ratio=0.5, expand_prob=0.0, max_timespan=1),
        "cosmooth": dict(mode="co-smooth", ratio=0.0, zero_ratio=0.9, random_ratio=0.2, expand_prob=0.0, max_timespan=1,
                         channels=[1, 4, 9]),
    }
    for name, mc in modes.items():
        full = dict(active=True, regions=None, channels=None)
        full.update(mc)
        for seed in (0, 1, 2, 7):
            mkr = Masker(DictConfig(full))
            mkr.train()
            torch.manual_seed(seed)
            so, mo = mkr(base.clone())
            torch.manual_seed(seed)
            ts = 1
            if full["mode"] == "temporal":
                if torch.bernoulli(torch.tensor(full["expand_prob"]).float()):
                    ts = int(torch.randint(1, full["max_timespan"] + 1, (1,)).item())
                probs = torch.full((B, T), full["ratio"] / ts)
            elif full["mode"] == "neuron":
                probs = torch.full((B, N), full["ratio"])
            elif full["mode"] == "random":
                probs = torch.full((B, T, N), full["ratio"])
            else:
                probs = torch.zeros(N)
                for c in full["channels"]:
                    probs[c] = 1
            md = torch.bernoulli(probs)
            zd = torch.bernoulli(torch.full((B, T, N), full["zero_ratio"]))
            rd = torch.bernoulli(torch.full((B, T, N), full["random_ratio"]))
            rn = torch.rand((B, T, N))
            key = f"{name}/{seed}"
            d.update({f"{key}/mask_draw": md.numpy(), f"{key}/zero": zd.numpy().astype(np.uint8),
                      f"{key}/random": rd.numpy().astype(np.uint8), f"{key}/rand": rn.numpy(), f"{key}/timespan": np.array(ts),
                      f"{key}/out_spikes": so.numpy(), f"{key}/out_mask": mo.numpy()})
    d["base"] = base.numpy()
    d["modes"] = np.array(list(modes.keys()))
    d["mode_names"] = np.array([modes[k]["mode"] for k in modes])
    np.savez_compressed(os.path.join(HERE, "masker.npz"), **d)

    # ------------------------------------------------------------------ 5. collate
    rng = np.random.default_rng(0)
    rows = []
    for L, S in ((50, 5), (37, 9), (44, 3)):
        rows.append({"spikes": rng.standard_normal((L, 6)).astype(np.float32), "spikes_mask": np.ones(L, dtype=np.int64),
                     "spikes_timestamp": np.arange(L, dtype=np.int64), "spikes_lengths": np.asarray(L, dtype=np.int64),
                     "targets": rng.integers(1, 41, S).astype(np.int64), "targets_lengths": np.asarray(S, dtype=np.int64),
                     "sentence": "abc", "extra": rng.standard_normal((L, 2)).astype(np.float32)})
    d = {}
    pads = {
        "right": {k: dict(dim=0, side="right", value=0, truncate=None, min_length=None) for k in ("spikes", "spikes_mask", "spikes_timestamp", "targets")},
        "left_trunc": {k: dict(dim=0, side="left", value=-1, truncate=40, min_length=None) for k in ("spikes", "spikes_mask", "spikes_timestamp", "targets")},
        "minlen": {k: dict(dim=0, side="right", value=0, truncate=64, min_length=60) for k in ("spikes", "spikes_mask", "spikes_timestamp")},
    }
    model_inputs = ["spikes", "spikes_mask", "spikes_timestamp", "spikes_lengths", "targets", "targets_lengths"]
    for name, pd in pads.items():
        padded, unused = R["pad_collate_fn"](copy.deepcopy(rows), model_inputs, pd)
        for k, v in padded.items():
            if torch.is_tensor(v):
                d[f"{name}/{k}"] = v.numpy()
            elif isinstance(v, list) and torch.is_tensor(v[0]):
                for i, vi in enumerate(v):
                    d[f"{name}/{k}/{i}"] = vi.numpy()
        d[f"{name}/unused_keys"] = np.array(sorted(unused.keys()))
    for i, r in enumerate(rows):
        for k, v in r.items():
            if isinstance(v, np.ndarray):
                d[f"row{i}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "collate.npz"), **d)

    # ------------------------------------------------------------------ 6. context band + greedy collapse
    d = {}
    for cf, cb in ((-2, -2), (0, -2), (-2, 0), (-1, -1), (3, 1), (0, 0), (-1, 4), (2, -1)):
        d[f"band/{cf}/{cb}"] = R["create_context_mask"](cf, cb, 12).numpy()
    seqs = [[0, 0, 3, 3, 0, 3, 4, 4, 0, 0, 5, 3], [1, 0, 1, 0, 1], [0, 0, 0], [7, 7, 7, 2, 2, 7], []]
    for i, s in enumerate(seqs):
        d[f"ctc_in/{i}"] = np.array(s, dtype=np.int64)
        d[f"ctc_out/{i}"] = np.array(R["format_ctc"](s, list(range(41)), 0), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "index_ops.npz"), **d)


VARIANTS = {
    # name: (model overrides on top of small_ctc_overrides(), needs day_idx)
    "rope": ({"encoder": {"transformer": {"use_rope": True, "rope_theta": 10000.0}}}, False),
    "adapt": ({"encoder": {"embedder": {"adapt": True, "n_days": 4}}}, True),
    "gelu_factors": ({"encoder": {"embedder": {"act": "gelu"},
                                  "factors": {"active": True, "size": 48, "act": "relu", "bias": True, "dropout": 0.0,
                                              "fixup_init": True, "init_range": 0.1}}}, False),
    "tokens_ctx": ({"encoder": {"embedder": {"block_token": True, "day_token": True, "n_blocks": 5, "n_days": 4},
                                "context": {"forward": 3, "backward": 7}}}, "both"),
    "rope_adapt_gelu_factors": ({"encoder": {"transformer": {"use_rope": True, "rope_theta": 500.0},
                                             "embedder": {"adapt": True, "n_days": 3, "act": "gelu", "pos": False},
                                             "factors": {"active": True, "size": 40, "act": "gelu", "bias": False, "dropout": 0.0,
                                                         "fixup_init": False, "init_range": 0.1}}}, True),
}


def variant_cases(R):
    """Options of the path that the shipped yaml leaves off (models/ndt1.py:118-127 adapt, :262-266, 285-286 RoPE, :146 embedder
    activation, :358-366 factors projection): one small CTC model per combination, outputs of the unmodified reference."""
    update_config, NDT1 = R["update_config"], R["NDT1"]
    trainer = update_config("configs/trainer_ctc_ndt1.yaml", None)
    base = update_config(copy.deepcopy(dict(trainer.model)), small_ctc_overrides())
    batch = make_ctc_batch(3, 120, 16, seed=5)
    d = flat("batch", batch)
    for name, (over, days) in VARIANTS.items():
        cfg = update_config(copy.deepcopy(dict(base)), over)
        torch.manual_seed(21)
        model = NDT1(cfg, **trainer.method.model_kwargs)
        b = dict(batch)
        if days:
            b["day_idx"] = torch.tensor([2, 0, 2])
            d[f"{name}/day_idx"] = b["day_idx"].numpy()
        if days == "both":
            b["block_idx"] = torch.tensor([4, 1, 0])
            d[f"{name}/block_idx"] = b["block_idx"].numpy()
        out, grads = run_ref(model, b, train=True)
        d.update(flat(f"{name}/param", dict(model.state_dict())))
        d.update(flat(f"{name}/grad", grads))
        d[f"{name}/out/loss"] = out.loss.detach().numpy()
        d[f"{name}/out/preds"] = out.preds.detach().numpy()
        print("variant", name, "loss", float(out.loss))
    np.savez_compressed(os.path.join(HERE, "ctc_variants.npz"), **d)

    # autoregressive next-bin prediction (models/ndt1.py:563-578) with the MSE loss and ReLU rates (:508-515)
    ar_over = {"encoder": {
        "embedder": {"n_channels": 24, "input_dim": 16, "max_F": 64, "dropout": 0.0, "stack": {"active": False}},
        "transformer": {"n_layers": 2, "hidden_size": 64, "n_heads": 4, "inter_size": 64, "dropout": 0.0},
        "smooth_and_noise": {"noise": False},
        "context": {"forward": 0, "backward": -2},
    }}
    cfg_a = update_config("configs/ndt1.yaml", ar_over)
    g = torch.Generator().manual_seed(19)
    B, T, N = 4, 40, 24
    sp = torch.poisson(torch.full((B, T, N), 0.6), generator=g)
    lens = torch.tensor([40, 31, 40, 22])
    msk = (torch.arange(T)[None] < lens[:, None]).to(torch.int64)
    sp = sp * msk[:, :, None]
    abatch = dict(spikes=sp, spikes_mask=msk, spikes_timestamp=torch.arange(T)[None].expand(B, T) * msk, spikes_lengths=lens)
    d = flat("batch", abatch)
    for name, kw in (("mse", dict(loss="mse", log_input=False)), ("poisson_rate", dict(loss="poisson_nll", log_input=False)),
                     ("poisson_log", dict(loss="poisson_nll", log_input=True))):
        torch.manual_seed(31)
        model = NDT1(update_config("configs/ndt1.yaml", ar_over), method_name="autoregressive", **kw)
        out, grads = run_ref(model, abatch, train=True)
        d.update(flat(f"{name}/param", dict(model.state_dict())))
        d.update(flat(f"{name}/grad", grads))
        d[f"{name}/out/loss"] = out.loss.detach().numpy()
        d[f"{name}/out/preds"] = out.preds.detach().numpy()
        d[f"{name}/out/n_examples"] = out.n_examples.numpy()
        print("autoregressive", name, "loss", float(out.loss), "n", int(out.n_examples))
    np.savez_compressed(os.path.join(HERE, "autoregressive_small.npz"), **d)


def full_case(R):
    update_config, NDT1 = R["update_config"], R["NDT1"]
    # ------------------------------------------------------------------ 7. full-size model (config 2 architecture), B=4
    trainer = update_config("configs/trainer_ctc_ndt1.yaml", None)
    cfg_f = update_config(copy.deepcopy(dict(trainer.model)), {"encoder": {
        "embedder": {"dropout": 0.0}, "transformer": {"dropout": 0.0}, "smooth_and_noise": {"noise": False}}})
    torch.manual_seed(1)
    model_f = NDT1(cfg_f, **trainer.method.model_kwargs)
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.ndt1_oracle import synthetic_ctc_batch
    fb = synthetic_ctc_batch(B=4, T=1000, N=256, seed=1)
    out_f, grads_f = run_ref(model_f, fb, train=True)
    d = {"out/loss": out_f.loss.detach().numpy(), "out/preds_rows": out_f.preds.detach().numpy()[:, ::40, :],
         "out/preds_sum": out_f.preds.detach().double().sum().numpy(),
         "out/argmax": out_f.preds.detach().argmax(-1).numpy()}
    names = list(dict(model_f.named_parameters()).keys())
    d["names"] = np.array(names)
    d["param_sum"] = np.array([float(p.detach().double().sum()) for p in model_f.parameters()])
    d["param_abs"] = np.array([float(p.detach().double().abs().sum()) for p in model_f.parameters()])
    d["grad_norm"] = np.array([float(grads_f[n].double().norm()) for n in names])
    d["grad_absmax"] = np.array([float(grads_f[n].abs().max()) for n in names])
    # a few full gradients that are small enough to keep
    for n in ("decoder.0.weight", "decoder.0.bias", "encoder.out_norm.weight", "encoder.layers.0.ln1.weight",
              "encoder.layers.4.mlp.down_proj.bias", "encoder.layers.2.attn.query.bias", "encoder.embedder.embed_spikes.weight",
              "encoder.embedder.stack_projection.bias"):
        d[f"grad/{n}"] = grads_f[n].numpy()
    d["grad_slice/encoder.layers.0.attn.value.weight"] = grads_f["encoder.layers.0.attn.value.weight"][:8, :].numpy()
    d["grad_slice/encoder.embedder.stack_projection.weight"] = grads_f["encoder.embedder.stack_projection.weight"][:4, :].numpy()
    d["grad_slice/encoder.embedder.embed_pos.weight"] = grads_f["encoder.embedder.embed_pos.weight"][:4, :].numpy()
    # the same model and batch in float64 (the reference module runs in double): the
    # higher-precision target for the 1e-4 check of the fp32 CUDA mode
    model_d = NDT1(cfg_f, **trainer.method.model_kwargs).double()
    model_d.load_state_dict({k: v.double() for k, v in model_f.state_dict().items()})
    fbd = {k: (v.double() if v.is_floating_point() else v) for k, v in fb.items()}
    out_d, grads_d = run_ref(model_d, fbd, train=True)
    d["out64/loss"] = out_d.loss.detach().numpy()
    d["grad_norm64"] = np.array([float(grads_d[n].norm()) for n in names])
    for n in ("decoder.0.weight", "decoder.0.bias", "encoder.out_norm.weight", "encoder.layers.0.ln1.weight",
              "encoder.layers.4.mlp.down_proj.bias", "encoder.layers.2.attn.query.bias", "encoder.embedder.embed_spikes.weight",
              "encoder.embedder.stack_projection.bias"):
        d[f"grad64/{n}"] = grads_d[n].numpy().astype(np.float32)   # fp64 result rounded once to fp32 for storage
    d["grad64_slice/encoder.layers.0.attn.value.weight"] = grads_d["encoder.layers.0.attn.value.weight"][:8, :].numpy().astype(np.float32)
    d["grad64_slice/encoder.embedder.stack_projection.weight"] = grads_d["encoder.embedder.stack_projection.weight"][:4, :].numpy().astype(np.float32)
    d["grad64_slice/encoder.embedder.embed_pos.weight"] = grads_d["encoder.embedder.embed_pos.weight"][:4, :].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "ctc_full_b4.npz"), **d)
    print("ctc_full_b4 loss", float(out_f.loss), "fp64", float(out_d.loss))


BCI_OVER = {"projector": {"stacking": 2, "inter_size": 48, "bias": True, "act": "relu"}, "ndt1": {"encoder": {
    "embedder": {"n_channels": 16, "input_dim": 16, "max_F": 64, "dropout": 0.0, "stack": {"active": True, "size": 32, "stride": 4}},
    "transformer": {"n_layers": 2, "hidden_size": 64, "n_heads": 4, "inter_size": 64, "dropout": 0.0},
    "smooth_and_noise": {"noise": False}}}}


def bci_case():
    """BCI coupler (models/bci.py:88-96, 107-168) of the unmodified reference with its debug LLaMA (:51-53): outputs of
    prepare_embeds, gradients of a fixed linear functional of the spliced embeddings w.r.t. projector and encoder, and the
    end-to-end loss through the fp16 LLaMA.  `peft` is absent here: a stub lets models/bci.py import (LoRA is not exercised)."""
    os.chdir(REF)
    sys.path.insert(0, REF)
    import scipy.signal
    import scipy.signal.windows
    scipy.signal.gaussian = scipy.signal.windows.gaussian
    peft = types.ModuleType("peft")
    peft.LoraConfig, peft.get_peft_model = object, (lambda m, c: m)
    sys.modules["peft"] = peft
    from utils.config_utils import update_config
    from models.bci import BCI
    torch.set_num_threads(8)
    d = {}
    for name, stacking, act in (("s2_relu", 2, "relu"), ("s3_gelu", 3, "gelu")):      # 23 tokens: 2 and 3 both need the zero padding
        over = copy.deepcopy(BCI_OVER)
        over["projector"].update(stacking=stacking, act=act)
        cfg = update_config("configs/bci.yaml", over)
        torch.manual_seed(3)
        m = BCI(cfg, llm_path=None, debug=True, method_name="ctc", vocab_size=41, blank_id=0, zero_infinity=True)
        m.train()
        B, T, N, Lt = 3, 120, 16, 7
        g = torch.Generator().manual_seed(5)
        spikes = torch.randn(B, T, N, generator=g)
        lens = torch.tensor([120, 100, 90])
        mask = (torch.arange(T)[None] < lens[:, None]).long()
        spikes = spikes * mask[:, :, None]
        ts = torch.arange(T)[None].expand(B, T) * mask
        ids = torch.randint(0, 1000, (B, Lt), generator=g)
        am = torch.ones(B, Lt, dtype=torch.long)
        am[1, -2:] = 0
        split = torch.tensor([2, 0, 7])
        tg = ids.clone()
        tg[:, :2] = -100
        embeds, amask, tgo = m.prepare_embeds(ids, am, split, spikes.clone(), mask, ts, lens, None, None, tg)
        R = torch.randn(embeds.shape, generator=g)
        m.zero_grad()
        (embeds * R).sum().backward()
        text = m.llm.get_input_embeddings()(ids).detach().float()
        sd = {k: v for k, v in m.state_dict().items() if k.startswith("ndt1.") or k.startswith("projector.")}
        d.update(flat(f"{name}/param", sd))
        d.update(flat(f"{name}/grad", {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in m.named_parameters()
                                        if n.startswith("ndt1.encoder.") or n.startswith("projector.")}))
        d.update(flat(f"{name}/in", dict(spikes=spikes, spikes_mask=mask, spikes_timestamp=ts, spikes_lengths=lens, input_ids=ids,
                                         attention_mask=am, input_split=split, targets=tg, text_embeds=text, R=R)))
        d[f"{name}/out/embeds"] = embeds.detach().float().numpy()
        d[f"{name}/out/attention_mask"] = amask.numpy()
        d[f"{name}/out/targets"] = tgo.numpy()
        # (the reference cannot run prepare_embeds under CPU bf16 autocast -- torch.cat of its fp16 prompt embeddings with bf16 features
        #  raises "Unexpected floating ScalarType in at::autocast::prioritize" -- so the bf16 yardstick of the ReLU projector is the
        #  ReLU-factors case of autocast_error_cases, the same mechanism)
        # end to end through the fp16 LLaMA (CPU half arithmetic: a loose yardstick only)
        m.zero_grad()
        out = m(ids, am, split, spikes.clone(), mask, ts, lens, None, None, tg)
        d[f"{name}/out/loss"] = out.loss.detach().float().numpy()
        d[f"{name}/out/n_examples"] = out.n_examples.numpy()
        print("bci", name, "embeds", tuple(embeds.shape), "loss", float(out.loss), "n", int(out.n_examples))
    np.savez_compressed(os.path.join(HERE, "bci_coupler.npz"), **d)


KEEP_FULL = ("decoder.0.weight", "decoder.0.bias", "encoder.out_norm.weight", "encoder.layers.0.ln1.weight",
             "encoder.layers.4.mlp.down_proj.bias", "encoder.layers.2.attn.query.bias", "encoder.embedder.embed_spikes.bias",
             "encoder.layers.0.attn.value.bias", "encoder.layers.3.ln2.bias")
KEEP_SLICE = {"encoder.layers.0.attn.value.weight": 8, "encoder.layers.4.attn.out_proj.weight": 8, "encoder.layers.2.mlp.up_proj.weight": 8,
              "encoder.embedder.embed_pos.weight": 4, "encoder.layers.1.attn.query.weight": 4}


def grad_error_metrics(got, ref):
    """The metrics of tests/test_gpu_parity.py::check_grads: per tensor relative L2 and max-abs error, both floored at 1e-3 of
    the global scale; returns the worst of each over the tensors."""
    gscale = max(float(v.abs().max()) for v in ref.values())
    nscale = max(float(v.double().norm()) for v in ref.values())
    l2 = {n: float((got[n].double() - r.double()).norm() / max(float(r.double().norm()), 1e-3 * nscale)) for n, r in ref.items()}
    mx = {n: float((got[n].double() - r.double()).abs().max() / max(float(r.abs().max()), 1e-3 * gscale)) for n, r in ref.items()}
    return l2, mx


def run_ref_autocast(model, batch):
    """The reference under bf16 autocast, the way Accelerate(mixed_precision='bf16') runs it (deepspeed/*.yaml presets)."""
    model.train(True)
    model.zero_grad()
    b = {k: v.clone() for k, v in batch.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out = model(**b)
    out.loss.float().backward()
    grads = {n: (p.grad.clone().float() if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()}
    return out, grads


def autocast_error_cases(R):
    """For every case whose bf16 tolerance tests/test_gpu_parity.py widens beyond the nominal 2e-2 (ReLU heads / factors):
    the error the REFERENCE ITSELF makes under bf16 autocast against its own fp32 run, in the metrics of check_grads.  The CUDA
    bf16 path is then bounded by a stated multiple of this figure instead of a hand-picked number."""
    update_config, NDT1 = R["update_config"], R["NDT1"]
    trainer = update_config("configs/trainer_ctc_ndt1.yaml", None)
    base = update_config(copy.deepcopy(dict(trainer.model)), small_ctc_overrides())
    batch = make_ctc_batch(3, 120, 16, seed=5)
    d = {}

    def record(name, model, b):
        out32, g32 = run_ref(model, b, train=True)
        out16, g16 = run_ref_autocast(model, b)
        l2, mx = grad_error_metrics(g16, g32)
        d[f"{name}/loss_rel"] = np.array(abs(float(out16.loss) - float(out32.loss)) / abs(float(out32.loss)))
        d[f"{name}/grad_l2_max"] = np.array(max(l2.values()))
        d[f"{name}/grad_maxabs_max"] = np.array(max(mx.values()))
        d[f"{name}/grad_l2_median"] = np.array(float(np.median(list(l2.values()))))
        worst = max(l2, key=l2.get)
        print(f"autocast {name}: loss rel {float(d[f'{name}/loss_rel']):.2e}, grad rel-L2 max {max(l2.values()):.3e} ({worst}), "
              f"median {float(d[f'{name}/grad_l2_median']):.2e}, max-abs {max(mx.values()):.3e}")

    for name in ("gelu_factors", "rope"):                 # "rope": a case that is NOT waived, as the yardstick
        over, days = VARIANTS[name]
        torch.manual_seed(21)
        model = NDT1(update_config(copy.deepcopy(dict(base)), over), **trainer.method.model_kwargs)
        record(f"ctc_variants/{name}", model, dict(batch))
    ar_over = {"encoder": {
        "embedder": {"n_channels": 24, "input_dim": 16, "max_F": 64, "dropout": 0.0, "stack": {"active": False}},
        "transformer": {"n_layers": 2, "hidden_size": 64, "n_heads": 4, "inter_size": 64, "dropout": 0.0},
        "smooth_and_noise": {"noise": False},
        "context": {"forward": 0, "backward": -2},
    }}
    g = torch.Generator().manual_seed(19)
    B, T, N = 4, 40, 24
    sp = torch.poisson(torch.full((B, T, N), 0.6), generator=g)
    lens = torch.tensor([40, 31, 40, 22])
    msk = (torch.arange(T)[None] < lens[:, None]).to(torch.int64)
    sp = sp * msk[:, :, None]
    abatch = dict(spikes=sp, spikes_mask=msk, spikes_timestamp=torch.arange(T)[None].expand(B, T) * msk, spikes_lengths=lens)
    for name, kw in (("mse", dict(loss="mse", log_input=False)), ("poisson_rate", dict(loss="poisson_nll", log_input=False)),
                     ("poisson_log", dict(loss="poisson_nll", log_input=True))):
        torch.manual_seed(31)
        model = NDT1(update_config("configs/ndt1.yaml", ar_over), method_name="autoregressive", **kw)
        record(f"autoregressive/{name}", model, abatch)
    np.savez_compressed(os.path.join(HERE, "bf16_autocast_error.npz"), **d)


def ssl_batch(B=16, T=100, N=668, seed=1):
    """BASELINE.json configs[0] inputs: Poisson(0.1) counts, full-length trials."""
    g = torch.Generator().manual_seed(seed)
    sp = torch.poisson(torch.full((B, T, N), 0.1), generator=g)
    msk = torch.ones(B, T, dtype=torch.int64)
    return dict(spikes=sp, spikes_mask=msk, spikes_timestamp=torch.arange(T)[None].expand(B, T).contiguous(),
                spikes_lengths=torch.full((B,), T, dtype=torch.int64))


def ssl_full_case(R):
    """BASELINE.json configs[0] at FULL size: NDT1 masked-spike SSL, 16 x 100 bins x 668 neurons, temporal masker 0.3, Poisson-NLL
    on log rates, 5 x 1024 encoder without stacking (33.7 M parameters: kept as per-tensor sums; the test re-creates them from
    the same torch seed, see test_init_matches_reference_rng_order)."""
    update_config, NDT1 = R["update_config"], R["NDT1"]
    mk = {"active": True, "mode": "temporal", "ratio": 0.3, "zero_ratio": 1.0, "random_ratio": 1.0, "expand_prob": 0.0,
          "max_timespan": 1, "regions": None, "channels": None}
    over = {"encoder": {"masker": {"active": mk}, "embedder": {"n_channels": 668, "dropout": 0.0, "stack": {"active": False}},
                        "transformer": {"dropout": 0.0}, "smooth_and_noise": {"noise": False}}}
    cfg = update_config("configs/ndt1.yaml", over)
    del cfg["encoder"]["masker"]["neuron"]
    torch.manual_seed(1)
    model = NDT1(cfg, method_name="mlm", loss="poisson_nll", log_input=True)
    batch = ssl_batch()
    out, grads = run_ref(model, batch, train=True, seed=77)
    torch.manual_seed(77)                                   # replay the CPU draws
    expand = bool(torch.bernoulli(torch.tensor(mk["expand_prob"]).float()))
    assert not expand
    m_draw = torch.bernoulli(torch.full((16, 100), mk["ratio"]))
    names = [n for n, _ in model.named_parameters()]
    d = {"names": np.array(names), "draw/mask": m_draw.numpy().astype(np.uint8), "draw/timespan": np.array(1),
         "out/loss": out.loss.detach().numpy(), "out/n_examples": out.n_examples.numpy(),
         "out/mask_bt": out.mask[:, :, 0].numpy().astype(np.uint8), "out/mask_sum": np.array(int(out.mask.sum())),
         "out/preds_rows": out.preds.detach().numpy()[:, ::10, ::4], "out/preds_sum": out.preds.detach().double().sum().numpy()}
    assert bool((out.mask == out.mask[:, :, :1]).all())     # temporal: the mask is constant over neurons
    d["param_sum"] = np.array([float(p.detach().double().sum()) for p in model.parameters()])
    d["grad_norm"] = np.array([float(grads[n].double().norm()) for n in names])
    d["grad_absmax"] = np.array([float(grads[n].abs().max()) for n in names])
    for n in KEEP_FULL:
        if n in grads:
            d[f"grad/{n}"] = grads[n].numpy()
    d["grad/encoder.embedder.projection.bias"] = grads["encoder.embedder.projection.bias"].numpy()
    for n, rows in KEEP_SLICE.items():
        d[f"grad_slice/{n}"] = grads[n][:rows].numpy()
    d["grad_slice/decoder.0.weight"] = grads["decoder.0.weight"][:16].numpy()
    d["grad_slice/encoder.embedder.embed_spikes.weight"] = grads["encoder.embedder.embed_spikes.weight"][:8].numpy()
    d["grad_slice/encoder.embedder.projection.weight"] = grads["encoder.embedder.projection.weight"][:8].numpy()
    # the reference's own bf16-autocast error on this very case (the yardstick for the bf16 CUDA mode at this size)
    out16, g16 = run_ref_autocast_seeded(model, batch, 77)
    l2, mx = grad_error_metrics(g16, grads)
    d["autocast/loss_rel"] = np.array(abs(float(out16.loss) - float(out.loss)) / abs(float(out.loss)))
    d["autocast/grad_l2_max"] = np.array(max(l2.values()))
    d["autocast/grad_maxabs_max"] = np.array(max(mx.values()))
    np.savez_compressed(os.path.join(HERE, "ssl_full_b16.npz"), **d)
    print("ssl_full_b16 loss", float(out.loss), "n", int(out.n_examples), "autocast loss rel", float(d["autocast/loss_rel"]),
          "grad l2 max", float(d["autocast/grad_l2_max"]))


def run_ref_autocast_seeded(model, batch, seed):
    torch.manual_seed(seed)
    return run_ref_autocast(model, batch)


def full_b32_case(R):
    """BASELINE.json configs[1] at FULL size and FULL batch: 32 x 1000 x 256, the parity variant (dropout 0, noise off).
    Parameters come from torch.manual_seed(1) (bit-identical init, checked through param_sum)."""
    update_config, NDT1 = R["update_config"], R["NDT1"]
    trainer = update_config("configs/trainer_ctc_ndt1.yaml", None)
    cfg_f = update_config(copy.deepcopy(dict(trainer.model)), {"encoder": {
        "embedder": {"dropout": 0.0}, "transformer": {"dropout": 0.0}, "smooth_and_noise": {"noise": False}}})
    torch.manual_seed(1)
    model = NDT1(cfg_f, **trainer.method.model_kwargs)
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.ndt1_oracle import synthetic_ctc_batch
    fb = synthetic_ctc_batch(B=32, T=1000, N=256, seed=1)
    out, grads = run_ref(model, fb, train=True)
    names = [n for n, _ in model.named_parameters()]
    preds = out.preds.detach()
    d = {"names": np.array(names), "out/loss": out.loss.detach().numpy(), "out/n_examples": out.n_examples.numpy(),
         "out/preds_rows": preds.numpy()[:, ::40, :], "out/preds_sum": preds.double().sum().numpy(),
         "out/argmax": preds.argmax(-1).numpy().astype(np.int16),
         "out/top2_margin_min": np.array(float((preds.topk(2, -1).values[..., 0] - preds.topk(2, -1).values[..., 1]).min()))}
    d["param_sum"] = np.array([float(p.detach().double().sum()) for p in model.parameters()])
    d["grad_norm"] = np.array([float(grads[n].double().norm()) for n in names])
    d["grad_absmax"] = np.array([float(grads[n].abs().max()) for n in names])
    for n in KEEP_FULL:
        d[f"grad/{n}"] = grads[n].numpy()
    d["grad/encoder.embedder.stack_projection.bias"] = grads["encoder.embedder.stack_projection.bias"].numpy()
    for n, rows in KEEP_SLICE.items():
        d[f"grad_slice/{n}"] = grads[n][:rows].numpy()
    d["grad_slice/encoder.embedder.stack_projection.weight"] = grads["encoder.embedder.stack_projection.weight"][:4].numpy()
    d["grad_slice/encoder.embedder.embed_spikes.weight"] = grads["encoder.embedder.embed_spikes.weight"][:16].numpy()
    # the same in float64: the target of the strict (fp32) CUDA mode
    model_d = NDT1(cfg_f, **trainer.method.model_kwargs).double()
    model_d.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    fbd = {k: (v.double() if v.is_floating_point() else v) for k, v in fb.items()}
    out_d, grads_d = run_ref(model_d, fbd, train=True)
    d["out64/loss"] = out_d.loss.detach().numpy()
    d["grad_norm64"] = np.array([float(grads_d[n].norm()) for n in names])
    for n in KEEP_FULL:
        d[f"grad64/{n}"] = grads_d[n].numpy().astype(np.float32)
    d["grad64/encoder.embedder.stack_projection.bias"] = grads_d["encoder.embedder.stack_projection.bias"].numpy().astype(np.float32)
    for n, rows in KEEP_SLICE.items():
        d[f"grad64_slice/{n}"] = grads_d[n][:rows].numpy().astype(np.float32)
    del model_d, grads_d
    # the reference's own bf16-autocast error at this size
    out16, g16 = run_ref_autocast(model, fb)
    l2, mx = grad_error_metrics(g16, grads)
    d["autocast/loss_rel"] = np.array(abs(float(out16.loss) - float(out.loss)) / abs(float(out.loss)))
    d["autocast/grad_l2_max"] = np.array(max(l2.values()))
    d["autocast/grad_maxabs_max"] = np.array(max(mx.values()))
    d["autocast/argmax_agree"] = np.array(float((out16.preds.detach().float().argmax(-1) == preds.argmax(-1)).float().mean()))
    np.savez_compressed(os.path.join(HERE, "ctc_full_b32.npz"), **d)
    print("ctc_full_b32 loss", float(out.loss), "fp64", float(out_d.loss), "autocast loss rel", float(d["autocast/loss_rel"]),
          "grad l2 max", float(d["autocast/grad_l2_max"]), "argmax agree", float(d["autocast/argmax_agree"]))




# --------------------------------------------------------------------------- iTransformer (models/itransformer.py)
ITR_MASKER = {"active": True, "force_active": True, "mode": "neuron", "ratio": 0.25, "zero_ratio": 1.0, "random_ratio": 1.0,
              "expand_prob": 0.0, "max_timespan": 1, "regions": None, "channels": None}
ITR_SMALL = {"masker": {"main": ITR_MASKER},
             "encoder": {"embedder": {"dropout": 0.0, "max_n_bins": 20}, "hidden_size": 64, "n_heads": 4, "n_layers": 2, "dropout": 0.0,
                         "max_n_channels": 32, "embed_region": False}}
# BASELINE.json configs[3] / SURVEY 8 f4: the shipped yaml (768 hidden, 8 heads of 96, 5 post-LN layers, FFN 3072, MLP embedder
# 100 -> 768 -> 768, cls token) with the masker keys the yaml lacks (`active`, `regions`: SURVEY component 8) and dropout off
ITR_FULL = {"masker": {"main": dict(ITR_MASKER, ratio=0.1)},
            "encoder": {"embedder": {"dropout": 0.0}, "dropout": 0.0, "embed_region": False}}


def itr_batch(B, T, N, seed, rate=0.3):
    g = torch.Generator().manual_seed(seed)
    sp = torch.poisson(torch.full((B, T, N), rate), generator=g)
    return dict(spikes=sp, spikes_mask=torch.ones(B, T, dtype=torch.int64), spikes_timestamp=torch.arange(T)[None].expand(B, T).contiguous())


def itransformer_cases(R):
    """Outputs of the unmodified reference iTransformer (mlm method, Poisson-NLL on log rates, MLP embedder, channel embeddings,
    cls token) in train mode with dropout 0: a small case with every gradient, and the config-3 size (16 x 100 bins x 669 neurons)
    with per-parameter norms, a gradient subset and the reference's own bf16-autocast error."""
    from models.itransformer import iTransformer
    uc = R["update_config"]
    # ---- small
    cfg = uc("configs/itransformer.yaml", ITR_SMALL)
    torch.manual_seed(1)
    m = iTransformer(cfg, method_name="mlm", loss="poisson_nll", log_input=True)
    b = itr_batch(3, 20, 24, 3)
    out, grads = run_ref(m, b, train=True, seed=5)
    d = {"names": np.array([n for n, _ in m.named_parameters()])}
    d.update(flat("param", dict(m.named_parameters())))
    d.update(flat("grad", grads))
    d.update(flat("batch", b))
    d.update({"out/loss": out.loss.detach().numpy(), "out/n_examples": out.n_examples.numpy(), "out/preds": out.preds.detach().numpy(),
              "out/mask": out.mask.numpy().astype(np.uint8), "seed": np.array(5)})
    # the reference's own bf16-autocast error on this case, per tensor (the yardstick for the bf16 CUDA mode on a 64-wide model)
    out16, g16 = run_ref_autocast_seeded(m, b, 5)
    l2, mx = grad_error_metrics(g16, grads)
    d["autocast/loss_rel"] = np.array(abs(float(out16.loss) - float(out.loss)) / abs(float(out.loss)))
    d["autocast/grad_l2_max"] = np.array(max(l2.values()))
    d["autocast/grad_l2_worst"] = np.array(max(l2, key=l2.get))
    # second method with the same weights: dyn_behaviour (cls token -> MLP decoder -> one value per bin, MSE over valid bins)
    torch.manual_seed(1)
    m2 = iTransformer(cfg, method_name="dyn_behaviour")
    b2 = dict(b, targets=torch.randn(3, 20, generator=torch.Generator().manual_seed(9)))
    b2["spikes_mask"] = (torch.arange(20)[None] < torch.tensor([20, 14, 17])[:, None]).to(torch.int64)
    out2, grads2 = run_ref(m2, b2, train=True, seed=5)
    main_params = {n: p.detach().clone() for n, p in m.named_parameters()}

    def own_params(model):      # the variants share every initial value with the main model except the decoder's last layer
        return {n: p for n, p in model.named_parameters() if n not in main_params or p.shape != main_params[n].shape or not torch.equal(p, main_params[n])}

    d.update(flat("dyn/param", own_params(m2)))
    d.update(flat("dyn/grad", grads2))
    d.update({"dyn/targets": b2["targets"].numpy(), "dyn/spikes_mask": b2["spikes_mask"].numpy(), "dyn/loss": out2.loss.detach().numpy(),
              "dyn/n_examples": out2.n_examples.numpy(), "dyn/preds": out2.preds.detach().numpy()})
    # third method: stat_behaviour (one label / value per trial from the cls token), cross-entropy over 3 labels and MSE
    for tag, kw, tg in (("xent", dict(method_name="stat_behaviour", loss="xent", n_labels=3), torch.tensor([[2.0], [0.0], [1.0]])),
                        ("smse", dict(method_name="stat_behaviour", loss="mse"), torch.tensor([[0.3], [-1.2], [0.7]]))):
        torch.manual_seed(1)
        m3 = iTransformer(cfg, **kw)
        b3 = dict(b, targets=tg)
        out3, grads3 = run_ref(m3, b3, train=True, seed=5)
        d.update(flat(f"{tag}/param", own_params(m3)))
        d.update(flat(f"{tag}/grad", grads3))
        d.update({f"{tag}/targets": tg.numpy(), f"{tag}/loss": out3.loss.detach().numpy(), f"{tag}/n_examples": out3.n_examples.numpy(),
                  f"{tag}/preds": out3.preds.detach().numpy()})
    np.savez_compressed(os.path.join(HERE, "itransformer_small.npz"), **d)
    print("itransformer_small loss", float(out.loss), "n", int(out.n_examples), "dyn loss", float(out2.loss))
    # ---- config 3 size
    cfg = uc("configs/itransformer.yaml", ITR_FULL)
    torch.manual_seed(1)
    m = iTransformer(cfg, method_name="mlm", loss="poisson_nll", log_input=True)
    b = itr_batch(16, 100, 669, 1, rate=0.1)
    out, grads = run_ref(m, b, train=True, seed=77)
    names = [n for n, _ in m.named_parameters()]
    d = {"names": np.array(names), "seed": np.array(77), "out/loss": out.loss.detach().numpy(), "out/n_examples": out.n_examples.numpy(),
         "out/mask_bn": out.mask[:, 0, :].numpy().astype(np.uint8), "out/preds_rows": out.preds.detach().numpy()[:, ::10, ::8],
         "out/preds_sum": out.preds.detach().double().sum().numpy()}
    assert bool((out.mask == out.mask[:, :1, :]).all())     # neuron mode: the mask is constant over time
    d["param_sum"] = np.array([float(p.detach().double().sum()) for p in m.parameters()])
    d["grad_norm"] = np.array([float(grads[n].double().norm()) for n in names])
    d["grad_absmax"] = np.array([float(grads[n].abs().max()) for n in names])
    for n in names:
        if grads[n].numel() <= 4096:
            d[f"grad/{n}"] = grads[n].numpy()
        else:
            d[f"grad_slice/{n}"] = grads[n][:8].numpy()
    out16, g16 = run_ref_autocast_seeded(m, b, 77)
    l2, mx = grad_error_metrics(g16, grads)
    d["autocast/loss_rel"] = np.array(abs(float(out16.loss) - float(out.loss)) / abs(float(out.loss)))
    d["autocast/grad_l2_max"] = np.array(max(l2.values()))
    d["autocast/grad_maxabs_max"] = np.array(max(mx.values()))
    np.savez_compressed(os.path.join(HERE, "itransformer_config3.npz"), **d)
    print("itransformer_config3 loss", float(out.loss), "n", int(out.n_examples), "autocast loss rel", float(d["autocast/loss_rel"]),
          "grad l2 max", float(d["autocast/grad_l2_max"]))

if __name__ == "__main__":
    main()
