"""CPU tests of the host logic: config system, C-ABI surface, host collate, init/
checkpoint compatibility, LR schedule, and the data-parallel semantics on gloo."""
import os
import re
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llm_bci_b200 as lb  # noqa: E402
from llm_bci_b200 import _C  # noqa: E402
from llm_bci_b200.trainer import onecycle_cos_lr, linear_warmup_lr, shard_batch, get_model_inputs  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def test_config_merge_and_include():
    tr = lb.default_trainer_config()
    assert tr.model.model_class == "NDT1" and tr.model.encoder.embedder.stack.size == 32
    cfg = lb.update_config(tr.model, {"encoder": {"embedder": {"n_channels": 99}, "new": {"x": 1}}})
    assert cfg.encoder.embedder.n_channels == 99 and cfg.encoder.new.x == 1
    assert tr.model.encoder.embedder.n_channels == 256      # the default was not mutated
    cfg.encoder.embedder.n_channels = 7                     # main.py:230-231 style assignment sticks
    assert cfg["encoder"]["embedder"]["n_channels"] == 7
    kw = lb.config_from_kwargs({"a.b": "1.e-3", "a.c": "true", "d": "[1,2]", "e": "none", "f": "-3"})
    assert kw.a.b == 1e-3 and kw.a.c is True and kw.d == [1, 2] and kw.e is None and kw.f == -3
    with pytest.raises(KeyError):
        _ = cfg.encoder.masker.active


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ndt1_b200.h")).read()
    declared = set(re.findall(r"\b(ndt1_[a-z0-9_]+)\s*\(", header))
    declared -= {"ndt1_engine", "ndt1_config", "ndt1_tensors", "ndt1_batch", "ndt1_outputs"}
    assert declared, "no declarations parsed"
    L = _C.lib()                                             # loads without a GPU
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/ndt1_b200.h but not exported"
        assert name in _C.PROTOTYPES, f"{name} has no ctypes prototype"
    assert L.ndt1_abi_version() == _C.ABI_VERSION


def test_forward_signature_is_the_api():
    tr = lb.default_trainer_config()
    model = lb.NAME2MODEL[tr.model.model_class](tr.model, **tr.method.model_kwargs)
    assert get_model_inputs(model) == ["spikes", "spikes_mask", "spikes_timestamp", "spikes_lengths", "targets", "targets_lengths",
                                       "block_idx", "day_idx"]
    with pytest.raises(RuntimeError):                        # no CPU fallback
        model(torch.zeros(1, 64, 256), torch.ones(1, 64, dtype=torch.int64), torch.zeros(1, 64, dtype=torch.int64), torch.tensor([64]),
              torch.ones(1, 2, dtype=torch.int64), torch.tensor([2]))


def test_unknown_method_and_ssl_asserts():
    tr = lb.default_trainer_config()
    with pytest.raises(Exception):
        lb.NDT1(tr.model, method_name="nope")
    with pytest.raises(AssertionError):                      # stacked inputs can't pretrain (ndt1.py:482)
        mk = dict(active=True, mode="temporal", ratio=0.3, zero_ratio=1.0, random_ratio=1.0, expand_prob=0.0, max_timespan=1, regions=None, channels=None)
        lb.NDT1(lb.update_config(tr.model, {"encoder": {"masker": {"active": mk}}}), method_name="mlm", loss="poisson_nll", log_input=True)


def test_init_matches_reference_rng_order():
    g = dict(np.load(os.path.join(G, "ctc_small.npz")))
    cfg = lb.update_config(lb.default_model_config(), {"encoder": {
        "embedder": {"n_channels": 16, "input_dim": 16, "max_F": 64, "dropout": 0.0, "stack": {"active": True, "size": 32, "stride": 4}},
        "transformer": {"n_layers": 2, "hidden_size": 64, "n_heads": 4, "inter_size": 64, "dropout": 0.0},
        "smooth_and_noise": {"noise": False}}})
    torch.manual_seed(11)
    model = lb.NDT1(cfg, method_name="ctc", vocab_size=41, blank_id=0, zero_infinity=True)
    sd = model.state_dict()
    ref = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    assert list(sd.keys()) == list(ref.keys())
    for k in sd:
        assert np.array_equal(sd[k].numpy(), ref[k]), k     # bit-identical initialisation incl. fixup scaling


def test_checkpoint_files_and_round_trip():
    cfg = lb.update_config(lb.default_model_config(), {"encoder": {
        "embedder": {"n_channels": 16, "input_dim": 16, "max_F": 64}, "transformer": {"n_layers": 1, "hidden_size": 32, "n_heads": 2, "inter_size": 32}}})
    kw = dict(method_name="ctc", vocab_size=41, blank_id=0, zero_infinity=True)
    torch.manual_seed(0)
    a = lb.NDT1(cfg, **kw)
    with tempfile.TemporaryDirectory() as d:
        a.save_checkpoint(d)
        assert sorted(os.listdir(d)) == ["decoder.bin", "encoder.bin", "encoder_config.pth"]
        torch.manual_seed(5)
        b = lb.NDT1(cfg, **kw)
        b.load_checkpoint(d)
        for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
            assert ka == kb and torch.equal(va, vb)
        warm = lb.update_config(cfg, {"encoder": {"from_pt": d}})
        c = lb.NDT1(warm, **kw)                               # warm start, ndt1.py:468-476,503-504
        assert torch.equal(c.decoder[0].weight, a.decoder[0].weight)
        enc_cfg = torch.load(os.path.join(d, "encoder_config.pth"), weights_only=False)
        assert isinstance(enc_cfg, dict) and enc_cfg["embedder"]["n_channels"] == 16


def test_host_collate_matches_reference_golden():
    g = dict(np.load(os.path.join(G, "collate.npz")))
    rows = []
    order = ["spikes", "spikes_mask", "spikes_timestamp", "spikes_lengths", "targets", "targets_lengths", "sentence", "extra"]
    for i in range(3):
        r = {k[len(f"row{i}/"):]: v for k, v in g.items() if k.startswith(f"row{i}/")}
        r["sentence"] = "abc"
        rows.append({k: r[k] for k in order})
    mi = ["spikes", "spikes_mask", "spikes_timestamp", "spikes_lengths", "targets", "targets_lengths"]
    pads = {
        "right": {k: dict(dim=0, side="right", value=0, truncate=None, min_length=None) for k in ("spikes", "spikes_mask", "spikes_timestamp", "targets")},
        "left_trunc": {k: dict(dim=0, side="left", value=-1, truncate=40, min_length=None) for k in ("spikes", "spikes_mask", "spikes_timestamp", "targets")},
        "minlen": {k: dict(dim=0, side="right", value=0, truncate=64, min_length=60) for k in ("spikes", "spikes_mask", "spikes_timestamp")},
    }
    for name, pd in pads.items():
        padded, unused = lb.pad_collate_fn(rows, mi, pd)
        assert sorted(unused.keys()) == list(g[f"{name}/unused_keys"])
        for k, v in padded.items():
            if torch.is_tensor(v):
                assert np.array_equal(v.numpy(), g[f"{name}/{k}"]) and v.numpy().dtype == g[f"{name}/{k}"].dtype, (name, k)
            else:
                for i, vi in enumerate(v):
                    assert np.array_equal(vi.numpy(), g[f"{name}/{k}/{i}"])
    with pytest.raises(AssertionError):                      # min_length above truncate (datasets.py:201-205)
        lb.padded_array([np.zeros(3)], truncate=2, min_length=5)
    with pytest.raises(Exception):
        lb.padded_array([np.zeros(3)], side="up")


def test_format_ctc_quirk():
    g = dict(np.load(os.path.join(G, "index_ops.npz")))
    vocab = list(range(41))
    for i in range(5):
        assert lb.format_ctc(g[f"ctc_in/{i}"].tolist(), vocab, 0) == g[f"ctc_out/{i}"].tolist()
    assert lb.format_ctc([1, 0, 1], vocab, 0) == [1]        # `last` only moves on emission


def test_context_mask_matches_reference():
    g = dict(np.load(os.path.join(G, "index_ops.npz")))
    for key in [k for k in g if k.startswith("band/")]:
        _, cf, cb = key.split("/")
        assert np.array_equal(lb.create_context_mask(int(cf), int(cb), 12).numpy(), g[key]), key


def test_onecycle_matches_torch():
    lin = torch.nn.Linear(1, 1)
    opt = torch.optim.AdamW(lin.parameters(), lr=1e-3)
    sch = torch.optim.lr_scheduler.OneCycleLR(opt, total_steps=50, max_lr=1e-3, pct_start=0.2, anneal_strategy="cos", div_factor=25)
    from llm_bci_b200.trainer import onecycle_cos_beta1
    for step in range(50):
        assert abs(opt.param_groups[0]["lr"] - onecycle_cos_lr(step, 50, 1e-3, 0.2, 25)) < 1e-12
        # OneCycleLR's default cycle_momentum=True also drives AdamW's beta1 (0.95 -> 0.85 -> 0.95)
        assert abs(opt.param_groups[0]["betas"][0] - onecycle_cos_beta1(step, 50, 0.2)) < 1e-12
        opt.step()
        if step < 49:
            sch.step()


def test_shard_batch_is_contiguous_split():
    b = {"x": torch.arange(10), "y": torch.arange(20).view(10, 2), "s": "keep"}
    parts = [shard_batch(b, r, 2) for r in range(2)]
    assert torch.equal(torch.cat([p["x"] for p in parts]), b["x"]) and parts[0]["x"].tolist() == [0, 1, 2, 3, 4]
    assert parts[1]["s"] == "keep"


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import ndt1_oracle as O
    from test_oracle_golden import small_ctc_cfg, CTC_KW, sub, load
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = load("ctc_small.npz")
    params = {k: torch.from_numpy(v) for k, v in sub(g, "param").items()}
    batch = {k: torch.from_numpy(v) for k, v in sub(g, "batch").items()}
    batch = {k: torch.cat([v, v[:1]]) for k, v in batch.items()}           # 4 trials -> 2 per rank
    shard = shard_batch(batch, rank, world)
    _, grads = O.ndt1_loss_and_grads(params, small_ctc_cfg(), CTC_KW, shard, training=True)
    names = sorted(grads)
    flat = torch.cat([grads[n].reshape(-1) for n in names])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                             # bucketed sum ...
    flat /= world                                                          # ... then DDP's mean over ranks
    if rank == 0:
        _, full = O.ndt1_loss_and_grads(params, small_ctc_cfg(), CTC_KW, batch, training=True)
        ref = torch.cat([full[n].reshape(-1) for n in names]) / world       # (1/world) * grad of the global SUM loss
        q.put(float((flat - ref).abs().max() / ref.abs().max()))
    dist.destroy_process_group()


def test_data_parallel_semantics_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    err = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert err < 1e-5


def test_linear_schedule_matches_transformers():
    """optimizer.scheduler == "linear" (models/trainer.py:233-238)."""
    from transformers import get_linear_schedule_with_warmup
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=2e-3)
    sch = get_linear_schedule_with_warmup(opt, num_warmup_steps=round(0.15 * 40), num_training_steps=40)
    for step in range(40):
        assert abs(opt.param_groups[0]["lr"] - linear_warmup_lr(step, 40, 2e-3, 0.15)) < 1e-12, step
        opt.step(); sch.step()


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in llm_bci_b200/_C.py have the size and field offsets the C compiler gives include/ndt1_b200.h."""
    import ctypes
    import subprocess
    from llm_bci_b200 import _C
    probes = {
        "ndt1_config": (_C.Config, ["abi_version", "rope_theta", "context_forward", "factors_active", "method", "p_embed", "max_targets"]),
        "ndt1_tensors": (_C.Tensors, ["embed_w", "day_emb", "embed_w_day", "embed_b_day", "layer", "out_norm_w", "dec_b"]),
        "ndt1_batch": (_C.Batch, ["spikes", "targets_mask", "B", "encoder_only", "seed", "seed_ptr"]),
        "ndt1_outputs": (_C.Outputs, ["loss", "features"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "ndt1_b200.h")}"', "int main(void) {"]
    for cname, (_, fields) in probes.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  printf("abi %d layers %d days %d\\n", NDT1_ABI_VERSION, NDT1_MAX_LAYERS, NDT1_MAX_DAYS);', "  return 0;", "}"]
    src = tmp_path / "abi_probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_probe"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = dict(l.split(" ", 1) for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines())
    for cname, (ctype, fields) in probes.items():
        assert int(out[cname]) == ctypes.sizeof(ctype), cname
        for f in fields:
            assert int(out[f"{cname}.{f}"]) == getattr(ctype, f).offset, (cname, f)
    assert out["abi"] == f"{_C.ABI_VERSION} layers {_C.MAX_LAYERS} days {_C.MAX_DAYS}"


def test_arena_layout_big_and_small_regions_per_stage():
    """The flat parameter / gradient arena: one contiguous span per gradient stage, big (GEMM weight) region first and aligned to
    1024 floats at both ends, q|k|v weights and biases adjacent (the engine's single 3H x H weight-gradient GEMM relies on it)."""
    import llm_bci_b200.ndt1 as N
    tr = lb.default_trainer_config()
    m = lb.NDT1(tr.model, **tr.method.model_kwargs)
    m._ptable, m._pstruct = N._flat_param_table(m), None           # (CPU: skip the CUDA-only parameter checks of _params())
    offs, total = m._grad_offsets()
    lay = sorted(m._arena_layout, key=lambda e: e["stage"])
    assert [e["stage"] for e in lay] == list(range(7))
    covered = 0
    for e in lay:
        (blo, bhi), (slo, shi) = e["big"], e["small"]
        assert blo % 1024 == 0 and bhi % 1024 == 0 and bhi == slo and shi >= slo
        covered += shi - blo
    assert covered <= total and total - covered < 7 * 1024         # only alignment gaps are outside the stages
    H = 1024
    for layer in m.encoder.layers:
        a = layer.attn
        assert offs[id(a.key.weight)] == offs[id(a.query.weight)] + H * H and offs[id(a.value.weight)] == offs[id(a.key.weight)] + H * H
        assert offs[id(a.key.bias)] == offs[id(a.query.bias)] + H and offs[id(a.value.bias)] == offs[id(a.key.bias)] + H
    spans = sorted((offs[id(p)], offs[id(p)] + p.numel()) for p in m.parameters())
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))     # no two parameters overlap
    emb = next(e for e in lay if e["stage"] == 6)
    assert emb["big"][0] <= offs[id(m.encoder.embedder.stack_projection.weight)] < emb["big"][1]
    assert emb["small"][0] <= offs[id(m.encoder.embedder.embed_pos.weight)] < emb["small"][1]    # the position table is read in fp32: replicated region


def test_reference_install_manifest_matches_the_tree():
    """baseline/reference_manifest.json (committed) lists the sha256 of every reference file bench.py's reference arm runs; when the
    copy exists (build container, GPU box) it must match, i.e. the arm really times the UNMODIFIED reference."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import install_reference as ir
    man = json.load(open(ir.MANIFEST))["files"]
    assert "models/ndt1.py" in man and "models/masker.py" in man and "configs/trainer_ctc_ndt1.yaml" in man and len(man) >= 20
    if os.path.isdir(ir.DST):
        assert ir.verify(ir.DST)


def test_itransformer_plugin_surface_and_no_cpu_fallback():
    """SURVEY 8 f4 host side: registry entry, the reference's state_dict keys (names of tests/golden/itransformer_small.npz, produced
    by the unmodified reference), unbuilt variants raise, and a CPU tensor raises instead of falling back."""
    import numpy as np
    import pytest
    import torch
    import llm_bci_b200 as lb
    over = {"masker": {"main": {"ratio": 0.25}}, "encoder": {"embedder": {"dropout": 0.0, "max_n_bins": 20}, "hidden_size": 64, "n_heads": 4,
                                                                "n_layers": 2, "dropout": 0.0, "max_n_channels": 32, "embed_region": False}}
    model = lb.NAME2MODEL["iTransformer"](over, method_name="mlm", loss="poisson_nll", log_input=True)
    g = np.load(os.path.join(ROOT, "tests", "golden", "itransformer_small.npz"))
    assert [n for n, _ in model.named_parameters()] == list(g["names"])
    assert isinstance(model, lb.iTransformer) and set(lb.iTransformerOutput.__dataclass_fields__) == {"loss", "n_examples", "mask", "preds", "targets"}
    with pytest.raises(NotImplementedError):
        lb.iTransformer(over, method_name="ctc", vocab_size=41, blank_id=0, zero_infinity=True)
    assert lb.iTransformer(over, method_name="stat_behaviour", loss="xent", n_labels=3).decoder[2].out_features == 3
    with pytest.raises(NotImplementedError):
        lb.iTransformer({**over, "encoder": {**over["encoder"], "embedder": {"mode": "transformer", "max_n_bins": 20}}}, method_name="mlm",
                        loss="poisson_nll", log_input=True)
    x = torch.zeros(2, 20, 24)
    with pytest.raises(RuntimeError, match="GPU only"):
        model(spikes=x, spikes_mask=torch.ones(2, 20, dtype=torch.int64), spikes_timestamp=torch.arange(20)[None].expand(2, 20))
