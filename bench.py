#!/usr/bin/env python
"""NDT1 CTC training throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

A "step" is one pass of the hot path over one synthetic batch: NDT1 forward +
backward (+ gradient all-reduce for N > 1) + AdamW, train mode (dropout 0.4/0.2,
white/offset noise), on BASELINE.json configs[1]: 32 trials x 1000 bins x 256
channels per GPU (weak scaling; N = 8 is the global batch 256 of configs[3]).
Prints ONE JSON line (see the keys below).  `value` has the inputs resident in
HBM; `e2e` goes through the public API with pinned host inputs copied every
step and the loss read back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, T_BINS, N_CH = 32, 1000, 256
FLOP_PER_TRIAL = 62.18e9          # fwd+bwd algorithmic FLOPs per trial, SURVEY.md 8(d)
METRIC = "ndt1_ctc_train_trials_per_sec"


def load_peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written for this pod's B200s), else the profiling recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_burst": d.get("bf16_tflops", 1620.8), "bf16_sustained": d.get("bf16_tflops_sustained", 1324.9),
                "hbm": d.get("hbm_gbs", 6551.7), "source": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every 5 ms; nvidia-smi as a fallback)."""

    def __init__(self, index: int):
        self.index, self.sm, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        self._stop.set()
        self.thread.join(timeout=1)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(sm)}

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm"
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": ["sampled after the timed region (no NVML)"], "samples": 1}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


# --------------------------------------------------------------------------- the synthetic workload (shared by every arm)
def make_host_batch(B, T, N, seed, fixed_length=False):
    """BASELINE.md section 3 inputs for configs[1], generated WITHOUT the oracle (the product path never imports oracle/)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    spikes = torch.randn(B, T, N, generator=g)
    lens = torch.randint(int(0.6 * T), T + 1, (B,), generator=g)
    lens[0] = T
    if fixed_length:                      # the all-full-length variant (no padded rows)
        lens[:] = T
    t = torch.arange(T)[None, :]
    mask = (t < lens[:, None]).to(torch.int64)
    spikes = spikes * mask[:, :, None]
    ts = t.expand(B, T) * mask
    tl = torch.randint(20, 61, (B,), generator=g)
    S = int(tl.max())
    tg = torch.randint(1, 41, (B, S), generator=g) * (torch.arange(S)[None, :] < tl[:, None])
    return dict(spikes=spikes, spikes_mask=mask, spikes_timestamp=ts.contiguous(), spikes_lengths=lens, targets=tg, targets_lengths=tl)


# --------------------------------------------------------------------------- the UNMODIFIED reference (baseline/_ref) on CPU or under torch-CUDA
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_status():
    """(usable, why): baseline/_ref must hold the reference files byte for byte (tools/install_reference.py manifest)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import install_reference as ir
        if not os.path.isdir(REF_DIR):
            return False, "baseline/_ref is absent (run tools/install_reference.py in the build container)"
        if not ir.verify(REF_DIR):
            return False, "baseline/_ref does not match baseline/reference_manifest.json"
        return True, "baseline/_ref verified against the sha256 manifest of /root/reference"
    except Exception as e:                                     # pragma: no cover
        return False, f"reference check failed: {e}"


_REF = {}


def import_reference():
    """The reference's own modules, imported from baseline/_ref (CWD must be the tree: DEFAULT_CONFIG is CWD-relative, models/ndt1.py:17)."""
    if _REF:
        return _REF
    os.chdir(REF_DIR)
    sys.path.insert(0, REF_DIR)
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, "gaussian"):                   # scipy >= 1.13 dropped the alias used at models/ndt1.py:87
        scipy.signal.gaussian = scipy.signal.windows.gaussian
    from utils.config_utils import update_config
    from models.ndt1 import NDT1
    _REF.update(update_config=update_config, NDT1=NDT1)
    return _REF


def reference_train_steps(device, n_trials, steps, warmup, autocast=None, budget_s=None, seed=1):
    """The reference's training step (models/trainer.py:336-343: forward, backward, AdamW step, zero_grad) on its own NDT1 built
    from its own configs/trainer_ctc_ndt1.yaml, train mode (dropout 0.4 / 0.2, noise on), on the synthetic batch of this bench.
    Returns (trials/s from the MEDIAN step, median seconds, steps actually timed)."""
    import torch
    R = import_reference()
    os.chdir(REF_DIR)                                           # (configs/*.yaml are CWD-relative in the reference)
    cfg = R["update_config"]("configs/trainer_ctc_ndt1.yaml", None)
    torch.manual_seed(1)
    model = R["NDT1"](cfg.model, **cfg.method.model_kwargs).to(device).train()
    os.chdir(ROOT)
    o = cfg.optimizer
    opt = torch.optim.AdamW(model.parameters(), lr=o.lr, weight_decay=o.wd, eps=o.eps)
    batch = {k: v.to(device) for k, v in make_host_batch(n_trials, T_BINS, N_CH, seed).items()}
    cuda = torch.device(device).type == "cuda"
    times, t_start = [], time.perf_counter()
    for i in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = {k: v.clone() for k, v in batch.items()}            # (the module may mutate its input)
        if autocast is not None:
            with torch.autocast(torch.device(device).type, dtype=autocast):
                out = model(**b)
        else:
            out = model(**b)
        out.loss.backward()
        opt.step()
        opt.zero_grad()
        if cuda:
            torch.cuda.synchronize()
        else:
            float(out.loss)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
            if budget_s is not None and len(times) >= 3 and time.perf_counter() - t_start > budget_s:
                break
    times.sort()
    med = times[len(times) // 2]
    return n_trials / med, med, len(times)


def port_train_steps(n_trials: int, steps: int, warmup: int):
    """Fallback when baseline/_ref is absent: the oracle port of the reference algorithm (oracle/ndt1_oracle.py) on the host CPU."""
    import torch
    from oracle import ndt1_oracle as O
    from llm_bci_b200.config import default_trainer_config
    from llm_bci_b200.ndt1 import NDT1
    tr = default_trainer_config()
    torch.manual_seed(1)
    shell = NDT1(tr.model, **tr.method.model_kwargs)          # parameter container only (CPU); no kernels involved
    params = {k: v.detach().clone().requires_grad_(True) for k, v in shell.state_dict().items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-3, weight_decay=5e-5, eps=1e-8)
    batch = make_host_batch(n_trials, T_BINS, N_CH, 1)
    ds = {"torch_dropout": {"embed": 0.2, "transformer": 0.4}}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        noise = {"white": torch.randn(n_trials, T_BINS, N_CH), "offset": torch.randn(n_trials, 1, N_CH)}
        out = O.ndt1_forward(params, tr.model, tr.method.model_kwargs, **batch, training=True, noise=noise, drop_scales=ds)
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return n_trials / med, med, len(times)


def cpu_baseline(steps, warmup, budget_s):
    """The reference's CPU implementation of the path on ALL host cores, on the bench's own 32-trial batch."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ok, why = reference_status()
    if ok:
        v, sec, n = reference_train_steps("cpu", B_PER_GPU, steps, warmup, budget_s=budget_s)
        kind, what = "reference", "unmodified reference module from baseline/_ref (models/ndt1.py:523-589 + AdamW, models/trainer.py:336-343)"
    else:
        v, sec, n = port_train_steps(B_PER_GPU, steps, warmup)
        kind, what = "port", f"oracle port ({why})"
    return {"value": v, "unit": "trials/s", "cores": cores, "kind": kind, "ms_per_step": sec * 1e3,
            "sample": f"{B_PER_GPU} trials/step (the full configs[1] batch), {warmup} warm-up + {n} timed steps, median; {what}; torch CPU fp32, train mode"}


WORKLOAD = ("NDT1 CTC train step (fwd+bwd+AdamW), BASELINE configs[1]: 32 trials x 1000 bins x 256 channels per GPU, "
            "5x1024 encoder, stack 32/4, 41 phonemes, dropout 0.4/0.2, noise on")


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, same config / metric / unit as
    this repo's arm, every step the SAME 32-trial batch (bounded to ~4 minutes: fewer timed steps are reported as such)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args.steps, max(1, min(args.warmup, 2)), budget_s=200.0)
    import re
    n_timed = int(re.search(r"\+ (\d+) timed", cb["sample"]).group(1))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "trials/s", "n_gpus": args.gpus, "steps": n_timed,
        "steps_requested": args.steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU, "parallelism": "host CPU, one process",
                   "lengths": "U{600..1000} bins, right-padded (padded rows are computed, as in the reference)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def gpu_eager_leg(dev):
    """SURVEY 2.2 / BASELINE.md section 3: the kernel to beat is the reference module under PyTorch-eager on the SAME B200
    (cuBLAS / SDPA / ATen; none of this repo's kernels), fp32 and bf16 autocast, same batch, fwd + bwd + AdamW."""
    import torch
    ok, why = reference_status()
    if not ok:
        return {"unavailable": why}
    out = {"what": "unmodified reference NDT1 + torch.optim.AdamW under torch-CUDA eager on this GPU, same 32-trial batch, 3 warm-up + 10 timed steps, median"}
    for name, ac in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        try:
            v, sec, n = reference_train_steps(str(dev), B_PER_GPU, 10, 3, autocast=ac)
            out[name] = {"value": v, "unit": "trials/s", "ms_per_step": sec * 1e3}
        except Exception as e:                                 # the eager leg must never take the bench line down
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    return out


def bind_to_gpu_numa_node(local: int):
    """Pin this rank to the CPU cores NVML reports as local to its GPU BEFORE any pinned host memory is allocated (first touch
    then places the staging buffers on the GPU's NUMA node): with 8 ranks copying 33 MB per step each, host-to-device copies
    that cross the socket interconnect are what bounds the end-to-end number.  Returns the cores, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------- this repo's arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import llm_bci_b200 as lb
    from llm_bci_b200 import _C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = bind_to_gpu_numa_node(local) if (world > 1 and not args.no_numa_bind) else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")       # (see llm_bci_b200/trainer.py: NVLS all-reduce serialises with the input copies)
        dist.init_process_group("nccl", device_id=dev)
    L = _C.lib()

    B, T, N = B_PER_GPU, T_BINS, N_CH
    host = make_host_batch(B, T, N, 1 + rank, args.fixed_length)
    lens = host["spikes_lengths"]
    valid_row_frac = float(((lens - 32) // 4 + 1).sum()) / float(B * ((T - 32) // 4 + 1))    # stacked rows that are not padding
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    tr = lb.default_trainer_config()
    torch.manual_seed(1)
    model = lb.NAME2MODEL[tr.model.model_class](tr.model, **tr.method.model_kwargs, precision="bf16", max_batch=B, max_T=T).to(dev)
    opt = tr.optimizer
    trainer = lb.DataParallelTrainer(model, lr=opt.lr, wd=opt.wd, eps=opt.eps, scheduler=opt.scheduler, total_steps=10000,
                                     warmup_pct=opt.warmup_pct, div_factor=opt.div_factor, use_graph=not args.no_graph)
    resident = {k: v.to(dev) for k, v in host.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()                                          # e.g. read the last step's result on the host
        trainer.synchronize()                                # the last step's optimizer tail (side stream) is inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- value: inputs resident in HBM
    step_resident = lambda: trainer.train_step(resident)
    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = L.ndt1_launch_counter() + trainer.replayed_launches
    ms_total = timed(step_resident, args.steps)
    launches = L.ndt1_launch_counter() + trainer.replayed_launches - l0        # kernels of this library executed in the timed region (eager or replayed)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- e2e: pinned host batch -> H2D (prefetched on a copy stream) -> train_step -> loss to host
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "loss": 0.0}

    copy_t0 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]     # how long the copies take INSIDE the loop (next to the
    copy_t1 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]     # step's kernels and collectives), read one use later
    copy_ms = []

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            if copy_t1[slot].query() and state["i"] > 2:
                copy_ms.append(copy_t0[slot].elapsed_time(copy_t1[slot]))
            copy_t0[slot].record(copy_stream)
            for k, v in host.items():
                if args.e2e_copy_frac < 1.0 and k == "spikes":   # diagnosis only (invalid as a result): is the loop bound by the BYTES copied?
                    nb = max(1, int(v.shape[0] * args.e2e_copy_frac))
                    bufs[slot][k][:nb].copy_(v[:nb], non_blocking=True)
                else:
                    bufs[slot][k].copy_(v, non_blocking=True)
            copy_t1[slot].record(copy_stream)
            ready[slot].record(copy_stream)

    for e in consumed:
        e.record()
    prefetch(0)

    # The loss of every step is copied to pinned host memory inside the timed region; the host READS it one step late
    # (after enqueueing the next step), the way an asynchronous logger does, so the device never idles on the host.
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]

    diag = {"copy": True, "sync": True}                       # (--e2e-diagnose switches parts of the loop off to name the limiter)

    def step_e2e():
        i = state["i"]
        slot = i & 1
        if diag["copy"]:
            if not args.prefetch_after:
                prefetch(slot ^ 1)                           # next step's inputs fly while this step computes
            torch.cuda.current_stream().wait_event(ready[slot])
        out = trainer.train_step(bufs[slot])
        consumed[slot].record()
        if diag["copy"] and args.prefetch_after:
            prefetch(slot ^ 1)
        loss_host[slot].copy_(out.loss, non_blocking=True)   # device -> host read of the step's result
        loss_ready[slot].record()
        if i > 0 and diag["sync"]:
            loss_ready[slot ^ 1].synchronize()
            state["loss"] = float(loss_host[slot ^ 1])
        state["i"] += 1

    def drain_e2e():
        last = (state["i"] - 1) & 1
        loss_ready[last].synchronize()
        state["loss"] = float(loss_host[last])

    for _ in range(6):                                       # (both input buffer sets: seen once, captured, replayed)
        step_e2e()
    drain_e2e()
    copy_ms.clear()
    ms_e2e = timed(step_e2e, args.steps, after=drain_e2e) / args.steps
    e2e_value = world * B / (ms_e2e * 1e-3)
    h2d_in_loop = float(sorted(copy_ms)[len(copy_ms) // 2]) if copy_ms else None
    e2e_diag = None
    if args.e2e_diagnose:
        e2e_diag = {}
        for name, c, sy in (("no_h2d_copy", False, True), ("no_loss_sync", True, False), ("neither", False, False)):
            diag["copy"], diag["sync"] = c, sy
            for _ in range(4):
                step_e2e()
            drain_e2e()
            e2e_diag[name + "_ms_per_step"] = timed(step_e2e, args.steps, after=drain_e2e) / args.steps
        diag["copy"], diag["sync"] = True, True
    # diagnosis: the host-to-device copies of a step ALONE (all ranks at once, nothing computing): when this approaches the step
    # time the end-to-end number is bound by the host's memory / PCIe path, not by anything on the GPU
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        h0.record()
        for _ in range(10):
            for k, v in host.items():
                bufs[0][k].copy_(v, non_blocking=True)
        h1.record()
    barrier()
    h2d_ms = torch.tensor([h0.elapsed_time(h1) / 10], device=dev)
    if world > 1:
        dist.all_reduce(h2d_ms, op=dist.ReduceOp.MAX)
    h2d_ms = float(h2d_ms)

    # ---- roofline: CUDA events around EVERY launch of the library, on its own stream, over extra steps.  The weight-gradient
    # stream and the overlapped optimizer are switched off here so that every launch is timed alone (its share of a real,
    # overlapped step is what the ncu launch list under profiles/ shows).
    peaks = load_peaks()
    L.ndt1_engine_set_overlap(model._engine, 0)
    trainer.serialize = True
    graph_was, trainer.use_graph = trainer.use_graph, False      # (events between launches: eager steps)
    step_resident()
    torch.cuda.synchronize()
    prof_steps = 3
    _C.profile_begin()
    for _ in range(prof_steps):
        step_resident()
    trainer.synchronize()
    torch.cuda.synchronize()
    prof = _C.profile_end()
    L.ndt1_engine_set_overlap(model._engine, 1)
    trainer.serialize = False
    trainer.use_graph = graph_was

    def family(pred):
        rows = [r for r in prof if pred(r["name"])]
        return {"launches": sum(r["launches"] for r in rows), "ms": sum(r["ms"] for r in rows), "flops": sum(r["flops"] for r in rows),
                "bytes": sum(r["bytes"] for r in rows)}

    def entry(kernel, f, bound):
        if f["launches"] == 0 or f["ms"] <= 0:
            return None
        sec = f["ms"] * 1e-3
        if bound == "tensor":
            ach, peak, unit = f["flops"] / sec / 1e12, peaks["bf16_burst"], "TFLOP/s"
        else:
            ach, peak, unit = f["bytes"] / sec / 1e9, peaks["hbm"], "GB/s"
        return {"kernel": kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "launches_per_step": f["launches"] / prof_steps, "us_per_launch": 1e3 * f["ms"] / f["launches"],
                "ms_per_step": f["ms"] / prof_steps}

    gemm = family(lambda n: "gemm_tc_kernel" in n)
    achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    traffic, traffic_src = None, None     # DRAM bytes per launch: from the committed ncu --set full capture of this same command (tools/gemm_traffic.py)
    for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            traffic, traffic_src = json.load(open(tpath)).get("dram_bytes_per_launch"), "profiles/" + name
            break
    gemm_ms_step = gemm["ms"] / prof_steps
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05/TMA bf16 GEMM, all shapes of the step)", "achieved": achieved,
                "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_burst"],
                "peak_source": peaks["source"] + " burst (each launch is timed alone between events)",
                "peak_sustained": peaks["bf16_sustained"], "frac_sustained": achieved / peaks["bf16_sustained"],
                "traffic": traffic, "traffic_source": traffic_src,
                "launches_per_step": gemm["launches"] / prof_steps, "gemm_ms_per_step": gemm_ms_step,
                "gemm_share_of_step": gemm_ms_step / ms_step,
                "step_tensor_frac": (B * FLOP_PER_TRIAL / (ms_step * 1e-3) / 1e12) / peaks["bf16_burst"],
                "step_tensor_frac_sustained": (B * FLOP_PER_TRIAL / (ms_step * 1e-3) / 1e12) / peaks["bf16_sustained"]}
    more = [entry("attn_tc_fwd_kernel (tcgen05 attention forward, algorithmic 2 contractions)", family(lambda n: "attn_tc_fwd" in n), "tensor"),
            entry("attn_tc_bwd_q_kernel (dP, dQ; recomputed S not counted)", family(lambda n: "attn_tc_bwd_q" in n), "tensor"),
            entry("attn_tc_bwd_kv3_kernel (dV, dK; recomputed S, dP not counted)", family(lambda n: "attn_tc_bwd_kv" in n), "tensor"),
            entry("attention, all three kernels", family(lambda n: "attn_tc_" in n), "tensor"),
            entry("ln_fwd_rows_kernel", family(lambda n: "ln_fwd" in n), "hbm"),
            entry("ln_bwd_rows_kernel", family(lambda n: "ln_bwd" in n), "hbm"),
            entry("adamw_fused_kernel", family(lambda n: "adamw_fused" in n), "hbm"),
            entry("smooth_noise_vec_kernel", family(lambda n: "smooth_noise" in n), "hbm"),
            entry("ctc_kernel + ctc_posterior_kernel (latency-bound sweeps)", family(lambda n: n.startswith("ctc_") or "::ctc_" in n), "hbm"),
            entry("colsum8_kernel", family(lambda n: "colsum8" in n), "hbm"),
            entry("grad_prep_kernel", family(lambda n: "grad_prep" in n), "hbm")]
    more = [m for m in more if m is not None]
    all_ms = sum(r["ms"] for r in prof) / prof_steps
    kernel_table = sorted(({"kernel": r["name"][:90], "launches_per_step": r["launches"] / prof_steps, "ms_per_step": r["ms"] / prof_steps}
                           for r in prof), key=lambda r: -r["ms_per_step"])[:12]

    if rank == 0:
        cpu, eager = None, None
        if world == 1 and not args.no_cpu_baseline:
            # free the GPU arm before the host legs: the eager leg needs HBM, the CPU leg the cores
            cpu = cpu_baseline(steps=3, warmup=1, budget_s=120.0)
            if not args.no_gpu_eager:
                eager = gpu_eager_leg(dev)
        line = {
            "metric": METRIC, "value": value, "unit": "trials/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "global_batch": world * B, "bins_per_sec": value * T, "parallelism": f"dp{world}",
                       "lengths": "all 1000 bins" if args.fixed_length else "U{600..1000} bins, right-padded (padded rows are computed, as in the reference)",
                       "valid_row_frac": valid_row_frac,
                       "l2": "per-step working set ~1.4 GB >> 126 MB L2; no flush needed", "loss": state["loss"],
                       "cuda_graph": bool(trainer.use_graph)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "trials/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "h2d_ms_per_step_in_loop": h2d_in_loop, "h2d_ms_per_step_alone": h2d_ms, "h2d_gbs_per_gpu_alone": h2d_bytes / (h2d_ms * 1e-3) / 1e9,
                    "numa_bound_cores": None if numa_cpus is None else len(numa_cpus), "diagnose": e2e_diag,
                    "nccl_nvls": os.environ.get("NCCL_NVLS_ENABLE") if world > 1 else None,
                    **({"INVALID_copy_frac": args.e2e_copy_frac} if args.e2e_copy_frac < 1.0 else {})},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_more": more,
            "kernel_ms_per_step_alone": {"sum": all_ms, "top": kernel_table},
            "cpu_baseline": cpu,
            "gpu_eager": eager,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-under-torch-CUDA comparison leg")
    ap.add_argument("--fixed-length", action="store_true", help="every trial 1000 bins long (no padding)")
    ap.add_argument("--no-graph", action="store_true", help="eager steps (no whole-step CUDA graph)")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not pin the ranks to their GPU's NUMA node")
    ap.add_argument("--prefetch-after", action="store_true", help="e2e: submit the next step's H2D copies after this step's work instead of before")
    ap.add_argument("--e2e-copy-frac", type=float, default=1.0, help="diagnosis only: copy this fraction of the spikes per step (the result is then not a valid e2e number)")
    ap.add_argument("--e2e-diagnose", action="store_true", help="also time the end-to-end loop without its H2D copies / without its per-step loss read")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
