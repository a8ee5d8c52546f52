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
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", 1324.9), d.get("hbm_gbs", 6551.7), "measured"
    return 1400.0, 6650.0, "fallback"


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


# --------------------------------------------------------------------------- reference arm / cpu baseline (oracle port)
def cpu_train_steps(n_trials: int, steps: int, warmup: int):
    """The reference algorithm (oracle/ndt1_oracle.py, validated against the unmodified reference) on the
    host CPU: train-mode forward + autograd backward + AdamW on `n_trials` trials per step."""
    import torch
    from oracle import ndt1_oracle as O
    from llm_bci_b200.config import default_trainer_config
    from llm_bci_b200.ndt1 import NDT1
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = default_trainer_config()
    torch.manual_seed(1)
    shell = NDT1(tr.model, **tr.method.model_kwargs)          # parameter container only (CPU); no kernels involved
    params = {k: v.detach().clone().requires_grad_(True) for k, v in shell.state_dict().items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-3, weight_decay=5e-5, eps=1e-8)
    batch = O.synthetic_ctc_batch(B=n_trials, T=T_BINS, N=N_CH, seed=1)
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
    sec = sum(times) / len(times)
    return n_trials / sec, sec, cores


WORKLOAD = ("NDT1 CTC train step (fwd+bwd+AdamW), BASELINE configs[1]: 32 trials x 1000 bins x 256 channels per GPU, "
            "5x1024 encoder, stack 32/4, 41 phonemes, dropout 0.4/0.2, noise on")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 4
    value, sec, cores = cpu_train_steps(n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "trials/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n} trials per step on the host CPU (a bounded sample of the same workload)"},
        "cpu_baseline": {"value": value, "unit": "trials/s", "cores": cores, "kind": "port",
                         "sample": f"{n} trials/step x {args.steps} steps, oracle port of the reference algorithm, torch CPU fp32"},
        "e2e": {"value": value, "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- this repo's arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import llm_bci_b200 as lb
    from llm_bci_b200 import _C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _C.lib()

    # synthetic batch generated WITHOUT the oracle (product path never imports oracle/)
    g = torch.Generator().manual_seed(1 + rank)
    B, T, N = B_PER_GPU, T_BINS, N_CH
    spikes = torch.randn(B, T, N, generator=g)
    lens = torch.randint(int(0.6 * T), T + 1, (B,), generator=g)
    lens[0] = T
    if args.fixed_length:                      # the all-full-length variant (no padded rows)
        lens[:] = T
    valid_row_frac = float(((lens - 32) // 4 + 1).sum()) / float(B * ((T - 32) // 4 + 1))    # stacked rows that are not padding
    t = torch.arange(T)[None, :]
    mask = (t < lens[:, None]).to(torch.int64)
    spikes = spikes * mask[:, :, None]
    ts = t.expand(B, T) * mask
    tl = torch.randint(20, 61, (B,), generator=g)
    S = int(tl.max())
    tg = torch.randint(1, 41, (B, S), generator=g) * (torch.arange(S)[None, :] < tl[:, None])
    host = dict(spikes=spikes, spikes_mask=mask, spikes_timestamp=ts.contiguous(), spikes_lengths=lens, targets=tg, targets_lengths=tl)
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    tr = lb.default_trainer_config()
    torch.manual_seed(1)
    model = lb.NAME2MODEL[tr.model.model_class](tr.model, **tr.method.model_kwargs, precision="bf16", max_batch=B, max_T=T).to(dev)
    opt = tr.optimizer
    trainer = lb.DataParallelTrainer(model, lr=opt.lr, wd=opt.wd, eps=opt.eps, scheduler=opt.scheduler, total_steps=10000,
                                     warmup_pct=opt.warmup_pct, div_factor=opt.div_factor)
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
    l0 = L.ndt1_launch_counter()
    ms_total = timed(step_resident, args.steps)
    launches = L.ndt1_launch_counter() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- e2e: pinned host batch -> H2D (prefetched on a copy stream) -> train_step -> loss to host
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "loss": 0.0}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for k, v in host.items():
                bufs[slot][k].copy_(v, non_blocking=True)
            ready[slot].record(copy_stream)

    for e in consumed:
        e.record()
    prefetch(0)

    # The loss of every step is copied to pinned host memory inside the timed region; the host READS it one step late
    # (after enqueueing the next step), the way an asynchronous logger does, so the device never idles on the host.
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]

    def step_e2e():
        i = state["i"]
        slot = i & 1
        prefetch(slot ^ 1)                                   # next step's inputs fly while this step computes
        torch.cuda.current_stream().wait_event(ready[slot])
        out = trainer.train_step(bufs[slot])
        consumed[slot].record()
        loss_host[slot].copy_(out.loss, non_blocking=True)   # device -> host read of the step's result
        loss_ready[slot].record()
        if i > 0:
            loss_ready[slot ^ 1].synchronize()
            state["loss"] = float(loss_host[slot ^ 1])
        state["i"] += 1

    def drain_e2e():
        last = (state["i"] - 1) & 1
        loss_ready[last].synchronize()
        state["loss"] = float(loss_host[last])

    for _ in range(3):
        step_e2e()
    drain_e2e()
    ms_e2e = timed(step_e2e, args.steps, after=drain_e2e) / args.steps
    e2e_value = world * B / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel family (tcgen05 GEMM), events around every launch, extra steps
    # (the weight-gradient stream is switched off here so that every launch is timed alone)
    peak_tf, peak_gbs, peak_src = load_peaks()
    L.ndt1_engine_set_overlap(model._engine, 0)
    step_resident()
    torch.cuda.synchronize()
    L.ndt1_profile_gemm_begin()
    prof_steps = 3
    for _ in range(prof_steps):
        step_resident()
    fl, pms, pn = _C.C.c_double(), _C.C.c_double(), _C.C.c_int64()
    torch.cuda.synchronize()
    L.ndt1_profile_gemm_end(_C.C.byref(fl), _C.C.byref(pms), _C.C.byref(pn))
    L.ndt1_engine_set_overlap(model._engine, 1)
    achieved = fl.value / (pms.value * 1e-3) / 1e12 if pms.value > 0 else 0.0
    traffic = None                       # DRAM bytes per launch from the committed ncu capture of this same command (tools/gemm_traffic.py)
    tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05/TMA bf16 GEMM, all shapes of the step)", "achieved": achieved,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "peak_source": peak_src + " sustained",
                "launches_per_step": pn.value / prof_steps, "gemm_ms_per_step": pms.value / prof_steps,
                "gemm_share_of_step": (pms.value / prof_steps) / ms_step,
                "step_tensor_frac": (B * FLOP_PER_TRIAL / (ms_step * 1e-3) / 1e12) / peak_tf}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, sec, cores = cpu_train_steps(8, 2, 1)
            cpu = {"value": v, "unit": "trials/s", "cores": cores, "kind": "port",
                   "sample": "8 trials/step, 1 warm-up + 2 timed steps of the oracle port (torch CPU fp32, train mode)"}
        line = {
            "metric": METRIC, "value": value, "unit": "trials/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "global_batch": world * B, "bins_per_sec": value * T, "parallelism": f"dp{world}",
                       "lengths": "all 1000 bins" if args.fixed_length else "U{600..1000} bins, right-padded (padded rows are computed, as in the reference)",
                       "valid_row_frac": valid_row_frac,
                       "l2": "per-step working set ~1.4 GB >> 126 MB L2; no flush needed", "loss": state["loss"]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "trials/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
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
    ap.add_argument("--fixed-length", action="store_true", help="every trial 1000 bins long (no padding)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
