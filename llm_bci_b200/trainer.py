"""The caller side of the hot path: model registry and the data-parallel step.

Mirrors what the reference's ``Trainer`` does around ``model(**model_inputs)``
(models/trainer.py:32-36, 149-150, 161-171, 227-262, 332-354), without Accelerate:

* ``NAME2MODEL`` -- the registry the trainer config's ``model_class`` selects from;
* ``DataParallelTrainer`` -- one process per GPU, trials sharded across ranks
  (``split_batches=True``: the global batch is split contiguously), every rank
  holds a full fp32 replica, gradients are all-reduced in buckets over NCCL on
  a side stream that waits on the engine's per-stage events (so communication
  overlaps the rest of the backward) and divided by the world size (DDP's mean
  of per-rank gradients of the per-rank SUM loss), then one fused AdamW kernel
  updates the flat parameter buffer.  OneCycle-cosine / linear / step learning
  rate as in models/trainer.py:239-253.
"""
from __future__ import annotations

import inspect
import math
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _C
from .ndt1 import NDT1

# NVLS (in-switch reduction) all-reduce and a concurrent host-to-device copy serialise on an 8-GPU NVSwitch box: with the next
# batch's 33 MB copy in flight the step takes 4.89 ms instead of 3.76 (the copy itself still takes its 1.43 ms); with NVLS off NCCL
# reduces over the same NVLinks with its ring / tree kernels, the resident step is unchanged (68.3 k vs 68.1 k trials/s) and the
# end-to-end step is 3.84 ms (profiles/r02_bench_8gpu_nvls_{on,off}.json).  So this package asks for NVLS off unless the caller's
# environment says otherwise; it has to be in the environment before the NCCL communicator is created.
os.environ.setdefault("NCCL_NVLS_ENABLE", "0")

def _bci(*args, **kwargs):
    from .bci import BCI
    return BCI(*args, **kwargs)


def _itransformer(*args, **kwargs):
    from .itransformer import iTransformer
    return iTransformer(*args, **kwargs)


NAME2MODEL = {"NDT1": NDT1, "BCI": _bci, "iTransformer": _itransformer}     # (models/trainer.py:36; PatchTST: DESIGN.md section 7)


def get_model_inputs(model) -> List[str]:
    """Names of the forward parameters = collate keys (models/trainer.py:161-171)."""
    return [k for k in inspect.signature(model.forward).parameters.keys() if k not in ("noise", "masker_draws")]


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Contiguous split of every tensor along dim 0 (Accelerate ``split_batches=True``, trainer.py:79)."""
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            n = v.shape[0]
            per = n // world
            out[k] = v[rank * per:(rank + 1) * per] if rank < world - 1 else v[rank * per:]
        else:
            out[k] = v
    return out


def onecycle_cos_lr(step: int, total_steps: int, max_lr: float, pct_start: float, div_factor: float,
                    final_div_factor: float = 1e4) -> float:
    """torch OneCycleLR(anneal_strategy='cos') value at optimizer step ``step`` (trainer.py:239-246)."""
    initial, min_lr = max_lr / div_factor, max_lr / div_factor / final_div_factor
    up_end = float(pct_start * total_steps) - 1.0
    down_end = float(total_steps) - 1.0

    def cos(a, b, pct):
        return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1.0)

    if step <= up_end or up_end >= down_end:
        return cos(initial, max_lr, step / up_end) if up_end > 0 else max_lr
    return cos(max_lr, min_lr, (step - up_end) / (down_end - up_end))


def onecycle_cos_beta1(step: int, total_steps: int, pct_start: float, base_momentum: float = 0.85, max_momentum: float = 0.95) -> float:
    """torch OneCycleLR's default ``cycle_momentum=True`` ALSO drives Adam's beta1 (trainer.py:239-246 builds it with the
    defaults): max_momentum -> base_momentum while the rate climbs, back to max_momentum while it anneals (cosine)."""
    up_end = float(pct_start * total_steps) - 1.0
    down_end = float(total_steps) - 1.0

    def cos(a, b, pct):
        return b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1.0)

    if step <= up_end or up_end >= down_end:
        return cos(max_momentum, base_momentum, step / up_end) if up_end > 0 else base_momentum
    return cos(base_momentum, max_momentum, (step - up_end) / (down_end - up_end))


def linear_warmup_lr(step: int, total_steps: int, lr: float, warmup_pct: float) -> float:
    """transformers.get_linear_schedule_with_warmup value at optimizer step ``step`` (trainer.py:233-238): linear ramp over
    round(warmup_pct * total_steps) steps, then linear decay to 0 at total_steps."""
    warmup = round(warmup_pct * total_steps)
    if step < warmup:
        return lr * float(step) / float(max(1, warmup))
    return lr * max(0.0, float(total_steps - step) / float(max(1, total_steps - warmup)))


class DataParallelTrainer:
    """Owns the flat parameter / gradient / AdamW-state buffers of one rank."""

    def __init__(self, model: NDT1, lr: float = 1e-3, wd: float = 5e-5, eps: float = 1e-8, betas=(0.9, 0.999),
                 scheduler: Optional[str] = None, total_steps: int = 1, warmup_pct: float = 0.0, div_factor: float = 25.0,
                 gamma: float = 0.95, process_group=None, bucket_layers: int = 1,
                 gradient_accumulation_steps: int = 1, loss_scale: Optional[float] = None, use_graph: Optional[bool] = None,
                 shard_optimizer: Optional[bool] = None):
        self.model = model
        self.lr, self.wd, self.eps, self.betas = lr, wd, eps, betas
        self.scheduler, self.total_steps, self.warmup_pct, self.div_factor, self.gamma = scheduler, total_steps, warmup_pct, div_factor, gamma
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.step_count = 0        # optimizer updates done so far
        self.epoch = 0             # epochs ended so far (end_epoch(); only the "step" schedule reads it)
        # gradient accumulation as the reference runs it (models/trainer.py:333-349): the loss is scaled by 1 / steps, the optimizer
        # steps on micro-batch 1, 1 + steps, ... and the micro-batches in between only accumulate (no_sync: no all-reduce)
        self.accum = max(1, int(gradient_accumulation_steps))
        self.global_step = 1
        self._loss_scale = float(loss_scale) if loss_scale is not None else 1.0 / self.accum
        self.bucket_layers = max(1, bucket_layers)
        offs, total = model._grad_offsets()
        table, _ = model._params()
        dev = next(model.parameters()).device
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for slot, p in table:
                if p is None:
                    continue
                o = offs[id(p)]
                self.flat_param[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat_param[o:o + p.numel()].view_as(p)     # parameters become views of the flat buffer
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        layout = sorted(model._arena_layout, key=lambda e: e["stage"])      # gradient stages in completion order (head first)
        model.invalidate_param_cache()
        model._grad_offsets()                                             # (re-derives the layout the cache reset dropped)
        # bf16 mode: the fused AdamW keeps a bf16 copy of the arena current, and the engine reads its weights from it
        self.shadow = None
        if dev.type == "cuda" and getattr(model, "precision", "") == "bf16":
            self.shadow = torch.empty(total, dtype=torch.bfloat16, device=dev)
            self.refresh_shadow()
            model.set_weight_shadow(self.flat_param, self.shadow)
        # one bucket per gradient stage: (stage, lo, hi) = [big region | small region] of that stage, contiguous in the arena
        self.buckets = [(e["stage"], e["big"][0], e["small"][1]) for e in layout]
        self._big = {e["stage"]: e["big"] for e in layout}
        # Sharded optimizer (world > 1, bf16 mode): the BIG region of every stage (the GEMM weights, read by the engine through the
        # bf16 shadow) is reduce-scattered instead of all-reduced, each rank runs AdamW on its 1 / world slice only (optimizer HBM
        # traffic / world) and the bf16 shadow slices are all-gathered (NVLink bytes 2 x 4 B -> 4 B + 2 B per parameter).  The small
        # region (biases, LayerNorm affines, position table: read in fp32) stays all-reduced and replicated.  The fp32 masters of
        # the big regions are then current on their owner rank only: gather_parameters() (called by NDT1.save_checkpoint through
        # model._gather_hook) all-gathers them before anything reads parameters outside the engine.
        # Measured on 8 B200s (profiles/r02_bench_8gpu*.json): 65.9 k trials/s sharded against 67.2 k with the plain all-reduce and the
        # replicated AdamW -- three collectives per stage cost more than the optimizer traffic they save while AdamW already
        # hides under the backward -- so it is OFF unless asked for (shard_optimizer=True or NDT1_SHARD_OPTIMIZER=1).
        if shard_optimizer is None:
            shard_optimizer = os.environ.get("NDT1_SHARD_OPTIMIZER", "0") == "1"
        self.shard = bool(shard_optimizer) and self.world > 1 and self.shadow is not None and self.world <= 16
        self.rank = dist.get_rank(process_group) if self.world > 1 else 0
        self._params_gathered = True
        if self.shard:
            model._gather_hook = self.gather_parameters
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.serialize = False     # measurement aid: run the exchange and the optimizer AFTER the backward instead of under it
        # Whole-step CUDA graph (the loop at models/trainer.py:332-349): prologue + forward + backward of a batch
        # geometry are captured once and replayed; the per-step Philox keys live in device memory (self._seeds_dev), the gradient
        # exchange and the bucketed optimizer stay eager on the side stream, ordered against the graph by external event nodes.
        if use_graph is None:
            use_graph = os.environ.get("NDT1_GRAPH", "1") != "0"
        self.use_graph = bool(use_graph) and dev.type == "cuda"
        self._graphs: Dict[tuple, dict] = {}
        self.replayed_launches = 0          # kernels executed through graph replays (bench.py gpu_launches)
        self._params_ev = None
        if dev.type == "cuda":
            ev = _C._p()
            _C.check(_C.lib().ndt1_event_create(_C.C.byref(ev)), "ndt1_event_create")
            self._params_ev = ev
            model._param_event = ev
            _C.check(_C.lib().ndt1_event_record(ev, torch.cuda.current_stream().cuda_stream), "ndt1_event_record")
            self._seeds_dev = torch.zeros(2, dtype=torch.int64, device=dev)
            self._seeds_ring = torch.zeros(256, 2, dtype=torch.int64).pin_memory()
            self._seed_i = 0
        model._param_stream = self.comm_stream
        self._ones = torch.full((), self._loss_scale, dtype=torch.float32, device=dev)     # d(loss) handed to the backward

    # ------------------------------------------------------------------
    def current_lr(self) -> float:
        """Learning rate of the NEXT optimizer update.  The reference calls ``lr_scheduler.step()`` after ``optimizer.step()``
        (models/trainer.py:340-342), so its k-th update (k = 1, 2, ...) runs at schedule(k - 1): the schedules are evaluated
        at the number of updates already done.  "step" is StepLR(step_size=1, gamma) stepped once per EPOCH
        (models/trainer.py:253, 418-419): lr * gamma ** epochs_ended, see ``end_epoch``."""
        if self.scheduler == "cosine":
            return onecycle_cos_lr(self.step_count, self.total_steps, self.lr, self.warmup_pct, self.div_factor)
        if self.scheduler == "linear":
            return linear_warmup_lr(self.step_count, self.total_steps, self.lr, self.warmup_pct)
        if self.scheduler == "step":
            return self.lr * (self.gamma ** self.epoch)
        return self.lr

    def current_beta1(self) -> float:
        """Adam's beta1 of the NEXT update: cycled by the OneCycle ("cosine") schedule like the reference's, constant otherwise."""
        if self.scheduler == "cosine":
            return onecycle_cos_beta1(self.step_count, self.total_steps, self.warmup_pct)
        return self.betas[0]

    def end_epoch(self) -> None:
        """Call once after every pass over the training set (models/trainer.py:418-419)."""
        self.epoch += 1

    def all_reduce_gradients(self) -> None:
        """Bucketed all-reduce (sum) on the side stream; each bucket waits for its backward stage."""
        if self.world == 1:
            return
        L = _C.lib()
        with torch.cuda.stream(self.comm_stream):
            for stage, lo, hi in self.buckets:
                _C.check(L.ndt1_engine_wait_stage(self.model._engine, stage, self.comm_stream.cuda_stream), "ndt1_engine_wait_stage")
                dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
        torch.cuda.current_stream().wait_stream(self.comm_stream)

    def synchronize(self) -> None:
        """Make the current stream see every pending parameter update (before reading parameters outside the model)."""
        if getattr(self, "comm_stream", None) is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def refresh_shadow(self) -> None:
        """Call after changing parameters behind the trainer's back (e.g. load_checkpoint)."""
        self.synchronize()
        if self.shadow is not None:
            self.shadow.copy_(self.flat_param)

    def train_step(self, batch: Dict[str, torch.Tensor]):
        """forward + backward + gradient all-reduce + AdamW on this rank's shard.  Returns the NDT1Output.
        (The gradient arena starts at zero and the fused optimizer step clears it again, like zero_grad.)

        The optimizer does not wait for the end of the backward: every gradient bucket (head, layers L-1..0, embedder) is
        all-reduced and then UPDATED on the side stream as soon as the engine signals that its gradients are complete, so
        the HBM-bound AdamW pass runs under the tensor-bound rest of the backward.  Safe because a layer's weights are no
        longer read once its gradients are complete."""
        m = self.model
        m.train()
        sync = (self.global_step - 1) % self.accum == 0
        self.global_step += 1
        if self._params_ev is not None:
            # this step's Philox keys go to device memory (eager and captured steps alike: same draws for the same torch seed)
            self._use_device_seeds(True)
            self._push_seeds()
        try:
            if self.use_graph and self._graphable(batch):
                out = self._step_graphed(batch)
            else:
                out = m.forward_backward(batch, self.flat_grad, self._ones)
        finally:
            self._use_device_seeds(False)       # (forwards outside the trainer draw their own keys again)
        if not sync:                      # accumulation micro-step: gradients stay in the arena, nothing is exchanged
            return out
        if self.comm_stream is None:
            self.all_reduce_gradients()
            self.optimizer_step()
            return out
        lr = float(self.current_lr())      # schedule(updates done), then count this update (1-based for Adam's bias correction)
        self._beta1 = float(self.current_beta1())
        self.step_count += 1
        L = _C.lib()
        if self.serialize:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            for stage, lo, hi in self.buckets:
                _C.check(L.ndt1_engine_wait_stage(m._engine, stage, self.comm_stream.cuda_stream), "ndt1_engine_wait_stage")
                if self.shard:
                    self._sharded_bucket(stage, lo, hi, lr)
                else:
                    if self.world > 1:
                        dist.all_reduce(self.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
                    self._adamw(lo, hi, lr, self.comm_stream.cuda_stream)
            _C.check(L.ndt1_event_record(self._params_ev, self.comm_stream.cuda_stream), "ndt1_event_record")   # parameters final
        # The caller's stream is NOT joined here: the last buckets' all-reduce and AdamW then run under the next step's
        # prologue (smoothing / noise / input cast, which read no parameter); the model joins before its first
        # parameter-dependent kernel (NDT1.wait_for_parameters, also called by save_checkpoint).
        return out

    # ------------------------------------------------------------------ sharded optimizer
    def _sharded_bucket(self, stage: int, lo: int, hi: int, lr: float) -> None:
        """One gradient stage on the side stream: reduce-scatter + owner-slice AdamW + bf16 all-gather over its big region,
        all-reduce + replicated AdamW over its small region (DDP's mean folded into the update as 1 / world, models/trainer.py:258-262)."""
        blo, bhi = self._big[stage]
        st = self.comm_stream.cuda_stream
        if bhi > blo:
            n = (bhi - blo) // self.world
            mine = slice(blo + self.rank * n, blo + (self.rank + 1) * n)
            g = self.flat_grad
            if dist.get_backend(self.pg) == "nccl":
                dist.reduce_scatter_tensor(g[mine], g[blo:bhi], op=dist.ReduceOp.SUM, group=self.pg)
            else:                                   # (gloo has no reduce-scatter: the CPU / single-GPU test path)
                dist.all_reduce(g[blo:bhi], op=dist.ReduceOp.SUM, group=self.pg)
            self._adamw(mine.start, mine.stop, lr, st)                  # (clears its slice of the gradient)
            if mine.start > blo:
                g[blo:mine.start].zero_()
            if mine.stop < bhi:
                g[mine.stop:bhi].zero_()
            if dist.get_backend(self.pg) == "nccl":
                dist.all_gather_into_tensor(self.shadow[blo:bhi], self.shadow[mine], group=self.pg)
            else:
                self._gather_by_sum(self.shadow, blo, bhi, mine)
            self._params_gathered = False
        if hi > bhi:
            dist.all_reduce(self.flat_grad[bhi:hi], op=dist.ReduceOp.SUM, group=self.pg)
            self._adamw(bhi, hi, lr, st)

    def _gather_by_sum(self, buf: torch.Tensor, lo: int, hi: int, mine: slice) -> None:
        """all-gather for back-ends without one: zero the slices of the other ranks, sum."""
        tmp = torch.zeros(hi - lo, dtype=torch.float32, device=buf.device)
        tmp[mine.start - lo:mine.stop - lo] = buf[mine].float()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.pg)
        buf[lo:hi] = tmp.to(buf.dtype)

    def gather_parameters(self) -> None:
        """Sharded optimizer: bring the fp32 masters of the big regions up to date on every rank (checkpoints, evaluation in
        fp32, anything that reads parameters outside the bf16 engine).  Collective: every rank must call it."""
        if not self.shard or self._params_gathered:
            return
        self.synchronize()
        for stage, lo, hi in self.buckets:
            blo, bhi = self._big[stage]
            if bhi > blo:
                n = (bhi - blo) // self.world
                mine = slice(blo + self.rank * n, blo + (self.rank + 1) * n)
                if dist.get_backend(self.pg) == "nccl":
                    dist.all_gather_into_tensor(self.flat_param[blo:bhi], self.flat_param[mine], group=self.pg)
                else:
                    self._gather_by_sum(self.flat_param, blo, bhi, mine)
        self._params_gathered = True

    # ------------------------------------------------------------------ whole-step CUDA graph
    def _graphable(self, batch) -> bool:
        m = self.model
        if any(mk.is_active() for mk in m.encoder.masker):       # the maskers draw on the CPU generator (models/masker.py:56-101)
            return False
        if m.encoder.smooth_and_noise.rng != "device" or batch.get("noise") is not None:
            return False
        return all((not torch.is_tensor(v)) or (v.is_cuda and v.is_contiguous()) for v in batch.values())

    def _use_device_seeds(self, on: bool) -> None:
        m = self.model
        t = self._seeds_dev if (on and self._params_ev is not None) else None
        m._seed_tensor = t
        m.encoder.smooth_and_noise.seed_tensor = t

    def _push_seeds(self) -> None:
        """This step's (noise key, dropout key) -> device memory, through a ring of pinned slots (the host runs ahead of the GPU)."""
        slot = self._seeds_ring[self._seed_i % self._seeds_ring.shape[0]]
        self._seed_i += 1
        slot.copy_(torch.randint(0, 2 ** 62, (2,)))
        self._seeds_dev.copy_(slot, non_blocking=True)

    def _step_graphed(self, batch):
        m = self.model
        key = (m._engine_epoch,
               tuple((k, v.data_ptr(), tuple(v.shape), str(v.dtype)) for k, v in sorted(batch.items()) if torch.is_tensor(v)))
        ent = self._graphs.get(key)
        if ent is None:
            # first sight of this batch geometry / these buffers: run eagerly (creates or grows the engine, sets every kernel
            # attribute, fills the tensor-map cache); the next step with the same key captures
            if len(self._graphs) >= 16:
                self._graphs.clear()
            out = m.forward_backward(batch, self.flat_grad, self._ones)
            key = (m._engine_epoch, key[1])          # (the engine may have been created, or grown, by this very call)
            self._graphs[key] = {}
            return out
        if "graph" not in ent:
            L = _C.lib()
            g = torch.cuda.CUDAGraph()
            n0 = L.ndt1_launch_counter()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                out = m.forward_backward(batch, self.flat_grad, self._ones)
            ent.update(graph=g, out=out, launches=int(L.ndt1_launch_counter() - n0), keep=dict(batch))
        ent["graph"].replay()
        self.replayed_launches += ent["launches"]
        return ent["out"]

    def _adamw(self, lo: int, hi: int, lr: float, stream: int) -> None:
        b1, b2 = getattr(self, "_beta1", self.betas[0]), self.betas[1]
        fp, bf = 4 * lo, 2 * lo
        _C.check(_C.lib().ndt1_adamw_step_fused(self.flat_param.data_ptr() + fp, self.flat_grad.data_ptr() + fp, self.exp_avg.data_ptr() + fp,
                                                self.exp_avg_sq.data_ptr() + fp, hi - lo, lr, b1, b2, self.eps, self.wd, self.step_count,
                                                1.0 / self.world, None if self.shadow is None else self.shadow.data_ptr() + bf, 1, stream),
                 "ndt1_adamw_step_fused")

    def optimizer_step(self) -> None:
        """AdamW over the whole arena on the current stream (for callers that drive forward_backward themselves)."""
        lr = float(self.current_lr())
        self._beta1 = float(self.current_beta1())
        self.step_count += 1
        self._adamw(0, self.flat_param.numel(), lr, _C.stream_ptr())
