"""Multi-rank parity of DataParallelTrainer.train_step against a single process under DDP-mean semantics.

What the reference does (models/trainer.py:77-80, 258-262, 335-349): Accelerate(split_batches=True) splits the GLOBAL batch
contiguously over the ranks, every rank back-propagates the SUM loss of its shard, DDP averages the gradients over ranks, AdamW
steps.  So after k steps every rank holds the parameters a single process reaches on the concatenated batch with the loss scaled
by 1 / world.  This script runs both and compares the flat fp32 parameter arena:

  * under pytest (tests/test_gpu_parity.py::test_data_parallel_trainer_world2_matches_single_process) as TWO processes sharing
    cuda:0 over gloo (so the driver's single-GPU test box exercises the bucket spans, the 1 / world fold into AdamW, the stage
    events and the deferred join of llm_bci_b200/trainer.py);
  * under torchrun on N GPUs over NCCL:
        python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 \
            tests/dp_parity.py --backend nccl
    (output kept in profiles/r02_dp_parity_*gpu.txt).

Dropout and noise are off (they are keyed by per-process seeds).  AdamW eps is 1e-3: with the default 1e-8 a parameter whose
gradient is analytically zero (attn.key.bias) moves by +-lr on rounding noise alone, which says nothing about
the exchange.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases():
    import llm_bci_b200 as lb
    from test_oracle_golden import small_ctc_cfg
    # fp32: the CUDA-core path (64-wide model); bf16: head size 128 -> the tcgen05 GEMM and attention kernels
    tc_cfg = lb.update_config(small_ctc_cfg(), {"encoder": {
        "embedder": {"n_channels": 64, "input_dim": 64, "max_F": 256},
        "transformer": {"n_layers": 2, "hidden_size": 256, "n_heads": 2, "inter_size": 256}}})
    return [("fp32", small_ctc_cfg(), 16, 120, 1e-5), ("bf16", tc_cfg, 64, 400, 2e-3)]


def run(rank: int, world: int, dev: torch.device, steps: int = 3, per_rank: int = 2):
    import llm_bci_b200 as lb
    from oracle import ndt1_oracle as O
    from test_oracle_golden import CTC_KW
    solo = dist.new_group(ranks=[0])           # (collective: every rank creates it) a world of one for the single-process run
    results = []
    for precision, cfg, N, T, tol in cases():
        Bg = per_rank * world
        batches = []
        for s in range(steps):
            b = O.synthetic_ctc_batch(B=Bg, T=T, N=N, seed=10 + s)
            batches.append({k: v.to(dev) for k, v in b.items()})
        S = max(int(b["targets"].shape[1]) for b in batches)
        for b in batches:
            b["targets"] = torch.nn.functional.pad(b["targets"], (0, S - b["targets"].shape[1]))

        def make():
            torch.manual_seed(7)
            return lb.NDT1(cfg, **CTC_KW, precision=precision, max_batch=Bg, max_T=T).to(dev)

        m = make()
        t = lb.DataParallelTrainer(m, lr=1e-3, wd=5e-5, eps=1e-3, scheduler="cosine", total_steps=10, warmup_pct=0.3, div_factor=25.0,
                                   shard_optimizer=(precision == "bf16"))      # bf16 case: also the (optional) sharded optimizer
        assert t.world == world and len(t.buckets) == cfg.encoder.transformer.n_layers + 2
        losses = []
        for b in batches:
            out = t.train_step(lb.trainer.shard_batch(b, rank, world))
            losses.append(out.loss.detach().clone())
        sharded = bool(t.shard)
        t.gather_parameters()          # (sharded optimizer: the fp32 masters of the big regions live on their owner until gathered)
        t.synchronize()
        torch.cuda.synchronize()
        # every rank must hold the same replica bit for bit (the all-reduced gradients are identical on all ranks)
        cdev = torch.device("cpu") if dist.get_backend() == "gloo" else dev      # (gloo gathers / broadcasts host tensors only)
        mine = t.flat_param.to(cdev).clone()
        ref0 = mine.clone()
        dist.broadcast(ref0, src=0)
        same = bool(torch.equal(mine, ref0))
        ls = torch.stack(losses).to(cdev).double()
        dist.all_reduce(ls, op=dist.ReduceOp.SUM)
        loss_sum = ls.tolist()
        flags = torch.tensor([1.0 if same else 0.0], device=cdev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            m1 = make()
            t1 = lb.DataParallelTrainer(m1, lr=1e-3, wd=5e-5, eps=1e-3, scheduler="cosine", total_steps=10, warmup_pct=0.3, div_factor=25.0,
                                        process_group=solo, loss_scale=1.0 / world)
            assert t1.world == 1
            init = t1.flat_param.clone()
            l1 = [float(t1.train_step(b).loss) for b in batches]
            t1.synchronize()
            torch.cuda.synchronize()
            moved = float((t1.flat_param - init).abs().max())
            err = float((t.flat_param - t1.flat_param).abs().max() / t1.flat_param.abs().max())
            lerr = max(abs(a - b) / abs(b) for a, b in zip(loss_sum, l1))
            results.append(dict(precision=precision, world=world, steps=steps, param_rel_err=err, tol=tol, loss_rel_err=lerr,
                                replicas_identical=bool(flags.item() == 1.0), max_param_update=moved, sharded_optimizer=sharded,
                                shadow_current=bool(t.shadow is None or torch.equal(t.shadow, t.flat_param.bfloat16())),
                                grads_cleared=float(t.flat_grad.abs().max()) == 0.0))
        dist.barrier()
    return results


def check(results):
    for r in results:
        assert r["replicas_identical"], r
        assert r["max_param_update"] > 1e-4, r           # the comparison is not vacuous
        assert r["param_rel_err"] <= r["tol"], r
        assert r["loss_rel_err"] <= (1e-5 if r["precision"] == "fp32" else 2e-2), r
        assert r["shadow_current"] and r["grads_cleared"], r


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = run(rank, world, torch.device("cuda", 0))
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group(args.backend)
    res = run(rank, world, dev)
    if rank == 0:
        for r in res:
            print(json.dumps(r))
        check(res)
        print("dp_parity ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
