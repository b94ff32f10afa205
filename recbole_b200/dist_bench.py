"""bench.py arm for N > 1 GPUs (one process per GPU, NCCL): user-partitioned, item-sharded BPR
training step + sharded full-sort evaluation.  Weak scaling: every rank processes `train_batch`
triples per step; `value` = all ranks' triples / max-over-ranks device time."""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def _max_over_ranks(x, dev):
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run(args, rank, world, local_rank, load_peaks, ClockSampler):
    import bench_workloads as bw
    from . import ops
    from .dist import Comm, ShardedBPR, ShardedEvalIndex
    from .evaluator import FusedTopKEvaluator

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = Comm()
    peaks = load_peaks()
    big = args.workload == "cfg3"
    if big:
        w = bw.Cfg3Device(rank, world, dev, batch=args.batch or (1 << 20), n_batches=args.n_batches,
                          scale=args.scale)
        B, d = w.batch, w.dim
        model = ShardedBPR(w.n_users, w.n_items, d, comm, dev, exchange=args.exchange)
        resident = w.batches
        host = [tuple(x.cpu().pin_memory() for x in b) for b in resident]
    else:
        w = bw.BprWorkload(args.workload, batch=args.batch, n_batches=1)
        B, d = w.batch, w.dim
        batches = w.rank_batches(rank, world, args.n_batches)
        model = ShardedBPR(w.n_users, w.n_items, d, comm, dev, U_full=w.U0, V_full=w.V0, exchange=args.exchange)
        host = [tuple(torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in b) for b in batches]
        resident = [tuple(x.to(dev) for x in b) for b in host]
    model.build_optimizer("adam", 1e-3, 0.0)
    model.ids_ready = True            # the batches are resident: next-step plans need not wait for the training stream
    nb = len(resident)
    GB = B * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def nxt(i):
        return resident[(i + 1) % nb]

    for i in range(args.warmup):
        model.train_step(*resident[i % nb], global_batch=GB, next_batch=nxt(i))
    barrier()
    ops.profile_enable(True)
    ops.profile_read()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        model.train_step(*resident[(args.warmup + i) % nb], global_batch=GB, next_batch=nxt(args.warmup + i))
    e1.record()
    barrier()
    ms_total = _max_over_ranks(e0.elapsed_time(e1), dev)
    stages = ops.profile_read()
    ops.profile_enable(False)
    phases = None
    if world > 1 and getattr(model, "last_exchange", "") != "dense":
        # second, instrumented pass (events between the phases perturb nothing but are kept out of the timed one)
        model.phase_timing = True
        for i in range(min(args.steps, 6)):
            model.train_step(*resident[(args.warmup + i) % nb], global_batch=GB, next_batch=nxt(args.warmup + i))
        barrier()
        phases = model.phase_report()
        model.phase_timing = False
    value = GB * args.steps / (ms_total / 1e3)
    final_loss = float(model.loss_out.item())

    # e2e: batches from pinned host memory, loss read back every step
    loss_host = torch.zeros(args.steps + args.warmup, dtype=torch.float32).pin_memory()

    def e2e_loop(n, offset):
        for i in range(n):
            hb = host[(offset + i) % nb]
            u, p, ng = (x.to(dev, non_blocking=True) for x in hb)
            lo = model.train_step(u, p, ng, global_batch=GB)
            loss_host[i:i + 1].copy_(lo, non_blocking=True)

    e2e_loop(min(args.warmup, 2), 0)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps, args.warmup)
    barrier()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, dev)
    clk = clocks.stop()

    # evaluation
    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k)

    ev = FusedTopKEvaluator(Cfg(metrics=["Recall", "MRR", "NDCG", "Hit", "Precision"], topk=[10],
                                metric_decimal_place=4))
    if big:
        nu = w.u_hi - w.u_lo
        first = 1 if rank == 0 else 0                      # user 0 is [PAD]
        uid_own = torch.arange(w.u_lo + first, w.u_hi, device=dev, dtype=torch.int64)
        n_all = w.n_users - 1
        pos_ptr = torch.arange(0, (nu - first + 1) * w.n_test, w.n_test, device=dev, dtype=torch.int64)
        hist_ptr = (w.used_indptr[first:] - w.used_indptr[first]).contiguous()
        index = ShardedEvalIndex(torch.empty(n_all, dtype=torch.int8, device="meta"), (None, None),
                                 (pos_ptr, w.test_items[first:].reshape(-1).contiguous()), None, nu - first,
                                 uid_own=uid_own, hist_own=(hist_ptr, w.used_indices[first * w.per_user:].contiguous()))
        nq = n_all
        if args.eval_layout == "sharded":
            raise SystemExit("cfg3 bench evaluates with --eval-layout replicate (own users x all items)")
        args.eval_layout = "replicate"
    else:
        index = ShardedEvalIndex.from_global(w.uid_list, w.hist, w.pos, model.user_bounds, model.item_bounds, rank,
                                             dev)
        nq = len(w.uid_list)
    model.evaluate(index, ev, mode=args.scorer, layout=args.eval_layout)
    barrier()
    e0.record()
    for _ in range(args.eval_reps):
        result = model.evaluate(index, ev, mode=args.scorer, layout=args.eval_layout)
    e1.record()
    barrier()
    eval_ms = _max_over_ranks(e0.elapsed_time(e1), dev) / args.eval_reps

    if rank == 0:
        def stage_ms(name):
            return stages[name][0] / stages[name][1] if name in stages else 0.0
        us_ms = stage_ms("user_side")
        train_ms = ms_total / args.steps
        alg_step = GB * (72 * d + 24)
        roofline = {"bound": "hbm", "kernel": "k_user_side", "achieved": B * (32 * d + 24) / (us_ms * 1e-3) / 1e9 if us_ms else None,
                    "peak": peaks["hbm"], "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peaks["source"],
                    "ms_per_launch": us_ms,
                    "step": {"achieved_all_gpus": alg_step / (train_ms * 1e-3) / 1e9,
                             "frac_of_n_gpu_peak": alg_step / (train_ms * 1e-3) / 1e9 / (peaks["hbm"] * world),
                             "bytes_per_sample": 72 * d + 24},
                    "exchange_phases_ms_rank0": phases, "stages_ms_per_step_rank0": {k: v[0] / args.steps for k, v in stages.items()}}
        if roofline["achieved"]:
            roofline["frac"] = roofline["achieved"] / peaks["hbm"]
        line = {
            "metric": "bpr_train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": train_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(w.describe(), train_batch_per_gpu=B, global_batch=GB, optimizer="adam(row-sparse) lr=1e-3",
                           scorer=args.scorer, eval_layout=args.eval_layout, exchange=getattr(model, "last_exchange", model.exchange),
                           parallelism="users range-partitioned, item table row-sharded x%d, NCCL all-to-all" % world,
                           l2="no flush: every step reads a different batch"),
            "clocks": clk, "roofline": roofline,
            "cpu_baseline": None,
            "e2e": {"value": GB * args.steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": 24 * B * world,
                    "d2h_bytes_per_step": 4 * world},
            "gpu_launches": int(sum(v[2] for v in stages.values())) * world, "loss": final_loss,
            "eval": {"metric": "fullsort_eval_users_per_s", "value": nq / (eval_ms * 1e-3), "unit": "users/s",
                     "users": nq, "ms": eval_ms, "topk": 10, "result": result,
                     "tc_pass2_rows_rank0": getattr(model, "last_eval_pass2_rows", None),
                     "tc_fallback_rows_rank0": getattr(model, "last_eval_fallback_rows", None)},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
