"""bench.py arm for the BPR workloads at N >= 1 GPUs (one process per GPU): users range-partitioned, item table
row-sharded, training step = the single-GPU fused step (N = 1) or the peer-memory step over NVLink (N > 1),
sharded full-sort evaluation.  Weak scaling: every rank processes `train_batch` triples per step;
`value` = all ranks' triples / max-over-ranks device time.  Bench support, not product.

Default workload: BASELINE.json configs[2] (cfg3, 10M users x 2M items x d=128), generated on the device.
"""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))


def _max_over_ranks(x, dev):
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class _Cfg(dict):
    def __getitem__(self, k):
        return self.get(k)


def parity_check(comm, dev, exchange):
    """A small problem through the SAME multi-rank path the timed run uses (real NCCL / NVLink peer memory),
    checked on rank 0 against the oracle's single-device step on the union of the ranks' batches: losses and
    tables to 1e-5 relative, top-10 ids bit-exact.  (The oracle is the checker here, never the thing measured.)"""
    from recbole_b200.dist import ShardedBPR, ShardedEvalIndex
    from recbole_b200.evaluator import FusedTopKEvaluator
    rank, world = comm.rank, comm.world
    rng = np.random.default_rng(7)
    n_users, n_items, d, B, steps = 4001, 3001, 128, 8192, 3
    U0 = (rng.standard_normal((n_users, d)) * 0.3).astype(np.float32)
    V0 = (rng.standard_normal((n_items, d)) * 0.3).astype(np.float32)
    batches = [(rng.integers(1, n_users, B),
                np.minimum(np.exp(rng.random(B) * np.log(n_items - 1)).astype(np.int64), n_items - 1).clip(1),
                rng.integers(1, n_items, B)) for _ in range(steps)]
    pairs = [(rng.integers(1, n_users, 20000), rng.integers(1, n_items, 20000)) for _ in range(3)]
    m = ShardedBPR(n_users, n_items, d, comm, dev, U_full=U0, V_full=V0, exchange=exchange)
    m.build_optimizer("adam", lr=2e-3)
    losses = []
    for (u, p, n) in batches:
        mine = (u >= m.u_lo) & (u < m.u_hi)
        t = lambda a: torch.from_numpy(a[mine]).to(dev)      # noqa: E731
        losses.append(float(m.train_step(t(u), t(p), t(n), global_batch=B).item()))
    m.check_flags()
    sd = m.state_dict()                                     # all-gathers the tables (NCCL)
    U, V = sd["user_embedding.weight"].cpu().numpy(), sd["item_embedding.weight"].cpu().numpy()
    # evaluation of every user that has positives in phase 2 (history = phases 0, 1), merged across ranks
    key = lambda a, b: np.unique(a.astype(np.int64) * n_items + b)     # noqa: E731
    hist_k, pos_k = np.union1d(key(*pairs[0]), key(*pairs[1])), key(*pairs[2])
    uid = np.unique(pos_k // n_items)
    remap = -np.ones(n_users, dtype=np.int64)
    remap[uid] = np.arange(len(uid))

    def csr(keys):
        r, c = remap[keys // n_items], keys % n_items
        keep = r >= 0
        r, c = r[keep], c[keep]
        ptr = np.zeros(len(uid) + 1, dtype=np.int64)
        ptr[1:] = np.cumsum(np.bincount(r, minlength=len(uid)))
        return ptr, c

    hist, pos = csr(hist_k), csr(pos_k)
    ev = FusedTopKEvaluator(_Cfg(metrics=["Recall", "MRR", "NDCG", "Hit", "Precision"], topk=[10],
                                 metric_decimal_place=4))
    idx = ShardedEvalIndex.from_global(uid, hist, pos, m.user_bounds, m.item_bounds, rank, dev)
    res = m.evaluate(idx, ev, mode="tc", layout="replicate")
    counts = comm.all_gather_object(int(m.last_topk.shape[0]))
    topk = comm.all_gather_rows(m.last_topk, counts).cpu().numpy() if world > 1 else m.last_topk.cpu().numpy()
    out = {"exchange": m.last_exchange, "world": world, "shape": "%dx%dx%d, B=%d, %d adam steps" % (n_users, n_items, d, B, steps)}
    if rank == 0:
        from oracle import bpr as obpr              # the checker (test infrastructure)
        from oracle import fullsort as ofs
        st = obpr.new_state(U0, V0)
        ref = [obpr.bpr_train_step(st, u, p, n, s + 1, optimizer="adam", lr=2e-3, dense=False)
               for s, (u, p, n) in enumerate(batches)]
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())      # noqa: E731
        o_ids, _ = ofs.full_sort_topk(U, V, uid, hist[0], hist[1], 10)
        o_res = ofs.evaluate(o_ids, pos[0], pos[1], ["recall", "mrr", "ndcg", "hit", "precision"], [10])
        out.update(loss_rel_err=float(max(abs(a - b) / abs(b) for a, b in zip(losses, ref))),
                   user_table_rel_err=rel(U, st["U"]), item_table_rel_err=rel(V, st["V"]),
                   topk_mismatches=int((topk != o_ids).sum()), metrics_equal=bool(res == o_res), users=int(len(uid)))
        out["ok"] = bool(out["loss_rel_err"] <= 1e-5 and out["user_table_rel_err"] <= 1e-5 and
                         out["item_table_rel_err"] <= 1e-5 and out["topk_mismatches"] == 0 and out["metrics_equal"])
    comm.barrier()
    del m
    # the parity mode (adam_lazy == the reference's DENSE torch.optim.Adam) through the same exchange: small batches,
    # so that most rows are NOT touched in a step and keep moving on their momentum
    if exchange != "sparse":
        torch.cuda.empty_cache()
        Bl, steps_l = 512, 5
        m = ShardedBPR(n_users, n_items, d, comm, dev, U_full=U0, V_full=V0, exchange=exchange)
        m.build_optimizer("adam_lazy", lr=2e-3)
        lb = [(rng.integers(1, n_users, Bl), rng.integers(1, n_items, Bl), rng.integers(1, n_items, Bl))
              for _ in range(steps_l)]
        losses = []
        for (u, p, n) in lb:
            mine = (u >= m.u_lo) & (u < m.u_hi)
            t = lambda a: torch.from_numpy(a[mine]).to(dev)      # noqa: E731
            losses.append(float(m.train_step(t(u), t(p), t(n), global_batch=Bl).item()))
        m.check_flags()
        sd = m.state_dict()
        U, V = sd["user_embedding.weight"].cpu().numpy(), sd["item_embedding.weight"].cpu().numpy()
        if rank == 0:
            st = obpr.new_state(U0, V0)
            ref = [obpr.bpr_train_step(st, u, p, n, s + 1, optimizer="adam", lr=2e-3, dense=True)
                   for s, (u, p, n) in enumerate(lb)]
            # an element whose gradient is ~1e-8 (|g| ~ Adam's eps) is known to ~1e-3 relative only, and the normalised
            # step lr * g / (|g| + eps) carries that into the parameter at every later zero-gradient step: such
            # elements (a handful per million) are counted, not held to 1e-5
            bad = lambda a, b: int((np.abs(a - b) > 1e-5 * np.abs(b).max()).sum())      # noqa: E731
            la = {"shape": "B=%d, %d dense-Adam steps" % (Bl, steps_l),
                  "loss_rel_err": float(max(abs(a - b) / abs(b) for a, b in zip(losses, ref))),
                  "user_table_rel_err": rel(U, st["U"]), "item_table_rel_err": rel(V, st["V"]),
                  "eps_conditioned_elements": bad(U, st["U"]) + bad(V, st["V"])}
            la["ok"] = bool(la["loss_rel_err"] <= 1e-5 and la["eps_conditioned_elements"] <= 3 and
                            max(la["user_table_rel_err"], la["item_table_rel_err"]) <= 2e-4)
            out["dense_adam"] = la
            out["ok"] = bool(out["ok"] and la["ok"])
        comm.barrier()
        del m
    torch.cuda.empty_cache()
    return out


def _traffic(kernel, workload):
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get("%s@%s" % (kernel, workload))
    return None


def run(args, rank, world, local_rank, load_peaks, ClockSampler, cpu_reference=None):
    import bench_workloads as bw
    from recbole_b200 import ops
    from recbole_b200.dist import Comm, ShardedBPR, ShardedEvalIndex
    from recbole_b200.evaluator import FusedTopKEvaluator

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = Comm()
    peaks = load_peaks()
    big = args.workload == "cfg3"
    # the parity problem goes through the exchange the TIMED model will use (a small table would pick "dense")
    shape = (2000001, 128) if big else {"cfg2": (26745, 64)}.get(args.workload, (26745, 64))
    n_it = int(shape[0] * (args.scale if big else 1.0))
    timed_exchange = ShardedBPR.resolve_exchange(args.exchange, n_it, shape[1], comm, dev) if world > 1 else args.exchange
    parity = None if args.skip_parity else parity_check(comm, dev, timed_exchange)
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
    # single occurrences per batch (the share of the item updates the user-side kernel finishes itself)
    singles = []
    for (_, p, n) in resident:
        _, cnt = torch.unique(torch.cat([p, n]), return_counts=True)
        singles.append(int((cnt == 1).sum().item()))
    single_occ = float(np.mean(singles))

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
    model.check_flags()
    phases = None
    if world > 1 and getattr(model, "last_exchange", "") == "sparse":
        model.phase_timing = True
        for i in range(min(args.steps, 6)):
            model.train_step(*resident[(args.warmup + i) % nb], global_batch=GB, next_batch=nxt(args.warmup + i))
        barrier()
        phases = model.phase_report()
        model.phase_timing = False
    value = GB * args.steps / (ms_total / 1e3)
    final_loss = float(model.loss_out.item())

    # The headline optimizer is ROW-SPARSE Adam (rows a batch does not touch keep parameters and moments:
    # torch.optim.SparseAdam's contract).  The reference arm runs the Trainer's default, DENSE torch.optim.Adam, whose
    # untouched rows keep moving on their momentum; the fused kind that reproduces THAT trajectory is 'adam_lazy'
    # (same row-sparse memory traffic + a catch-up replay per touched row).  Same workload, fewer steps:
    lazy = None
    if world == 1 and not getattr(args, "no_extra", False):
        st = model.state
        st["lastU"] = torch.zeros(model.U.shape[0], dtype=torch.int32, device=dev)
        st["lastV"] = torch.zeros(model.V.shape[0], dtype=torch.int32, device=dev)
        opt_lazy = ops.Optim("adam_lazy", 1e-3, 0.0)
        opt_lazy.step = model.optim.step
        st["lastU"].fill_(opt_lazy.step)
        st["lastV"].fill_(opt_lazy.step)
        ws = model._workspace(B)
        n_lazy = max(min(args.steps, 50), 5)
        for i in range(3):
            ops.bpr_train_step(model.U, model.V, st, *resident[i % nb], opt_lazy, model.loss_out, None, ws)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n_lazy):
            ops.bpr_train_step(model.U, model.V, st, *resident[(3 + i) % nb], opt_lazy, model.loss_out, None, ws)
        e1.record()
        torch.cuda.synchronize()
        lazy_ms = e0.elapsed_time(e1) / n_lazy
        ops.adam_lazy_flush(model.U, st["mU"], st["vU"], st["lastU"], opt_lazy)
        ops.adam_lazy_flush(model.V, st["mV"], st["vV"], st["lastV"], opt_lazy)
        model.optim.step = opt_lazy.step
        del st["lastU"], st["lastV"]
        lazy = {"optimizer": "adam_lazy (trajectory of the reference's dense torch.optim.Adam)", "steps": n_lazy,
                "ms_per_step": lazy_ms, "value": B / (lazy_ms * 1e-3), "unit": "samples/s",
                "frac_of_hbm_roofline": B * (72 * d + 24) / (lazy_ms * 1e-3) / 1e9 / peaks["hbm"]}
        ops.profile_read()

    elif world > 1 and getattr(model, "last_exchange", "") == "p2p" and not getattr(args, "no_extra", False):
        # the same on the peer-memory exchange: user rows caught up locally, every owner steps its whole item shard
        st, sparse_opt = model.state, model.optim
        st["lastU"] = torch.full((model.U.shape[0],), sparse_opt.step, dtype=torch.int32, device=dev)
        model.optim = ops.Optim("adam_lazy", 1e-3, 0.0)
        model.optim.step = sparse_opt.step
        n_lazy = max(min(args.steps, 50), 5)
        for i in range(3):
            model.train_step(*resident[i % nb], global_batch=GB, next_batch=nxt(i))
        barrier()
        e0.record()
        for i in range(n_lazy):
            model.train_step(*resident[(3 + i) % nb], global_batch=GB, next_batch=nxt(3 + i))
        e1.record()
        barrier()
        lazy_ms = _max_over_ranks(e0.elapsed_time(e1), dev) / n_lazy
        model.flush()
        sparse_opt.step = model.optim.step
        model.optim = sparse_opt
        del st["lastU"]
        lazy = {"optimizer": "adam_lazy (trajectory of the reference's dense torch.optim.Adam; peer-memory exchange)",
                "steps": n_lazy, "ms_per_step": lazy_ms, "value": GB / (lazy_ms * 1e-3), "unit": "samples/s",
                "frac_of_hbm_roofline": GB * (72 * d + 24) / (lazy_ms * 1e-3) / 1e9 / (peaks["hbm"] * world)}
        ops.profile_read()

    # ---- e2e: batches in pinned HOST memory, double-buffered H2D on a copy stream, loss read back every step -------
    copy_stream = torch.cuda.Stream()
    slots = [[torch.empty(B, dtype=torch.int64, device=dev) for _ in range(3)] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(args.steps + args.warmup + 8, dtype=torch.float32).pin_memory()

    def e2e_loop(n, offset):
        main = torch.cuda.current_stream()
        for i in range(n):
            s = i % 2
            hb = host[(offset + i) % nb]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[s])
                for k in range(3):
                    slots[s][k].copy_(hb[k], non_blocking=True)
                ready[s].record(copy_stream)
            main.wait_event(ready[s])
            lo = model.train_step(*slots[s], global_batch=GB)
            loss_host[i:i + 1].copy_(lo, non_blocking=True)
            free[s].record(main)

    for s in range(2):
        free[s].record(torch.cuda.current_stream())
    model.ids_ready = False
    e2e_loop(min(args.warmup, 3), 0)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps, args.warmup)
    barrier()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, dev)
    clk = clocks.stop()

    # ---- evaluation --------------------------------------------------------------------------------------------------
    ev = FusedTopKEvaluator(_Cfg(metrics=["Recall", "MRR", "NDCG", "Hit", "Precision"], topk=[10],
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
        # warm-up on a slice of the users (a full pass costs seconds on one GPU)
        warm = ShardedEvalIndex(index.uid_all, (None, None), (pos_ptr[:65537].contiguous(), index.pos_indices), None,
                                min(65536, nu - first), uid_own=uid_own[:65536].contiguous(),
                                hist_own=(hist_ptr[:65537].contiguous(), index.own_hist_indices))
        model.evaluate(warm, ev, mode=args.scorer, layout="replicate")
    else:
        index = ShardedEvalIndex.from_global(w.uid_list, w.hist, w.pos, model.user_bounds, model.item_bounds, rank,
                                             dev)
        nq = len(w.uid_list)
        model.evaluate(index, ev, mode=args.scorer, layout=args.eval_layout)
    barrier()
    ops.profile_enable(True)
    ops.profile_read()
    e0.record()
    for _ in range(args.eval_reps):
        result = model.evaluate(index, ev, mode=args.scorer, layout=args.eval_layout)
    e1.record()
    barrier()
    eval_ms = _max_over_ranks(e0.elapsed_time(e1), dev) / args.eval_reps
    estages = ops.profile_read()
    ops.profile_enable(False)
    # e2e evaluation: the user ids come from pinned host memory, the result dict lands on the host
    uid_host = index.uid_own.cpu().pin_memory()
    barrier()
    t0 = time.perf_counter()
    index.uid_own = uid_host.to(dev, non_blocking=True)
    result2 = model.evaluate(index, ev, mode=args.scorer, layout=args.eval_layout)
    barrier()
    eval_e2e_s = _max_over_ranks(time.perf_counter() - t0, dev)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return None

    def stage_ms(st, name):
        return st[name][0] / st[name][1] if name in st else 0.0

    us_ms = stage_ms(stages, "user_side")
    train_ms = ms_total / args.steps
    alg_step = GB * (72 * d + 24)
    # the user-side kernel's share of the algorithmic bytes: user row read + step (24 d), two item rows gathered
    # (8 d), ids (24), and the whole step (m, v read; p, m, v written: 20 d) of every single-occurrence item
    exch = getattr(model, "last_exchange", model.exchange)
    if exch == "local":
        alg_user = B * (32 * d + 24) + single_occ * 20 * d
    else:
        # sharded step: the kernel steps the user row and gathers two item rows; the gradient rows it writes (into
        # the owners' slots over NVLink or into gu) are traffic of the exchange and NOT counted as algorithmic
        alg_user = B * (32 * d + 24)
    kernel = "k_user_fused"
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": alg_user / (us_ms * 1e-3) / 1e9 if us_ms else None,
                "peak": peaks["hbm"], "unit": "GB/s", "frac": None,
                "traffic": _traffic(kernel, args.workload) if exch == "local" else None,
                "algorithmic_bytes_per_launch": alg_user, "single_item_occurrences_per_batch": single_occ,
                "peak_source": peaks["source"], "ms_per_launch": us_ms,
                "share_of_step": us_ms / (sum(v[0] for v in stages.values()) / args.steps) if us_ms else None,
                "step": {"achieved_all_gpus": alg_step / (train_ms * 1e-3) / 1e9,
                         "frac_of_n_gpu_peak": alg_step / (train_ms * 1e-3) / 1e9 / (peaks["hbm"] * world),
                         "bytes_per_sample": 72 * d + 24,
                         "traffic_per_step": _traffic("step", args.workload) if exch == "local" else None},
                "exchange_phases_ms_rank0": phases,
                "stages_ms_per_step_rank0": {k: v[0] / args.steps for k, v in stages.items()}}
    if roofline["achieved"]:
        roofline["frac"] = roofline["achieved"] / peaks["hbm"]
    fs_ms = sum(estages[k][0] for k in ("tc_score",) if k in estages) / max(args.eval_reps, 1)
    flops = 2.0 * (nq / world) * w.n_items * d            # rank 0's share
    eval_roof = {"bound": "tensor", "kernel": "k_fullsort_tc", "peak": peaks["tf"], "unit": "TFLOP/s",
                 "achieved": flops / (fs_ms * 1e-3) / 1e12 if fs_ms else None,
                 "frac": flops / (fs_ms * 1e-3) / 1e12 / peaks["tf"] if fs_ms else None,
                 "whole_eval_all_gpus": 2.0 * nq * w.n_items * d / (eval_ms * 1e-3) / 1e12,
                 "whole_eval_frac_of_n_gpu_peak": 2.0 * nq * w.n_items * d / (eval_ms * 1e-3) / 1e12 / (peaks["tf"] * world),
                 "traffic": None, "ms_in_kernel_rank0": fs_ms,
                 "stages_ms_rank0": {k: v[0] / max(args.eval_reps, 1) for k, v in estages.items()}}
    how = {"local": "single GPU: fused step on the whole tables",
           "p2p": "users range-partitioned, item table row-sharded x%d; rows read from / gradients pushed into the "
                  "owners' memory by the kernels over NVLink (cudaIpc peer mappings, flag barriers; no NCCL in the "
                  "step)" % world,
           "sparse": "users range-partitioned, item table row-sharded x%d, NCCL all-to-all of de-duplicated rows" % world,
           "dense": "users range-partitioned, item table row-sharded x%d, NCCL all-gather + reduce-scatter" % world}
    cpu_baseline = None
    if cpu_reference is not None and world == 1 and not args.skip_cpu:
        cpu_baseline = cpu_reference(args)
    line = {
        "metric": "bpr_train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": train_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(w.describe(), train_batch_per_gpu=B, global_batch=GB, optimizer="adam(row-sparse: SparseAdam semantics; the dense-Adam-equivalent kind is under "
                                 "dense_adam_equivalent) lr=1e-3",
                       scorer=args.scorer, eval_layout=args.eval_layout, exchange=exch, parallelism=how.get(exch, exch),
                       l2="no flush: every step reads a different batch; tables + Adam state (%.1f GB per GPU) exceed "
                          "the 126 MB L2" % (3 * 4 * d * (w.n_users / world + w.n_items / world) / 1e9)),
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity_check": parity,
        "dense_adam_equivalent": lazy,
        "e2e": {"value": GB * args.steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": 24 * B * world,
                "d2h_bytes_per_step": 4 * world},
        "gpu_launches": int(sum(v[2] for v in stages.values())) * world, "loss": final_loss,
        "eval": {"metric": "fullsort_eval_users_per_s", "value": nq / (eval_ms * 1e-3), "unit": "users/s",
                 "users": nq, "ms": eval_ms, "topk": 10, "result": result, "roofline": eval_roof,
                 "e2e": {"value": nq / eval_e2e_s, "unit": "users/s", "h2d_bytes_per_step": 8 * nq,
                         "d2h_bytes_per_step": 6 * 10 * 8 * world},
                 "tc_pass2_rows_rank0": getattr(model, "last_eval_pass2_rows", None),
                 "tc_fallback_rows_rank0": getattr(model, "last_eval_fallback_rows", None)},
    }
    assert result2 == result
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line
