"""bench.py arms for BASELINE configs 4 (SASRec-shaped CE head) and 5 (FM, Criteo shape); single GPU.
Same JSON contract as bench.py (value / e2e / roofline / cpu_baseline / clocks)."""
import json
import os
import time

import numpy as np

# Criteo-Kaggle categorical cardinalities, 26 fields, sum = 33 762 577
CRITEO_CARD = np.array([1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                        5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572], dtype=np.int64)


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def fm_batches(card, B, n_batches, seed=2020):
    """ids per field ~ Zipf(1.05)-like via the inverse-CDF trick on ranks, labels ~ Bernoulli(0.256)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_batches):
        u = rng.random((B, len(card)))
        ids = np.minimum(np.exp(u * np.log(card)[None, :]).astype(np.int64), card[None, :] - 1)
        lab = (rng.random(B) < 0.256).astype(np.float32)
        out.append((ids, lab))
    return out


def run_cfg5(args, load_peaks, ClockSampler):
    import torch
    from recbole_b200 import ops
    from oracle import torch_port
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    card, d, F = CRITEO_CARD, 16, 26
    B = args.batch or (1 << 18)
    rows = int(card.sum())
    off = np.concatenate([[0], np.cumsum(card)[:-1]]).astype(np.int64)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2020)
    std = (2.0 / (rows + d)) ** 0.5
    E = torch.randn(rows, d, device=dev, generator=gen) * std
    W = torch.randn(rows, device=dev, generator=gen) * (2.0 / (rows + 1)) ** 0.5
    bias3 = torch.zeros(3, device=dev)
    st = dict(mE=torch.zeros_like(E), vE=torch.zeros_like(E), mW=torch.zeros_like(W), vW=torch.zeros_like(W))
    batches = fm_batches(card, B, args.n_batches)
    host = [(torch.from_numpy(i).pin_memory(), torch.from_numpy(l).pin_memory()) for i, l in batches]
    res = [(i.to(dev), l.to(dev)) for i, l in host]
    offs = torch.from_numpy(off).to(dev)
    opt = ops.Optim("adam", lr=1e-3)
    loss = torch.zeros(1, device=dev)
    ws = ops.fm_workspace(B, F, d, dev)
    nb = len(res)

    def step(ids, lab):
        ops.fm_train_step(E, W, bias3, st, ids, offs, lab, opt, loss, None, ws)

    for i in range(args.warmup):
        step(*res[i % nb])
    torch.cuda.synchronize()
    ops.profile_enable(True)
    ops.profile_read()
    clocks = ClockSampler(0)
    clocks.start()
    e0, e1 = _events(torch)
    e0.record()
    for i in range(args.steps):
        step(*res[(args.warmup + i) % nb])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    stages = ops.profile_read()
    ops.profile_enable(False)
    # e2e: ids + labels from pinned host memory, loss back
    loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ids, lab = host[(args.warmup + i) % nb]
        step(ids.to(dev, non_blocking=True), lab.to(dev, non_blocking=True))
        loss_host[i:i + 1].copy_(loss, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()
    # CPU baseline: the reference's torch calls, dense Adam over the whole 33.8M-row table
    cpu = None
    if not args.skip_cpu:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        m = torch_port.RefFM(card.tolist(), d)
        o = torch.optim.Adam(m.parameters(), lr=1e-3)
        cb = [(torch.from_numpy(i), torch.from_numpy(l)) for i, l in batches[:2]]
        torch_port.fm_train_steps(m, o, cb[:1])
        t0 = time.perf_counter()
        torch_port.fm_train_steps(m, o, [cb[i % 2] for i in range(args.cpu_steps)])
        dt = time.perf_counter() - t0
        cpu = {"value": args.cpu_steps * B / dt, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "1 warm-up + %d timed steps of %d rows (torch CPU: FM forward, BCELoss, dense Adam over %d rows)"
                         % (args.cpu_steps, B, rows)}
    alg = B * (F * (24 * d + 24) + 8 * F + 4)
    upd_ms = stages["fm_update"][0] / stages["fm_update"][1]
    alg_upd = B * F * (24 * d + 24)
    line = {
        "metric": "fm_train_samples_per_s", "value": B / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg5", "desc": "FM CTR, synthetic Criteo shape: 26 categorical fields, 33.76M total vocab, d=16, Adam",
                   "fields": F, "rows": rows, "dim": d, "train_batch": B,
                   "l2": "no flush: tables + Adam state 6.9 GB and every step reads a different batch"},
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "k_fm_rows(+fixup)", "achieved": alg_upd / (upd_ms * 1e-3) / 1e9,
                     "peak": peaks["hbm"], "unit": "GB/s", "frac": alg_upd / (upd_ms * 1e-3) / 1e9 / peaks["hbm"],
                     "traffic": None, "peak_source": peaks["source"], "ms_per_launch": upd_ms,
                     "step": {"achieved": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm"],
                              "bytes_per_sample": F * (24 * d + 24) + 8 * F + 4},
                     "stages_ms_per_step": {k: v[0] / args.steps for k, v in stages.items()}},
        "cpu_baseline": cpu,
        "e2e": {"value": B * args.steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": B * (8 * F + 4),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(sum(v[2] for v in stages.values())), "loss": float(loss.item()),
    }
    print(json.dumps(line), flush=True)


def run_cfg4(args, load_peaks, ClockSampler):
    import torch
    from recbole_b200 import ops
    from oracle import torch_port
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    nq, N, d, K = args.batch or 4096, 1_000_001, 64, 10
    rng = np.random.default_rng(2020)
    X = rng.standard_normal((nq, d)).astype(np.float32)
    X = ((X - X.mean(1, keepdims=True)) / X.std(1, keepdims=True)).astype(np.float32)   # LayerNorm output
    E = (rng.standard_normal((N, d)) * 0.02).astype(np.float32)                         # SASRec init std
    E[0] = 0
    tgt = rng.integers(1, N, nq)
    Xh, th = torch.from_numpy(X).pin_memory(), torch.from_numpy(tgt).pin_memory()
    Xd, Ed, td = Xh.to(dev), torch.from_numpy(E).to(dev), th.to(dev)
    for _ in range(max(args.warmup // 2, 3)):
        out = ops.ce_head(Xd, Ed, td, K)
        ids_tc, _ = ops.fullsort_topk(Xd, None, Ed, K, mode="tc")
    out32 = ops.ce_head(Xd, Ed, td, K, scorer="fp32")
    torch.cuda.synchronize()
    assert torch.equal(out["ids"], ids_tc) and torch.equal(out["ids"], out32["ids"])    # all scorers agree bit for bit
    assert abs(out["loss"].item() - out32["loss"].item()) <= 1e-5 * abs(out32["loss"].item())
    steps = max(min(args.steps, 20), 3)
    ops.profile_enable(True)
    ops.profile_read()
    clocks = ClockSampler(0)
    clocks.start()
    e0, e1 = _events(torch)
    e0.record()
    for _ in range(steps):
        out = ops.ce_head(Xd, Ed, td, K)
    e1.record()
    torch.cuda.synchronize()
    ms_ce = e0.elapsed_time(e1) / steps
    st_ce = ops.profile_read()
    e0.record()
    for _ in range(3):
        ops.ce_head(Xd, Ed, td, K, scorer="fp32")
    e1.record()
    torch.cuda.synchronize()
    ms_fp32 = e0.elapsed_time(e1) / 3
    ops.profile_read()
    # backward of the head (rb2_ce_head_backward: dX and dE, logits and softmax - onehot never materialised) and one dense
    # Adam step on the item table with its gradient = the training step of the head (sasrec.py:137-141, trainer.py:170-173)
    bws = ops.Workspace(ops.lib.rb2_ce_head_backward_workspace_bytes(nq, N, d), dev)
    mE, vE, opt = torch.zeros_like(Ed), torch.zeros_like(Ed), ops.Optim("adam", 1e-3)
    Etrain = Ed.clone()
    for _ in range(2):
        ops.ce_head_backward(Xd, Etrain, td, out["lse"], ws=bws)
    torch.cuda.synchronize()
    ops.profile_read()
    e0.record()
    for i in range(steps):
        o_tr = ops.ce_head(Xd, Etrain, td, 1)
        dx, de = ops.ce_head_backward(Xd, Etrain, td, o_tr["lse"], ws=bws)
        opt.step += 1
        ops.dense_step(Etrain, mE, vE, de, opt)
    e1.record()
    torch.cuda.synchronize()
    ms_train = e0.elapsed_time(e1) / steps
    st_bw = ops.profile_read()
    e0.record()
    for i in range(steps):
        ops.ce_head_backward(Xd, Etrain, td, o_tr["lse"], ws=bws)
    e1.record()
    torch.cuda.synchronize()
    ms_bwd = e0.elapsed_time(e1) / steps
    ops.profile_read()
    del Etrain, mE, vE, dx, de
    ops.ce_head(Xd[:8], Ed[:64], td[:8] % 64, K, scorer="auto")
    ops.profile_enable(False)
    t0 = time.perf_counter()
    for _ in range(steps):
        o = ops.ce_head(Xh.to(dev, non_blocking=True), Ed, th.to(dev, non_blocking=True), K)
        host_ids = o["ids"].cpu()
        host_loss = o["loss"].item()
    e2e_s = (time.perf_counter() - t0) / steps
    clk = clocks.stop()
    cpu = None
    if not args.skip_cpu:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        rows_cpu = 256
        Ec = torch.from_numpy(E)
        torch_port.ce_head(torch.from_numpy(X[:32]), Ec, torch.from_numpy(tgt[:32]), K)
        t0 = time.perf_counter()
        torch_port.ce_head(torch.from_numpy(X[:rows_cpu]), Ec, torch.from_numpy(tgt[:rows_cpu]), K)
        dt = time.perf_counter() - t0
        cpu = {"value": rows_cpu / dt, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": "%d of the %d rows: torch.matmul + cross_entropy + masked topk on CPU (the full batch would "
                         "materialise 16.4 GB of logits)" % (rows_cpu, nq)}
    flops = 2.0 * nq * N * d                       # the algorithmic GEMM; the split-precision kernel executes 3x that
    tc_ms = st_ce["tc_score"][0] / st_ce["tc_score"][1]
    line = {
        "metric": "ce_head_rows_per_s", "value": nq / (ms_ce * 1e-3), "unit": "rows/s", "n_gpus": 1, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms_ce, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x2 split operands, f32 accumulate (logits ~2^-16), f32 logsumexp", "data": "synthetic",
        "config": {"workload": "cfg4", "desc": "SASRec-shaped full-sort CE head: fused GEMM + logsumexp + top-10, logits never "
                   "materialised", "rows": nq, "n_items": N, "dim": d, "topk": K,
                   "l2": "item table 256 MB (fp32) / 384 MB (split bf16) > 126 MB L2"},
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "k_fullsort_tc<KB=3, LSE> (hi.hi + hi.lo + lo.hi in one K=192 GEMM, online "
                     "logsumexp + candidate lists in the epilogue)",
                     "achieved": flops / (tc_ms * 1e-3) / 1e12, "executed": 3 * flops / (tc_ms * 1e-3) / 1e12,
                     "peak": peaks["tf"], "unit": "TFLOP/s", "frac": flops / (tc_ms * 1e-3) / 1e12 / peaks["tf"],
                     "frac_executed": 3 * flops / (tc_ms * 1e-3) / 1e12 / peaks["tf"], "traffic": None, "ms_per_launch": tc_ms,
                     "stages_ms": {k: v[0] / max(v[1], 1) for k, v in st_ce.items()},
                     "note": "achieved counts the algorithmic 2*rows*items*d flops; the kernel executes 3x (split precision) "
                             "and one exp per logit (4.1e9 MUFU.EX2 per call)"},
        "fp32_cuda_core_kernel_ms": ms_fp32,
        "backward": {"metric": "ce_head_backward_rows_per_s", "value": nq / (ms_bwd * 1e-3), "unit": "rows/s",
                     "ms_per_call": ms_bwd, "kernel": "k_ce_bwd<dX> + k_ce_bwd<dE> (tcgen05, FP16 two-way splits, logits "
                     "recomputed in both passes)", "achieved": 3 * flops / (ms_bwd * 1e-3) / 1e12, "unit_roofline": "TFLOP/s",
                     "frac": 3 * flops / (ms_bwd * 1e-3) / 1e12 / peaks["tf"],
                     "note": "algorithmic flops = 3 x 2*rows*items*d (logits, dX, dE); the two passes execute 12 such "
                             "units of MMA work (3 split terms x (2 x logits + dX + dE))",
                     "train_step_ms": ms_train, "train_step": "forward (loss, lse) + backward + dense Adam on the item table",
                     "stages_ms": {k: v[0] / max(v[1], 1) for k, v in st_bw.items() if v[1]}},
        "cpu_baseline": cpu,
        "e2e": {"value": nq / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": nq * (4 * d + 8),
                "d2h_bytes_per_step": nq * K * 8 + 4},
        "gpu_launches": int(sum(v[2] for v in st_ce.values())), "loss": float(host_loss),
    }
    print(json.dumps(line), flush=True)


def run_cfg5_multi(args, rank, world, local_rank, load_peaks, ClockSampler):
    """cfg5 across N GPUs: the 33.76M-row table row-sharded, the batch split by rows (every rank feeds
    `train_batch` samples per step: weak scaling), NCCL all-to-all of rows and gradients."""
    import torch
    import torch.distributed as dist
    from recbole_b200 import ops
    from recbole_b200.dist import Comm, ShardedFM
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    comm = Comm()
    peaks = load_peaks()
    card, d, F = CRITEO_CARD, 16, 26
    B = args.batch or (1 << 18)
    m = ShardedFM(card.tolist(), d, comm, dev, seed=2020)
    m.build_optimizer("adam", 1e-3)
    m.ids_ready = True                  # resident batches
    batches = fm_batches(card, B, args.n_batches, seed=2020 + rank)
    host = [(torch.from_numpy(i).pin_memory(), torch.from_numpy(l).pin_memory()) for i, l in batches]
    res = [(i.to(dev), l.to(dev)) for i, l in host]
    nb, GB = len(res), B * world

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def mx(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(args.warmup):
        m.train_step(*res[i % nb], global_batch=GB, next_batch=res[(i + 1) % nb][0])
    barrier()
    ops.profile_enable(True)
    ops.profile_read()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = _events(torch)
    barrier()
    e0.record()
    for i in range(args.steps):
        m.train_step(*res[(args.warmup + i) % nb], global_batch=GB, next_batch=res[(args.warmup + i + 1) % nb][0])
    e1.record()
    barrier()
    ms = mx(e0.elapsed_time(e1)) / args.steps
    stages = ops.profile_read()
    ops.profile_enable(False)
    loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ids, lab = host[(args.warmup + i) % nb]
        lo = m.train_step(ids.to(dev, non_blocking=True), lab.to(dev, non_blocking=True), global_batch=GB)
        loss_host[i:i + 1].copy_(lo.reshape(1), non_blocking=True)
    barrier()
    e2e_s = mx(time.perf_counter() - t0)
    clk = clocks.stop()
    if rank == 0:
        bytes_per_sample = F * (24 * d + 24) + 8 * F + 4
        upd = stages.get("fm_update", (0.0, 1, 0))
        line = {
            "metric": "fm_train_samples_per_s", "value": GB / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg5", "desc": "FM CTR, synthetic Criteo shape: 26 categorical fields, 33.76M total vocab, "
                       "d=16, Adam; table row-sharded x%d, batch split by rows, NCCL all-to-all of rows and gradients" % world,
                       "fields": F, "rows": int(card.sum()), "dim": d, "train_batch_per_gpu": B, "global_batch": GB,
                       "l2": "no flush: every step reads a different batch"},
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "k_fm_rows(+fixup) on the fetched rows",
                         "achieved": None, "peak": peaks["hbm"], "unit": "GB/s", "frac": None, "traffic": None,
                         "peak_source": peaks["source"], "ms_per_launch": upd[0] / max(upd[1], 1),
                         "step": {"achieved_all_gpus": GB * bytes_per_sample / (ms * 1e-3) / 1e9,
                                  "frac_of_n_gpu_peak": GB * bytes_per_sample / (ms * 1e-3) / 1e9 / (world * peaks["hbm"]),
                                  "bytes_per_sample": bytes_per_sample},
                         "stages_ms_per_step_rank0": {k: v[0] / args.steps for k, v in stages.items()}},
            "cpu_baseline": None,
            "e2e": {"value": GB * args.steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": GB * (8 * F + 4),
                    "d2h_bytes_per_step": 4 * world},
            "gpu_launches": int(sum(v[2] for v in stages.values())), "loss": float(m.loss2[0].item()),
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_reference(args):
    """`bench.py --impl reference --workload cfg4|cfg5`: the reference's own CPU path (oracle/torch_port.py: the
    torch calls the reference makes) on the host cores, a bounded sample of the workload, same metric and unit."""
    import torch
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    if args.workload == "cfg5":
        card, d = CRITEO_CARD, 16
        B = args.batch or (1 << 18)
        steps = max(min(args.steps, args.cpu_steps), 1)
        m = torch_port.RefFM(card.tolist(), d)
        o = torch.optim.Adam(m.parameters(), lr=1e-3)
        cb = [(torch.from_numpy(i), torch.from_numpy(l)) for i, l in fm_batches(card, B, 2)]
        torch_port.fm_train_steps(m, o, cb[:1])                     # warm-up
        t0 = time.perf_counter()
        torch_port.fm_train_steps(m, o, [cb[i % 2] for i in range(steps)])
        dt = time.perf_counter() - t0
        value, unit, metric = steps * B / dt, "samples/s", "fm_train_samples_per_s"
        sample = "1 warm-up + %d timed steps of %d rows (FM forward, BCELoss, autograd, dense Adam over %d rows)" % (
            steps, B, int(card.sum()))
        config = {"workload": "cfg5", "fields": 26, "rows": int(card.sum()), "dim": d, "train_batch": B}
        ms = dt / steps * 1e3
    else:
        nq, N, d, K = 4096, 1_000_001, 64, 10
        rows_cpu = 256
        rng = np.random.default_rng(2020)
        X = rng.standard_normal((nq, d)).astype(np.float32)
        X = ((X - X.mean(1, keepdims=True)) / X.std(1, keepdims=True)).astype(np.float32)
        E = (rng.standard_normal((N, d)) * 0.02).astype(np.float32)
        E[0] = 0
        tgt = rng.integers(1, N, nq)
        Ec = torch.from_numpy(E)
        torch_port.ce_head(torch.from_numpy(X[:32]), Ec, torch.from_numpy(tgt[:32]), K)
        steps = max(min(args.steps, 3), 1)
        t0 = time.perf_counter()
        for s in range(steps):
            lo = (s * rows_cpu) % (nq - rows_cpu)
            torch_port.ce_head(torch.from_numpy(X[lo:lo + rows_cpu]), Ec, torch.from_numpy(tgt[lo:lo + rows_cpu]), K)
        dt = time.perf_counter() - t0
        value, unit, metric = steps * rows_cpu / dt, "rows/s", "ce_head_rows_per_s"
        sample = "%d x %d of the %d rows: torch.matmul + cross_entropy + masked topk (the full batch would materialise " \
                 "16.4 GB of logits)" % (steps, rows_cpu, nq)
        config = {"workload": "cfg4", "rows": nq, "n_items": N, "dim": d, "topk": K}
        ms = dt / steps * 1e3
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
