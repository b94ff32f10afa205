#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native embedding hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

Metric (BASELINE.json): BPR train samples/s (`value`) + full-sort eval users/s at top-10 (`eval`).
Default workload at every N: BASELINE.json configs[2] = cfg3 (10M users x 2M items x d=128, 2^20 triples per GPU
and step), the configuration the metric's "at 1/2/4/8 B200" and north_star's 60 % / 50 % targets are quoted on; it
fits one GPU (18.4 GB of tables + Adam state).  At N = 1 the line also carries `extra.cfg2` = the same measurement on
configs[1] (the synthetic ml-20m shape; `--no-extra` skips it).  One JSON line on stdout (rank 0).

A "step" is one fused training step over one batch of `train_batch` (user, pos, neg) triples.
  value      inputs resident in HBM, device-timed with CUDA events, every step a different batch
  e2e        the same through the public host API (FusedBPR.train_step) with the batch in pinned
             HOST memory: H2D of the ids and D2H of the loss inside the timed region, every step
  roofline   dominant kernel (k_user_fused) timed live by the library's per-stage CUDA events
  parity_check  a small problem through the same (multi-rank) path, checked against the oracle before timing
  cpu_baseline  oracle/torch_port (the reference's own torch calls) on the host cores, bounded sample
`--impl reference` times that CPU path alone, with every host thread, on the same workload.
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

import numpy as np  # noqa: E402


def env_int(k, d):
    return int(os.environ.get(k, d))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d["bf16_tflops"]), tf_sustained=float(d["bf16_tflops_sustained"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        late = False
        if not self.lines:
            # the timed region was shorter than nvidia-smi's start-up: one query right after it
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
                self.lines = [ln.strip() for ln in out.splitlines() if ln.strip()]
                late = True
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        out = dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                   reasons=sorted(reasons), samples=len(sm))
        if late:
            out["note"] = "timed region shorter than the sampler's start-up: sampled right after it"
        return out


# -------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/torch_port.py) -- cpu_baseline leg and --impl reference arm
# -------------------------------------------------------------------------------------------------

def cpu_reference(w, train_steps, warmup_steps, eval_users, threads, budget_s=240.0):
    """The reference's own torch calls on the host cores.  Every step is a FULL batch (dense gradients and dense
    Adam cost O(table) per step whatever the batch size, so truncating batches would understate the reference);
    what is bounded is the NUMBER of steps: if `warmup_steps + train_steps` would exceed `budget_s`, fewer timed
    steps are run (at least 3) and the line says so."""
    import torch
    from oracle import torch_port
    torch.set_num_threads(threads)
    torch.manual_seed(2020)
    model = torch_port.RefBPR(w.n_users, w.n_items, w.dim)
    if getattr(w, "U0", None) is not None:
        with torch.no_grad():
            model.user_embedding.weight.copy_(torch.from_numpy(w.U0))
            model.item_embedding.weight.copy_(torch.from_numpy(w.V0))
    opt = torch_port.build_optimizer(model, "adam", 1e-3, 0.0)
    batches = [tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in b) for b in w.batches]
    nb = len(batches)
    t0 = time.perf_counter()
    torch_port.train_steps(model, opt, [batches[0]])          # first step: allocates the dense grads / Adam state
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    torch_port.train_steps(model, opt, [batches[1 % nb]])
    t1 = time.perf_counter() - t0
    warm_left = max(warmup_steps - 2, 0)
    steps = train_steps
    if t1 * (warm_left + steps) > budget_s:
        warm_left = min(warm_left, 1)
        steps = max(3, min(train_steps, int(budget_s / t1) - warm_left))
    torch_port.train_steps(model, opt, [batches[(2 + i) % nb] for i in range(warm_left)])
    t0 = time.perf_counter()
    loss = torch_port.train_steps(model, opt, [batches[(2 + warm_left + i) % nb] for i in range(steps)])
    t_train = time.perf_counter() - t0
    ne = min(eval_users, len(w.uid_list))
    hist = (w.hist[0][:ne + 1], w.hist[1])
    pos = (w.pos[0][:ne + 1], w.pos[1])
    t0 = time.perf_counter()
    res, _ = torch_port.full_sort_eval(model, w.uid_list[:ne], hist, pos, w.n_items, topk=(10,))
    t_eval = time.perf_counter() - t0
    return dict(train_samples_per_s=steps * w.batch / t_train, train_s=t_train, train_steps=steps,
                warmup_steps=2 + warm_left, per_step=w.batch, first_step_s=t_first,
                eval_users_per_s=ne / t_eval, eval_s=t_eval, eval_users=ne, loss=loss, result=res)


def _host_workload(args):
    import bench_workloads as bw
    if args.workload == "cfg3":
        return bw.Cfg3Host(batch=args.batch or (1 << 20), n_batches=4, eval_users=args.ref_eval_users, scale=args.scale)
    return bw.BprWorkload(args.workload, batch=args.batch, n_batches=args.n_batches)


def _ref_sample_text(r, w):
    return ("%d warm-up + %d timed steps of %d triples each (full batches; torch CPU: dense grads + dense Adam over "
            "the %d x %d and %d x %d tables%s); eval of %d users (%.0f users/s)" % (
                r["warmup_steps"], r["train_steps"], r["per_step"], w.n_users, w.dim, w.n_items, w.dim,
                "; throughput extrapolated from this bounded number of steps" if w.name == "cfg3" else "",
                r["eval_users"], r["eval_users_per_s"]))


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = _host_workload(args)
    threads = os.cpu_count() or 1
    r = cpu_reference(w, args.steps, args.warmup, args.ref_eval_users, threads)
    line = {
        "impl": "reference", "metric": "bpr_train_samples_per_s", "value": r["train_samples_per_s"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": r["train_steps"], "warmup": r["warmup_steps"],
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * r["train_s"] / r["train_steps"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": w.describe(),
        "cpu_baseline": {"value": r["train_samples_per_s"], "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": _ref_sample_text(r, w)},
        "e2e": {"value": r["train_samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "eval": {"metric": "fullsort_eval_users_per_s", "value": r["eval_users_per_s"], "unit": "users/s",
                 "users": r["eval_users"], "result": r["result"]},
        "gpu_launches": 0,
    }
    print(json.dumps(line, default=float), flush=True)


def cpu_baseline_leg(args):
    """cpu_baseline of the GPU arm's line (rank 0, N = 1): a bounded sample of the same workload."""
    w = _host_workload(args)
    threads = os.cpu_count() or 1
    cb = cpu_reference(w, args.cpu_steps, 2, args.ref_eval_users, threads, budget_s=60.0)
    return {"value": cb["train_samples_per_s"], "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": _ref_sample_text(cb, w), "eval_users_per_s": cb["eval_users_per_s"]}


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------

def run_ours(args, rank, world, local_rank, emit=True):
    import torch
    import torch.distributed as dist

    import bench_workloads as bw
    from recbole_b200 import FusedBPR, Interaction, ops
    from recbole_b200.data import EvalIndex
    from recbole_b200.evaluator import FusedTopKEvaluator

    if world > 1 or args.workload == "cfg3":
        line = run_ours_multi(args, rank, world, local_rank)
        if line is not None:
            if world == 1 and args.workload == "cfg3" and not args.no_extra:
                # configs[1] (the synthetic ml-20m shape) beside the headline, same process, same GPU
                import copy
                a2 = copy.copy(args)
                a2.workload, a2.batch, a2.eval_reps = "cfg2", None, 3
                torch.cuda.empty_cache()
                line["extra"] = {"cfg2": run_ours(a2, rank, world, local_rank, emit=False)}
            print(json.dumps(line, default=float), flush=True)
        return line

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peaks = load_peaks()
    w = bw.BprWorkload(args.workload, batch=args.batch, n_batches=args.n_batches)
    B, d = w.batch, w.dim

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k)

    class DS:
        def num(self, f):
            return {"user_id": w.n_users, "item_id": w.n_items}[f]

    cfg = Cfg(USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", NEG_PREFIX="neg_", device=dev, embedding_size=d,
              metrics=["Recall", "MRR", "NDCG", "Hit", "Precision"], topk=[10], metric_decimal_place=4)
    model = FusedBPR(cfg, DS()).to(dev)
    with torch.no_grad():
        model.user_embedding.weight.copy_(torch.from_numpy(w.U0))
        model.item_embedding.weight.copy_(torch.from_numpy(w.V0))
    model.build_optimizer("adam", 1e-3, 0.0)

    host = [Interaction({"user_id": torch.from_numpy(np.ascontiguousarray(u)).pin_memory(),
                         "item_id": torch.from_numpy(np.ascontiguousarray(p)).pin_memory(),
                         "neg_item_id": torch.from_numpy(np.ascontiguousarray(n)).pin_memory()})
            for (u, p, n) in w.batches]
    resident = [b.to(dev) for b in host]
    nb = len(resident)
    torch.cuda.synchronize()

    # ---- value: K steps, inputs resident, CUDA events -------------------------------------------------
    for i in range(args.warmup):
        model.train_step(resident[i % nb])
    torch.cuda.synchronize()
    ops.profile_enable(True)
    ops.profile_read()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        model.train_step(resident[(args.warmup + i) % nb])
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    stages = ops.profile_read()
    ops.profile_enable(False)
    train_ms = ms_total / args.steps
    value = B * args.steps / (ms_total / 1e3)
    final_loss = float(model._loss_out.item())

    # ---- e2e: host batches, H2D + D2H inside the timed region ------------------------------------------
    copy_stream = torch.cuda.Stream()
    slots = [{k: torch.empty(B, dtype=torch.int64, device=dev) for k in ("user_id", "item_id", "neg_item_id")}
             for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(args.steps + args.warmup, dtype=torch.float32).pin_memory()

    def e2e_loop(n, offset):
        main = torch.cuda.current_stream()
        for i in range(n):
            s = i % 2
            hb = host[(offset + i) % nb]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[s])
                for k in slots[s]:
                    slots[s][k].copy_(hb[k], non_blocking=True)
                ready[s].record(copy_stream)
            main.wait_event(ready[s])
            lo = model.train_step(Interaction(slots[s]))
            loss_host[i:i + 1].copy_(lo, non_blocking=True)
            free[s].record(main)

    for s in range(2):
        free[s].record(torch.cuda.current_stream())
    e2e_loop(args.warmup, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_loop(args.steps, args.warmup)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = B * args.steps / e2e_s
    clk = clocks.stop()

    # ---- device-resident batch pipeline: one EPOCH through DeviceTrainLoader -- the train
    # interactions live in HBM, the epoch is a device-side permutation, negatives come from the sampler kernel
    # (rejecting the user's train positives from a device CSR), every batch feeds the fused step; no host round trip
    from recbole_b200 import DeviceSampler, DeviceTrainLoader
    from recbole_b200.data import build_csr
    tu, ti = (torch.from_numpy(x).to(dev) for x in w.phases[0])
    used = build_csr(w.n_users, tu, ti, w.n_items, dev)
    loader = DeviceTrainLoader(tu, ti, DeviceSampler(w.n_items, used[0], used[1], mode="hash", seed=2020), B)
    n_train = int(tu.numel())
    it = iter(loader)
    for _ in range(2):
        model.train_step(next(it))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    n_seen = 0
    for inter in loader:                                       # one full pass over the train interactions
        model.train_step(inter)
        n_seen += len(inter)
    e1.record()
    torch.cuda.synchronize()
    pipe_ms, pipe_wall = e0.elapsed_time(e1), time.perf_counter() - t0
    pipeline = {"metric": "bpr_epoch_samples_per_s", "value": n_seen / (pipe_ms * 1e-3), "unit": "samples/s",
                "interactions": n_train, "batches": len(loader), "epoch_ms": pipe_ms, "wall_s": pipe_wall,
                "what": "DeviceTrainLoader (device permutation + gather) + rb2_neg_sample_hash + fused step, per batch; "
                        "replaces GeneralNegSampleDataLoader._next_batch_data + Interaction.to + Sampler.sample_by_user_ids "
                        "(general_dataloader.py:212-241, trainer.py:158-159, sampler.py:103-154)"}
    del loader, tu, ti, used

    # ---- evaluation: all test users, top-10, metrics on device ------------------------------------------
    index = EvalIndex(torch.from_numpy(w.uid_list).to(dev), (torch.from_numpy(w.hist[0]).to(dev),
                                                             torch.from_numpy(w.hist[1]).to(dev)),
                      (torch.from_numpy(w.pos[0]).to(dev), torch.from_numpy(w.pos[1]).to(dev)), w.n_items)
    ev = FusedTopKEvaluator(cfg)
    nq = index.n_eval_users
    from recbole_b200._lib import lib as _lib
    _lib.rb2_fullsort_tc_set_kprime(args.tc_kprime)

    def eval_once():
        ids, _ = model.full_sort_topk(index.uid_list, 10, index.hist_indptr, index.hist_indices, mode=args.scorer)
        return ev.sums(ids, index)

    eval_once()
    torch.cuda.synchronize()
    ops.profile_enable(True)
    ops.profile_read()
    e0.record()
    for _ in range(args.eval_reps):
        sums = eval_once()
    e1.record()
    torch.cuda.synchronize()
    eval_ms = e0.elapsed_time(e1) / args.eval_reps
    estages = ops.profile_read()
    ops.profile_enable(False)
    result = ev.result(sums, nq)
    # e2e eval: user ids from pinned host memory, result dict on the host
    uid_host = torch.from_numpy(w.uid_list).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.eval_reps):
        uid_dev = uid_host.to(dev, non_blocking=True)
        ids, _ = model.full_sort_topk(uid_dev, 10, index.hist_indptr, index.hist_indices, mode=args.scorer)
        res2 = ev.result(ev.sums(ids, index), nq)  # D2H of 6*K doubles
    eval_e2e_s = (time.perf_counter() - t0) / args.eval_reps

    # ---- roofline objects ---------------------------------------------------------------------------------
    def stage_ms(st, name):
        return st[name][0] / st[name][1] if name in st else 0.0

    us_ms = stage_ms(stages, "user_side")
    alg_step = B * (72 * d + 24)                      # bytes per triple, row-sparse Adam
    alg_user = B * (32 * d + 24)                      # the share k_user_side must move
    step_kernels_ms = sum(v[0] for v in stages.values()) / args.steps
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tp) and args.workload == "cfg2" and B == (1 << 20):
        traffic = json.load(open(tp)).get("%s@cfg2" % ("k_user_side" if B >= w.n_users else "k_user_fused"))
    dense_batch = B >= w.n_users and 2 * B >= 4 * w.n_items        # train_bpr.cu launch_step: register-path kernels
    kname = "k_user_side" if dense_batch else "k_user_fused"
    roofline = {"bound": "hbm", "kernel": kname, "achieved": alg_user / (us_ms * 1e-3) / 1e9 if us_ms else None,
                "peak": peaks["hbm"], "unit": "GB/s", "frac": None, "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_user,
                "peak_source": peaks["source"], "ms_per_launch": us_ms,
                "share_of_step": us_ms / step_kernels_ms if step_kernels_ms else None,
                "step": {"achieved": alg_step / (train_ms * 1e-3) / 1e9, "frac": alg_step / (train_ms * 1e-3) / 1e9 / peaks["hbm"],
                         "bytes_per_sample": 72 * d + 24,
                         "note": "whole fused step vs the no-duplicate algorithmic bytes; at this shape a batch of "
                                 "2^20 triples touches each user row ~8x and each item row ~78x, so the bytes that "
                                 "really reach DRAM are far below the algorithmic count (see traffic)"},
                "stages_ms_per_step": {k: v[0] / args.steps for k, v in stages.items()}}
    if roofline["achieved"]:
        roofline["frac"] = roofline["achieved"] / peaks["hbm"]
    fs_stage = "tc_score" if args.scorer == "tc" else "fullsort"
    fs_ms = stage_ms(estages, fs_stage)
    flops = 2.0 * nq * w.n_items * d
    eval_roof = {"bound": "tensor", "kernel": "k_fullsort_%s" % ("tc" if args.scorer == "tc" else "fp32"),
                 "achieved": flops / (fs_ms * 1e-3) / 1e12 if fs_ms else None, "peak": peaks["tf"],
                 "unit": "TFLOP/s", "frac": (flops / (fs_ms * 1e-3) / 1e12 / peaks["tf"]) if fs_ms else None,
                 "traffic": None, "ms_per_launch": fs_ms, "stages_ms": {k: v[0] / max(v[1], 1) for k, v in estages.items()}}
    launches = sum(v[2] for v in stages.values())

    # ---- CPU baseline (bounded sample, rank 0) ------------------------------------------------------------
    threads = os.cpu_count() or 1
    if args.skip_cpu:
        cpu_baseline = None
    else:
        cb = cpu_reference(w, args.cpu_steps, 2, args.ref_eval_users, threads, budget_s=60.0)
        cpu_baseline = {"value": cb["train_samples_per_s"], "unit": "samples/s", "cores": threads, "kind": "port",
                        "sample": _ref_sample_text(cb, w), "eval_users_per_s": cb["eval_users_per_s"]}

    line = {
        "metric": "bpr_train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": train_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(w.describe(), optimizer="adam(row-sparse) lr=1e-3", scorer=args.scorer,
                       l2="no flush: every step reads a different batch; per-step working set (tables + Adam "
                          "state %.0f MB, ids %.0f MB, scratch %.0f MB) exceeds the 126 MB L2"
                          % (3 * 4 * d * (w.n_users + w.n_items) / 1e6, 24 * B / 1e6, 4 * d * B / 1e6)),
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": 24 * B, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "loss": final_loss,
        "pipeline": pipeline,
        "eval": {"metric": "fullsort_eval_users_per_s", "value": nq / (eval_ms * 1e-3), "unit": "users/s",
                 "users": nq, "ms": eval_ms, "topk": 10, "roofline": eval_roof,
                 "e2e": {"value": nq / eval_e2e_s, "unit": "users/s", "h2d_bytes_per_step": 8 * nq,
                         "d2h_bytes_per_step": 6 * 10 * 8},
                 "result": result, "gpu_launches": int(sum(v[2] for v in estages.values())),
                 "tc_pass2_rows": int(_lib.rb2_fullsort_tc_last_pass2_rows()) if args.scorer == "tc" else None,
                 "tc_fallback_rows": int(_lib.rb2_fullsort_tc_last_fallback_rows()) if args.scorer == "tc" else None},
    }
    assert res2 == result
    if emit:
        print(json.dumps(line, default=float), flush=True)
    return line


def run_ours_multi(args, rank, world, local_rank):
    import bench_cfg3
    return bench_cfg3.run(args, rank, world, local_rank, load_peaks, ClockSampler, cpu_reference=cpu_baseline_leg)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--n-batches", type=int, default=8, dest="n_batches")
    ap.add_argument("--scorer", default="tc", choices=["fp32", "tc"])
    ap.add_argument("--eval-reps", type=int, default=None, dest="eval_reps")
    ap.add_argument("--cpu-steps", type=int, default=3, dest="cpu_steps")
    ap.add_argument("--ref-eval-users", type=int, default=None, dest="ref_eval_users")
    ap.add_argument("--eval-layout", default="auto", choices=["auto", "replicate", "sharded"], dest="eval_layout")
    ap.add_argument("--exchange", default="auto", choices=["auto", "dense", "sparse", "p2p"])
    ap.add_argument("--no-extra", action="store_true", dest="no_extra", help="N=1: skip the cfg2 measurement")
    ap.add_argument("--skip-parity", action="store_true", dest="skip_parity", help="profiling runs only")
    ap.add_argument("--scale", type=float, default=1.0, help="cfg3 only: shrink users/items by this factor")
    ap.add_argument("--skip-cpu", action="store_true", dest="skip_cpu", help="profiling runs only")
    ap.add_argument("--tc-kprime", type=int, default=0, dest="tc_kprime", choices=[0, 16, 32],
                    help="candidates per list of the tensor-core scorer (0 = automatic)")
    args = ap.parse_args()
    if args.eval_reps is None:
        args.eval_reps = 1 if args.workload == "cfg3" else 3     # one cfg3 pass is 5 PFLOP: seconds on one GPU
    if args.ref_eval_users is None:
        args.ref_eval_users = 512 if args.workload == "cfg3" else 4096
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.workload == "cfg5" and world > 1 and args.impl != "reference":
        import bench_extra
        bench_extra.run_cfg5_multi(args, rank, world, local_rank, load_peaks, ClockSampler)
        return
    if args.workload in ("cfg4", "cfg5") and args.impl == "reference":
        if rank == 0:
            import bench_extra
            bench_extra.run_reference(args)
        return
    if args.workload in ("cfg4", "cfg5"):
        if rank == 0 and args.impl != "reference":
            import bench_extra
            (bench_extra.run_cfg4 if args.workload == "cfg4" else bench_extra.run_cfg5)(args, load_peaks, ClockSampler)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
