"""Multi-GPU path (the reference has no multi-device code at all).

One process per GPU.  Users are range-partitioned together with their interactions, their rows of
the user table and their optimizer state, so the user side of a step is local.  The ITEM table
(+ m, v) is row-sharded in contiguous blocks.  Per training step each rank

  1. de-duplicates the item ids of its local batch (sorted unique => already grouped by owner),
  2. all-to-all #1: unique ids out, item rows back  -> compact table C [n_uniq, d],
  3. runs the fused user side + per-compact-row gradient sums on C (rb2_bpr_train_step_sharded),
  4. all-to-all #2: gradient rows back to the owners,
  5. owners sum what every rank sent and take one optimizer step per touched row
     (rb2_sparse_rows_update),
  6. all-reduces one scalar (the loss, a mean over the GLOBAL batch).

The result equals the single-GPU step on the union of the ranks' batches up to fp32 summation
order.  Evaluation: the user table is all-gathered once, every rank scores every evaluated user
against its item shard (mask from its column-slice of the history CSR), the per-shard top-K lists
go to the users' owners (all-to-all), are merged (rb2_topk_merge) and reduced to metric sums
(rb2_topk_metrics); one all-reduce of 6*K doubles finishes the job.

Collectives go through torch.distributed (NCCL over NVLink on the GPU box).  `Comm(staged=True)`
stages them through host memory over gloo so that several ranks can share ONE GPU in tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous equal blocks of ceil(n / world) rows (the last ones may be short or empty): rank g
    owns [b[g], b[g+1]).  Equal blocks let all-gather / reduce-scatter run without padding copies."""
    s = (n + world - 1) // world
    return np.array([min(g * s, n) for g in range(world + 1)], dtype=np.int64)


class Comm:
    def __init__(self, group=None, staged=False):
        self.group = group
        self.staged = staged
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def _stage(self, t):
        return t.cpu() if (self.staged and t.is_cuda) else t

    def exchange_counts(self, send_counts):
        """send_counts: python list [world] -> what every rank sends me."""
        if self.world == 1:
            return list(send_counts)
        on_gpu = (not self.staged) and dist.get_backend(self.group) == "nccl"
        dev = torch.cuda.current_device() if on_gpu else "cpu"
        s = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        r = torch.empty_like(s)
        dist.all_to_all_single(r, s, group=self.group)
        return r.tolist()

    def all_to_all(self, inp, send_counts, recv_counts):
        """Rows of `inp` (first dim) split by send_counts -> concatenation of what each rank sent me."""
        out_shape = (int(sum(recv_counts)),) + tuple(inp.shape[1:])
        if self.world == 1:
            return inp.clone()
        if self.staged:
            src = inp.cpu().contiguous()
            out = torch.empty(out_shape, dtype=inp.dtype)
            dist.all_to_all_single(out, src, list(recv_counts), list(send_counts), group=self.group)
            return out.to(inp.device)
        out = torch.empty(out_shape, dtype=inp.dtype, device=inp.device)
        dist.all_to_all_single(out, inp.contiguous(), list(recv_counts), list(send_counts), group=self.group)
        return out

    def all_gather_equal(self, local):
        """all-gather of equally sized blocks -> [world * rows, ...]."""
        if self.world == 1:
            return local
        shape = (self.world * local.shape[0],) + tuple(local.shape[1:])
        if self.staged:
            out = torch.empty(shape, dtype=local.dtype)
            dist.all_gather_into_tensor(out, local.cpu().contiguous(), group=self.group)
            return out.to(local.device)
        out = torch.empty(shape, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out

    def reduce_scatter_sum(self, full):
        """[world * rows, ...] summed over ranks -> my block [rows, ...]."""
        if self.world == 1:
            return full
        rows = full.shape[0] // self.world
        shape = (rows,) + tuple(full.shape[1:])
        if self.staged:
            # gloo has no reduce_scatter: all-reduce on the host, keep my block
            c = full.cpu().contiguous()
            dist.all_reduce(c, group=self.group)
            return c[self.rank * rows:(self.rank + 1) * rows].to(full.device)
        out = torch.empty(shape, dtype=full.dtype, device=full.device)
        dist.reduce_scatter_tensor(out, full.contiguous(), group=self.group)
        return out

    def all_reduce_sum(self, t):
        if self.world == 1:
            return t
        if self.staged and t.is_cuda:
            c = t.cpu()
            dist.all_reduce(c, group=self.group)
            t.copy_(c)
            return t
        dist.all_reduce(t, group=self.group)
        return t

    def all_gather_rows(self, local, counts):
        """Concatenate every rank's rows (uneven counts allowed)."""
        if self.world == 1:
            return local
        mx = int(max(counts))
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        if self.staged:
            src = pad.cpu()
            out = torch.empty((self.world * mx,) + tuple(local.shape[1:]), dtype=local.dtype)
            dist.all_gather_into_tensor(out, src, group=self.group)
            out = out.to(local.device)
        else:
            out = torch.empty((self.world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(out, pad, group=self.group)
        if all(int(c) == mx for c in counts):
            return out
        return torch.cat([out[g * mx: g * mx + int(c)] for g, c in enumerate(counts)], dim=0)

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)

    def all_gather_object(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out


# ---- exchange plumbing (pure index work; unit-tested on CPU with gloo) ------------------------------

def plan_item_exchange(items, item_bounds):
    """items int64[n] (global ids).  Returns (uniq sorted, inverse, send_counts list): uniq is
    grouped by owner because shards are contiguous id ranges."""
    uniq, inv = torch.unique(items, sorted=True, return_inverse=True)
    cut = torch.searchsorted(uniq, torch.as_tensor(item_bounds, device=uniq.device))
    send_counts = (cut[1:] - cut[:-1]).tolist()
    return uniq, inv, send_counts


def fetch_rows(comm, uniq, send_counts, gather_fn, row_base):
    """all-to-all #1.  gather_fn(local_row_idx) -> rows [n, d] of the local shard.
    Returns (C [n_uniq, d], requested local idx, recv_counts)."""
    recv_counts = comm.exchange_counts(send_counts)
    req = comm.all_to_all(uniq, send_counts, recv_counts)          # ids other ranks want from me
    local_idx = req - int(row_base)
    rows = gather_fn(local_idx)
    C = comm.all_to_all(rows, recv_counts, send_counts)            # rows come back in uniq order
    return C, local_idx, recv_counts


def return_grads(comm, G, send_counts, recv_counts):
    """all-to-all #2: gradient rows (uniq order) to their owners; arrives aligned with the
    `local_idx` returned by fetch_rows."""
    return comm.all_to_all(G, send_counts, recv_counts)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ShardedBPR:
    """BPR-MF with range-partitioned users and a row-sharded item table."""

    def __init__(self, n_users, n_items, dim, comm, device, U_full=None, V_full=None, seed=2020, exchange="auto"):
        from . import ops
        self.ops = ops
        self.comm, self.device, self.dim = comm, device, dim
        self.n_users, self.n_items = n_users, n_items
        self.user_bounds = shard_bounds(n_users, comm.world)
        self.item_bounds = shard_bounds(n_items, comm.world)
        self.u_lo, self.u_hi = int(self.user_bounds[comm.rank]), int(self.user_bounds[comm.rank + 1])
        self.i_lo, self.i_hi = int(self.item_bounds[comm.rank]), int(self.item_bounds[comm.rank + 1])
        # every rank allocates the full block size (tail ranks pad), so collectives need no re-packing
        self.u_block = int(self.user_bounds[1] - self.user_bounds[0])
        self.i_block = int(self.item_bounds[1] - self.item_bounds[0])
        self.U = torch.zeros(self.u_block, dim, device=device)
        # "p2p": the kernels read the owners' rows and write the owners' gradient slots over NVLink themselves
        #        (rb2_bpr_train_step_p2p); the item shard lives in an arena every rank maps with cudaIpc.
        self.arena = None
        if self.resolve_exchange(exchange, n_items, dim, comm, device) == "p2p":
            exchange = "p2p"
            self.arena = ops.PeerArena(comm, device, self.i_block, dim)
            self.V = self.arena.item_p
        else:
            self.V = torch.zeros(self.i_block, dim, device=device)
        if U_full is not None:
            self.U[: self.u_hi - self.u_lo] = torch.as_tensor(U_full[self.u_lo:self.u_hi]).to(device)
            self.V[: self.i_hi - self.i_lo] = torch.as_tensor(V_full[self.i_lo:self.i_hi]).to(device)
        else:
            g = torch.Generator(device=device)
            g.manual_seed(seed + comm.rank)
            self.U[: self.u_hi - self.u_lo] = torch.randn(self.u_hi - self.u_lo, dim, device=device, generator=g) * (2.0 / (n_users + dim)) ** 0.5
            self.V[: self.i_hi - self.i_lo] = torch.randn(self.i_hi - self.i_lo, dim, device=device, generator=g) * (2.0 / (n_items + dim)) ** 0.5
        # "dense": all-gather the item rows, reduce-scatter their gradients (no data-dependent sizes,
        #          no host sync; right when the table is small next to the batch).
        # "sparse": all-to-all of the de-duplicated rows the batch touches (large tables).
        # "auto": dense (all-gather rows + reduce-scatter gradients, no plan, no host sync) for small tables and
        # whenever a rank's batch covers most of the item table anyway (B >= n_items: >= 86 % of the rows are
        # touched, measured at cfg3 on 8 GPUs: 7.9 ms dense vs 15.7 ms sparse at B = 2^22, 4.8 vs 4.1 at 2^20)
        self.exchange_auto = exchange == "auto" and self.arena is None
        if exchange == "auto":
            exchange = "dense" if n_items * dim * 4 <= (64 << 20) else "sparse"
        self.exchange = exchange
        self.optim = None
        self.state = {}
        self.loss_out = torch.zeros(1, dtype=torch.float32, device=device)
        self.loss_accum = torch.zeros(1, dtype=torch.float64, device=device)
        self._ws = {}
        self._rows_ws = None
        self._p2p_ws = None

    @staticmethod
    def resolve_exchange(exchange, n_items, dim, comm, device):
        """The exchange a model of this shape runs when asked for `exchange` ("auto": peer memory for big tables on
        2-8 CUDA ranks, else dense for small tables / sparse for big ones; the dense-vs-sparse choice can still flip
        per batch, see train_step)."""
        if exchange == "p2p" or (exchange == "auto" and device.type == "cuda" and not comm.staged
                                 and n_items * dim * 4 > (64 << 20) and 1 < comm.world <= 8):
            return "p2p"
        if exchange == "auto":
            return "dense" if n_items * dim * 4 <= (64 << 20) else "sparse"
        return exchange

    def check_flags(self):
        """Raise for sticky device-side errors (id out of range, peer barrier timeout); synchronises."""
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)      # (also drains the library's side stream of the peer-memory step)
        for ws in list(self._ws.values()) + ([self._p2p_ws[1]] if self._p2p_ws else []):
            ws.check_flags()

    def __del__(self):
        # the peer-memory step may have left the next batch's sorts running on the library's side stream: they write
        # into this model's workspace, which must not go back to the allocator before they are done
        try:
            if getattr(self, "_p2p_ws", None) is not None and self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
        except Exception:
            pass

    def build_optimizer(self, kind="adam", lr=1e-3, weight_decay=0.0):
        """'adam' (row-sparse), 'sgd', or 'adam_lazy' = the trajectory of the reference's DENSE torch.optim.Adam
        (trainer.py:116,173) on the peer-memory and the dense exchange: the batch's (local) user rows replay the steps
        they missed, every owner takes the zero-gradient step of the untouched rows of its item shard."""
        if kind not in ("adam", "sgd", "adam_lazy"):
            raise ValueError("the sharded path implements learner in {adam, adam_lazy, sgd}")
        if kind == "adam_lazy" and self.comm.world > 1 and self.exchange == "sparse":
            raise ValueError("adam_lazy is not available on the sparse (all-to-all) exchange: rows are fetched before "
                             "their owner could bring them up to date; use exchange='p2p' or 'dense'")
        self.optim = self.ops.Optim(kind, lr, weight_decay)
        if kind != "sgd":
            z = torch.zeros_like
            self.state = dict(mU=z(self.U), vU=z(self.U), mV=z(self.V), vV=z(self.V))
        if kind == "adam_lazy":
            self.state["lastU"] = torch.zeros(self.U.shape[0], dtype=torch.int32, device=self.device)
            if self.comm.world == 1 and self.exchange != "p2p":      # the single-GPU step keeps `last` for items too
                self.state["lastV"] = torch.zeros(self.V.shape[0], dtype=torch.int32, device=self.device)

    def flush(self):
        """adam_lazy: bring every local row to the current step (before the tables are read: evaluation, checkpoints).
        Item shards are always current on the sharded exchanges (the owner steps every row)."""
        o = self.optim
        if o is None or o.kind_name != "adam_lazy" or o.step == 0:
            return
        st = self.state
        self.ops.adam_lazy_flush(self.U, st["mU"], st["vU"], st["lastU"], o)
        if "lastV" in st:
            self.ops.adam_lazy_flush(self.V, st["mV"], st["vV"], st["lastV"], o)

    # ---- checkpoint interop (trainer.py:191-232) ------------------------------------------
    def state_dict(self):
        """The reference's parameter names with the FULL tables (every rank gets the same dict; the user table is
        all-gathered: 5 GB at cfg3 -- save from one rank)."""
        self.flush()
        U = self.comm.all_gather_equal(self.U)
        V = self.comm.all_gather_equal(self.V)
        # blocks are padded to equal size: rank r's rows sit at [r * block, r * block + its count)
        return {"user_embedding.weight": self._unpad(U, self.user_bounds, self.u_block),
                "item_embedding.weight": self._unpad(V, self.item_bounds, self.i_block)}

    def load_state_dict(self, sd):
        U, V = sd["user_embedding.weight"], sd["item_embedding.weight"]
        if tuple(U.shape) != (self.n_users, self.dim) or tuple(V.shape) != (self.n_items, self.dim):
            raise ValueError("state dict shapes %s / %s do not match the model (%d x %d, %d x %d)" % (
                tuple(U.shape), tuple(V.shape), self.n_users, self.dim, self.n_items, self.dim))
        self.U.zero_()
        self.V.zero_()
        self.U[: self.u_hi - self.u_lo] = U[self.u_lo:self.u_hi].to(self.device)
        self.V[: self.i_hi - self.i_lo] = V[self.i_lo:self.i_hi].to(self.device)

    def optimizer_state_dict(self):
        """`torch.optim.Adam.state_dict()` layout for params [user_embedding.weight, item_embedding.weight]
        (what `Trainer._save_checkpoint` stores, trainer.py:198-206)."""
        o = self.optim
        state = {}
        self.flush()
        if o.kind_name != "sgd":
            step = torch.tensor(float(o.step))
            g = lambda t, b, blk: self._unpad(self.comm.all_gather_equal(t), b, blk)   # noqa: E731
            state = {0: dict(step=step, exp_avg=g(self.state["mU"], self.user_bounds, self.u_block),
                             exp_avg_sq=g(self.state["vU"], self.user_bounds, self.u_block)),
                     1: dict(step=step.clone(), exp_avg=g(self.state["mV"], self.item_bounds, self.i_block),
                             exp_avg_sq=g(self.state["vV"], self.item_bounds, self.i_block))}
        group = dict(lr=o.lr, betas=o.betas, eps=o.eps, weight_decay=o.weight_decay, params=[0, 1])
        return dict(state=state, param_groups=[group], fused_kind=o.kind_name)

    def load_optimizer_state_dict(self, sd):
        o = self.optim
        grp = sd["param_groups"][0]
        o.set_hyper(lr=grp["lr"], weight_decay=grp["weight_decay"], betas=tuple(grp.get("betas", o.betas)),
                    eps=grp.get("eps", o.eps))          # (drops adam_lazy's bias-correction tables of the old lr)
        if sd["state"]:
            o.step = int(sd["state"][0]["step"])
            for k, idx, key, lo, hi in (("mU", 0, "exp_avg", self.u_lo, self.u_hi), ("vU", 0, "exp_avg_sq", self.u_lo, self.u_hi),
                                        ("mV", 1, "exp_avg", self.i_lo, self.i_hi), ("vV", 1, "exp_avg_sq", self.i_lo, self.i_hi)):
                self.state[k].zero_()
                self.state[k][: hi - lo] = sd["state"][idx][key][lo:hi].to(self.device)
            for k in ("lastU", "lastV"):          # a loaded state is current at its step
                if k in self.state:
                    self.state[k].fill_(o.step)

    @staticmethod
    def _unpad(full, bounds, block):
        world = len(bounds) - 1
        if world == 1:
            return full[: int(bounds[1])]
        return torch.cat([full[r * block: r * block + int(bounds[r + 1] - bounds[r])] for r in range(world)])

    def _workspace(self, batch):
        return self.ops.grow_workspace(self._ws, batch, lambda b: self.ops.bpr_workspace(b, self.dim, self.device))

    def plan(self, user, pos, neg, ids_ready=False):
        """Everything of a sparse-exchange step that depends only on the batch IDS (not on the
        parameters): de-duplication, bucketing by owner, the count exchange and all-to-all of the
        requested ids.  It contains the step's only host synchronisations, so train_step() runs it for
        the NEXT batch on a side stream while the current step's kernels execute."""
        comm = self.comm
        if not hasattr(self, "_plan_stream"):
            self._plan_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        st = self._plan_stream
        ctx = torch.cuda.stream(st) if st is not None else _NullCtx()
        if not hasattr(self, "_bounds_dev"):
            self._bounds_dev = torch.as_tensor(self.item_bounds, device=self.device)
            self._plan_ws = [None, None]
            self._plan_free = [None, None]            # event: the step that used this buffer has finished
            self._plan_flip = 0
        self._plan_flip ^= 1                          # two plans are alive at a time (current + next)
        if st is not None:
            if not ids_ready:
                st.wait_stream(torch.cuda.current_stream())
            elif self._plan_free[self._plan_flip] is not None:
                # the ids are complete already: only the step that read this plan buffer (two plans ago) must be over
                st.wait_event(self._plan_free[self._plan_flip])
        with ctx:
            B = int(user.numel())
            ip = self.ops.item_plan(pos, neg, self.n_items, self._bounds_dev, comm.world,
                                    self._plan_ws[self._plan_flip])
            self._plan_ws[self._plan_flip] = ip["ws"]
            cuts = ip["cuts"].tolist()                # the plan's host sync
            n_uniq = cuts[-1]
            send_counts = [cuts[g + 1] - cuts[g] for g in range(comm.world)]
            uniq = ip["uniq"][:n_uniq]
            recv_counts = comm.exchange_counts(send_counts)
            req = comm.all_to_all(uniq, send_counts, recv_counts)      # ids other ranks want from me
            p = dict(B=B, user_local=(user - self.u_lo).contiguous(), pos_c=ip["pos_c"], neg_c=ip["neg_c"],
                     n_uniq=n_uniq, send_counts=send_counts, recv_counts=recv_counts,
                     local_idx=(req - self.i_lo).contiguous(), ids=(user, pos, neg), plan_ws=ip["ws"],
                     flip=self._plan_flip)
            p["event"] = torch.cuda.Event() if st is not None else None
            if st is not None:
                p["event"].record(st)
        return p

    def train_step(self, user, pos, neg, global_batch=None, next_batch=None):
        """user/pos/neg: int64 device vectors with GLOBAL ids; every user must belong to this rank.
        Returns the device scalar holding the global mean loss.  next_batch=(user, pos, neg) lets the
        id-only part of the following step overlap with this one (sparse exchange)."""
        ops, comm = self.ops, self.comm
        B = int(user.numel())
        if global_batch is None:
            global_batch = B * comm.world
        self.optim.step += 1
        t = self.optim.step
        if comm.world == 1 and self.exchange != "p2p":
            # nothing to exchange: the single-GPU fused step on the whole tables
            self.last_exchange = "local"
            ops.bpr_train_step(self.U, self.V, self.state, user, pos, neg, self.optim, self.loss_out, self.loss_accum,
                               self._workspace(B), step=t)
            return self.loss_out
        if self.exchange == "p2p":
            self.last_exchange = "p2p"
            if self._p2p_ws is None or self._p2p_ws[0] < B:
                if self._p2p_ws is not None:
                    # the library's side stream may still be sorting the next batch's keys into the old workspace
                    torch.cuda.synchronize(self.device)
                self._p2p_ws = (B, ops.bpr_p2p_workspace(B, self.dim, self.device))
            # next_batch: its keys and sorts are computed inside this call, while it waits for the slower peers;
            # only sound when those id tensors are complete already (resident batches: ids_ready)
            ops.bpr_train_step_p2p(self.U, self.state, self.arena, user, self.u_lo, pos, neg, self.n_items,
                                   global_batch, self.optim, self.loss_out, self.loss_accum, self._p2p_ws[1], step=t,
                                   next_batch=next_batch if self.ids_ready else None)
            return self.loss_out
        if self.exchange == "dense" or (self.exchange_auto and B >= self.n_items):
            self.last_exchange = "dense"
            return self._train_step_dense(user, pos, neg, global_batch, t)
        self.last_exchange = "sparse"
        p = getattr(self, "_next_plan", None)
        if p is None or p["ids"][0] is not user:
            p = self.plan(user, pos, neg)
        self._next_plan = None
        mark = self._phase_mark
        mark("start")
        if p["event"] is not None:
            torch.cuda.current_stream().wait_event(p["event"])
            for k in ("user_local", "local_idx"):     # allocated on the plan stream, read on this one
                p[k].record_stream(torch.cuda.current_stream())
        mark("wait_plan")
        # all-to-all #1 (rows): parameters as they are after the previous step
        rows = self.V.index_select(0, p["local_idx"])
        mark("gather_rows")
        C = comm.all_to_all(rows, p["recv_counts"], p["send_counts"])
        mark("a2a_rows")
        G = torch.empty_like(C)
        ops.bpr_train_step_sharded(self.U, self.state, C, p["user_local"], p["pos_c"], p["neg_c"], global_batch,
                                   self.optim, self.loss_out, None, G, self._workspace(B), step=t,
                                   item_plan=p["plan_ws"])
        mark("kernels")
        if next_batch is not None:      # enqueue the next plan now: its host syncs overlap with the kernels above
            self._next_plan = self.plan(*next_batch, ids_ready=self.ids_ready)
        grads = return_grads(comm, G, p["send_counts"], p["recv_counts"])
        mark("a2a_grads")
        self._rows_ws = ops.sparse_rows_update(self.V, self.state.get("mV"), self.state.get("vV"), None,
                                               p["local_idx"], grads, self.optim, self._rows_ws, step=t)
        mark("rows_update")
        comm.all_reduce_sum(self.loss_out)
        self.loss_accum += self.loss_out.double()
        mark("loss")
        if self.device.type == "cuda":
            done = torch.cuda.Event()
            done.record()
            self._plan_free[p["flip"]] = done
        return self.loss_out

    # ---- optional per-phase timing of the sparse-exchange step (bench diagnostics) -----------------------------
    phase_timing = False
    ids_ready = False     # True: next_batch tensors are already complete (resident batches): the plan stream need not
                          # wait for the training stream

    def _phase_mark(self, name):
        if not self.phase_timing:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._phase_events = getattr(self, "_phase_events", [])
        self._phase_events.append((name, ev))

    def phase_report(self):
        """ms per phase averaged over the recorded steps (call after a synchronize)."""
        evs = getattr(self, "_phase_events", [])
        tot, cnt = {}, {}
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            if n1 == "start":
                continue
            tot[n1] = tot.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        self._phase_events = []
        return {k: tot[k] / cnt[k] for k in tot}

    overlap_gather = True

    def _train_step_dense(self, user, pos, neg, global_batch, t):
        ops, comm = self.ops, self.comm
        B = int(user.numel())
        # (measured on 8 GPUs: +2 % at cfg3 / 2^22 triples per GPU, but the 0.7 ms cfg2 step doubled -- the extra
        # stream and event traffic costs more than the 50 us all-gather it hides; so only for big tables)
        cuda = (self.device.type == "cuda" and not comm.staged and self.overlap_gather
                and self.n_items * self.dim * 4 > (64 << 20))
        rows_ready = None
        if cuda:
            # the all-gather of the item rows runs on its own stream; the id-only part of the step (keys, two radix
            # sorts) and the clearing of the gradient buffers overlap with it
            if not hasattr(self, "_comm_stream"):
                self._comm_stream = torch.cuda.Stream(device=self.device)
            cs = self._comm_stream
            cs.wait_stream(torch.cuda.current_stream())               # the previous step's row update
            with torch.cuda.stream(cs):
                V_all = comm.all_gather_equal(self.V)                   # [world * i_block, d]; global id == row
                rows_ready = torch.cuda.Event()
                rows_ready.record(cs)
            V_all.record_stream(torch.cuda.current_stream())
        else:
            V_all = comm.all_gather_equal(self.V)
        if not hasattr(self, "_G_all") or self._G_all.shape != V_all.shape:
            self._G_all = torch.empty_like(V_all)
            self._touched_all = torch.empty(V_all.shape[0], dtype=torch.int32, device=self.device)
        self._G_all.zero_()
        self._touched_all.zero_()
        user_local = (user - self.u_lo).contiguous()
        ops.bpr_train_step_sharded(self.U, self.state, V_all, user_local, pos, neg, global_batch, self.optim,
                                   self.loss_out, None, self._G_all, self._workspace(B), step=t,
                                   item_touched=self._touched_all, rows_ready=rows_ready)
        G = comm.reduce_scatter_sum(self._G_all)
        touched = comm.reduce_scatter_sum(self._touched_all)
        ops.dense_rows_update(self.V, self.state.get("mV"), self.state.get("vV"), G, touched, self.optim, step=t)
        comm.all_reduce_sum(self.loss_out)
        self.loss_accum += self.loss_out.double()
        return self.loss_out

    # ---- evaluation ------------------------------------------------------------------------------------
    def gather_user_table(self):
        self.flush()
        return self.comm.all_gather_equal(self.U)

    @torch.no_grad()
    def evaluate(self, index, evaluator, mode="tc", user_tile=1 << 20, layout="auto"):
        """index: ShardedEvalIndex.  Returns the metric dict over ALL evaluated users (identical on
        every rank).

        layout="replicate": the item table is all-gathered once (p only; 1 GB at 2M x 128) and every
            rank scores ITS OWN users against all items -- no user-table all-gather, no candidate
            exchange, and the scorer sees the full item range (long streams keep the tensor pipe fed).
        layout="sharded": items stay sharded; every rank scores every user against its shard and the
            per-shard top-K lists are merged at the users' owners (north_star's all-gather merge);
            needed only when the item table does not fit beside the rest.
        """
        ops, comm = self.ops, self.comm
        K = evaluator.max_k
        self.flush()
        if layout == "auto":
            layout = "replicate" if self.n_items * self.dim * 4 <= (8 << 30) else "sharded"
        if layout == "replicate":
            V_all = comm.all_gather_equal(self.V)[: self.n_items]
            n = index.n_own
            ids = torch.empty((n, K), dtype=torch.int64, device=self.device)
            local_users = (index.uid_own - self.u_lo).contiguous()
            self.last_eval_fallback_rows = 0
            self.last_eval_pass2_rows = 0
            for lo in range(0, n, user_tile):
                hi = min(lo + user_tile, n)
                ptr = index.own_hist_indptr[lo:hi + 1].contiguous()
                if not hasattr(self, "_scorer_state"):
                    self._scorer_state = ops.ScorerState()      # this model's own knobs + adaptive statistics
                i, _ = ops.fullsort_topk(self.U, local_users[lo:hi].contiguous(), V_all, K, ptr,
                                         index.own_hist_indices, mode=mode, state=self._scorer_state)
                ids[lo:hi] = i
                if mode == "tc":
                    self.last_eval_fallback_rows += self._scorer_state.last_fallback_rows
                    self.last_eval_pass2_rows += self._scorer_state.last_pass2_rows
            if n > 0:
                sums = ops.topk_metrics(ids, index.pos_indptr, index.pos_indices, self.n_items)["sums"]
            else:
                sums = torch.zeros((6, K), dtype=torch.float64, device=self.device)
            comm.all_reduce_sum(sums)
            self.last_topk = ids
            return evaluator.result(sums, int(index.uid_all.numel()))
        U_all = self.gather_user_table()
        n = int(index.uid_all.numel())
        ids = torch.empty((n, K), dtype=torch.int64, device=self.device)
        sc = torch.empty((n, K), dtype=torch.float32, device=self.device)
        n_local = self.i_hi - self.i_lo
        if n_local <= 0:                      # more ranks than item blocks: nothing to contribute
            ids.fill_(-1)
            sc.fill_(float("-inf"))
        V_local = self.V[:n_local]
        for lo in range(0, n if n_local > 0 else 0, user_tile):
            hi = min(lo + user_tile, n)
            ptr = index.hist_indptr[lo:hi + 1].contiguous()
            i, s = ops.fullsort_topk(U_all, index.uid_all[lo:hi].contiguous(), V_local, K, ptr, index.hist_indices,
                                     item_base=self.i_lo, mode=mode)
            ids[lo:hi], sc[lo:hi] = i, s
        # per-shard lists -> the users' owners
        send_counts = index.owner_counts
        recv_counts = [index.n_own] * comm.world
        r_ids = comm.all_to_all(ids, send_counts, recv_counts).view(comm.world, index.n_own, K)
        r_sc = comm.all_to_all(sc, send_counts, recv_counts).view(comm.world, index.n_own, K)
        if index.n_own > 0:
            m_ids, _ = ops.topk_merge(r_ids.contiguous(), r_sc.contiguous())
            sums = ops.topk_metrics(m_ids, index.pos_indptr, index.pos_indices, self.n_items)["sums"]
        else:
            m_ids = torch.empty((0, K), dtype=torch.int64, device=self.device)
            sums = torch.zeros((6, K), dtype=torch.float64, device=self.device)
        comm.all_reduce_sum(sums)
        self.last_topk = m_ids
        return evaluator.result(sums, n)


class ShardedEvalIndex:
    """Per-rank evaluation index: every evaluated user's history restricted to this rank's item
    shard, and the positives of the users this rank owns."""

    def __init__(self, uid_all, hist, pos_own, owner_counts, n_own, uid_own=None, hist_own=None):
        self.uid_all = uid_all
        self.hist_indptr, self.hist_indices = hist            # all users x my item shard
        self.pos_indptr, self.pos_indices = pos_own
        self.owner_counts, self.n_own = owner_counts, n_own
        self.uid_own = uid_own                                # my users (global ids)
        if hist_own is not None:
            self.own_hist_indptr, self.own_hist_indices = hist_own   # my users x all items

    @classmethod
    def from_global(cls, uid_list, hist, pos, user_bounds, item_bounds, rank, device):
        """uid_list (sorted), hist / pos numpy CSRs over uid_list rows (what every rank can derive
        from the shared dataset description)."""
        uid_list = np.asarray(uid_list)
        i_lo, i_hi = int(item_bounds[rank]), int(item_bounds[rank + 1])
        hp, hi = np.asarray(hist[0]), np.asarray(hist[1])
        keep = (hi >= i_lo) & (hi < i_hi)
        rows = np.repeat(np.arange(len(uid_list)), np.diff(hp))[keep]
        nhp = np.zeros(len(uid_list) + 1, dtype=np.int64)
        nhp[1:] = np.cumsum(np.bincount(rows, minlength=len(uid_list)))
        cuts = np.searchsorted(uid_list, user_bounds)
        owner_counts = np.diff(cuts).tolist()
        a, b = int(cuts[rank]), int(cuts[rank + 1])
        pp, pi = np.asarray(pos[0]), np.asarray(pos[1])
        own_ptr = (pp[a:b + 1] - pp[a]).astype(np.int64)
        own_idx = pi[pp[a]:pp[b]]
        oh_ptr = (hp[a:b + 1] - hp[a]).astype(np.int64)
        oh_idx = hi[hp[a]:hp[b]]
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(device)  # noqa: E731
        return cls(t(uid_list.astype(np.int64)), (t(nhp), t(hi[keep].astype(np.int64))),
                   (t(own_ptr), t(own_idx.astype(np.int64))), owner_counts, b - a,
                   uid_own=t(uid_list[a:b].astype(np.int64)), hist_own=(t(oh_ptr), t(oh_idx.astype(np.int64))))


class ShardedFM:
    """Row-sharded FM (BASELINE config 5 "across 8 B200"): the [sum(field_dims), d] table, the d = 1
    first-order table and their Adam moments are row-sharded in equal contiguous blocks; the batch is split by
    rows; the scalar bias is replicated.  One step = de-duplicate the batch's rows (rb2_item_plan) -> all-to-all
    ids out / rows back -> rb2_fm_grad_step on the fetched copies (forward, BCE loss, per-row gradient sums) ->
    all-to-all gradients to the owners -> rb2_sparse_rows_update / rb2_scalar_rows_update step exactly the touched
    rows -> all-reduce of (loss, d bias) and the bias step.  Same arithmetic as the single-GPU rb2_fm_train_step
    on the concatenated batch (mean loss over the GLOBAL batch); duplicate rows are summed in a fixed order."""

    def __init__(self, field_dims, dim, comm, device, E_full=None, W_full=None, bias=0.0, seed=2020):
        from . import ops
        self.ops, self.comm, self.device, self.dim = ops, comm, device, int(dim)
        self.field_dims = [int(x) for x in field_dims]
        self.n_fields = len(self.field_dims)
        self.rows = int(sum(self.field_dims))
        off = np.concatenate([[0], np.cumsum(self.field_dims)[:-1]]).astype(np.int64)
        self.offsets = torch.from_numpy(off).to(device)
        self.zero_offsets = torch.zeros(self.n_fields, dtype=torch.int64, device=device)
        self.bounds = shard_bounds(self.rows, comm.world)
        self.blk = self.bounds[1] - self.bounds[0]
        self.lo, self.hi = self.bounds[comm.rank], self.bounds[comm.rank + 1]
        n_loc = self.hi - self.lo
        self.E = torch.zeros((self.blk, self.dim), dtype=torch.float32, device=device)
        self.W = torch.zeros(self.blk, dtype=torch.float32, device=device)
        if E_full is not None:
            self.E[:n_loc] = torch.as_tensor(E_full[self.lo:self.hi]).to(device)
            self.W[:n_loc] = torch.as_tensor(W_full[self.lo:self.hi]).to(device)
        elif n_loc > 0:
            g = torch.Generator(device=device)
            g.manual_seed(seed + 7919 * comm.rank)
            self.E[:n_loc] = torch.randn(n_loc, self.dim, device=device, generator=g) * (2.0 / (self.rows + self.dim)) ** 0.5
            self.W[:n_loc] = torch.randn(n_loc, device=device, generator=g) * (2.0 / (self.rows + 1)) ** 0.5
        self.bias3 = torch.tensor([float(bias), 0.0, 0.0], dtype=torch.float32, device=device)
        self.state = {}
        self.loss2 = torch.zeros(2, dtype=torch.float32, device=device)
        self._bounds_dev = torch.as_tensor(self.bounds, device=device)
        self._plan_ws = self._ws = self._rows_ws = self._w_ws = None
        self.optim = None

    def build_optimizer(self, kind="adam", lr=1e-3, weight_decay=0.0):
        self.optim = self.ops.Optim(kind, lr=lr, weight_decay=weight_decay)
        if kind != "sgd":
            self.state = dict(mE=torch.zeros_like(self.E), vE=torch.zeros_like(self.E), mW=torch.zeros_like(self.W),
                              vW=torch.zeros_like(self.W))
        return self.optim

    ids_ready = False     # True: next_batch tensors are complete when passed (resident batches)

    def plan(self, ids, ids_ready=False):
        """The part of a step that depends only on the batch ids (de-duplication, bucketing by owner, count
        exchange, all-to-all of the requested row ids): run for the NEXT batch on a side stream, its host
        synchronisations overlap with the current step's kernels."""
        ops, comm = self.ops, self.comm
        cuda = self.device.type == "cuda"
        if not hasattr(self, "_plan_stream"):
            self._plan_stream = torch.cuda.Stream(device=self.device) if cuda else None
            self._plan_bufs = [None, None]
            self._plan_free = [None, None]
            self._plan_flip = 0
        st = self._plan_stream
        self._plan_flip ^= 1
        if st is not None:
            if not ids_ready:
                st.wait_stream(torch.cuda.current_stream())
            elif self._plan_free[self._plan_flip] is not None:
                st.wait_event(self._plan_free[self._plan_flip])
        with (torch.cuda.stream(st) if st is not None else _NullCtx()):
            B, F = int(ids.shape[0]), int(ids.shape[1])
            rows = (ids + self.offsets).reshape(-1)
            M = int(rows.numel())
            if M % 2:                                           # rb2_item_plan takes two id vectors of equal length
                rows = torch.cat([rows, rows[-1:]])
            half = int(rows.numel()) // 2
            ip = ops.item_plan(rows[:half], rows[half:], self.rows, self._bounds_dev, comm.world,
                               self._plan_bufs[self._plan_flip])
            self._plan_bufs[self._plan_flip] = ip["ws"]
            cuts = ip["cuts"].tolist()                          # host sync
            n_uniq = cuts[-1]
            send_counts = [cuts[g + 1] - cuts[g] for g in range(comm.world)]
            recv_counts = comm.exchange_counts(send_counts)
            req = comm.all_to_all(ip["uniq"][:n_uniq], send_counts, recv_counts)      # rows other ranks want from me
            p = dict(ids=ids, B=B, F=F, send_counts=send_counts, recv_counts=recv_counts,
                     local_idx=(req - self.lo).contiguous(),
                     compact=torch.cat([ip["pos_c"], ip["neg_c"]])[:M].view(B, F).contiguous(), flip=self._plan_flip,
                     event=None)
            if st is not None:
                p["event"] = torch.cuda.Event()
                p["event"].record(st)
        return p

    def train_step(self, ids, label, global_batch=None, next_batch=None):
        """ids int64 [B, F] raw per-field ids of THIS rank's samples, label fp32 [B].  Returns the device scalar
        holding the global mean loss.  next_batch = the ids of the following step (its plan overlaps this step)."""
        ops, comm = self.ops, self.comm
        B, F = int(ids.shape[0]), int(ids.shape[1])
        if global_batch is None:
            global_batch = B * comm.world
        self.optim.step += 1
        t = self.optim.step
        p = getattr(self, "_next_plan", None)
        if p is None or p["ids"] is not ids:
            p = self.plan(ids)
        self._next_plan = None
        if p["event"] is not None:
            torch.cuda.current_stream().wait_event(p["event"])
            for k in ("local_idx", "compact"):
                p[k].record_stream(torch.cuda.current_stream())
        send_counts, recv_counts, local_idx = p["send_counts"], p["recv_counts"], p["local_idx"]
        C_E = comm.all_to_all(self.E.index_select(0, local_idx), recv_counts, send_counts).contiguous()   # uniq order
        C_W = comm.all_to_all(self.W.index_select(0, local_idx), recv_counts, send_counts).contiguous()
        if self._ws is None or self._ws_key != (B, F):
            self._ws, self._ws_key = ops.fm_workspace(B, F, self.dim, self.device), (B, F)
        ops.fm_grad_step(C_E, C_W, self.bias3, p["compact"], self.zero_offsets, label, global_batch, self.loss2, self._ws)
        if next_batch is not None:
            self._next_plan = self.plan(next_batch, ids_ready=self.ids_ready)
        gE = comm.all_to_all(C_E, send_counts, recv_counts)
        gW = comm.all_to_all(C_W, send_counts, recv_counts)
        self._rows_ws = ops.sparse_rows_update(self.E, self.state.get("mE"), self.state.get("vE"), None, local_idx, gE,
                                               self.optim, self._rows_ws, step=t)
        self._w_ws = ops.scalar_rows_update(self.W, self.state.get("mW"), self.state.get("vW"), local_idx, gW, self.optim,
                                            self._w_ws, step=t)
        comm.all_reduce_sum(self.loss2)
        ops.scalar_step(self.bias3, self.loss2[1:], self.optim, step=t)
        if self.device.type == "cuda":
            done = torch.cuda.Event()
            done.record()
            self._plan_free[p["flip"]] = done
        return self.loss2[0]

    def gather_tables(self):
        """(E [rows, d], W [rows]) assembled on every rank (tests)."""
        E = self.comm.all_gather_equal(self.E)[: self.rows]
        W = self.comm.all_gather_equal(self.W.unsqueeze(1))[: self.rows, 0]
        return E, W
