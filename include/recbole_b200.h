/*
 * recbole_b200.h -- C ABI of the B200-native embedding hot path for ghazalehnt/RecBole.
 *
 * This is the drop-in boundary (last row).  The reference is pure
 * Python/PyTorch and has no FFI of its own; each entry point below names the reference
 * Python call it replaces (file:line relative to the reference repo) and is bound from Python
 * with ctypes in recbole_b200/_lib.py (the stub a maintainer would add is shown in
 * INTEGRATION.md).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless its name starts with `h_`.  All buffers
 *     (tables, optimizer state, id vectors, outputs, workspace) are owned by the caller
 *     (PyTorch); the library allocates nothing persistent and frees nothing.
 *   - Embedding tables are row-major fp32 [rows, dim]; ids are int64 with 0 = [PAD]
 *     (reference: recbole/data/dataset/dataset.py:908-928,1699-1700).
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *     Calls are asynchronous on that stream unless stated otherwise and re-entrant per
 *     (device, stream); the only process-wide mutable state is the last-error string, the stage
 *     profiler (rb2_profile_*), and the tuning knobs / diagnostics of the tensor-core scorer
 *     (rb2_fullsort_tc_set_*, rb2_ce_head_set_scorer, rb2_fullsort_tc_last_*_rows and the failure
 *     statistic that picks its first pass): they change speed, never results.
 *   - Return value: 0 on success, otherwise a cudaError_t or one of the RB2_E* codes;
 *     rb2_last_error() describes the failure.  There is no CPU fallback anywhere.
 *   - Workspace: query the size with the matching *_workspace_bytes() and pass a device buffer
 *     of at least that many bytes, 256-byte aligned.
 */
#ifndef RECBOLE_B200_H_
#define RECBOLE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB2_ABI_VERSION 1

#define RB2_EINVAL 10001   /* bad argument (dim not supported, null pointer, ...) */
#define RB2_EWORKSPACE 10002 /* workspace too small */
#define RB2_ERANGE 10003   /* an id is outside its table (reference: IndexError in F.embedding) */

int rb2_abi_version(void);
const char *rb2_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py).  With profiling on, every stage of the calls below is bracketed
 * by a pair of cudaEvents on the caller's stream; rb2_profile_read synchronises the device and
 * returns, per stage, the summed device time (ms), the number of times the stage ran and the
 * number of kernel launches it issued since the last read, then resets the counters.  The
 * reference has no counterpart (it wraps Trainer.fit in torch.autograd.profiler,
 * recbole/quick_start/quick_start.py:57-61).
 * ---------------------------------------------------------------------------------------- */
enum {
  RB2_ST_KEYS = 0, RB2_ST_SORT_USER = 1, RB2_ST_SORT_ITEM = 2, RB2_ST_USER_SIDE = 3, RB2_ST_USER_FIXUP = 4,
  RB2_ST_ITEM_SIDE = 5, RB2_ST_ITEM_FIXUP = 6, RB2_ST_LOSS = 7, RB2_ST_FULLSORT = 8, RB2_ST_TOPK_MERGE = 9,
  RB2_ST_METRICS = 10, RB2_ST_SAMPLER = 11, RB2_ST_GATHER_DOT = 12, RB2_ST_TC_CONVERT = 13, RB2_ST_TC_SCORE = 14,
  RB2_ST_TC_REFINE = 15, RB2_ST_FM_FWD = 16, RB2_ST_FM_UPDATE = 17, RB2_ST_MISC = 18, RB2_ST_PLAN = 19,
  RB2_ST_BARRIER = 20, RB2_ST_OWNER = 21, RB2_ST_FETCH = 22, RB2_ST_BARRIER_B = 23, RB2_NUM_STAGES = 24
};
int rb2_profile_enable(int on);
int rb2_profile_read(float *h_ms /* [RB2_NUM_STAGES] */, int64_t *h_calls /* [RB2_NUM_STAGES] */,
                     int64_t *h_launches /* [RB2_NUM_STAGES] */);

/* ------------------------------------------------------------------------------------------
 * Optimizer description.  Replaces the torch.optim objects built by
 * Trainer._build_optimizer (recbole/trainer/trainer.py:109-130) and stepped at trainer.py:173.
 * Scalars are computed by the host in double exactly as torch/optim/adam.py does
 * (bias_correction1 = 1 - beta1**t, step_size = lr / bias_correction1,
 *  bias_correction2_sqrt = sqrt(1 - beta2**t)) and rounded to fp32 once.
 * ---------------------------------------------------------------------------------------- */
enum {
  RB2_OPT_SGD = 0,        /* p -= lr * (g + wd*p)                  (torch.optim.SGD, momentum 0)     */
  RB2_OPT_ADAM = 1,       /* row-sparse Adam: the dense formula on the rows touched by the batch    */
  RB2_OPT_ADAM_LAZY = 2   /* row-sparse Adam that replays the zero-gradient steps a row missed, so
                             the trajectory equals the reference's dense Adam (needs *_last arrays).
                             A batch's rows are caught up ONCE per distinct row before the step reads them;
                             the sharded entry points catch up the (local) user rows and let every owner
                             take the zero-gradient step of the untouched rows of its item shard
                             (rb2_dense_rows_update, rb2_bpr_train_step_p2p); refused with an item plan
                             (sparse all-to-all exchange) */
};

typedef struct rb2_optim {
  int32_t kind;
  int32_t step;            /* t, 1-based, one counter per training run (adam.py: state['step'])      */
  float lr;
  float weight_decay;      /* L2 added to the gradient of touched rows (adam.py:416-429)             */
  float beta1, beta2;
  float one_minus_beta1;   /* (float)(1 - beta1) as torch passes to lerp_                            */
  float one_minus_beta2;
  float eps;
  float step_size;         /* lr / (1 - beta1**t)                                                    */
  float bc2_sqrt;          /* sqrt(1 - beta2**t)                                                     */
  /* RB2_OPT_ADAM_LAZY only: DEVICE tables indexed by step j = 1..step (entry 0 unused) holding
     lr/(1-beta1**j) and sqrt(1-beta2**j); NULL otherwise. */
  const float *lazy_step_size;
  const float *lazy_bc2_sqrt;
} rb2_optim;

/* ------------------------------------------------------------------------------------------
 * (1) Fused BPR training step.
 * Replaces, for one batch: BPR.calculate_loss (recbole/model/general_recommender/bpr.py:74-83),
 * BPRLoss.forward (recbole/model/loss.py:43-49), loss.backward() (trainer.py:170) and
 * optimizer.step() (trainer.py:173).  All gradients are evaluated at the parameters as they
 * are on entry, duplicates of a row inside the batch are summed, then each touched row takes
 * exactly one optimizer step.
 *
 * user_m/user_v/item_m/item_v may be NULL for RB2_OPT_SGD.  user_last/item_last (int32 per row,
 * zero-initialised) are only read for RB2_OPT_ADAM_LAZY.
 * loss_out[0]   = mean loss of this batch (fp32), as calculate_loss returns it.
 * loss_accum[0] += loss_out[0] if loss_accum != NULL (fp64; the Trainer's running total_loss,
 *                  trainer.py:168, kept on the device so the epoch needs one sync).
 * ---------------------------------------------------------------------------------------- */
size_t rb2_bpr_workspace_bytes(int64_t batch, int32_t dim);

int rb2_bpr_train_step(float *user_p, float *user_m, float *user_v, int32_t *user_last,
                       float *item_p, float *item_m, float *item_v, int32_t *item_last,
                       int64_t n_users, int64_t n_items, int32_t dim,
                       const int64_t *user, const int64_t *pos, const int64_t *neg, int64_t batch,
                       const rb2_optim *h_opt, float *loss_out, double *loss_accum,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * (1c) The same step when the item table is row-sharded over several GPUs (the
 * reference has no multi-device code).  The caller fetched the item rows its local batch touches
 * from their owners into the compact table `item_rows` [n_item_rows, dim] and rewrote pos / neg as
 * indices into it.  The user side runs exactly as above (users are partitioned with their
 * interactions, so it is local); the item side only SUMS the gradient per compact row into
 * item_grad_out [n_item_rows, dim] (every compact row must occur in the batch), to be sent back to
 * the owners, who apply it with rb2_sparse_rows_update.  The loss and its gradient are scaled by
 * 1/global_batch (mean over the union of all ranks' batches); all-reduce loss_out with SUM.
 * ---------------------------------------------------------------------------------------- */
int rb2_bpr_train_step_sharded(float *user_p, float *user_m, float *user_v, int32_t *user_last,
                               const float *item_rows, int64_t n_users, int64_t n_item_rows, int32_t dim,
                               const int64_t *user, const int64_t *pos, const int64_t *neg, int64_t batch,
                               int64_t global_batch, const rb2_optim *h_opt, float *loss_out, double *loss_accum,
                               float *item_grad_out, int32_t *item_touched /* nullable: [n_item_rows], set to 1 for
                               every row written */, const void *item_plan /* nullable: the plan workspace
                               rb2_item_plan filled for this batch; its sorted occurrences are reused */,
                               void *workspace, size_t workspace_bytes, void *stream);
/* Same; `rows_ready_event` (a cudaEvent_t, nullable) is waited for on `stream` after the id-only part of the
 * step (keys, sorts) and before the first kernel that reads item_rows, so that the collective delivering
 * item_rows on another stream overlaps with that part. */
int rb2_bpr_train_step_sharded_ev(float *user_p, float *user_m, float *user_v, int32_t *user_last,
                                  const float *item_rows, int64_t n_users, int64_t n_item_rows, int32_t dim,
                                  const int64_t *user, const int64_t *pos, const int64_t *neg, int64_t batch,
                                  int64_t global_batch, const rb2_optim *h_opt, float *loss_out, double *loss_accum,
                                  float *item_grad_out, int32_t *item_touched, const void *item_plan, void *workspace,
                                  size_t workspace_bytes, void *stream, void *rows_ready_event);

/* The id-only half of a sharded step (what the caller needs before any parameter is read): unique
 * item ids of the batch in ascending order (= grouped by owner shard), pos / neg rewritten as indices
 * into that list, and cuts[g] = number of unique ids below shard_bounds[g] (g = 0..world),
 * cuts[world+1] = number of unique ids.  uniq must hold 2*batch entries.  All outputs on the device;
 * the caller copies `cuts` (world+2 values) to the host to size the all-to-all. */
size_t rb2_item_plan_workspace_bytes(int64_t batch);
int rb2_item_plan(const int64_t *pos, const int64_t *neg, int64_t batch, int64_t n_items,
                  const int64_t *shard_bounds, int32_t world, int64_t *uniq, int64_t *pos_c, int64_t *neg_c,
                  int64_t *cuts, void *plan_workspace, size_t plan_workspace_bytes, void *stream);

/* Owner side of the replicated-small-table exchange (all-gather rows, reduce-scatter gradients):
 * rows with touched[row] > 0 take one optimizer step with grads[row, :].  RB2_OPT_ADAM_LAZY: the other rows
 * take the zero-gradient step of the reference's dense Adam (rows with zero moments and no weight decay do
 * not move and are not written), so the shard is always current and needs no `last` array. */
int rb2_dense_rows_update(float *p, float *m, float *v, int64_t n_rows, int32_t dim, const float *grads,
                          const int32_t *touched, const rb2_optim *h_opt, void *stream);

/* Row-sparse optimizer step from explicit gradient rows: grads[j, :] belongs to row ids[j]; duplicate
 * ids are summed (fixed order), then every touched row takes one step.  Owner side of the sharded
 * training step; also the generic replacement for "dense grad + dense optimizer.step()"
 * (trainer.py:170-173) for any embedding table. */
size_t rb2_sparse_rows_update_workspace_bytes(int64_t count, int32_t dim);
int rb2_sparse_rows_update(float *p, float *m, float *v, int32_t *last, int64_t n_rows, int32_t dim,
                           const int64_t *ids, const float *grads, int64_t count, const rb2_optim *h_opt,
                           void *workspace, size_t workspace_bytes, void *stream);

/* Forward only: loss_out[0] = BPR loss of the batch, parameters untouched (calculate_loss under
 * torch.no_grad, bpr.py:74-83 + loss.py:48). */
int rb2_bpr_loss(const float *user_p, const float *item_p, int64_t n_users, int64_t n_items, int32_t dim,
                 const int64_t *user, const int64_t *pos, const int64_t *neg, int64_t batch,
                 float *loss_out, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * (1e) The same step over PEER MEMORY: the item table is row-sharded in equal contiguous blocks over
 * the GPUs of one NVLink / NVSwitch domain, one process per GPU, and the kernels read the owners'
 * rows and write the owners' gradient slots themselves (BASELINE north_star "row-sharded
 * across the 8 GPUs ... exchanges rows and gradients ... over NVLink").  No NCCL call and no host
 * synchronisation inside a step:
 *   keys, two sorts (local) -> flag barrier A (wait half) -> per occurrence source / destination (an item that
 *   occurs ONCE in this rank's batch is read from its owner directly and its gradient is pushed
 *   straight into the owner's slot by the user-side kernel; an item that occurs several times is
 *   fetched once into item_cache and its summed gradient is pushed by the item-side walk) ->
 *   flag barrier B (the ranks' loss sums travel with it) -> every owner sums, in rank order, the slots
 *   stamped with this step and takes ONE optimizer step per touched local row.
 * Result == rb2_bpr_train_step on the union of the ranks' batches (up to fp32 summation order).
 *
 * rb2_peers: pointers valid IN THIS PROCESS for every rank's buffers (own: plain device pointers;
 * peers: mapped with rb2_ipc_open or any other peer mapping).  Per rank r:
 *   item_p[r]      fp32 [item_block, dim]             its shard of the item table
 *   grad_slots[r]  fp32 [world, item_block, dim]      slot s = the gradient rows rank s pushed
 *   stamps[r]      int32 [world, item_block]          zero-initialised; == step where the slot row is valid
 *   flags[r]       uint32 [2, RB2_MAX_PEERS]          zero-initialised barrier flags
 *   loss_slots[r]  fp64 [2, RB2_MAX_PEERS]
 * Every rank must call rb2_bpr_train_step_p2p the same number of times with h_peers->seq = 1, 2, 3, ...
 * (the barrier sequence number; independent of h_opt->step, which a resumed run restores from its checkpoint).
 * The "owner update done" half of barrier A of call seq + 1 is signalled at the END of call seq, so a rank's keys
 * and sorts never hold its peers up.  user holds global user ids, all inside THIS rank's block
 * [user_base, user_base + n_users_local) of the user table; pos / neg are global item ids.  item_m / item_v: Adam moments of the local shard.
 * next_user / next_pos / next_neg (all NULL, or the ids of the batch the NEXT call will train on, same batch size): their
 * keys and sorts are computed into the workspace's other batch slot while this call waits for its peers in barrier B;
 * the next call then passes prepared = 1 (same workspace, same ids) and skips that work.  That work is issued on a
 * side stream owned by the library (one per device and caller stream) right after this rank's barrier-B signal, so it
 * also overlaps the owner update and barrier A; the next call on the same caller stream waits for it.  Keep the
 * workspace and the next_* id arrays alive (and unchanged) until that next call, or synchronise the device first.
 * item_cache: fp32 [world * item_block, dim] scratch.  A barrier that waits longer than 30 s gives up
 * and sets the workspace's peer_timeout flag (second int32 of the workspace) instead of hanging.
 * RB2_OPT_ADAM_LAZY (the trajectory of the reference's DENSE torch.optim.Adam, trainer.py:116,173): user_last =
 * int32 [n_users_local], zero-initialised (NULL for the other kinds) -- the batch's user rows replay the zero-gradient
 * steps they missed once per row before the step; every owner additionally takes the zero-gradient step of the rows
 * of its item shard that nobody touched (n_items / world rows: cheap), so item rows are always current.
 * ---------------------------------------------------------------------------------------- */
#define RB2_MAX_PEERS 8
typedef struct rb2_peers {
  int32_t world, me;
  int64_t item_block;
  const float *item_p[RB2_MAX_PEERS];
  float *grad_slots[RB2_MAX_PEERS];
  int32_t *stamps[RB2_MAX_PEERS];
  uint32_t *flags[RB2_MAX_PEERS];
  double *loss_slots[RB2_MAX_PEERS];
  int64_t seq;   /* call number on THESE buffers: 1 for the first step after they were zero-initialised, then
                    2, 3, ... (identical on every rank); it is the barrier sequence and the slot stamp */
} rb2_peers;

size_t rb2_bpr_p2p_workspace_bytes(int64_t batch, int32_t dim);
int rb2_bpr_train_step_p2p(float *user_p, float *user_m, float *user_v, float *item_m, float *item_v,
                           int64_t n_users_local, int64_t n_items, int32_t dim, const int64_t *user,
                           int64_t user_base, const int64_t *pos, const int64_t *neg, int64_t batch,
                           int64_t global_batch,
                           const rb2_optim *h_opt, const rb2_peers *h_peers, float *item_cache,
                           float *loss_out, double *loss_accum, void *workspace, size_t workspace_bytes,
                           void *stream, int32_t prepared, const int64_t *next_user, const int64_t *next_pos,
                           const int64_t *next_neg, int32_t *user_last);

/* Peer mapping helpers (cudaIpc*; legacy IPC handles work between processes on one GPU and across
 * NVLink peers).  rb2_ipc_export: 64-byte handle of the allocation that contains dev_ptr + the offset
 * of dev_ptr inside it (PyTorch's caching allocator sub-allocates).  rb2_ipc_open maps a handle
 * exported by ANOTHER process and returns base + offset; one mapping per (process, allocation) is
 * kept and shared by later opens.  rb2_ipc_close_all unmaps everything this process opened. */
#define RB2_IPC_HANDLE_BYTES 64
int rb2_ipc_export(const void *dev_ptr, void *h_handle /* [64] */, int64_t *h_offset);
int rb2_ipc_open(const void *h_handle /* [64] */, int64_t offset, void **h_mapped);
int rb2_ipc_close_all(void);

/* ------------------------------------------------------------------------------------------
 * The grouping primitive of the training steps: stable sort of (key, position) pairs for keys < 2^key_bits
 * (row ids of a table).  keys_sorted ascending, positions[i] = index in `keys` of the i-th smallest key, equal keys
 * in ascending position.  One persistent cooperative launch: LSD radix sort with 12-bit digits, warp-private
 * shared-memory histograms, hand-written grid barrier (csrc/bucket_sort.cuh); no library code.
 * ---------------------------------------------------------------------------------------- */
size_t rb2_sort_positions_workspace_bytes(int64_t count);
int rb2_sort_positions(const uint32_t *keys, int64_t count, int32_t key_bits, uint32_t *keys_sorted,
                       uint32_t *positions, void *workspace, size_t workspace_bytes, void *stream);

/* Flush for RB2_OPT_ADAM_LAZY: bring every row to step `h_opt->step` (replaying the zero-gradient
 * steps it missed) so the tables can be read by evaluation / checkpointing. */
int rb2_adam_lazy_flush(float *p, float *m, float *v, int32_t *last, int64_t rows, int32_t dim,
                        const rb2_optim *h_opt, void *stream);

/* ------------------------------------------------------------------------------------------
 * (1d) Fused FM (factorization machine, TOKEN fields) training step: multi-field embedding bag.
 * Replaces, per batch: ContextRecommender.embed_input_fields + FMEmbedding (abstract_recommender.py:
 * 220-224,361-412; layers.py:141-144), BaseFactorizationMachine (layers.py:164-171),
 * FMFirstOrderLinear (layers.py:1021-1061), FM.forward/calculate_loss with nn.BCELoss
 * (recbole/model/context_aware_recommender/fm.py:47-56), loss.backward() and optimizer.step().
 *   E [n_rows, dim]  = token_embedding_table.embedding.weight       (all token fields share it)
 *   W [n_rows]       = first_order_linear.token_embedding_table.embedding.weight  (output_dim 1)
 *   bias3 [3]        = first_order_linear.bias and its Adam moments (b, m, v)
 *   ids [batch, n_fields] raw per-field ids; row = ids[s, f] + offsets[f]   (layers.py:142)
 * TOKEN fields through E / W, FLOAT fields through rb2_fm_float, TOKEN_SEQ fields (mean pooling) through rb2_fm_seq.
 * dim in {16, 32, 64, 128}; optimizer RB2_OPT_SGD, RB2_OPT_ADAM (row-sparse) or RB2_OPT_ADAM_LAZY; the bias is dense.
 * rb2_fm_predict: y[s] = sigmoid(first_order + fm)  (FM.predict, fm.py:58-59).
 * ---------------------------------------------------------------------------------------- */
/* FLOAT fields (ContextRecommender.embed_float_fields abstract_recommender.py:236-258, FMFirstOrderLinear
 * layers.py:947-966): field f owns one row Ef[f] of float_embedding_table.weight [n_float, dim] and one scalar Wf[f]
 * of first_order_linear.float_embedding_table.weight; sample s contributes values[s, f] * Ef[f] as that field's
 * vector.  Every sample touches every float row, so their gradients are dense reductions over the batch and their
 * optimizer step is the reference's dense one at every step.  NULL = no float fields. */
#define RB2_FM_MAX_FLOAT 64
typedef struct rb2_fm_float {
  const float *values;        /* [batch, n_float] */
  int32_t n_float;
  float *Ef, *mEf, *vEf;      /* [n_float, dim] (m, v: Adam moments, NULL for SGD / predict / loss) */
  float *Wf, *mWf, *vWf;      /* [n_float] */
} rb2_fm_float;

/* TOKEN_SEQ fields (ContextRecommender.embed_token_seq_fields abstract_recommender.py:277-314, mode 'mean'; first order
 * layers.py:989-1019): the id matrix gets extra columns -- after the n_token_cols TOKEN columns, the padded sequence
 * of every TOKEN_SEQ field (field j: columns [seq_start[j], seq_start[j + 1]), id 0 = padding = masked) -- and
 * offsets[col] places a sequence column's ids in that field's own table, stored as rows >= seq_row_base of E / W.
 * Field j contributes e_j = sum over its unmasked ids of the row / (count + 1e-8).  n_fields counts ALL columns.
 * pooled [batch, n_seq, dim] and coef [batch, n_seq] keep e_j and 1 / (count + 1e-8) for the backward (training only).
 * NULL = no such fields. */
#define RB2_FM_MAX_SEQ 16
typedef struct rb2_fm_seq {
  int32_t n_seq, n_token_cols;
  const int32_t *seq_start;   /* device [n_seq + 1] */
  const int32_t *col_seq;     /* device [n_fields]: the TOKEN_SEQ field of a column, -1 for TOKEN columns */
  int64_t seq_row_base;
  float *pooled, *coef;
} rb2_fm_seq;

size_t rb2_fm_workspace_bytes(int64_t batch, int32_t n_fields, int32_t dim);
int rb2_fm_train_step(float *E, float *mE, float *vE, float *W, float *mW, float *vW, float *bias3,
                      int32_t *row_last, int64_t n_rows, int32_t dim, const int64_t *ids, const int64_t *offsets,
                      int32_t n_fields, const float *label, int64_t batch, const rb2_optim *h_opt, float *loss_out,
                      double *loss_accum, void *workspace, size_t workspace_bytes, void *stream,
                      const rb2_fm_float *h_float, const rb2_fm_seq *h_seq);
/* RB2_OPT_ADAM_LAZY (row_last: int32 [n_rows], zero-initialised, shared by E and W): the trajectory of the
 * reference's DENSE torch.optim.Adam, weight decay included (MFSimple.yaml:2 sets 1e-8: every row moves at every
 * step).  rb2_fm_lazy_flush brings all rows to step h_opt->step before predict / loss / a checkpoint read them. */
int rb2_fm_lazy_flush(float *E, float *mE, float *vE, float *W, float *mW, float *vW, int32_t *row_last,
                      int64_t n_rows, int32_t dim, const rb2_optim *h_opt, void *stream);
int rb2_fm_predict(const float *E, const float *W, const float *bias3, int64_t n_rows, int32_t dim,
                   const int64_t *ids, const int64_t *offsets, int32_t n_fields, int64_t batch, float *y_out,
                   void *workspace, size_t workspace_bytes, void *stream, const rb2_fm_float *h_float,
                   const rb2_fm_seq *h_seq);
/* forward + mean nn.BCELoss only (FM.calculate_loss fm.py:52-56 / MFSimple.calculate_loss mfsimple.py:48-57 as a
 * VALUE: what an unmodified Trainer reads with loss.item(), trainer.py:168); nothing is kept for a backward. */
int rb2_fm_loss(const float *E, const float *W, const float *bias3, int64_t n_rows, int32_t dim,
                const int64_t *ids, const int64_t *offsets, int32_t n_fields, const float *label, int64_t batch,
                float *loss_out, void *workspace, size_t workspace_bytes, void *stream, const rb2_fm_float *h_float,
                const rb2_fm_seq *h_seq);

/* Row-sharded FM (one 33M-row table sharded over the GPUs, batch split by rows; no reference
 * counterpart).  rb2_fm_grad_step is the local part of a step: rows_e [n_rows, dim] / rows_w [n_rows] are the
 * FETCHED copies of the rows this rank's samples touch (ids index them; same workspace as rb2_fm_train_step); on
 * return every row holds its summed gradient for loss = sum over the GLOBAL batch / global_batch, loss2[0] = this
 * rank's share of the loss, loss2[1] = its share of d loss / d bias.  The owners then step their rows:
 * rb2_sparse_rows_update for the [rows, dim] table, rb2_scalar_rows_update for the d = 1 table (duplicate ids
 * summed in a fixed order), rb2_scalar_step for the replicated (bias, m, v) block after an all-reduce. */
int rb2_fm_grad_step(float *rows_e, float *rows_w, const float *bias3, int64_t n_rows, int32_t dim,
                     const int64_t *ids, const int64_t *offsets, int32_t n_fields, const float *label, int64_t batch,
                     int64_t global_batch, float *loss2, void *workspace, size_t workspace_bytes, void *stream);
size_t rb2_scalar_rows_update_workspace_bytes(int64_t m);
int rb2_scalar_rows_update(float *p, float *m, float *v, int64_t n_rows, const int64_t *ids, const float *grads,
                           int64_t M, const rb2_optim *h_opt, void *workspace, size_t workspace_bytes, void *stream);
int rb2_scalar_step(float *p3, const float *grad, const rb2_optim *h_opt, void *stream);

/* ------------------------------------------------------------------------------------------
 * (1b) Gather-dot for explicit (user, item) pairs.  Replaces BPR.predict (bpr.py:85-89).
 * ---------------------------------------------------------------------------------------- */
int rb2_gather_dot(const float *user_p, const float *item_p, int64_t n_users, int64_t n_items, int32_t dim,
                   const int64_t *user, const int64_t *item, int64_t n, float *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * (2) Full-sort scorer with fused mask and top-K; the [users, n_items] score matrix is never
 * written.  Replaces BPR.full_sort_predict (bpr.py:91-96) / SASRec.full_sort_predict
 * (recbole/model/sequential_recommender/sasrec.py:152-158), the masking of
 * Trainer._full_sort_batch_eval (trainer.py:342-345) and TopKEvaluator.collect's topk
 * (recbole/evaluator/evaluators.py:68-72).
 *
 * Row r of the query is query_p[query_ids[r]] (query_ids == NULL: row r itself, e.g. sequence
 * embeddings).  Candidates for row r: item ids item_base+1.. (id 0 = [PAD] is never a candidate)
 * minus hist_indices[hist_indptr[r] .. hist_indptr[r+1]) (sorted ascending, global item ids;
 * hist_indptr == NULL: no history).  Scores are the canonical fp32 chain
 * s = fmaf(q[k], v[k], s), k ascending.  Order: score descending, ties by ascending item id.
 * out_ids int64 [nq, k] (missing slots -1), out_scores fp32 [nq, k] (missing -inf).
 * item_p points at the local shard [n_items_local, dim] whose first row has global id item_base.
 *
 * mode: RB2_SCORER_FP32   exact CUDA-core path.
 *       RB2_SCORER_TC     tcgen05 bf16 tensor-core filter + exact fp32 re-score of the survivors,
 *                         certified per row; rows whose certificate fails are redone in fp32.
 * ---------------------------------------------------------------------------------------- */
enum { RB2_SCORER_FP32 = 0, RB2_SCORER_TC = 1 };

/* Scorer state, owned by the CALLER (host memory; zero-initialise): the knobs, the adaptive statistics that choose the
 * tensor-core scorer's first pass, and what the last call did.  The *_s entry points below read and update the state
 * they are given and nothing else, so a process can run one state per (device, stream, model) from as many threads
 * as it likes.  The entry points without a state argument use a default state private to the CALLING THREAD; the
 * rb2_fullsort_tc_set_* / rb2_ce_head_set_scorer / rb2_fullsort_tc_last_* calls address that thread-default state.
 * (The tensor-core scorer synchronises the stream once or twice per call to read how many rows failed their
 * certificate: it is not capturable in a CUDA graph.) */
typedef struct rb2_scorer_state {
  int32_t variant;             /* 0 = default; 1 = bf16 / fp32 accumulators; 2 = CTA-pair MMAs; 3 = fp16 / FP16 accumulators */
  int32_t kprime;              /* candidates per list: 0 = automatic, 16, 32 */
  int32_t ce_scorer;           /* rb2_ce_head: 0 = tensor cores where covered, 1 = CUDA-core kernel */
  float fail_ema;              /* recent fraction of rows failing the first certificate */
  int32_t calls;
  int32_t last_fallback_rows;  /* rows the last call redid with the exact CUDA-core kernel */
  int32_t last_pass2_rows;     /* rows the last call sent through the second (fp32-accumulator) tensor pass */
  int32_t reserved;
  void *trace;                 /* diagnostics: device buffer (tools/tc_trace.py) or NULL */
} rb2_scorer_state;

size_t rb2_fullsort_workspace_bytes(int64_t nq, int64_t n_items_local, int32_t dim, int32_t k, int32_t mode);

/* Compatibility path: the score matrix itself, out_scores fp32 [nq, n_items] = BPR.full_sort_predict (bpr.py:91-96)
 * for an UNMODIFIED reference Trainer._full_sort_batch_eval (trainer.py:328-352), which masks and top-k's it with
 * ATen, 1-2 users per call (general_dataloader.py:330-334).  Every score is the canonical fp32 chain of the top-K
 * kernels.  query_ids may be NULL (row r of query_p); ids outside [0, n_query_rows) are clamped. */
int rb2_fullsort_scores(const float *query_p, const int64_t *query_ids, int64_t nq, int64_t n_query_rows,
                        const float *item_p, int64_t n_items, int32_t dim, float *out_scores, void *stream);

int rb2_fullsort_topk(const float *query_p, const int64_t *query_ids, int64_t nq,
                      const float *item_p, int64_t n_items_local, int64_t item_base, int32_t dim,
                      const int64_t *hist_indptr, const int64_t *hist_indices, int32_t k, int32_t mode,
                      int64_t *out_ids, float *out_scores,
                      void *workspace, size_t workspace_bytes, void *stream);
int rb2_fullsort_topk_s(const float *query_p, const int64_t *query_ids, int64_t nq,
                      const float *item_p, int64_t n_items_local, int64_t item_base, int32_t dim,
                      const int64_t *hist_indptr, const int64_t *hist_indices, int32_t k, int32_t mode,
                      int64_t *out_ids, float *out_scores,
                      void *workspace, size_t workspace_bytes, void *stream,
                        rb2_scorer_state *h_state);

/* ------------------------------------------------------------------------------------------
 * (2c) Full-sort cross-entropy head (SASRec-style, loss_type='CE').  Replaces
 *   logits = seq_output @ item_emb.weight.T ; nn.CrossEntropyLoss()(logits, pos)
 * (recbole/model/sequential_recommender/sasrec.py:137-141) and, in the same pass, full_sort_predict +
 * pad mask + top-k (sasrec.py:152-158, trainer.py:343, evaluators.py:68-72): the [nq, n_items] logits
 * (16.4 GB at 4096 x 1M) are never written.  The logsumexp runs online over EVERY item incl. the
 * padding row 0 (it is a class of the CE), the top-k excludes id 0.  Scores are the canonical fp32
 * chain.  loss_out[0] = mean_r(lse[r] - logit[r, target[r]]) (target == NULL: no loss);
 * lse_out [nq] optional.
 * ---------------------------------------------------------------------------------------- */
size_t rb2_ce_head_workspace_bytes(int64_t nq, int64_t n_items, int32_t dim, int32_t k);
int rb2_ce_head(const float *x, int64_t nq, const float *item_p, int64_t n_items, int32_t dim,
                const int64_t *target, int32_t k, float *loss_out, float *lse_out, int64_t *topk_ids,
                float *topk_scores, void *workspace, size_t workspace_bytes, void *stream);
int rb2_ce_head_s(const float *x, int64_t nq, const float *item_p, int64_t n_items, int32_t dim,
                const int64_t *target, int32_t k, float *loss_out, float *lse_out, int64_t *topk_ids,
                float *topk_scores, void *workspace, size_t workspace_bytes, void *stream,
                  rb2_scorer_state *h_state);

/* rb2_ce_head runs on the tensor cores where covered (dim == 64, k <= 16): bf16 hi/lo split operands, one
 * K = 192 GEMM with fp32 accumulators (logits good to ~2^-16 ||x|| ||e||), online logsumexp in the epilogue,
 * certified exact top-k as RB2_SCORER_TC.  mode 1 forces the CUDA-core fp32 kernel, 0 = automatic. */
int rb2_ce_head_set_scorer(int32_t mode);

/* Backward of the same head (autograd of sasrec.py:137-141: G = (softmax(logits) - onehot(target)) * grad_scale,
 * dx = G @ E, de = G^T @ x) on the tensor cores, hidden size 64.  Neither the logits nor G ([nq, n_items] fp32 each in
 * the reference) are materialised: the logits are recomputed tile by tile from `lse` (the row logsumexp rb2_ce_head
 * returned), G tiles live in shared memory as split-bf16 MMA operands.  grad_scale = upstream gradient / nq for the
 * mean loss.  dx_out [nq, dim] and/or de_out [n_items, dim] (either may be NULL); both are complete gradients --
 * every class has one, so the item table's optimizer step is dense: rb2_dense_step (torch.optim.Adam / SGD on the
 * whole tensor, as trainer.py:173 does).  Deterministic (no float atomics). */
size_t rb2_ce_head_backward_workspace_bytes(int64_t nq, int64_t n_items, int32_t dim);
int rb2_ce_head_backward(const float *x, int64_t nq, const float *item_p, int64_t n_items, int32_t dim,
                         const int64_t *target, const float *lse, float grad_scale, float *dx_out, float *de_out,
                         void *workspace, size_t workspace_bytes, void *stream);
int rb2_dense_step(float *p, float *m, float *v, const float *grad, int64_t count, const rb2_optim *h_opt,
                   void *stream);

/* Diagnostic: how many rows of the last RB2_SCORER_TC call failed the certificate and were redone by
 * the fp32 kernel (or nq if the shape is not covered by the MMA tiling: dim not in {64,128}, k > 16). */
int32_t rb2_fullsort_tc_last_fallback_rows(void);
/* Diagnostic: rows of the last RB2_SCORER_TC call whose first (FP16-accumulator) certificate failed and that
 * were re-scored by the second tensor-core pass (same fp16 operands, fp32 accumulators, ~10x tighter error
 * bound); only what fails there too reaches the fp32 kernel (rb2_fullsort_tc_last_fallback_rows). */
int32_t rb2_fullsort_tc_last_pass2_rows(void);
/* Tuning knob of RB2_SCORER_TC: candidates kept per (row, list): 16, 32, or 0 = automatic (32, or 16
 * when k <= 8).  More candidates = looser certificate, more epilogue work.  The result is exact either
 * way. */
int rb2_fullsort_tc_set_kprime(int32_t kprime);
/* MMA variant of RB2_SCORER_TC: 0 = default (= 3); 1 = bf16 operands, fp32 accumulators, per-CTA M = 128
 * MMAs (cta_group::1) on item slots that a CTA pair loads by halves and multicasts; 3 = fp16 operands (rows
 * rescaled by exact powers of two) with FP16 accumulators read back two per register (tcgen05.ld
 * .pack::16b), per-CTA MMAs; 2 = as 3 with CTA-pair MMAs (cta_group::2, M = 256 across two SMs, each SM holds
 * half of every item slot).  The accumulate error joins the certificate: the result is exact in every
 * variant. */
int rb2_fullsort_tc_set_variant(int32_t variant);
/* Diagnostic: `device_buffer` (16 int64 per CTA, 148 CTAs at most; NULL = off) receives the cycles the
 * producer / MMA / epilogue roles of the RB2_SCORER_TC kernel spent waiting on each of their barriers. */
int rb2_fullsort_tc_set_trace(void *device_buffer);

/* Merge `parts` per-shard top-K lists ([parts, nq, k], each sorted) into the global top-K
 * (multi-GPU all-gather merge). */
int rb2_topk_merge(const int64_t *ids, const float *scores, int32_t parts, int64_t nq, int32_t k,
                   int64_t *out_ids, float *out_scores, void *stream);

/* ------------------------------------------------------------------------------------------
 * (2b) Metrics on device.  Replaces TopKEvaluator.evaluate/_calculate_metrics
 * (evaluators.py:78-105,122-141) and recbole/evaluator/metrics.py:27-164.
 * pos CSR: the evaluated phase's positives per row (sorted).  discount[j] = 1/log2(j+2) and
 * idcg[j] = sum_{i<=j} discount[i] (fp64, length k) are supplied by the host so that the
 * per-user values are bit-identical to numpy's.
 * sums[6*k] (fp64): metric-major sums over rows in the order
 *   RB2_M_RECALL, RB2_M_MRR, RB2_M_NDCG, RB2_M_HIT, RB2_M_PRECISION, RB2_M_MAP; value at rank j+1.
 * Optional outputs (NULL to skip): hit uint8 [nq, k]; ref_idx int64 [nq, k+1] = the reference's
 * own `topk_idx | shape` matrix in its swapped+flipped coordinates (evaluators.py:68-75), so
 * that an unmodified TopKEvaluator.evaluate can consume it.
 * ---------------------------------------------------------------------------------------- */
enum { RB2_M_RECALL = 0, RB2_M_MRR = 1, RB2_M_NDCG = 2, RB2_M_HIT = 3, RB2_M_PRECISION = 4, RB2_M_MAP = 5,
       RB2_NUM_METRICS = 6 };

size_t rb2_topk_metrics_workspace_bytes(int64_t nq, int32_t k);

int rb2_topk_metrics(const int64_t *topk_ids, int64_t nq, int32_t k, int64_t n_items,
                     const int64_t *pos_indptr, const int64_t *pos_indices,
                     const double *discount, const double *idcg,
                     double *sums, uint8_t *hit, int64_t *ref_idx,
                     void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * (3) Negative sampler against a device-resident CSR of used ids.
 * Replaces Sampler.sample_by_user_ids -> AbstractSampler.sample_by_key_ids
 * (recbole/sampler/sampler.py:103-154,246-265).  Output layout as the reference:
 * out[k*n_keys + i] is the k-th negative of key_ids[i].
 *
 * rb2_neg_sample_ref: the reference's own stream -- values are taken from `random_list`
 *   (the once-shuffled candidate list, sampler.py:54-57) starting at *h_random_pr, re-drawing
 *   rejected slots in order from the following entries (sampler.py:144-153); *h_random_pr is
 *   advanced exactly as the reference advances it.  Synchronous (one host sync per round).
 * rb2_neg_sample_hash: counter-based stream (seed, step, slot, attempt); one launch, async.
 * ---------------------------------------------------------------------------------------- */
size_t rb2_neg_sample_workspace_bytes(int64_t n_keys, int32_t num);

int rb2_neg_sample_ref(const int64_t *key_ids, int64_t n_keys, int32_t num,
                       const int64_t *random_list, int64_t random_list_length, int64_t *h_random_pr,
                       const int64_t *used_indptr, const int64_t *used_indices, int64_t n_rows,
                       int64_t *out, void *workspace, size_t workspace_bytes, void *stream);

int rb2_neg_sample_hash(const int64_t *key_ids, int64_t n_keys, int32_t num, int64_t n_items,
                        const int64_t *used_indptr, const int64_t *used_indices, int64_t n_rows,
                        uint64_t seed, uint64_t step, int64_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RECBOLE_B200_H_ */
