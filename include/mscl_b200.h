/*
 * mscl_b200.h -- C ABI of the B200-native MSCL contrastive hot path.
 *
 * One shared library (libmscl_b200.so, sm_100a only) exports the entry points below.
 * Plain pointers and sizes; no torch types.  Every pointer named d_* is a DEVICE
 * pointer owned by the caller; nothing is allocated or freed inside.  Every call
 * enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 * immediately: 0 on success, a negative MSCL_E* code otherwise, with the text
 * available from mscl_last_error() (thread-local).
 *
 * The reference (megvii-research/MSCL, an MMAction2 fork) has no native code on
 * this path: each entry point replaces a sequence of PyTorch/NumPy calls, cited
 * as `file:line` relative to the reference root.
 *
 * Device data layout:
 *   queue   float32 [K_local, C]   one key per ROW ("key-major"); the reference
 *                                  keeps the transpose (C, K) (moco.py:390).
 *   queue_tf32  same shape         the same keys rounded to nearest tf32: what the
 *                                  tensor-core pass reads; `queue` stays the bit-exact
 *                                  fp32 master for state_dict().
 *   birth   int32   [K_local]      enqueue number at which the row was written;
 *                                  reference count[j] == n_enq - birth[j]
 *                                  (moco.py:427,437).
 *   qstate  int64   [4]            {ptr, n_enq, block-done counter, reserved}.
 *   qpack   float32 [M, 132]       per query row: q[0:128] | pos2 | shift2 | dup slot (int bits) | dup decay
 *   part    float32 [P, M, 132]    per CTA slab p and query row: O[0:128] | sum-exp | #neg>pos | 0 | 0
 *                                  (P = mscl_infonce_num_partials); summed over p in a fixed order
 */
#ifndef MSCL_B200_H_
#define MSCL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSCL_ABI_VERSION 1
#define MSCL_DIM 128          /* feature dimension of the contrastive space (cfg ft_dim) */
#define MSCL_PACK_LD 132      /* row pitch (floats) of qpack / acc */

#define MSCL_OK 0
#define MSCL_EINVAL (-1)      /* bad argument (shape, alignment, null pointer) */
#define MSCL_ECUDA (-2)       /* CUDA runtime / driver error */
#define MSCL_EUNSUPPORTED (-3)/* not an sm_100 device, or driver too old for TMA */

typedef void *mscl_stream_t;

int mscl_abi_version(void);
const char *mscl_last_error(void);
/* 0 if device `dev` can run this library (compute capability 10.x). */
int mscl_device_check(int dev);

/* Small host -> device transfer WITHOUT the copy engine: one CTA reads nbytes (a multiple of 16, <= 1 MiB) from page-locked
 * host memory that is mapped into the device (cudaHostAlloc / cudaHostRegister under unified addressing) and stores them to
 * d_dst.  For the few hundred per-step parameters the host draws (the augmentation's decisions, ssl_aug_v2.py:31-48): as a
 * cudaMemcpyAsync they queue in the H2D engine behind the data loader's next batch (128 MB per step here) and stall the
 * compute stream for the whole of it.  The host buffer must stay untouched until the kernel has run. */
int mscl_fetch_host(void *d_dst, const void *h_src_pinned, int64_t nbytes, mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K5  ring-buffer enqueue.   replaces MoCoV2._dequeue_and_enqueue
 *     (mmaction/models/recognizers/moco.py:423-440) minus the all_gather.
 * keys [B_all, C] are the rank-major gathered keys.  Rows whose global slot
 * ptr+i falls in [shard_begin, shard_begin+K_local) are written; the age of the
 * written rows becomes 1 and every other age grows by one (kept implicitly as
 * n_enq - birth).  ptr <- (ptr + B_all) % K_total; n_enq <- n_enq + 1.
 * If d_saved != NULL the overwritten rows (those in this shard) are first copied
 * to d_saved[B_all, C] -- from d_queue_tf32 when it is given, i.e. in the form the
 * tensor-core pass reads, else from d_queue -- and their birth to d_saved_birth[B_all]
 * (snapshot support, moco.py:484-488; mscl_infonce_fused_multi_x streams them for the
 * rows that read the queue as it was before this enqueue).  If d_queue_tf32 != NULL the new rows are also written there
 * rounded to nearest tf32: the operand copy the tensor-core pass reads (the tensor
 * core itself would truncate fp32, a systematic -3.5e-4 relative bias per operand).
 */
int mscl_enqueue(float *d_queue, float *d_queue_tf32, int32_t *d_birth,
                 int64_t *d_qstate, const float *d_keys, int32_t B_all, int32_t C,
                 int64_t K_total, int64_t shard_begin, int64_t K_local,
                 float *d_saved, int32_t *d_saved_birth, mscl_stream_t stream);

/* Materialise the reference's buffers for state_dict(): queue_ck [C, K_local]
 * (transposed) and count int64 [K_local] = n_enq - birth.  (moco.py:390-396) */
int mscl_queue_export(const float *d_queue, const int32_t *d_birth,
                      const int64_t *d_qstate, float *d_queue_ck,
                      int64_t *d_count, int32_t C, int64_t K_local,
                      mscl_stream_t stream);
/* Inverse: load reference-layout buffers. birth = n_enq - count with
 * n_enq taken from d_qstate[1] (caller sets it to max(count) beforehand). */
int mscl_queue_import(float *d_queue, float *d_queue_tf32, int32_t *d_birth,
                      const int64_t *d_qstate,
                      const float *d_queue_ck, const int64_t *d_count, int32_t C,
                      int64_t K_local, mscl_stream_t stream);
/* The decayed snapshot the reference calls `weight` (moco.py:484-486):
 * weight_ck[c, j] = queue[j, c] * 0.99999^(n_enq - birth[j]), float32 [C, K_local]. */
int mscl_queue_weight(const float *d_queue, const int32_t *d_birth,
                      const int64_t *d_qstate, float *d_weight_ck, int32_t C,
                      int64_t K_local, mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K4  multi-tensor momentum EMA.   replaces MoCoV2._momentum_update_key_encoder
 *     (moco.py:408-421): k <- fl(fl(k*m) + fl(q*(1-m))) element-wise, bit-exact
 *     with the reference's two multiplies and one add (no FMA contraction).
 * d_k_ptrs/d_q_ptrs/d_sizes: device tables of n_tensors entries.
 * d_blk_tensor/d_blk_start: device tables of n_blocks entries mapping a CTA to
 * (tensor index, first element); each CTA handles up to chunk_elems elements.
 */
int mscl_ema_multi(float *const *d_k_ptrs, const float *const *d_q_ptrs,
                   const int64_t *d_sizes, const int32_t *d_blk_tensor,
                   const int64_t *d_blk_start, int32_t n_blocks,
                   int32_t chunk_elems, float m, float one_minus_m,
                   mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K3  FRA: flow rotation augmentation.   replaces NormFlowWithStidedAug.__call__
 *     + norm_flow (mmaction/datasets/pipelines/transforms_motion.py:103-142,7-29).
 * layout 0: planar      flow [N, 2, T, H*W]   (u plane then v plane per clip)
 * layout 1: interleaved flow [N, T, H*W, 2]   (the reference's per-frame H,W,2)
 * d_cid int32 [N] chunk id per clip; d_cs float [2*num_chunks] = cos,sin pairs of
 * beta = (start + stride*cid)*pi.  Output planar [N, 2, 2T, H*W]: frames [0,T) =
 * base / (max|base| + 1e-5), frames [T,2T) = rotated / (max|rotated| + 1e-5),
 * maxima per frame.  d_maxrad float [N, T, 2] is scratch (base, rotated).
 * mscl_fra_maxrad must run before mscl_fra_apply on the same stream.
 */
int mscl_fra_maxrad(const float *d_flow, const int32_t *d_cid, const float *d_cs,
                    float *d_maxrad, int32_t N, int32_t T, int32_t HW,
                    int32_t layout, mscl_stream_t stream);
int mscl_fra_apply(const float *d_flow, const int32_t *d_cid, const float *d_cs,
                   const float *d_maxrad, float *d_out, int32_t N, int32_t T,
                   int32_t HW, int32_t layout, mscl_stream_t stream);
/* One-pass form of the two calls above for frames of up to 204800 pixels: a cluster of 8 CTAs per
 * frame stages the frame in (distributed) shared memory, so the flow is read once -- 24 B instead of
 * 32 B per (u,v) pair -- and no scratch is needed.  Same output, bit for bit. */
int mscl_fra_fused(const float *d_flow, const int32_t *d_cid, const float *d_cs,
                   float *d_out, int32_t N, int32_t T, int32_t HW, int32_t layout,
                   mscl_stream_t stream);
/* Rotation only on an already normalised planar clip [N,2,T,HW] -> [N,2,T,HW]
 * (Appendix A.10 of SURVEY.md: rotation preserves the per-frame maximum). */
int mscl_fra_rotate(const float *d_flow, const int32_t *d_cid, const float *d_cs,
                    float *d_out, int32_t N, int32_t T, int32_t HW,
                    mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K2  LMCL frame-level contrast.   replaces MSCLWithAugPosHeadV2.forward/.loss
 *     (mmaction/models/heads/local_cl_head.py:57-73,41-55).
 * (a) spatial mean over H*W of R = N*C*T rows (AdaptiveAvgPool3d((None,1,1)),
 *     local_cl_head.py:61-62) and its backward (broadcast of g/HW).
 * (b) the loss on pooled features xq [N, C, t], xf [N, C, t2] (t2 = 2t):
 *     L2-normalise over C (eps 1e-12), sim = xq^T xf / T, cross-entropy of row i
 *     against column i, top-1/top-5; forward and backward in one launch.
 *     d_out float [4] = {loss (mean over N*t rows), top1, top5, 0};
 *     d_gxq/d_gxf = d loss / d xq, d xf for unit upstream gradient.
 *     d_part float [N*4 + 4] is scratch; its last 4 floats must be ZERO before
 *     the first call (the kernel re-zeroes them for the next call).
 */
int mscl_hw_mean_fwd(const float *d_x, float *d_out, int64_t R, int32_t HW,
                     mscl_stream_t stream);
int mscl_hw_mean_bwd(const float *d_gout, float *d_gx, int64_t R, int32_t HW,
                     mscl_stream_t stream);
/* The same for a channels-last feature map (torch.channels_last_3d): x [N, T, H*W, C] in memory, C % 32 == 0, N*T <= 65535;
 * the pooled tensor (and its gradient) stays (N, C, T) row-major. */
int mscl_hw_mean_ndhwc_fwd(const float *d_x, float *d_out, int64_t N, int32_t C, int32_t T, int32_t HW, mscl_stream_t stream);
int mscl_hw_mean_ndhwc_bwd(const float *d_gout, float *d_gx, int64_t N, int32_t C, int32_t T, int32_t HW, mscl_stream_t stream);
int mscl_lmcl(const float *d_xq, const float *d_xf, int32_t N, int32_t C,
              int32_t t, int32_t t2, float inv_T, float *d_out, float *d_gxq,
              float *d_gxf, float *d_part, mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K1  fused InfoNCE: q.[k; queue] logits, temperature, age decay, log-sum-exp,
 *     cross-entropy, top-1/5 and d loss/d q in ONE pass over the queue.
 *     replaces MoCoV2.forward_train logits block (moco.py:481-498),
 *     MoCoHead.loss (heads/moco_head.py:38-77), CrossEntropyLoss_torch.forward
 *     (losses/cross_entropy_loss.py:134-138), top_k_accuracy
 *     (core/evaluation/accuracy.py:130-149) and
 *     MSCLWithAugMxHead._forward_moco_mx/.loss (heads/moco_head_v2.py:38-100).
 *
 * Three launches per pass (when the queue is sharded, reduce + a reduce-scatter sit
 * between 2 and 3):
 *  1 prep      qpack[i] = q_i | pos2 | shift2 | dup slot | dup decay ; dscale[j]
 *              pos2_i   = (q_i . kpos_i) / T * log2(e)
 *              shift2_i = |q_i| * key_norm_bound / T * log2(e)   (>= every logit)
 *              dscale_j = 0.99999^(n_enq - birth_j) / T * log2(e);  d_dscale holds
 *              ceil(K_local/128)*128 floats, the pad is written as 0 (d_dscale may be NULL when
 *              mscl_infonce_pass follows: it derives the scale from birth[] itself)
 *              d_dup_slot (int32 [M] or NULL): GLOBAL queue slot that holds a copy of
 *              row i's positive key (the cross-modal rf term reads the flow queue right
 *              after k_flow was enqueued into it, mscl.py:239-248), -1 for none; dup_age
 *              = age of those copies.  The reference scores such an entry at
 *              pos * 0.99999^age, i.e. above the positive iff pos < 0: a 1e-5 margin that
 *              tf32 operands cannot resolve, so `partial` skips the top-k hit test of that
 *              one column (its softmax mass is still summed) and `finalize` adds the exact
 *              comparison.
 *  2 partial   part[p][i] = ( sum_j p_ij dscale_j queue_j | sum_j p_ij | #{j: s_ij > pos2_i} )
 *              over the keys j of CTA slab p, with s_ij = (q_i . queue_j) dscale_j and
 *              p_ij = 2^(s_ij - shift2_i); tcgen05 (tf32 operands, fp32 accumulate in TMEM),
 *              queue and query tiles by TMA.  Plain stores, no atomics: every (p, i) row is
 *              written once, so the pass is bit-reproducible.  n_part CTAs along the keys
 *              (x ceil(M/128) row blocks); ask mscl_infonce_num_partials for n_part.
 *  2' reduce   acc[i] = sum_p part[p][i]  (only needed before a cross-GPU reduce-scatter)
 *  3 finalize  sums the n_part slabs, then per row: Z, lse, loss_i, p0, dq_unit_i; per group of
 *              rows_per_group consecutive rows: mean loss, top-1, top-5.
 * Row i belongs to group i / rows_per_group.  d_group_out float [n_groups, 4] =
 * {loss, top1, top5, 0}.  d_dq_unit [M, C] = d(group loss)/d q_i.
 * d_row_loss float [2*M]: loss_i for i<M, then #{j: s_ij > pos_i} (as float).
 * finalize uses part[0][0][130] as a CTA-done counter and leaves it zero.
 * mscl_infonce_bwd scales: d_dq[i] = d_dq_unit[i] * d_gout[group(i)].
 */
int mscl_infonce_prep(const float *d_q, const float *d_kpos, int32_t M,
                      const int32_t *d_birth, const int64_t *d_qstate,
                      int64_t K_local, float inv_T, float key_norm_bound,
                      float *d_qpack, float *d_dscale,
                      const int32_t *d_dup_slot, int32_t dup_age,
                      float *const *d_peer_qpack, int32_t n_peers, int32_t row_offset,
                      mscl_stream_t stream);
/* Sharded queue, exchange over NVLink peer memory (the reference replicates the queue and has no counterpart;
 * these replace an all_gather of the packed queries and a reduce-scatter of the partial results):
 *  - prep: with n_peers > 0, d_peer_qpack[p] (device table of n_peers pointers into every rank's gathered query
 *    table, this rank's own included) receives rows row_offset .. row_offset+M-1 by peer stores;
 *  - mscl_infonce_reduce_scatter: sums this rank's n_part CTA slabs of row i (i < M_all) and stores the 132 floats
 *    into d_peer_acc[i / M_local] + (my_rank * M_local + i % M_local) * 132, i.e. slot my_rank of the owner's
 *    accumulator [G, M_local, 132], which mscl_infonce_finalize then reads as G slabs.
 * The caller puts a device-side barrier across the ranks after each of the two (torch symmetric memory). */
int mscl_infonce_reduce_scatter(const float *d_part, int32_t n_part, int32_t M_all, int32_t M_local,
                                float *const *d_peer_acc, int32_t my_rank, mscl_stream_t stream);
/* Number of CTA slabs along the keys for M rows over K_local keys on num_sms SMs (> 0), or a
 * negative MSCL_E* code. */
int mscl_infonce_num_partials(int32_t M, int64_t K_local, int32_t num_sms);
/* shard_begin: global slot of d_queue[0] (0 when the queue is not sharded).
 * d_part: float [n_part, M, 132]. */
int mscl_infonce_partial(const float *d_qpack, int32_t M, const float *d_queue,
                         const float *d_dscale, int64_t K_local, int64_t shard_begin,
                         float *d_part, int32_t n_part, int32_t with_grad,
                         mscl_stream_t stream);
/* Same sums on CUDA cores in fp32 (validation twin): ACCUMULATES into d_acc float [M, 132]
 * (= one slab), which the caller must zero first. */
int mscl_infonce_partial_simt(const float *d_qpack, int32_t M, const float *d_queue,
                              const float *d_dscale, int64_t K_local, int64_t shard_begin,
                              float *d_acc, int32_t with_grad, mscl_stream_t stream);
int mscl_infonce_reduce(const float *d_part, int32_t n_part, int32_t M, float *d_acc,
                        mscl_stream_t stream);
int mscl_infonce_finalize(const float *d_qpack, const float *d_kpos,
                          float *d_part, int32_t n_part, int32_t M, int32_t rows_per_group,
                          float inv_T, int32_t with_grad, float *d_row_loss, float *d_dq_unit,
                          float *d_group_out, mscl_stream_t stream);
int mscl_infonce_bwd(const float *d_dq_unit, const float *d_gout, int32_t M,
                     int32_t rows_per_group, float *d_dq, mscl_stream_t stream);

/* K1, single-launch form (csrc/infonce_fused.cu): prep + pass + statistics reduction + finalize of the block above in ONE
 * kernel.  Same replaced reference lines (moco.py:481-498, heads/moco_head.py:38-77, heads/moco_head_v2.py:38-100,
 * losses/cross_entropy_loss.py:134-138, core/evaluation/accuracy.py:130-149) and the same formulas; the differences:
 *  - q / kpos are the raw fp32 rows [M, 128] (no qpack): pos2, shift2 and the tf32 rounding of q happen in the kernel;
 *  - the per-key scale 0.99999^(n_enq - birth_j) / T * log2(e) is computed from d_birth / d_qstate tile by tile (no dscale);
 *  - the row statistics (sum-exp, hit count) of all CTAs are added into d_ws with red.global.add (the float sum-exp in no
 *    fixed order: the loss is reproducible to rounding, the hit counts exactly); the last CTA to finish turns them into
 *    d_row_loss / d_group_out (same meaning as mscl_infonce_finalize) and d_rowaux;
 *  - the O partials stay per-CTA slabs d_part [n_part, M, 132] (written only when with_grad); mscl_infonce_bwd_slabs sums
 *    them in a fixed order when autograd asks for the gradient:
 *        d_dq[i] = d_gout[group(i)] * ( ck_i * kpos_i + co_i * sum_p part[p][i][0:128] ),   (ck_i, co_i) = d_rowaux[i][2:4].
 * d_ws: float [16*M*4 + 4] workspace (16 copies of the statistics accumulator + CTA counter).  It must be ZERO before the first call; the
 * kernel leaves it zero.  One workspace per stream: two calls that may run concurrently must not share it.
 * d_rowaux: float [M, 4] = pos2, shift2, ck, co per row.
 * n_part: CTAs along the keys (x ceil(M/128) row blocks), from mscl_infonce_fused_parts.
 * flags: MSCL_INFONCE_EARLY_PREFETCH -- the queue was NOT written by the launch immediately preceding this one on the
 * stream, so its tiles may be requested before the programmatic-dependent-launch wait.
 */
#define MSCL_INFONCE_EARLY_PREFETCH 1
int mscl_infonce_fused(const float *d_q, const float *d_kpos, int32_t M, const float *d_queue_tf32,
                       const int32_t *d_birth, const int64_t *d_qstate, int64_t K_local, float inv_T,
                       float key_norm_bound, const int32_t *d_dup_slot, int32_t dup_age, float *d_ws,
                       float *d_part, int32_t n_part, int32_t rows_per_group, int32_t with_grad, int32_t flags,
                       float *d_row_loss, float *d_rowaux, float *d_group_out, mscl_stream_t stream);
/* Several independent terms ("jobs": own queries, queue, workspace and outputs) in ONE launch: blockIdx.y runs over the row
 * blocks of all jobs, so they occupy disjoint SMs and pay the launch's ramp / drain / finalize once.  Every argument of
 * mscl_infonce_fused becomes a HOST array of n_jobs (<= 4) entries (d_dup_slot[j] / d_part[j] may be NULL as there);
 * n_part from mscl_infonce_fused_parts_multi, with_grad common to all jobs, one distinct workspace per job. */
int mscl_infonce_fused_multi(int32_t n_jobs, const float *const *d_q, const float *const *d_kpos, const int32_t *M,
                             const float *const *d_queue_tf32, const int32_t *const *d_birth,
                             const int64_t *const *d_qstate, const int64_t *K_local, const float *inv_T,
                             const float *key_norm_bound, const int32_t *const *d_dup_slot, const int32_t *dup_age,
                             float *const *d_ws, float *const *d_part, int32_t n_part, const int32_t *rows_per_group,
                             int32_t with_grad, const int32_t *flags, float *const *d_row_loss, float *const *d_rowaux,
                             float *const *d_group_out, mscl_stream_t stream);
int mscl_infonce_fused_parts_multi(int32_t n_jobs, const int32_t *M, const int64_t *K_local, int32_t num_sms);
/* The same launch with an "epoch split" on ONE of its jobs (x_job; -1: none, then identical to mscl_infonce_fused_multi):
 * rows [row_split, M) of that job read its queue as it is, rows [0, row_split) read it as it was BEFORE its last
 * enqueue, which wrote rep_n (<= 128) keys into slots [rep_begin, rep_begin + rep_n) and saved what it overwrote:
 * d_xkeys [rep_n, 128] / d_xbirth [rep_n] = mscl_enqueue's d_saved / d_saved_birth.  Before that enqueue every age was
 * one lower and those slots held the saved keys (moco.py:423-440: count += 1, count[ptr:ptr+B] = 1).
 * That is the pair of queue states MSCLWithAug reads within one step (mscl.py:228-277: the base-flow call's logits
 * before its enqueue, the FRA call's and the cross-modal rf terms after it), so the step streams W_flow ONCE for both.
 * How: the pre rows' q is scaled by 1/0.99999 before its tf32 rounding (a uniform age - 1) and their gradient
 * coefficient by the same factor; in the queue's own tiles the freshly written slots are masked for them; the saved keys
 * are one extra 128-key tile (same key indices, the age they would have now) that only the pre rows see, streamed by
 * the CTA with the fewest queue units.  d_dup_slot / dup_age keep their meaning (rows at or after row_split whose
 * positive is one of the keys just enqueued: global slot rep_begin + i, dup_age = 1). */
int mscl_infonce_fused_multi_x(int32_t n_jobs, const float *const *d_q, const float *const *d_kpos, const int32_t *M,
                               const float *const *d_queue_tf32, const int32_t *const *d_birth,
                               const int64_t *const *d_qstate, const int64_t *K_local, const float *inv_T,
                               const float *key_norm_bound, const int32_t *const *d_dup_slot, const int32_t *dup_age,
                               float *const *d_ws, float *const *d_part, int32_t n_part, const int32_t *rows_per_group,
                               int32_t with_grad, const int32_t *flags, float *const *d_row_loss, float *const *d_rowaux,
                               float *const *d_group_out, int32_t x_job, const float *d_xkeys, const int32_t *d_xbirth,
                               int64_t rep_begin, int32_t rep_n, int32_t row_split, mscl_stream_t stream);
int mscl_infonce_bwd_slabs(const float *d_part, int32_t n_part, int32_t M, const float *d_kpos, const float *d_rowaux,
                           const float *d_gout, int32_t rows_per_group, float *d_dq, mscl_stream_t stream);
/* mscl_infonce_bwd_slabs for all jobs of one mscl_infonce_fused_multi(_x) launch (same n_part) in ONE launch: host arrays
 * of n_jobs (<= 4) entries. */
int mscl_infonce_bwd_slabs_multi(int32_t n_jobs, const float *const *d_part, int32_t n_part, const int32_t *M,
                                 const float *const *d_kpos, const float *const *d_rowaux, const float *const *d_gout,
                                 const int32_t *rows_per_group, float *const *d_dq, mscl_stream_t stream);
/* The pass alone in the same form, for the sharded queue: d_qpack [M, 132] is the gathered table mscl_infonce_prep fills
 * on every rank; d_part float [n_part, M, 132] receives the per-CTA slabs exactly as mscl_infonce_partial writes them
 * (mscl_infonce_reduce_scatter / mscl_infonce_finalize follow as before), n_part from mscl_infonce_fused_parts. */
int mscl_infonce_pass(const float *d_qpack, int32_t M, const float *d_queue_tf32, const int32_t *d_birth,
                      const int64_t *d_qstate, int64_t K_local, int64_t shard_begin, float inv_T, float *d_part,
                      int32_t n_part, int32_t with_grad, int32_t flags, mscl_stream_t stream);
int mscl_infonce_fused_parts(int32_t M, int64_t K_local, int32_t num_sms);

/* ---------------------------------------------------------------------------
 * K6  shuffle-BN row gather.   replaces x_gather[idx_this] in
 *     MoCo._batch_shuffle_ddp / _batch_unshuffle_ddp (moco.py:172,191).
 * out[r, :] = x[idx[r], :] for r < n_rows, rows of row_elems floats
 * (row_elems % 4 == 0), 128-bit copies.
 */
int mscl_gather_rows(const float *d_x, const int64_t *d_idx, float *d_out,
                     int32_t n_rows, int64_t row_elems, mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K7  trilinear up-sampling between pyramid levels of the TPN neck.   replaces
 *     F.interpolate(x, size=[T,H,W], mode="trilinear") (align_corners=False) in
 *     mmaction/models/necks/sepc.py:126-130.  x [NC, Ti, Hi, Wi] -> y [NC, To, Ho, Wo], fp32, contiguous.
 * The backward is a gather over the outputs that read each input element (no atomics).
 */
int mscl_upsample_trilinear_fwd(const float *d_x, float *d_y, int64_t NC, int32_t Ti, int32_t Hi,
                                int32_t Wi, int32_t To, int32_t Ho, int32_t Wo, mscl_stream_t stream);
int mscl_upsample_trilinear_bwd(const float *d_gy, float *d_gx, int64_t NC, int32_t Ti, int32_t Hi,
                                int32_t Wi, int32_t To, int32_t Ho, int32_t Wo, mscl_stream_t stream);
/* The same for channels-last tensors (torch.channels_last_3d): x [N, Ti, Hi, Wi, C] -> y [N, To, Ho, Wo, C], C % 4 == 0. */
int mscl_upsample_trilinear_ndhwc_fwd(const float *d_x, float *d_y, int64_t N, int32_t C, int32_t Ti, int32_t Hi,
                                      int32_t Wi, int32_t To, int32_t Ho, int32_t Wo, mscl_stream_t stream);
int mscl_upsample_trilinear_ndhwc_bwd(const float *d_gy, float *d_gx, int64_t N, int32_t C, int32_t Ti, int32_t Hi,
                                      int32_t Wi, int32_t To, int32_t Ho, int32_t Wo, mscl_stream_t stream);
/* One axis of that backward (the trilinear weights factorise): src [outer][out_size][inner4 float4s] ->
 * dst [outer][in_size][inner4], dst[.., i, ..] = sum over the outputs o that read input index i of w(o, i) * src[.., o, ..].
 * Three calls (W, H, T) give the channels-last input gradient with every intermediate read once. */
int mscl_linear_axis_bwd(const float *d_src, float *d_dst, int64_t outer, int32_t in_size, int32_t out_size, int64_t inner4,
                         mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K8  flow visualisation + flip.   replaces FlowVisualizer.__call__ / flow_uv_to_colors
 *     (mmaction/models/common/ssl_aug.py:87-136, colour wheel tools/RAFT/core/utils/flow_viz.py:20-67) and the
 *     mirror of SyncMoCoAugmentV5.forward_flip (common/ssl_aug_v2.py:109-117) for the flow images.
 * flow planar [N, 2, T, H, W] -> out [N, 3, T, H, W] in {0, 1/255, ..., 1}.  d_flip uint8 [N] (mirror along W;
 * may be NULL); d_norm float[6] = mean[3], std[3] applied after the colour lookup (NULL: normalize_flow=False).
 * Arithmetic follows the reference's dtypes: float32 up to f = fk - k0 and (1 - f), float64 for the interpolation.
 */
int mscl_flow_visualize(const float *d_flow, const uint8_t *d_flip, const float *d_norm, float *d_out,
                        int32_t N, int32_t T, int32_t H, int32_t W, mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K9  RGB colour pipeline of SyncMoCoAugmentV5 (common/ssl_aug_v2.py:31-48,66-68, common/ssl_aug.py:138-174):
 *     flip -> ColorJitter(brightness, contrast, saturation, hue) -> grayscale -> separable Gaussian blur
 *     (reflect border) -> Normalize, decisions and parameters per clip, drawn by the caller.
 * x [N, 3, T, H, W] in [0,1] -> out same shape.  d_params float [U][16]: flip, jitter?, brightness, contrast,
 * saturation, hue matrix[9] (row major, RGB -> RGB), gray?, blur?, with U = N rows (per_frame = 0: one set per clip, the
 * reference's 'params' sync level, toConsistentAug, ssl_aug.py:62-66) or U = N*T rows (per_frame = 1: one set per frame,
 * row n*T + t -- the 'batch' sync level, toVideoAug, :56-60, shares only the apply decisions between the frames of a
 * clip; the caller repeats those); the contrast step is taken about the mean luminance of the clip / of the frame.
 * d_taps float[n_taps] (odd, <= 31) the 1-D blur kernel; d_norm float[6] = mean[3], std[3]; d_gray_partial float
 * [U][n_chunks] scratch for the luminance sums.
 */
int mscl_color_pipeline(const float *d_x, const float *d_params, const float *d_taps, int32_t n_taps,
                        const float *d_norm, float *d_gray_partial, int32_t n_chunks, float *d_out,
                        int32_t N, int32_t T, int32_t H, int32_t W, int32_t per_frame, mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K10  gradient-norm clip + SGD-momentum step, multi-tensor.   replaces torch.nn.utils.clip_grad_norm_(max_norm=40)
 *      (mmcv OptimizerHook, optimizer_config.grad_clip) followed by torch.optim.SGD(lr, momentum, weight_decay).step()
 *      (configs/recognition/moco/mscl_r18_cosm_lr2e-2.py:112-119).  Tables as for mscl_ema_multi.
 * mscl_grad_norm_multi: d_stats[0] = ||g||_2 over all tensors, d_stats[1] = min(1, max_norm / (||g|| + 1e-6));
 *   d_partial float[n_blocks] scratch (per-CTA sums, reduced in a fixed order).
 * mscl_clip_sgd_multi:  g <- g * d_stats[1] (d_stats NULL: no clipping); d = g + weight_decay * p;
 *   buf <- momentum * buf + d (first_step != 0: buf <- d);  p <- p - lr * buf.
 */
int mscl_grad_norm_multi(const float *const *d_g_ptrs, const int64_t *d_sizes, const int32_t *d_blk_tensor,
                         const int64_t *d_blk_start, int32_t n_blocks, int32_t chunk_elems, float max_norm,
                         float *d_partial, float *d_stats, mscl_stream_t stream);
int mscl_clip_sgd_multi(float *const *d_g_ptrs, float *const *d_p_ptrs, float *const *d_buf_ptrs,
                        const int64_t *d_sizes, const int32_t *d_blk_tensor, const int64_t *d_blk_start,
                        int32_t n_blocks, int32_t chunk_elems, const float *d_stats, float weight_decay,
                        float momentum, float lr, int32_t first_step, mscl_stream_t stream);

/* ---------------------------------------------------------------------------
 * K11  nearest-neighbour retrieval evaluation.   replaces tools/test_retrival.py:286-304:
 *        feature -= feature.mean(dim=0); feature = F.normalize(feature, p=2, dim=1)           (train and test)
 *        sim = test @ train.T                    (a plain library GEMM: cuBLAS, done by the caller)
 *        for k in (1,5,10,20,50): acc = any(train_label[topk(sim, k).indices] == test_label[:, None], dim=1).mean()
 * mscl_center_normalize: x [N, D] row-major -> out [N, D]; d_partial double [n_chunks][D] and d_mean float [D] scratch
 *   (column sums in double, fixed order).
 * mscl_retrieval_rank:   sim [n_test, ld >= n_train] -> rank int32 [n_test] = number of train items scoring strictly
 *   above the test item's best same-label train item (n_train if the label never occurs), so that the reference's
 *   acc@k == mean(rank < k) for every k at once (ties at the k-th value aside, which torch.topk breaks arbitrarily).
 */
int mscl_center_normalize(const float *d_x, int64_t N, int32_t D, double *d_partial, int32_t n_chunks, float *d_mean,
                          float *d_out, mscl_stream_t stream);
int mscl_retrieval_rank(const float *d_sim, int64_t ld, const int64_t *d_train_label, const int64_t *d_test_label,
                        int32_t n_test, int32_t n_train, int32_t *d_rank, mscl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MSCL_B200_H_ */
