/*
 * links_b200.h -- C ABI of the B200-native LInKs lifting hot path.
 *
 * The reference (Aswarin/LInKs-3D-Human-Pose-Estimation) is pure Python/PyTorch and has no
 * FFI of its own: the drop-in boundary is its Python module surface
 * (utils/models_def.py, utils/helpers.py, utils/rotation_conversions.py, utils/metrics_batch.py,
 * utils/metrics.py, FrEIA SequenceINN/AllInOneBlock).  This header is what those Python modules
 * bind (ctypes); each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is DEVICE memory owned by the caller (PyTorch);
 *     the library never allocates, frees or synchronises; all launches are asynchronous on
 *     `stream` (a cudaStream_t passed as void*) and CUDA-graph capturable.
 *   - return 0 on success, negative = argument error (LINKS_E_*), positive = cudaError_t.
 *   - fp32 tensors are row-major and dense unless a leading dimension is given.
 *   - "bf16" tensors are raw 16-bit bfloat16.
 *   - 2D poses are [M,34] = (17 x, 17 y); 3D poses are [M,51] = (17 X, 17 Y, 17 Z)
 *     (reference utils/helpers.py:262-267 layout).
 */
#ifndef LINKS_B200_H
#define LINKS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LINKS_ABI_VERSION 1

#define LINKS_E_ARG (-1)      /* null pointer / bad size */
#define LINKS_E_ALIGN (-2)    /* pointer or leading dimension not 16-byte aligned */
#define LINKS_E_RANGE (-3)    /* value outside the supported range */
#define LINKS_E_DRIVER (-4)   /* cuTensorMapEncodeTiled unavailable / failed */

#define LINKS_MAX_GEMM_PROBLEMS 8
#define LINKS_MAX_PUSH_RANKS 8
#define LINKS_HEAD_LD 32      /* leading dimension of fp32 head outputs / part-gradient buffers */
#define LINKS_KPAD 64         /* bf16 operand rows are padded to a multiple of 64 in K */
#define LINKS_J 17

int links_abi_version(void);
/* 1 if the library was built for sm_100a and the current device is CC 10.x, else 0. */
int links_device_ok(void);
/* Kernels launched through this library since it was loaded (every entry point counts its own launches). */
size_t links_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Grouped GEMM with fused epilogue (tcgen05 / TMEM / TMA).   D = epi(A[M,K] * B[N,K]^T)
 * Replaces every nn.Linear (+LeakyReLU, +residual) of reference utils/models_def.py:10-39,
 * 111-327 in forward, and its autograd backward (dgrad / wgrad).  A and B are bf16, accumulate fp32.
 * Operand layouts: by default K-major (A stored [M, lda], B stored [N, ldb], K contiguous -- nn.Linear's
 * weight layout for forward).  LINKS_GEMM_A_MN: A is stored [K, lda] (M contiguous); LINKS_GEMM_B_MN: B is
 * stored [K, ldb] (N contiguous).  dgrad dX = G.W passes W [N_layer, K_layer] as an MN-major B; wgrad
 * dW = G^T.X passes the row-major G and X as MN-major A and B -- no transposed copies exist anywhere.
 *
 * Epilogue, applied per element in this order (null pointer = step skipped):
 *   v = acc (+ bias[n])
 *   sign_out bit (m, n) = !(v > 0)                       [1 bit / element, 32 columns / word]
 *   LEAKY_PRE : v = v > 0 ? v : 0.01 v      (RELU_PRE : v = max(v, 0))
 *   v += add0[m,n]  ; v += add1[m,n]                     [bf16]
 *   LEAKY_POST: v = v > 0 ? v : 0.01 v
 *   v *= (ymask[m,n] > 0 ? 1 : 0.01)                     [bf16 activation, leaky' of its producer; 0 with YMASK_ZERO]
 *   mid[m,n] = bf16(v)
 *   v *= (bits(m,n) ? 0.01 : 1)
 *   out[m,n] = bf16(v); out_f32[m,n] (+)= v
 * ------------------------------------------------------------------------------------------ */
#define LINKS_EPI_LEAKY_PRE 1u
#define LINKS_EPI_LEAKY_POST 2u
#define LINKS_EPI_RELU_PRE 4u
#define LINKS_EPI_ACCUM_F32 8u
#define LINKS_GEMM_A_MN 16u
#define LINKS_GEMM_B_MN 32u
#define LINKS_EPI_YMASK_ZERO 64u   /* ymask multiplies by 0 instead of 0.01 where ymask <= 0 (ReLU') */

typedef struct LinksGemmProblem {
  const void* A;      /* bf16 [M, lda]  (A_MN: [K, lda]) */
  const void* B;      /* bf16 [N, ldb]  (B_MN: [K, ldb]) */
  int M, N, K;        /* K = contraction length; lda, ldb multiples of 8 */
  int lda, ldb;
  uint32_t flags;
  const float* bias;  /* [N], 16-byte aligned */
  const void* add0; int ld_add0;   /* bf16 [M, ld] */
  const void* add1; int ld_add1;
  const void* ymask; int ld_ymask; /* bf16 [M, ld] */
  const uint32_t* bits; int ld_bits;   /* [M, ld_bits] words, bit (n & 31) of word n >> 5 */
  uint32_t* sign_out; int ld_sign;
  void* mid; int ld_mid;           /* bf16 [M, ld], ld multiple of 8 */
  void* out; int ld_out;           /* bf16 [M, ld], ld multiple of 8 */
  float* out_f32; int ld_f32;      /* fp32 [M, ld] */
  /* Fused optimiser (weight-gradient problems of a single-GPU step; all NULL otherwise).  When adam_p is set the
   * problem must have NO other epilogue step and no out_f32: instead of storing dW the epilogue applies
   * torch.optim.Adam (train_leg_torso_lifter.py:111-114) to its tile -- g = acc * grad_scale + weight_decay * p, then m,
   * v, p in place (fp32 [M, ld_f32], 16-byte aligned rows, N a multiple of 64) -- and refreshes the bf16 shadow
   * [M, ld_shadow] that the next step's GEMMs read.  adam_hyper: 8 device floats written by links_adam_prepare. */
  float* adam_p; float* adam_m; float* adam_v;
  void* adam_shadow; int ld_shadow;
  const float* adam_hyper;
  /* Data-parallel reduce-scatter fused into the weight-gradient GEMM (push_rows > 0; no other epilogue step, no out_f32):
   * rows [r * push_rows, (r + 1) * push_rows) of the result are stored as bf16 to push[r] ([push_rows, ld_push] row-major) --
   * rank r's staging buffer, mapped into this process over NVLink (torch symmetric memory / CUDA IPC).  push_rows is a
   * multiple of 128 that divides M, M / push_rows <= LINKS_MAX_PUSH_RANKS, N a multiple of 64.  The owner then reduces
   * the ranks' slots and updates its rows with links_adam_zero. */
  void* push[LINKS_MAX_PUSH_RANKS]; int push_rows; int ld_push;
} LinksGemmProblem;

int links_gemm_grouped(const LinksGemmProblem* problems, int n_problems, void* stream);

/* Chain launch: a whole DEPENDENT sequence of grouped GEMMs -- all layers of a lifter forward pass (reference
 * utils/models_def.py:133-152: upscale -> res_common -> res_pose1..3 / res_angle1..3 -> heads, for several networks at
 * once) or of its autograd backward (dgrad chain + weight gradients) -- as ONE persistent kernel.  Tiles of layer l+1
 * start as soon as the tiles of layer l that produce their operand rows have completed (completion counters in global
 * memory), so kernel setup / TMEM allocation / pipeline fill are paid once per pass and the epilogue of one layer
 * overlaps the main loop of the next.  Same arithmetic as links_gemm_grouped, problem by problem.
 *   level        problems with equal level are independent of each other; their tiles are interleaved row block by
 *                row block.  Levels must ascend along every dependency.
 *   dep[0..2]    index of the EARLIER problem of the chain that writes this problem's [A operand, add0, add1] (out /
 *                mid / out_f32 of that problem), or -1 when the operand comes from outside the chain.  A dependency is
 *                row-block-wise: rows [256 i, 256 i + 256) of the consumer need the same rows of the producer (equal
 *                M required) -- unless bit d of dep_all_rows is set: the consumer waits for ALL tiles of the producer
 *                (an A_MN operand contracted over the rows: weight gradients; or a write-after-read hazard: the fused
 *                optimiser may only overwrite a layer's bf16 shadow once the dgrad problem that reads it has finished --
 *                such ordering-only dependencies use a slot whose operand pointer is NULL).
 * The workspace (links_gemm_chain_ws_bytes, 256-byte aligned device memory owned by the caller) holds descriptors,
 * schedule and counters; links_gemm_chain_build fills it (host work + one synchronous copy: call it outside stream
 * capture), links_gemm_chain_run launches (capturable).  One plan must not run concurrently with itself. */
#define LINKS_MAX_CHAIN_PROBLEMS 512
typedef struct LinksChainProblem {
  LinksGemmProblem g;
  int level;
  int dep[3];
  int dep_all_rows;
} LinksChainProblem;
typedef struct LinksGemmChainPlan {
  void* ws; void* probs; void* sched; void* sched_cnt; void* counters;
  int grid, n_counters, sched_ld, n_problems, total_tiles;
  float sim_units;     /* makespan of the host's schedule simulation, in 64-deep k-block units (~0.44 us) */
  float ideal_units;   /* sum of the main loops / clusters: the dependency- and epilogue-free bound */
} LinksGemmChainPlan;
size_t links_gemm_chain_ws_bytes(const LinksChainProblem* problems, int n_problems);
int links_gemm_chain_build(const LinksChainProblem* problems, int n_problems, void* ws_dev, size_t ws_bytes,
                           LinksGemmChainPlan* plan, void* stream);
int links_gemm_chain_run(const LinksGemmChainPlan* plan, void* stream);
/* Number of GEMM kernel launches (grouped + chain) issued through this library since it was loaded. */
size_t links_gemm_launch_count(void);
/* Cap the persistent grid of links_gemm_grouped at n CTAs (0 = one per SM).  Data-parallel runs leave a few SMs free so
 * that the concurrently running NCCL all-reduce kernels never push GEMM CTAs into a second wave.  Returns the old cap. */
int links_gemm_set_max_ctas(int n);

/* ------------------------------------------------------------------------------------------
 * Operand packing / reductions around the GEMMs
 * ------------------------------------------------------------------------------------------ */
/* Gather a joint subset of fp32 rows into a zero-padded bf16 GEMM operand (+ transposed copy).
 * Replaces split_data_left_right (utils/helpers.py:55-65), the leg/torso slices
 * (train_leg_torso_lifter.py:147-148) and the occlusion input gathers
 * (train_occlusion_models.py:185-191) as integer index maps:
 *   dst[m, c] = bf16(src[m * ld_src + idx[c]]) for c < n_idx, 0 for n_idx <= c < 64
 *   dstT[c, colT0 + m] = same, for c < n_idx                   (dstT may be null)
 * period > 1: idx holds period*n_idx offsets relative to a group of `period` rows
 * (split_data_left_right_3d, utils/helpers.py:81-91, mixes row pairs; period = 2).          */
int links_pack_rows(const float* src, int ld_src, int M, const int* idx_dev, int n_idx, int period,
                    void* dst_bf16, void* dstT_bf16, int ldT, int colT0, void* stream);

/* Column sums of a bf16 matrix: out[n] (+)= sum_m G[m, n]  (bias gradients). */
int links_colsum_bf16(const void* G, int ldg, int M, int N, float* out, int accumulate, void* stream);

/* The same for up to LINKS_MAX_COLSUM_ITEMS matrices in two launches (all bias gradients of a step). */
#define LINKS_MAX_COLSUM_ITEMS 96
typedef struct LinksColsumItem {
  const void* G;   /* bf16 [M, ldg] */
  float* out;      /* [N] */
  int ldg, M, N, accumulate;
} LinksColsumItem;
int links_colsum_bf16_batched(const LinksColsumItem* items, int n_items, void* stream);

/* fp32 master weights -> bf16 shadow W[N, Kpad] and W^T[K, Npad] (zero padded). */
int links_cast_weight(const float* W, int N, int K, void* W_bf16, int ldw, void* WT_bf16, int ldwt,
                      void* stream);

/* The same for all layers of a network set in one launch (bf16 shadows W[N, ldw] only). */
#define LINKS_MAX_CAST_ITEMS 128
typedef struct LinksCastItem {
  const float* W;  /* fp32 [N, K] */
  void* Wb;        /* bf16 [N, ldw], columns >= K are written as zero */
  int N, K, ldw, pad_;
} LinksCastItem;
int links_cast_weight_batched(const LinksCastItem* items, int n_items, void* stream);

/* torch.optim.Adam step with coupled L2 decay (train_leg_torso_lifter.py:111-114), flat buffers.
 * The step number t (bias correction) is `step` (>= 1), or, when step_dev != NULL, *step_dev + 1 read on the
 * device; *step_dev is then incremented after the update (unless step == -1: used when one optimiser step is issued as
 * several per-bucket launches), so a captured CUDA graph replays correctly.
 * Gradients are multiplied by grad_scale first (1/world_size after a SUM all-reduce).
 * lr_dev (may be NULL): when given, the learning rate is read from *lr_dev on the device instead of `lr`, so that a
 * captured CUDA graph follows the ExponentialLR schedule (train_leg_torso_lifter.py:116-121) without re-capture. */
int links_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    int* step_dev, float grad_scale, const float* lr_dev, void* stream);

/* out[n_out] = mat[n_out, n_in] . in[n_in] (fp32, tiny sizes): the loss summary of a step (train_leg_torso_lifter.py:
 * 266-284 -- means, weights and the total of the device-side loss sums) as one launch. */
int links_small_matvec(const float* mat, const float* in, int n_in, int n_out, float* out, void* stream);

/* Step constants of the fused optimiser (LinksGemmProblem.adam_*), computed on the device so that captured graphs
 * replay correctly: hyper[0..7] = { lr / (1 - beta1^t), sqrt(1 - beta2^t), eps, beta1, beta2, weight_decay, grad_scale, t }
 * with t = *step_dev + 1 and lr = *lr_dev when lr_dev != NULL.  Does not advance *step_dev. */
int links_adam_prepare(const int* step_dev, const float* lr_dev, float lr, float beta1, float beta2, float eps,
                       float weight_decay, float grad_scale, float* hyper, void* stream);

/* Sharded optimiser step behind the push reduce-scatter (ZeRO-1 style; one launch for all big layers of a network set).
 * For layer l (rows x cols weights, row-major, at offset master_off in the flat fp32 buffers p / m / v) this rank owns rows
 * [rank * rows_per_owner, +rows_per_owner).  The gradient of an owned element is the fp32 sum over the `world` slots of
 * `stage` (bf16, slot s at stage + s * stage_slot_elems, the layer's owned block at stage_off inside a slot, written by
 * rank s's GEMM epilogue); Adam with the constants of links_adam_prepare (grad_scale = 1 / world) updates p, m, v in place
 * and the new bf16 value is stored into EVERY rank's shadow of the layer (shadow[r]: peer-mapped [rows, cols] bf16) --
 * the all-gather of the updated weights, also by peer stores. */
typedef struct LinksAdamZeroLayer {
  unsigned long long master_off;    /* element offset of the layer's [rows, cols] block in p / m / v */
  unsigned long long stage_off;     /* element offset of the layer's owned block inside one staging slot */
  void* shadow[LINKS_MAX_PUSH_RANKS];
} LinksAdamZeroLayer;
int links_adam_zero(float* p, float* m, float* v, const void* stage, size_t stage_slot_elems, const LinksAdamZeroLayer* layers_dev,
                    int n_layers, int rows_per_owner, int cols, int world, int rank, const float* hyper, void* stream);

/* Device-side barrier over the ranks of one node (one process per GPU): flag_ptrs[r] (HOST array of `world` device
 * pointers) is rank r's zero-initialised flag array (>= 64 uint32), mapped into this process (symmetric memory).  Peer
 * stores issued by earlier work of the stream are visible to every rank behind the barrier.  slot 0..7: independent
 * barriers.  A plain kernel launch (CUDA-graph capturable); traps after ~10 s if a rank never arrives. */
int links_peer_barrier(void* const* flag_ptrs, int world, int rank, int slot, void* stream);

/* Data-parallel gradient compression: grad_bf16[i] = bf16(grad[i]) before the NCCL all-reduce (half the NVLink
 * bytes), and the Adam step that consumes the reduced bf16 gradients directly (same arithmetic otherwise). */
int links_grad_compress_bf16(const float* grad, void* grad_bf16, size_t n, void* stream);
int links_adam_step_g16(float* param, const void* grad_bf16, float* exp_avg, float* exp_avg_sq, size_t n,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                        int* step_dev, float grad_scale, const float* lr_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Geometry + losses (train_leg_torso_lifter.py:153-272, train_left_right_lifter.py:150-423)
 * ------------------------------------------------------------------------------------------ */
typedef struct LinksGeomMaps {
  int V;                 /* pose variants: 1 (leg/torso) or 2 (left-choice, right-choice) */
  int n_joints[2];       /* joints per part net (7,10) or (11,11) */
  int src_net[2][17];    /* [v][j]: which net's depth head feeds joint j of variant v */
  int col[17];           /* column of that head (same for both variants) */
  int part_net[2][17];   /* [v][j]: part net whose pass-2 / flow input takes joint j of variant v, or -1 */
  int part_idx[2][17];   /* index of joint j inside that part */
  float bone_rel[16];    /* bone-length prior constants */
  float depth, w_likeli, w_2d, w_3d, w_vel, w_bl;
} LinksGeomMaps;

/* (mean, unbiased std) of props = (ang0 + ang1)/2 over N rows -> stats[2]
 * (train_leg_torso_lifter.py:153,168).  ang* are head outputs with leading dim LINKS_HEAD_LD. */
int links_elev_stats(const float* ang0, const float* ang1, int N, float* stats, void* stream);

/* lift, rotate, project (:159-199).  Writes per-part projected inputs qpart[p] fp32 [N, 2*n_joints[p]]
 * (x's then y's) and optionally the full rot_2d per variant q[v] [N,34]. */
int links_geom_forward(const LinksGeomMaps* maps, const float* u, const float* head0, const float* head1,
                       const float* ang0, const float* ang1, const float* eps_x, const float* u_y,
                       const float* stats, int N, float* qpart0, float* qpart1, float* q_full0,
                       float* q_full1, void* stream);

/* Re-lift consistency / reprojection / pairwise / bone losses and their gradient w.r.t. the pass-2
 * depth heads (:228-259).  loss_sums[4] += (sum L3d, sum rep_rot, sum pair, sum bl) (un-normalised);
 * g2_head[p] bf16 [N,64] = dLoss/d(pass-2 head p) (+ transposed copies at column colT0).
 * The rows leave the kernel as whole 16-byte chunks: inside the chunks that hold a fed column, columns that no
 * joint feeds (the root's included) are written as zero; chunks beyond the last used column are not touched.
 * u, the dense part tensors and the gradient rows must be 16-byte aligned (LINKS_E_ALIGN); N <= 2^25 - 8. */
int links_geom_loss(const LinksGeomMaps* maps, const float* u, const float* head0, const float* head1,
                    const float* ang0, const float* ang1, const float* eps_x, const float* u_y,
                    const float* stats, const float* head2_0, const float* head2_1, int N,
                    float* loss_sums, void* g2_head0, void* g2_head1, void* g2T_head0, void* g2T_head1,
                    int ldT, int colT0, void* stream);

/* Backward of the whole geometry given external gradients on the projected parts
 * (dpart_flow[p] [N, 2*n_joints[p]] from the flows, dpart_lift[p] [N, LINKS_HEAD_LD] from the pass-2
 * upscale dgrad).  Phase A: per-row gradients, writes g1_head[p] (bf16 [N,64] + T), dgamma_direct[N],
 * da[N] and accumulates red[2] += (sum da, sum eps*da).  Phase B (links_geom_backward_angles) adds the
 * batch-statistic terms and writes the angle-head gradients. */
int links_geom_backward(const LinksGeomMaps* maps, const float* u, const float* head0, const float* head1,
                        const float* ang0, const float* ang1, const float* eps_x, const float* u_y,
                        const float* stats, const float* head2_0, const float* head2_1,
                        const float* dpart_flow0, const float* dpart_flow1, const float* dpart_lift0,
                        const float* dpart_lift1, int N, void* g1_head0, void* g1_head1, void* g1T_head0,
                        void* g1T_head1, int ldT, int colT0, float* dgamma_direct, float* da, float* red,
                        void* stream);
int links_geom_backward_angles(const float* ang0, const float* ang1, const float* eps_x, const float* stats,
                               const float* dgamma_direct, const float* red, int N, void* g_ang0,
                               void* g_ang1, void* gT_ang0, void* gT_ang1, int ldT, int colT0, int n_stat, void* stream);
/* Global elevation statistics under data parallelism (the reference's props.mean() / props.std() over the WHOLE batch,
 * train_leg_torso_lifter.py:168): links_elev_sums writes this rank's (sum gamma, sum gamma^2) as doubles; after a SUM
 * all-reduce links_elev_finalize turns them into stats = (mean, unbiased std) over n_total rows.  In backward, red[2] of
 * links_geom_backward is all-reduced as well and links_geom_backward_angles gets n_stat = n_total (<= 0: N). */
int links_elev_sums(const float* ang0, const float* ang1, int N, double* sums, void* stream);
int links_elev_finalize(const double* sums, int n_total, float* stats, void* stream);

/* ------------------------------------------------------------------------------------------
 * Normalising flow (FrEIA SequenceINN of 8 AllInOneBlock, permute_soft=True; call sites
 * train_leg_torso_lifter.py:134-136,207-214; train_full_pose_norm_flow.py:75-90)
 * ------------------------------------------------------------------------------------------ */
/* Packed per-flow parameters are produced by links_flow_pack from FrEIA-layout tensors. */
size_t links_flow_packed_floats(int C, int n_blocks);
int links_flow_pack(int C, int n_blocks, const float* const* w0, const float* const* b0,
                    const float* const* w2, const float* const* b2, const float* const* gscale,
                    const float* const* goffset, const float* const* wperm, const float* const* wperm_inv,
                    float* packed, void* stream);
/* Flow batches of >= 48 rows run on the tensor-core kernel (csrc/flow_tc.cuh, 128-row tiles, bf16x3 operands);
 * smaller ones on the 32-row fp32 SIMT kernel (csrc/flow.cuh).  links_flow_set_simt_only(1) pins the SIMT kernel
 * (A/B measurements and tests only); returns the previous setting. */
int links_flow_set_simt_only(int on);
/* z, log_jac_det = inn(x, rev) for x [M,C] (ld = C). */
int links_flow_apply(const float* packed, int C, int n_blocks, const float* x, int M, int rev,
                     float* out, float* log_jac_det, void* stream);
/* nll[m] = 0.5*|z|^2 - log_jac_det, nll_sum += sum_m nll, dx = scale * d(nll)/dx (frozen flow).
 * stash (may be NULL): scratch of links_flow_stash_floats(C, n_blocks, M) floats, 16-byte aligned.  With it the
 * tensor-core kernel keeps every block's input and subnet output from the forward pass instead of reconstructing them
 * with the inverse coupling in the backward pass (2 instead of 3 subnet evaluations per block; same results up to the
 * round-off of the inverse). */
size_t links_flow_stash_floats(int C, int n_blocks, int M);
int links_flow_nll_fwdbwd(const float* packed, int C, int n_blocks, const float* x, int M, float scale,
                          float* nll_sum, float* dx, float* stash, void* stream);
/* Training a flow (train_full_pose_norm_flow.py:75-98): NLL forward + backward as above, and in addition, per coupling
 * block k, the operands of the parameter-gradient GEMMs -- ex_x1[k] = subnet input x1 (bf16 [M,64], first c1 columns)
 * and ex_dsub[k] = d nll / d subnet output (bf16 [M,64], first 2*c2 columns; the caller keeps the padding zero) -- and
 * the row-reduced gradients of the global affine, d_gscale / d_goffset fp32 [n_blocks, C] (accumulated).  With these
 *   dW2 = ex_dsub^T . relu(x1 W1^T + b1),  db2 = colsum(ex_dsub),
 *   dW1 = ((ex_dsub . W2) * relu')^T . x1, db1 = colsum of that,
 * which the host issues as grouped GEMMs (links_gemm_grouped).  Tensor-core kernel for every M. */
int links_flow_nll_train(const float* packed, int C, int n_blocks, const float* x, int M, float scale,
                         float* nll_sum, float* dx, void* ex_x1, void* ex_dsub, float* d_gscale, float* d_goffset,
                         float* stash, void* stream);
/* Vector-Jacobian product of the forward map (z, log_jac_det) = inn(x):
 * dx = (dz/dx)^T gz + (d log_jac_det/dx)^T gld  (gld may be NULL = 0).  Backs autograd of the FrEIA shim. */
int links_flow_vjp(const float* packed, int C, int n_blocks, const float* x, int M, const float* gz,
                   const float* gld, float* dx, void* stream);
/* The same vector-Jacobian product with the parameter-gradient exports of links_flow_nll_train (ex_x1, ex_dsub, d_gscale,
 * d_goffset; same layouts, d_gscale / d_goffset accumulated) for arbitrary seeds (gz, gld): what autograd needs when a flow
 * is trained THROUGH the reference's call `z, jac = inn(x)` (train_full_pose_norm_flow.py:75-98,
 * train_leg_torso_left_right_norm_flow.py:108-166) instead of through the fused NLL step.  dx may be NULL.
 * Tensor-core kernel; C in {14, 20, 22, 32, 34}. */
int links_flow_vjp_train(const float* packed, int C, int n_blocks, const float* x, int M, const float* gz,
                         const float* gld, float* dx, void* ex_x1, void* ex_dsub, float* d_gscale, float* d_goffset,
                         float* stash, void* stream);
/* Sampling block (train_leg_torso_lifter.py:133-142): out = [x ; s], s = inn^-1(z + 0.2*noise*z) with
 * the root joint zeroed; C must be 34.  out is [2M,34]. */
int links_flow_sample(const float* packed, int n_blocks, const float* x, const float* noise, int M,
                      float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input pipeline: normalize_head / normalize_head_test (utils/helpers.py:198-207, 222-230) fused with the dataset
 * classes' transpose-flatten (utils/h36m_dataset_class.py:25-27).
 *   raw  [n, 17, 2] key-points (transposed_input = 0) or already flattened [n, 34] = (17 x, 17 y) rows (= 1), fp32
 *   out  [n, 34]: root-centred, divided by the mean over ALL n poses of |joint 0 - joint 10| (fixed_scale <= 0; the mean
 *        is accumulated in *dist_sum, a device double) or by fixed_scale (> 0), times 1/10.
 * ------------------------------------------------------------------------------------------ */
int links_normalize_head(const float* raw, int n, int root_joint, int transposed_input, float fixed_scale, float* out,
                         double* dist_sum, void* stream);

/* ------------------------------------------------------------------------------------------
 * Metrics (utils/metrics_batch.py:8-159, utils/metrics.py:35-171)
 * per_pose / per_pose_max [M] and dist [M, num_joints] (each may be null) receive per-pose mean / max joint
 * distance and all joint distances; sum (may be null) accumulates sum_m per_pose[m] in double.
 * ------------------------------------------------------------------------------------------ */
int links_mpjpe(const float* p_ref, const float* p, int M, int num_joints, int root_joint, int use_scaling,
                float* per_pose, float* per_pose_max, float* dist, double* sum, void* stream);
/* counts[k] += #(values[i] < thresholds[k]) (strict != 0) or #(values[i] <= thresholds[k]); thresholds ascending,
 * n_thresh <= 512.  With dist / per_pose_max from links_mpjpe this gives PCK, AUC and CPS
 * (utils/metrics_batch.py:26-102) as exact integer counts. */
int links_threshold_counts(const float* values, size_t n, const float* thresholds, int n_thresh, int strict,
                           unsigned long long* counts, void* stream);
/* mode 0: metrics_batch.pmpjpe semantics (RMS scale match, R = diag(1,1,det)UV^T);
 * mode 1: metrics.pmpjpe(reflection='best') semantics (optimal scale, reflection allowed). */
int links_pmpjpe(const float* p_ref, const float* p, int M, int num_joints, int mode, float* per_pose,
                 float* aligned /* [M, 3*num_joints] Procrustes-aligned p, may be NULL */, double* sum, void* stream);
/* Eval fusion (eval_h36m.py:58-78): lift 2D poses with combined depths, score against GT. */
int links_eval_lift_score(const float* poses_2d, const float* depth_off, int ld_depth, const float* gt_3d,
                          int M, float depth, double* sums3, void* stream);

/* ------------------------------------------------------------------------------------------
 * Occlusion step pieces (train_occlusion_models.py:164-217)
 * ------------------------------------------------------------------------------------------ */
/* pose[M,51] = root-centred lift of x[M,34] with depth heads (no clamp), :164-174 */
int links_occ_lift(const float* x, const float* head_leg, const float* head_torso, int M, float depth,
                   float* pose, void* stream);
/* pose_out = Ry((u-0.5)*1.99*pi) @ pose, :213-217 */
int links_occ_rotate_y(const float* pose, const float* u, int M, float* pose_out, void* stream);
/* loss_sum += sum_m sum_c (pred[m,c]-pose_flat[m*51+tidx[c]])^2 ; g (bf16 [M,64] + T) = scale*2*(pred-target) */
int links_occ_mse(const float* pred, int ld_pred, const float* pose, const int* tidx_dev, int n_out, int M,
                  float scale, float* loss_sum, void* g_bf16, void* gT_bf16, int ldT, int colT0, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LINKS_B200_H */
