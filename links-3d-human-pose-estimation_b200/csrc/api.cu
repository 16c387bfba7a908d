// extern "C" launch wrappers (see include/links_b200.h) for the non-tensor-core kernels.
#include "common.cuh"
#include "elementwise.cuh"
#include "metrics.cuh"
#include "geom.cuh"
#include "flow.cuh"
#include "flow_tc.cuh"
#include "occ.cuh"
#include "dataprep.cuh"
#include <math.h>

using namespace links;

unsigned long long g_links_kernel_launches = 0;

extern "C" __attribute__((visibility("default"))) int links_abi_version(void) { return LINKS_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) size_t links_launch_count(void) { return static_cast<size_t>(g_links_kernel_launches); }

extern "C" __attribute__((visibility("default"))) int links_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default"))) int links_pack_rows(const float* src, int ld_src, int M, const int* idx_dev, int n_idx, int period,
                               void* dst_bf16, void* dstT_bf16, int ldT, int colT0, void* stream) {
  LINKS_CHECK_PTR(src); LINKS_CHECK_PTR(idx_dev); LINKS_CHECK_PTR(dst_bf16);
  if (M < 1 || n_idx < 1 || n_idx > 64 || period < 1 || (M % period) != 0) return LINKS_E_RANGE;
  const long long total = static_cast<long long>(M) * 64;
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads);
  pack_rows_kernel<<<blocks, threads, 0, links_stream(stream)>>>(
      src, ld_src, M, idx_dev, n_idx, period, static_cast<__nv_bfloat16*>(dst_bf16),
      static_cast<__nv_bfloat16*>(dstT_bf16), ldT, colT0);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_colsum_bf16(const void* G, int ldg, int M, int N, float* out, int accumulate, void* stream) {
  LINKS_CHECK_PTR(G); LINKS_CHECK_PTR(out);
  if (M < 1 || N < 1) return LINKS_E_RANGE;
  cudaStream_t s = links_stream(stream);
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * N, s);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const int rows_per_block = 256;
  dim3 grid((N + 31) / 32, (M + rows_per_block - 1) / rows_per_block);
  colsum_bf16_kernel<<<grid, dim3(32, 8), 0, s>>>(static_cast<const __nv_bfloat16*>(G), ldg, M, N, out, rows_per_block);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_colsum_bf16_batched(const LinksColsumItem* items, int n_items, void* stream) {
  LINKS_CHECK_PTR(items);
  if (n_items < 1 || n_items > LINKS_MAX_COLSUM_ITEMS) return LINKS_E_RANGE;
  ColsumBatch B;
  memset(&B, 0, sizeof(B));
  int maxN = 0, maxM = 0;
  for (int i = 0; i < n_items; ++i) {
    if (!items[i].G || !items[i].out || items[i].M < 1 || items[i].N < 1) return LINKS_E_ARG;
    B.it[i] = items[i];
    if (items[i].N > maxN) maxN = items[i].N;
    if (items[i].M > maxM) maxM = items[i].M;
  }
  B.n = n_items;
  cudaStream_t s = links_stream(stream);
  colsum_batched_zero_kernel<<<n_items, 256, 0, s>>>(B);
  const int rows_per_block = 256;
  dim3 grid((maxN + 31) / 32, n_items, (maxM + rows_per_block - 1) / rows_per_block);
  colsum_batched_kernel<<<grid, dim3(32, 8), 0, s>>>(B, rows_per_block);
  return links_launch_status(2);
}

extern "C" __attribute__((visibility("default"))) int links_cast_weight(const float* W, int N, int K, void* W_bf16, int ldw, void* WT_bf16, int ldwt,
                                 void* stream) {
  LINKS_CHECK_PTR(W);
  if (N < 1 || K < 1) return LINKS_E_RANGE;
  if (W_bf16 && ldw < K) return LINKS_E_RANGE;
  if (WT_bf16 && ldwt < N) return LINKS_E_RANGE;
  const int kx = ((W_bf16 ? (ldw > K ? ldw : K) : K) + 31) / 32;
  const int ny = ((WT_bf16 ? (ldwt > N ? ldwt : N) : N) + 31) / 32;
  cast_weight_kernel<<<dim3(kx, ny), dim3(32, 8), 0, links_stream(stream)>>>(
      W, N, K, static_cast<__nv_bfloat16*>(W_bf16), ldw, static_cast<__nv_bfloat16*>(WT_bf16), ldwt);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_cast_weight_batched(const LinksCastItem* items, int n_items, void* stream) {
  LINKS_CHECK_PTR(items);
  if (n_items < 1 || n_items > LINKS_MAX_CAST_ITEMS) return LINKS_E_RANGE;
  CastBatch B;
  memset(&B, 0, sizeof(B));
  for (int i = 0; i < n_items; ++i) {
    if (!items[i].W || !items[i].Wb || items[i].N < 1 || items[i].K < 1 || items[i].ldw < items[i].K) return LINKS_E_ARG;
    B.it[i] = items[i];
  }
  B.n = n_items;
  cast_weight_batched_kernel<<<dim3(64, n_items), 256, 0, links_stream(stream)>>>(B);
  return links_launch_status();
}

template <typename GradT>
static int adam_launch(float* param, const GradT* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, int* step_dev, float grad_scale, const float* lr_dev,
                       void* stream) {
  LINKS_CHECK_PTR(param); LINKS_CHECK_PTR(grad); LINKS_CHECK_PTR(exp_avg); LINKS_CHECK_PTR(exp_avg_sq);
  if (n == 0 || (step_dev == nullptr && step < 1)) return LINKS_E_RANGE;
  const int threads = 256;
  size_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  adam_kernel<GradT><<<static_cast<int>(blocks), threads, 0, links_stream(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_dev, step, grad_scale, lr_dev);
  if (step_dev != nullptr && step >= 0) adam_incr_kernel<<<1, 32, 0, links_stream(stream)>>>(step_dev);
  return links_launch_status(step_dev != nullptr && step >= 0 ? 2 : 1);
}

extern "C" __attribute__((visibility("default"))) int links_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int step, int* step_dev,
                               float grad_scale, const float* lr_dev, void* stream) {
  return adam_launch<float>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, step_dev,
                            grad_scale, lr_dev, stream);
}

extern "C" __attribute__((visibility("default"))) int links_adam_step_g16(float* param, const void* grad_bf16, float* exp_avg, float* exp_avg_sq, size_t n,
                                   float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                                   int* step_dev, float grad_scale, const float* lr_dev, void* stream) {
  return adam_launch<__nv_bfloat16>(param, static_cast<const __nv_bfloat16*>(grad_bf16), exp_avg, exp_avg_sq, n, lr, beta1,
                                    beta2, eps, weight_decay, step, step_dev, grad_scale, lr_dev, stream);
}

extern "C" __attribute__((visibility("default"))) int links_normalize_head(const float* raw, int n, int root_joint, int transposed_input, float fixed_scale,
                                    float* out, double* dist_sum, void* stream) {
  LINKS_CHECK_PTR(raw); LINKS_CHECK_PTR(out);
  if (n < 1 || root_joint < 0 || root_joint > 16) return LINKS_E_RANGE;
  if (fixed_scale <= 0.f && dist_sum == nullptr) return LINKS_E_ARG;
  cudaStream_t s = links_stream(stream);
  if (dist_sum != nullptr) {
    cudaError_t e = cudaMemsetAsync(dist_sum, 0, sizeof(double), s);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  normalize_head_center_kernel<<<(n + 255) / 256, 256, 0, s>>>(raw, n, root_joint, transposed_input, out, dist_sum);
  const size_t total = static_cast<size_t>(n) * 34;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  normalize_head_scale_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(out, total, n, fixed_scale, dist_sum);
  return links_launch_status(2);
}

extern "C" __attribute__((visibility("default"))) int links_small_matvec(const float* mat, const float* in, int n_in, int n_out, float* out, void* stream) {
  LINKS_CHECK_PTR(mat); LINKS_CHECK_PTR(in); LINKS_CHECK_PTR(out);
  if (n_in < 1 || n_out < 1 || n_in > 4096 || n_out > 4096) return LINKS_E_RANGE;
  small_matvec_kernel<<<(n_out + 63) / 64, 64, 0, links_stream(stream)>>>(mat, in, n_in, n_out, out);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_adam_prepare(const int* step_dev, const float* lr_dev, float lr, float beta1, float beta2,
                                  float eps, float weight_decay, float grad_scale, float* hyper, void* stream) {
  LINKS_CHECK_PTR(step_dev); LINKS_CHECK_PTR(hyper); LINKS_CHECK_ALIGN16(hyper);
  adam_prepare_kernel<<<1, 32, 0, links_stream(stream)>>>(step_dev, lr_dev, lr, beta1, beta2, eps, weight_decay, grad_scale, hyper);
  return links_launch_status();
}

// Device-side barrier over the ranks of one node through peer-mapped flag words (flags[r]: rank r's flag array, >= world
// uint32 zero-initialised words per slot): rank `rank` sets word [slot * 8 + rank] in every rank's array and consumes the
// words [slot * 8 + r] of its own.  System-scope fences on both sides: peer stores issued before the barrier (by earlier
// kernels of the stream) are visible to every rank after it.  A plain kernel launch: CUDA-graph capturable.
struct PeerFlags { unsigned int* f[LINKS_MAX_PUSH_RANKS]; };
__global__ void peer_barrier_kernel(PeerFlags P, int world, int rank, int slot) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();
  unsigned int* remote = P.f[t] + slot * LINKS_MAX_PUSH_RANKS + rank;
  long long t0 = clock64();
  while (atomicCAS_system(remote, 0u, 1u) != 0u) {
    if (clock64() - t0 > 20000000000LL) __trap();       // ~10 s: a rank never arrived
  }
  unsigned int* local = P.f[rank] + slot * LINKS_MAX_PUSH_RANKS + t;
  t0 = clock64();
  while (atomicCAS_system(local, 1u, 0u) != 1u) {
    if (clock64() - t0 > 20000000000LL) __trap();
  }
  __threadfence_system();
}

extern "C" __attribute__((visibility("default"))) int links_peer_barrier(void* const* flag_ptrs, int world, int rank, int slot, void* stream) {
  LINKS_CHECK_PTR(flag_ptrs);
  if (world < 1 || world > LINKS_MAX_PUSH_RANKS || rank < 0 || rank >= world || slot < 0 || slot > 7) return LINKS_E_RANGE;
  PeerFlags P;
  memset(&P, 0, sizeof(P));
  for (int r = 0; r < world; ++r) {
    if (flag_ptrs[r] == nullptr) return LINKS_E_ARG;
    P.f[r] = static_cast<unsigned int*>(flag_ptrs[r]);
  }
  peer_barrier_kernel<<<1, 32, 0, links_stream(stream)>>>(P, world, rank, slot);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_adam_zero(float* p, float* m, float* v, const void* stage, size_t stage_slot_elems,
                               const LinksAdamZeroLayer* layers_dev, int n_layers, int rows_per_owner, int cols, int world, int rank,
                               const float* hyper, void* stream) {
  LINKS_CHECK_PTR(p); LINKS_CHECK_PTR(m); LINKS_CHECK_PTR(v); LINKS_CHECK_PTR(stage); LINKS_CHECK_PTR(layers_dev); LINKS_CHECK_PTR(hyper);
  LINKS_CHECK_ALIGN16(stage);
  if (n_layers < 1 || rows_per_owner < 1 || cols < 8 || (cols & 7) || world < 1 || world > LINKS_MAX_PUSH_RANKS || rank < 0 ||
      rank >= world || (stage_slot_elems & 7))
    return LINKS_E_RANGE;
  const size_t n = static_cast<size_t>(rows_per_owner) * cols;
  size_t bx = (n / 8 + 255) / 256;
  if (bx > 64) bx = 64;
  adam_zero_kernel<<<dim3(static_cast<unsigned>(bx), n_layers), 256, 0, links_stream(stream)>>>(
      p, m, v, static_cast<const __nv_bfloat16*>(stage), stage_slot_elems, layers_dev, rows_per_owner, cols, world, rank, hyper);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_grad_compress_bf16(const float* grad, void* grad_bf16, size_t n, void* stream) {
  LINKS_CHECK_PTR(grad); LINKS_CHECK_PTR(grad_bf16);
  if (n == 0) return LINKS_E_RANGE;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  grad_compress_bf16_kernel<<<static_cast<int>(blocks), 256, 0, links_stream(stream)>>>(
      grad, static_cast<__nv_bfloat16*>(grad_bf16), n);
  return links_launch_status();
}

// ---------------------------------------------------------------------------------------------
// Launch a staged geometry kernel: plan the shared-memory staging, raise the kernel's dynamic shared-memory limit when
// this launch needs more than any earlier one, launch.
template <class Kernel>
static int geom_launch(Kernel kernel, GeomArgs& A, int level, int* smem_limit, void* stream) {
  int rc = geom_plan(A, level);
  if (rc) return rc;
  const size_t smem = geom_smem_bytes(A.st);
  if (smem > 200 * 1024) return LINKS_E_RANGE;
  if (static_cast<int>(smem) > *smem_limit) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    *smem_limit = static_cast<int>(smem);
  }
  // One resident wave: every block builds its tables and zero-fills its staging buffers once, then walks its rows with the
  // grid-stride loop (a grid of several waves repeats that set-up: ~5 % of the run time at 4 M rows).
  struct Occ { const void* k; size_t smem; int per_sm, n_sms; };
  static Occ cache[16];
  static int n_cache = 0;
  int per_sm = 0, n_sms = 148;
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < n_cache; ++i)
    if (cache[i].k == key && cache[i].smem == smem) { per_sm = cache[i].per_sm; n_sms = cache[i].n_sms; }
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kGeomWarps * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
    if (n_cache < 16) cache[n_cache++] = Occ{key, smem, per_sm, n_sms};
  }
  const int iters = (A.N + A.st.rows - 1) / A.st.rows;
  const int need = (iters + kGeomWarps - 1) / kGeomWarps;
  const int grid = need < n_sms * per_sm ? need : n_sms * per_sm;
  kernel<<<grid, kGeomWarps * 32, smem, links_stream(stream)>>>(A);
  return links_launch_status();
}

static int check_maps(const LinksGeomMaps* m) {
  if (!m) return LINKS_E_ARG;
  if (m->V < 1 || m->V > 2) return LINKS_E_RANGE;
  for (int p = 0; p < 2; ++p) if (m->n_joints[p] < 1 || m->n_joints[p] > 16) return LINKS_E_RANGE;
  for (int v = 0; v < m->V; ++v)
    for (int j = 0; j < 17; ++j) {
      if (m->src_net[v][j] < 0 || m->src_net[v][j] > 1) return LINKS_E_RANGE;
      if (m->part_net[v][j] < -1 || m->part_net[v][j] > 1) return LINKS_E_RANGE;
      if (m->col[j] < 0 || m->col[j] >= LINKS_HEAD_LD) return LINKS_E_RANGE;
    }
  return 0;
}

extern "C" __attribute__((visibility("default"))) int links_elev_stats(const float* ang0, const float* ang1, int N, float* stats, void* stream) {
  LINKS_CHECK_PTR(ang0); LINKS_CHECK_PTR(ang1); LINKS_CHECK_PTR(stats);
  if (N < 2) return LINKS_E_RANGE;
  elev_stats_kernel<<<1, 1024, 0, links_stream(stream)>>>(ang0, ang1, N, stats);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_geom_forward(const LinksGeomMaps* maps, const float* u, const float* head0, const float* head1,
                                  const float* ang0, const float* ang1, const float* eps_x, const float* u_y,
                                  const float* stats, int N, float* qpart0, float* qpart1, float* q_full0,
                                  float* q_full1, void* stream) {
  int rc = check_maps(maps);
  if (rc) return rc;
  LINKS_CHECK_PTR(u); LINKS_CHECK_PTR(head0); LINKS_CHECK_PTR(head1); LINKS_CHECK_PTR(ang0); LINKS_CHECK_PTR(ang1);
  LINKS_CHECK_PTR(eps_x); LINKS_CHECK_PTR(u_y); LINKS_CHECK_PTR(stats); LINKS_CHECK_PTR(qpart0); LINKS_CHECK_PTR(qpart1);
  if (N < 1 || N > kGeomMaxRows) return LINKS_E_RANGE;
  GeomArgs A;
  memset(&A, 0, sizeof(A));
  A.maps = *maps;
  A.u = u; A.head[0] = head0; A.head[1] = head1; A.ang[0] = ang0; A.ang[1] = ang1;
  A.eps_x = eps_x; A.u_y = u_y; A.stats = stats; A.N = N;
  A.qpart[0] = qpart0; A.qpart[1] = qpart1; A.qfull[0] = q_full0; A.qfull[1] = q_full1;
  static int lim[4] = {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024};
  const bool full = q_full0 != nullptr || q_full1 != nullptr;
  if (maps->V == 1) {
    if (full) return geom_launch(geom_forward_kernel<1, true>, A, 0, &lim[0], stream);
    return geom_launch(geom_forward_kernel<1, false>, A, 0, &lim[1], stream);
  }
  if (full) return geom_launch(geom_forward_kernel<2, true>, A, 0, &lim[2], stream);
  return geom_launch(geom_forward_kernel<2, false>, A, 0, &lim[3], stream);
}

extern "C" __attribute__((visibility("default"))) int links_geom_loss(const LinksGeomMaps* maps, const float* u, const float* head0, const float* head1,
                               const float* ang0, const float* ang1, const float* eps_x, const float* u_y,
                               const float* stats, const float* head2_0, const float* head2_1, int N,
                               float* loss_sums, void* g2_head0, void* g2_head1, void* g2T_head0, void* g2T_head1,
                               int ldT, int colT0, void* stream) {
  int rc = check_maps(maps);
  if (rc) return rc;
  LINKS_CHECK_PTR(u); LINKS_CHECK_PTR(head0); LINKS_CHECK_PTR(head1); LINKS_CHECK_PTR(ang0); LINKS_CHECK_PTR(ang1);
  LINKS_CHECK_PTR(eps_x); LINKS_CHECK_PTR(u_y); LINKS_CHECK_PTR(stats); LINKS_CHECK_PTR(head2_0); LINKS_CHECK_PTR(head2_1);
  LINKS_CHECK_PTR(loss_sums); LINKS_CHECK_PTR(g2_head0); LINKS_CHECK_PTR(g2_head1);
  if (N < 1 || N > kGeomMaxRows) return LINKS_E_RANGE;
  GeomArgs A;
  memset(&A, 0, sizeof(A));
  A.maps = *maps;
  A.u = u; A.head[0] = head0; A.head[1] = head1; A.ang[0] = ang0; A.ang[1] = ang1;
  A.eps_x = eps_x; A.u_y = u_y; A.stats = stats; A.N = N;
  A.head2[0] = head2_0; A.head2[1] = head2_1;
  A.loss_sums = loss_sums;
  A.g2[0] = static_cast<__nv_bfloat16*>(g2_head0); A.g2[1] = static_cast<__nv_bfloat16*>(g2_head1);
  A.g2T[0] = static_cast<__nv_bfloat16*>(g2T_head0); A.g2T[1] = static_cast<__nv_bfloat16*>(g2T_head1);
  A.ldT = ldT; A.colT0 = colT0;
  const bool tr = g2T_head0 != nullptr || g2T_head1 != nullptr;
  static int lim[4] = {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024};
  if (maps->V == 1) {
    if (tr) return geom_launch(geom_lossgrad_kernel<false, 1, true>, A, 1, &lim[0], stream);
    return geom_launch(geom_lossgrad_kernel<false, 1, false>, A, 1, &lim[1], stream);
  }
  if (tr) return geom_launch(geom_lossgrad_kernel<false, 2, true>, A, 1, &lim[2], stream);
  return geom_launch(geom_lossgrad_kernel<false, 2, false>, A, 1, &lim[3], stream);
}

extern "C" __attribute__((visibility("default"))) int links_geom_backward(const LinksGeomMaps* maps, const float* u, const float* head0, const float* head1,
                                   const float* ang0, const float* ang1, const float* eps_x, const float* u_y,
                                   const float* stats, const float* head2_0, const float* head2_1,
                                   const float* dpart_flow0, const float* dpart_flow1, const float* dpart_lift0,
                                   const float* dpart_lift1, int N, void* g1_head0, void* g1_head1, void* g1T_head0,
                                   void* g1T_head1, int ldT, int colT0, float* dgamma_direct, float* da, float* red,
                                   void* stream) {
  int rc = check_maps(maps);
  if (rc) return rc;
  LINKS_CHECK_PTR(u); LINKS_CHECK_PTR(head0); LINKS_CHECK_PTR(head1); LINKS_CHECK_PTR(ang0); LINKS_CHECK_PTR(ang1);
  LINKS_CHECK_PTR(eps_x); LINKS_CHECK_PTR(u_y); LINKS_CHECK_PTR(stats); LINKS_CHECK_PTR(head2_0); LINKS_CHECK_PTR(head2_1);
  LINKS_CHECK_PTR(dpart_flow0); LINKS_CHECK_PTR(dpart_flow1); LINKS_CHECK_PTR(dpart_lift0); LINKS_CHECK_PTR(dpart_lift1);
  LINKS_CHECK_PTR(g1_head0); LINKS_CHECK_PTR(g1_head1); LINKS_CHECK_PTR(dgamma_direct); LINKS_CHECK_PTR(da); LINKS_CHECK_PTR(red);
  if (N < 1 || N > kGeomMaxRows) return LINKS_E_RANGE;
  GeomArgs A;
  memset(&A, 0, sizeof(A));
  A.maps = *maps;
  A.u = u; A.head[0] = head0; A.head[1] = head1; A.ang[0] = ang0; A.ang[1] = ang1;
  A.eps_x = eps_x; A.u_y = u_y; A.stats = stats; A.N = N;
  A.head2[0] = head2_0; A.head2[1] = head2_1;
  A.dflow[0] = dpart_flow0; A.dflow[1] = dpart_flow1; A.dlift[0] = dpart_lift0; A.dlift[1] = dpart_lift1;
  A.g1[0] = static_cast<__nv_bfloat16*>(g1_head0); A.g1[1] = static_cast<__nv_bfloat16*>(g1_head1);
  A.g1T[0] = static_cast<__nv_bfloat16*>(g1T_head0); A.g1T[1] = static_cast<__nv_bfloat16*>(g1T_head1);
  A.ldT = ldT; A.colT0 = colT0;
  A.dgamma = dgamma_direct; A.da = da; A.red = red;
  const bool tr = g1T_head0 != nullptr || g1T_head1 != nullptr;
  static int lim[4] = {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024};
  if (maps->V == 1) {
    if (tr) return geom_launch(geom_lossgrad_kernel<true, 1, true>, A, 2, &lim[0], stream);
    return geom_launch(geom_lossgrad_kernel<true, 1, false>, A, 2, &lim[1], stream);
  }
  if (tr) return geom_launch(geom_lossgrad_kernel<true, 2, true>, A, 2, &lim[2], stream);
  return geom_launch(geom_lossgrad_kernel<true, 2, false>, A, 2, &lim[3], stream);
}

extern "C" __attribute__((visibility("default"))) int links_geom_backward_angles(const float* ang0, const float* ang1, const float* eps_x, const float* stats,
                                          const float* dgamma_direct, const float* red, int N, void* g_ang0,
                                          void* g_ang1, void* gT_ang0, void* gT_ang1, int ldT, int colT0, int n_stat,
                                          void* stream) {
  (void)eps_x;
  if (n_stat <= 0) n_stat = N;
  LINKS_CHECK_PTR(ang0); LINKS_CHECK_PTR(ang1); LINKS_CHECK_PTR(stats); LINKS_CHECK_PTR(dgamma_direct);
  LINKS_CHECK_PTR(red); LINKS_CHECK_PTR(g_ang0); LINKS_CHECK_PTR(g_ang1);
  if (N < 2) return LINKS_E_RANGE;
  geom_backward_angles_kernel<<<(N + 255) / 256, 256, 0, links_stream(stream)>>>(
      ang0, ang1, stats, dgamma_direct, red, N, static_cast<__nv_bfloat16*>(g_ang0),
      static_cast<__nv_bfloat16*>(g_ang1), static_cast<__nv_bfloat16*>(gT_ang0), static_cast<__nv_bfloat16*>(gT_ang1),
      ldT, colT0, n_stat);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_elev_sums(const float* ang0, const float* ang1, int N, double* sums, void* stream) {
  LINKS_CHECK_PTR(ang0); LINKS_CHECK_PTR(ang1); LINKS_CHECK_PTR(sums);
  if (N < 1) return LINKS_E_RANGE;
  elev_sums_kernel<<<1, 1024, 0, links_stream(stream)>>>(ang0, ang1, N, sums);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_elev_finalize(const double* sums, int n_total, float* stats, void* stream) {
  LINKS_CHECK_PTR(sums); LINKS_CHECK_PTR(stats);
  if (n_total < 2) return LINKS_E_RANGE;
  elev_finalize_kernel<<<1, 32, 0, links_stream(stream)>>>(sums, static_cast<double>(n_total), stats);
  return links_launch_status();
}

// ---------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default"))) size_t links_flow_packed_floats(int C, int n_blocks) {
  if (C < 2 || C > 34 || n_blocks < 1 || n_blocks > kFlowMaxBlocks) return 0;
  // [SIMT records | tensor-core operand images]
  return static_cast<size_t>(flow_block_floats(C)) * n_blocks + tc_block_bytes(C) / 4 * n_blocks;
}

extern "C" __attribute__((visibility("default"))) int links_flow_pack(int C, int n_blocks, const float* const* w0, const float* const* b0,
                               const float* const* w2, const float* const* b2, const float* const* gscale,
                               const float* const* goffset, const float* const* wperm, const float* const* wperm_inv,
                               float* packed, void* stream) {
  if (C < 2 || C > 34 || n_blocks < 1 || n_blocks > kFlowMaxBlocks) return LINKS_E_RANGE;
  LINKS_CHECK_PTR(packed); LINKS_CHECK_ALIGN16(packed);
  FlowPackArgs A;
  memset(&A, 0, sizeof(A));
  for (int k = 0; k < n_blocks; ++k) {
    if (!w0[k] || !b0[k] || !w2[k] || !b2[k] || !gscale[k] || !goffset[k] || !wperm[k] || !wperm_inv[k]) return LINKS_E_ARG;
    A.w0[k] = w0[k]; A.b0[k] = b0[k]; A.w2[k] = w2[k]; A.b2[k] = b2[k];
    A.gs[k] = gscale[k]; A.go[k] = goffset[k]; A.wp[k] = wperm[k]; A.wpi[k] = wperm_inv[k];
  }
  A.packed = packed; A.C = C; A.n_blocks = n_blocks;
  const dim3 pack_grid(n_blocks, kFlowPackSplit);
  flow_pack_kernel<<<pack_grid, 256, 0, links_stream(stream)>>>(A);
  flow_tc_pack_kernel<<<pack_grid, 256, 0, links_stream(stream)>>>(
      A, reinterpret_cast<unsigned char*>(packed + static_cast<size_t>(flow_block_floats(C)) * n_blocks));
  return links_launch_status();
}

template <int C, int MODE>
static int launch_flow_c(const FlowArgs& A, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(flow_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(FlowSmem<C>::bytes));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  const int blocks = (A.M + kFlowRows - 1) / kFlowRows;
  flow_kernel<C, MODE><<<blocks, kFlowWarps * 32, FlowSmem<C>::bytes, s>>>(A);
  return links_launch_status();
}

// Tensor-core path (128-row tiles) for M >= kFlowTcMinRows; small batches stay on the 32-row SIMT kernel.
constexpr int kFlowTcMinRows = 48;
static bool g_flow_force_simt = false;

template <int C, int MODE>
static int launch_flow_tc_c(const FlowArgs& A, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(flow_tc_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kTcSmemBytes));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  FlowTcArgs T;
  T.packed = reinterpret_cast<const unsigned char*>(A.packed + static_cast<size_t>(flow_block_floats(C)) * A.n_blocks);
  T.x = A.x; T.noise = A.noise; T.out = A.out; T.ld = A.ld; T.nll_sum = A.nll_sum; T.scale = A.scale;
  T.gz = A.gz; T.gld = A.gld; T.M = A.M; T.n_blocks = A.n_blocks;
  T.ex_x1 = A.ex_x1; T.ex_dsub = A.ex_dsub; T.d_gscale = A.d_gscale; T.d_goffset = A.d_goffset;
  T.stash = A.stash;
  const int blocks = (A.M + kTcRows - 1) / kTcRows;
  flow_tc_kernel<C, MODE><<<blocks, kTcThreads, kTcSmemBytes, s>>>(T);
  return links_launch_status();
}

template <int C, int MODE>
static int launch_flow_any(const FlowArgs& A, cudaStream_t s) {
  if (A.M >= kFlowTcMinRows && !g_flow_force_simt) return launch_flow_tc_c<C, MODE>(A, s);
  return launch_flow_c<C, MODE>(A, s);
}

template <int MODE>
static int launch_flow(int C, const FlowArgs& A, cudaStream_t s) {
  switch (C) {
    case 14: return launch_flow_any<14, MODE>(A, s);
    case 20: return launch_flow_any<20, MODE>(A, s);
    case 22: return launch_flow_any<22, MODE>(A, s);
    case 32: return launch_flow_any<32, MODE>(A, s);
    case 34: return launch_flow_any<34, MODE>(A, s);
    default: return LINKS_E_RANGE;
  }
}

extern "C" __attribute__((visibility("default"))) int links_flow_set_simt_only(int on) {
  const int prev = g_flow_force_simt ? 1 : 0;
  g_flow_force_simt = on != 0;
  return prev;
}

extern "C" __attribute__((visibility("default"))) int links_flow_apply(const float* packed, int C, int n_blocks, const float* x, int M, int rev, float* out,
                                float* log_jac_det, void* stream) {
  LINKS_CHECK_PTR(packed); LINKS_CHECK_PTR(x); LINKS_CHECK_PTR(out); LINKS_CHECK_ALIGN16(packed);
  if (M < 1 || n_blocks < 1 || n_blocks > kFlowMaxBlocks) return LINKS_E_RANGE;
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.out = out; A.ld = log_jac_det; A.M = M; A.n_blocks = n_blocks;
  return rev ? launch_flow<FLOW_REV>(C, A, links_stream(stream)) : launch_flow<FLOW_FWD>(C, A, links_stream(stream));
}

extern "C" __attribute__((visibility("default"))) size_t links_flow_stash_floats(int C, int n_blocks, int M) {
  if (C < 2 || n_blocks < 1 || M < 1) return 0;
  return static_cast<size_t>(M) * static_cast<size_t>(n_blocks) * static_cast<size_t>(C + 2 * flow_c2(C));
}

extern "C" __attribute__((visibility("default"))) int links_flow_nll_fwdbwd(const float* packed, int C, int n_blocks, const float* x, int M, float scale,
                                     float* nll_sum, float* dx, float* stash, void* stream) {
  LINKS_CHECK_PTR(packed); LINKS_CHECK_PTR(x); LINKS_CHECK_ALIGN16(packed);
  if (M < 1 || n_blocks < 1 || n_blocks > kFlowMaxBlocks) return LINKS_E_RANGE;
  if (stash != nullptr) LINKS_CHECK_ALIGN16(stash);
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.out = dx; A.nll_sum = nll_sum; A.scale = scale; A.M = M; A.n_blocks = n_blocks;
  A.stash = stash;
  return launch_flow<FLOW_NLL_FWDBWD>(C, A, links_stream(stream));
}

template <int MODE>
static int launch_flow_tc_only(int C, const FlowArgs& A, cudaStream_t s) {
  switch (C) {
    case 14: return launch_flow_tc_c<14, MODE>(A, s);
    case 20: return launch_flow_tc_c<20, MODE>(A, s);
    case 22: return launch_flow_tc_c<22, MODE>(A, s);
    case 32: return launch_flow_tc_c<32, MODE>(A, s);
    case 34: return launch_flow_tc_c<34, MODE>(A, s);
    default: return LINKS_E_RANGE;
  }
}

extern "C" __attribute__((visibility("default"))) int links_flow_nll_train(const float* packed, int C, int n_blocks, const float* x, int M, float scale,
                                    float* nll_sum, float* dx, void* ex_x1, void* ex_dsub, float* d_gscale,
                                    float* d_goffset, float* stash, void* stream) {
  LINKS_CHECK_PTR(packed); LINKS_CHECK_PTR(x); LINKS_CHECK_PTR(ex_x1); LINKS_CHECK_PTR(ex_dsub);
  LINKS_CHECK_PTR(d_gscale); LINKS_CHECK_PTR(d_goffset); LINKS_CHECK_ALIGN16(packed);
  if (M < 1 || n_blocks < 1 || n_blocks > kFlowMaxBlocks) return LINKS_E_RANGE;
  if (stash != nullptr) LINKS_CHECK_ALIGN16(stash);
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.out = dx; A.nll_sum = nll_sum; A.scale = scale; A.M = M; A.n_blocks = n_blocks;
  A.ex_x1 = ex_x1; A.ex_dsub = ex_dsub; A.d_gscale = d_gscale; A.d_goffset = d_goffset;
  A.stash = stash;
  return launch_flow_tc_only<FLOW_NLL_FWDBWD>(C, A, links_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int links_flow_vjp(const float* packed, int C, int n_blocks, const float* x, int M, const float* gz,
                              const float* gld, float* dx, void* stream) {
  LINKS_CHECK_PTR(packed); LINKS_CHECK_PTR(x); LINKS_CHECK_PTR(gz); LINKS_CHECK_PTR(dx); LINKS_CHECK_ALIGN16(packed);
  if (M < 1 || n_blocks < 1 || n_blocks > kFlowMaxBlocks) return LINKS_E_RANGE;
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.out = dx; A.gz = gz; A.gld = gld; A.M = M; A.n_blocks = n_blocks;
  return launch_flow<FLOW_NLL_FWDBWD>(C, A, links_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int links_flow_vjp_train(const float* packed, int C, int n_blocks, const float* x, int M, const float* gz,
                                    const float* gld, float* dx, void* ex_x1, void* ex_dsub, float* d_gscale,
                                    float* d_goffset, float* stash, void* stream) {
  LINKS_CHECK_PTR(packed); LINKS_CHECK_PTR(x); LINKS_CHECK_PTR(gz); LINKS_CHECK_PTR(ex_x1); LINKS_CHECK_PTR(ex_dsub);
  LINKS_CHECK_PTR(d_gscale); LINKS_CHECK_PTR(d_goffset); LINKS_CHECK_ALIGN16(packed);
  if (M < 1 || n_blocks < 1 || n_blocks > kFlowMaxBlocks) return LINKS_E_RANGE;
  if (stash != nullptr) LINKS_CHECK_ALIGN16(stash);
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.out = dx; A.gz = gz; A.gld = gld; A.M = M; A.n_blocks = n_blocks;
  A.ex_x1 = ex_x1; A.ex_dsub = ex_dsub; A.d_gscale = d_gscale; A.d_goffset = d_goffset;
  A.stash = stash;
  return launch_flow_tc_only<FLOW_NLL_FWDBWD>(C, A, links_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int links_flow_sample(const float* packed, int n_blocks, const float* x, const float* noise, int M,
                                 float* out, void* stream) {
  LINKS_CHECK_PTR(packed); LINKS_CHECK_PTR(x); LINKS_CHECK_PTR(noise); LINKS_CHECK_PTR(out); LINKS_CHECK_ALIGN16(packed);
  if (M < 1 || n_blocks < 1 || n_blocks > kFlowMaxBlocks) return LINKS_E_RANGE;
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.noise = noise; A.out = out; A.M = M; A.n_blocks = n_blocks;
  return launch_flow_any<34, FLOW_SAMPLE>(A, links_stream(stream));
}

// ---------------------------------------------------------------------------------------------
// grid of the chunk-walking metric kernels: at most 8 blocks (of 64 threads, ~27-36 KB shared) per SM
static int metric_grid(int M, int blocks_per_sm) {
  const int chunks = (M + kPosesPerBlock - 1) / kPosesPerBlock;
  return chunks < 148 * blocks_per_sm ? chunks : 148 * blocks_per_sm;
}

static int check_pose_args(const float* a, const float* b, int M, int J) {
  if (!a || !b) return LINKS_E_ARG;
  if (M < 1 || J < 2 || J > 17) return LINKS_E_RANGE;
  if ((reinterpret_cast<uintptr_t>(a) & 15u) || (reinterpret_cast<uintptr_t>(b) & 15u)) return LINKS_E_ALIGN;
  return 0;
}

extern "C" __attribute__((visibility("default"))) int links_mpjpe(const float* p_ref, const float* p, int M, int num_joints, int root_joint, int use_scaling,
                           float* per_pose, float* per_pose_max, float* dist, double* sum, void* stream) {
  int rc = check_pose_args(p_ref, p, M, num_joints);
  if (rc) return rc;
  if (root_joint < 0 || root_joint >= num_joints) return LINKS_E_RANGE;
  // light math per pose: two staging buffers keep the next chunk in flight (52 KB of shared memory -> 4 blocks per SM)
  const int grid = metric_grid(M, 4);
  cudaStream_t s = links_stream(stream);
  constexpr size_t smem = metric_smem_bytes(2);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(mpjpe_kernel<17, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(mpjpe_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaFuncSetAttribute(mpjpe_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    attr_done = true;
  }
  if (num_joints == 17)
    mpjpe_kernel<17, 2><<<grid, kPosesPerBlock, smem, s>>>(p_ref, p, M, 17, root_joint, use_scaling, per_pose, per_pose_max, dist, sum);
  else if (num_joints == 16)
    mpjpe_kernel<16, 2><<<grid, kPosesPerBlock, smem, s>>>(p_ref, p, M, 16, root_joint, use_scaling, per_pose, per_pose_max, dist, sum);
  else
    mpjpe_kernel<0, 2><<<grid, kPosesPerBlock, smem, s>>>(p_ref, p, M, num_joints, root_joint, use_scaling, per_pose, per_pose_max, dist, sum);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_threshold_counts(const float* values, size_t n, const float* thresholds, int n_thresh, int strict,
                                      unsigned long long* counts, void* stream) {
  LINKS_CHECK_PTR(values); LINKS_CHECK_PTR(thresholds); LINKS_CHECK_PTR(counts);
  if (n == 0 || n_thresh < 1 || n_thresh > 512) return LINKS_E_RANGE;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  threshold_counts_kernel<<<static_cast<int>(blocks), 256, 0, links_stream(stream)>>>(values, n, thresholds, n_thresh,
                                                                                    strict, counts);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_pmpjpe(const float* p_ref, const float* p, int M, int num_joints, int mode, float* per_pose,
                            float* aligned, double* sum, void* stream) {
  int rc = check_pose_args(p_ref, p, M, num_joints);
  if (rc) return rc;
  if (mode < 0 || mode > 1) return LINKS_E_RANGE;
  // SVD-heavy: one staging buffer, more resident blocks to hide the dependent chains
  const int grid = metric_grid(M, 8);
  cudaStream_t s = links_stream(stream);
  constexpr size_t smem = metric_smem_bytes(1);
  if (num_joints == 17)
    pmpjpe_kernel<17, 1><<<grid, kPosesPerBlock, smem, s>>>(p_ref, p, M, 17, mode, per_pose, aligned, sum);
  else if (num_joints == 16)
    pmpjpe_kernel<16, 1><<<grid, kPosesPerBlock, smem, s>>>(p_ref, p, M, 16, mode, per_pose, aligned, sum);
  else
    pmpjpe_kernel<0, 1><<<grid, kPosesPerBlock, smem, s>>>(p_ref, p, M, num_joints, mode, per_pose, aligned, sum);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_eval_lift_score(const float* poses_2d, const float* depth_off, int ld_depth, const float* gt_3d,
                                     int M, float depth, double* sums3, void* stream) {
  int rc = check_pose_args(poses_2d, gt_3d, M, 17);
  if (rc) return rc;
  LINKS_CHECK_PTR(depth_off); LINKS_CHECK_PTR(sums3);
  if (ld_depth < 17) return LINKS_E_RANGE;
  eval_lift_score_kernel<<<metric_grid(M, 6), kPosesPerBlock, 0, links_stream(stream)>>>(
      poses_2d, depth_off, ld_depth, gt_3d, M, depth, sums3);
  return links_launch_status();
}

// ---------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default"))) int links_occ_lift(const float* x, const float* head_leg, const float* head_torso, int M, float depth,
                              float* pose, void* stream) {
  LINKS_CHECK_PTR(x); LINKS_CHECK_PTR(head_leg); LINKS_CHECK_PTR(head_torso); LINKS_CHECK_PTR(pose);
  if (M < 1) return LINKS_E_RANGE;
  const long long total = static_cast<long long>(M) * 17;
  occ_lift_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, links_stream(stream)>>>(x, head_leg, head_torso, M,
                                                                                        depth, pose);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_occ_rotate_y(const float* pose, const float* u, int M, float* pose_out, void* stream) {
  LINKS_CHECK_PTR(pose); LINKS_CHECK_PTR(u); LINKS_CHECK_PTR(pose_out);
  if (M < 1) return LINKS_E_RANGE;
  const long long total = static_cast<long long>(M) * 17;
  occ_rotate_y_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, links_stream(stream)>>>(pose, u, M, pose_out);
  return links_launch_status();
}

extern "C" __attribute__((visibility("default"))) int links_occ_mse(const float* pred, int ld_pred, const float* pose, const int* tidx_dev, int n_out, int M,
                             float scale, float* loss_sum, void* g_bf16, void* gT_bf16, int ldT, int colT0,
                             void* stream) {
  LINKS_CHECK_PTR(pred); LINKS_CHECK_PTR(pose); LINKS_CHECK_PTR(tidx_dev); LINKS_CHECK_PTR(loss_sum); LINKS_CHECK_PTR(g_bf16);
  if (M < 1 || n_out < 1 || n_out > 64) return LINKS_E_RANGE;
  occ_mse_kernel<<<(M + 3) / 4, 128, 0, links_stream(stream)>>>(pred, ld_pred, pose, tidx_dev, n_out, M, scale, loss_sum,
                                                               static_cast<__nv_bfloat16*>(g_bf16),
                                                               static_cast<__nv_bfloat16*>(gT_bf16), ldT, colT0);
  return links_launch_status();
}
