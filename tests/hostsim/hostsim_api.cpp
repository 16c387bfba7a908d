// TEST INFRASTRUCTURE: the product's non-tensor-core CUDA kernels (csrc/*.cuh) compiled for the CPU with
// cuda_shim.h, exposed with the same argument lists as the C ABI (minus the stream) on HOST pointers.
// tests/test_hostsim_*.py compare these against the oracle, so kernel math is verified without a GPU.
#define LINKS_HOSTSIM 1
#include "cuda_shim.h"
using std::min;
using std::max;
#include "../../include/links_b200.h"
#include "../../links-3d-human-pose-estimation_b200/csrc/elementwise.cuh"
#include "../../links-3d-human-pose-estimation_b200/csrc/metrics.cuh"
#include "../../links-3d-human-pose-estimation_b200/csrc/geom.cuh"
#include "../../links-3d-human-pose-estimation_b200/csrc/flow.cuh"
#include "../../links-3d-human-pose-estimation_b200/csrc/occ.cuh"
#include "../../links-3d-human-pose-estimation_b200/csrc/dataprep.cuh"

using namespace links;
typedef __nv_bfloat16 bf16;
#define SIM extern "C" __attribute__((visibility("default")))

SIM int sim_pack_rows(const float* src, int ld_src, int M, const int* idx, int n_idx, int period, void* dst, void* dstT,
                      int ldT, int colT0) {
  const long long total = (long long)M * 64;
  hostsim::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [&] {
    pack_rows_kernel(src, ld_src, M, idx, n_idx, period, (bf16*)dst, (bf16*)dstT, ldT, colT0);
  });
  return 0;
}
SIM int sim_colsum_bf16(const void* G, int ldg, int M, int N, float* out, int accumulate) {
  if (!accumulate) for (int i = 0; i < N; ++i) out[i] = 0.f;
  const int rpb = 256;
  hostsim::launch(dim3((N + 31) / 32, (M + rpb - 1) / rpb), dim3(32, 8), 0,
                  [&] { colsum_bf16_kernel((const bf16*)G, ldg, M, N, out, rpb); });
  return 0;
}
SIM int sim_colsum_bf16_batched(const LinksColsumItem* items, int n_items) {
  ColsumBatch B;
  memset(&B, 0, sizeof(B));
  int maxN = 0, maxM = 0;
  for (int i = 0; i < n_items; ++i) { B.it[i] = items[i]; maxN = std::max(maxN, items[i].N); maxM = std::max(maxM, items[i].M); }
  B.n = n_items;
  hostsim::launch(dim3(n_items), dim3(256), 0, [&] { colsum_batched_zero_kernel(B); });
  const int rpb = 256;
  hostsim::launch(dim3((maxN + 31) / 32, n_items, (maxM + rpb - 1) / rpb), dim3(32, 8), 0, [&] { colsum_batched_kernel(B, rpb); });
  return 0;
}
SIM int sim_cast_weight(const float* W, int N, int K, void* Wb, int ldw, void* WT, int ldwt) {
  const int kx = ((Wb ? (ldw > K ? ldw : K) : K) + 31) / 32;
  const int ny = ((WT ? (ldwt > N ? ldwt : N) : N) + 31) / 32;
  hostsim::launch(dim3(kx, ny), dim3(32, 8), 0, [&] { cast_weight_kernel(W, N, K, (bf16*)Wb, ldw, (bf16*)WT, ldwt); });
  return 0;
}
SIM int sim_cast_weight_batched(const LinksCastItem* items, int n_items) {
  CastBatch B;
  memset(&B, 0, sizeof(B));
  for (int i = 0; i < n_items; ++i) B.it[i] = items[i];
  B.n = n_items;
  hostsim::launch(dim3(2, n_items), dim3(64), 0, [&] { cast_weight_batched_kernel(B); });
  return 0;
}
SIM int sim_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps,
                      float wd, int step, int* step_dev, float grad_scale, const float* lr_dev) {
  hostsim::launch(dim3(2), dim3(64), 0, [&] { adam_kernel<float>(p, g, m, v, n, lr, b1, b2, eps, wd, step_dev, step, grad_scale, lr_dev); });
  if (step_dev && step >= 0) hostsim::launch(dim3(1), dim3(32), 0, [&] { adam_incr_kernel(step_dev); });
  return 0;
}

SIM int sim_adam_step_g16(float* p, const void* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps,
                          float wd, int step, int* step_dev, float grad_scale, const float* lr_dev) {
  hostsim::launch(dim3(2), dim3(64), 0, [&] { adam_kernel<bf16>(p, (const bf16*)g, m, v, n, lr, b1, b2, eps, wd, step_dev, step, grad_scale, lr_dev); });
  if (step_dev && step >= 0) hostsim::launch(dim3(1), dim3(32), 0, [&] { adam_incr_kernel(step_dev); });
  return 0;
}
SIM int sim_normalize_head(const float* raw, int n, int root, int transposed, float fixed_scale, float* out, double* dist_sum) {
  if (dist_sum) *dist_sum = 0.0;
  hostsim::launch(dim3((n + 255) / 256), dim3(256), 0, [&] { normalize_head_center_kernel(raw, n, root, transposed, out, dist_sum); });
  hostsim::launch(dim3(2), dim3(256), 0, [&] { normalize_head_scale_kernel(out, (size_t)n * 34, n, fixed_scale, dist_sum); });
  return 0;
}
SIM int sim_small_matvec(const float* mat, const float* in, int n_in, int n_out, float* out) {
  hostsim::launch(dim3((n_out + 63) / 64), dim3(64), 0, [&] { small_matvec_kernel(mat, in, n_in, n_out, out); });
  return 0;
}
SIM int sim_adam_prepare(const int* step_dev, const float* lr_dev, float lr, float b1, float b2, float eps, float wd,
                         float grad_scale, float* hyper) {
  hostsim::launch(dim3(1), dim3(32), 0, [&] { adam_prepare_kernel(step_dev, lr_dev, lr, b1, b2, eps, wd, grad_scale, hyper); });
  return 0;
}
SIM int sim_grad_compress_bf16(const float* g, void* out, size_t n) {
  hostsim::launch(dim3(2), dim3(64), 0, [&] { grad_compress_bf16_kernel(g, (bf16*)out, n); });
  return 0;
}

SIM int sim_mpjpe(const float* r, const float* p, int M, int J, int root, int scaling, float* per_pose, float* per_max,
                  float* dist, double* sum) {
  hostsim::launch(dim3(std::min((M + kPosesPerBlock - 1) / kPosesPerBlock, 2)), dim3(kPosesPerBlock), metric_smem_bytes(2),
                  [&] {
                    if (J == 17) mpjpe_kernel<17, 2>(r, p, M, J, root, scaling, per_pose, per_max, dist, sum);
                    else if (J == 16) mpjpe_kernel<16, 2>(r, p, M, J, root, scaling, per_pose, per_max, dist, sum);
                    else mpjpe_kernel<0, 2>(r, p, M, J, root, scaling, per_pose, per_max, dist, sum);
                  });
  return 0;
}
SIM int sim_threshold_counts(const float* values, size_t n, const float* thr, int T, int strict, unsigned long long* counts) {
  hostsim::launch(dim3(2), dim3(256), 0, [&] { threshold_counts_kernel(values, n, thr, T, strict, counts); });
  return 0;
}
SIM int sim_pmpjpe(const float* r, const float* p, int M, int J, int mode, float* per_pose, float* aligned, double* sum) {
  hostsim::launch(dim3(std::min((M + kPosesPerBlock - 1) / kPosesPerBlock, 2)), dim3(kPosesPerBlock), metric_smem_bytes(1),
                  [&] {
                    if (J == 17) pmpjpe_kernel<17, 1>(r, p, M, J, mode, per_pose, aligned, sum);
                    else if (J == 16) pmpjpe_kernel<16, 1>(r, p, M, J, mode, per_pose, aligned, sum);
                    else pmpjpe_kernel<0, 1>(r, p, M, J, mode, per_pose, aligned, sum);
                  });
  return 0;
}
SIM int sim_eval_lift_score(const float* p2d, const float* doff, int ldd, const float* gt, int M, float depth, double* sums3) {
  hostsim::launch(dim3(std::min((M + kPosesPerBlock - 1) / kPosesPerBlock, 2)), dim3(kPosesPerBlock), 0,
                  [&] { eval_lift_score_kernel(p2d, doff, ldd, gt, M, depth, sums3); });
  return 0;
}

SIM int sim_elev_stats(const float* a0, const float* a1, int N, float* stats) {
  hostsim::launch(dim3(1), dim3(128), 0, [&] { elev_stats_kernel(a0, a1, N, stats); });
  return 0;
}
static GeomArgs base_args(const LinksGeomMaps* maps, const float* u, const float* h0, const float* h1, const float* a0,
                          const float* a1, const float* eps, const float* uy, const float* stats, int N) {
  GeomArgs A;
  memset(&A, 0, sizeof(A));
  A.maps = *maps;
  A.u = u; A.head[0] = h0; A.head[1] = h1; A.ang[0] = a0; A.ang[1] = a1; A.eps_x = eps; A.u_y = uy; A.stats = stats; A.N = N;
  return A;
}
SIM int sim_geom_forward(const LinksGeomMaps* maps, const float* u, const float* h0, const float* h1, const float* a0,
                         const float* a1, const float* eps, const float* uy, const float* stats, int N, float* qp0,
                         float* qp1, float* qf0, float* qf1) {
  GeomArgs A = base_args(maps, u, h0, h1, a0, a1, eps, uy, stats, N);
  A.qpart[0] = qp0; A.qpart[1] = qp1; A.qfull[0] = qf0; A.qfull[1] = qf1;
  if (int rc = geom_plan(A, 0)) return rc;
  hostsim::launch(dim3(2), dim3(kGeomWarps * 32), geom_smem_bytes(A.st), [&] { if (A.maps.V == 1) geom_forward_kernel<1, true>(A); else geom_forward_kernel<2, true>(A); });   // grid-stride over row blocks
  return 0;
}
SIM int sim_geom_loss(const LinksGeomMaps* maps, const float* u, const float* h0, const float* h1, const float* a0,
                      const float* a1, const float* eps, const float* uy, const float* stats, const float* h20,
                      const float* h21, int N, float* loss_sums, void* g20, void* g21, void* g2T0, void* g2T1, int ldT,
                      int colT0) {
  GeomArgs A = base_args(maps, u, h0, h1, a0, a1, eps, uy, stats, N);
  A.head2[0] = h20; A.head2[1] = h21; A.loss_sums = loss_sums;
  A.g2[0] = (bf16*)g20; A.g2[1] = (bf16*)g21; A.g2T[0] = (bf16*)g2T0; A.g2T[1] = (bf16*)g2T1; A.ldT = ldT; A.colT0 = colT0;
  const int pairs = (N + 1) / 2;
  (void)pairs;
  if (int rc = geom_plan(A, 1)) return rc;
  hostsim::launch(dim3(1), dim3(kGeomWarps * 32), geom_smem_bytes(A.st), [&] { if (A.maps.V == 1) geom_lossgrad_kernel<false, 1, true>(A); else geom_lossgrad_kernel<false, 2, true>(A); });
  return 0;
}
SIM int sim_geom_backward(const LinksGeomMaps* maps, const float* u, const float* h0, const float* h1, const float* a0,
                          const float* a1, const float* eps, const float* uy, const float* stats, const float* h20,
                          const float* h21, const float* df0, const float* df1, const float* dl0, const float* dl1, int N,
                          void* g10, void* g11, void* g1T0, void* g1T1, int ldT, int colT0, float* dgamma, float* da,
                          float* red) {
  GeomArgs A = base_args(maps, u, h0, h1, a0, a1, eps, uy, stats, N);
  A.head2[0] = h20; A.head2[1] = h21;
  A.dflow[0] = df0; A.dflow[1] = df1; A.dlift[0] = dl0; A.dlift[1] = dl1;
  A.g1[0] = (bf16*)g10; A.g1[1] = (bf16*)g11; A.g1T[0] = (bf16*)g1T0; A.g1T[1] = (bf16*)g1T1; A.ldT = ldT; A.colT0 = colT0;
  A.dgamma = dgamma; A.da = da; A.red = red;
  const int pairs = (N + 1) / 2;
  (void)pairs;
  if (int rc = geom_plan(A, 2)) return rc;
  hostsim::launch(dim3(2), dim3(kGeomWarps * 32), geom_smem_bytes(A.st), [&] { if (A.maps.V == 1) geom_lossgrad_kernel<true, 1, true>(A); else geom_lossgrad_kernel<true, 2, true>(A); });
  return 0;
}
SIM int sim_geom_backward_angles(const float* a0, const float* a1, const float* eps, const float* stats, const float* dgamma,
                                 const float* red, int N, void* g0, void* g1, void* gT0, void* gT1, int ldT, int colT0,
                                 int n_stat) {
  (void)eps;
  if (n_stat <= 0) n_stat = N;
  hostsim::launch(dim3((N + 255) / 256), dim3(256), 0, [&] {
    geom_backward_angles_kernel(a0, a1, stats, dgamma, red, N, (bf16*)g0, (bf16*)g1, (bf16*)gT0, (bf16*)gT1, ldT, colT0, n_stat);
  });
  return 0;
}
SIM int sim_elev_sums(const float* a0, const float* a1, int N, double* sums) {
  hostsim::launch(dim3(1), dim3(1024), 0, [&] { elev_sums_kernel(a0, a1, N, sums); });
  return 0;
}
SIM int sim_elev_finalize(const double* sums, int n_total, float* stats) {
  hostsim::launch(dim3(1), dim3(32), 0, [&] { elev_finalize_kernel(sums, (double)n_total, stats); });
  return 0;
}

SIM size_t sim_flow_packed_floats(int C, int nb) { return (size_t)flow_block_floats(C) * nb; }
SIM int sim_flow_pack(int C, int nb, const float* const* w0, const float* const* b0, const float* const* w2,
                      const float* const* b2, const float* const* gs, const float* const* go, const float* const* wp,
                      const float* const* wpi, float* packed) {
  FlowPackArgs A;
  memset(&A, 0, sizeof(A));
  for (int k = 0; k < nb; ++k) {
    A.w0[k] = w0[k]; A.b0[k] = b0[k]; A.w2[k] = w2[k]; A.b2[k] = b2[k]; A.gs[k] = gs[k]; A.go[k] = go[k]; A.wp[k] = wp[k]; A.wpi[k] = wpi[k];
  }
  A.packed = packed; A.C = C; A.n_blocks = nb;
  hostsim::launch(dim3(nb, kFlowPackSplit), dim3(256), 0, [&] { flow_pack_kernel(A); });
  return 0;
}
template <int C, int MODE>
static void sim_flow_launch(const FlowArgs& A) {
  hostsim::launch(dim3((A.M + kFlowRows - 1) / kFlowRows), dim3(kFlowWarps * 32), FlowSmem<C>::bytes, [&] { flow_kernel<C, MODE>(A); });
}
template <int MODE>
static int sim_flow_dispatch(int C, const FlowArgs& A) {
  switch (C) {
    case 14: sim_flow_launch<14, MODE>(A); return 0;
    case 20: sim_flow_launch<20, MODE>(A); return 0;
    case 22: sim_flow_launch<22, MODE>(A); return 0;
    case 32: sim_flow_launch<32, MODE>(A); return 0;
    case 34: sim_flow_launch<34, MODE>(A); return 0;
  }
  return LINKS_E_RANGE;
}
SIM int sim_flow_apply(const float* packed, int C, int nb, const float* x, int M, int rev, float* out, float* ld) {
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.out = out; A.ld = ld; A.M = M; A.n_blocks = nb;
  return rev ? sim_flow_dispatch<FLOW_REV>(C, A) : sim_flow_dispatch<FLOW_FWD>(C, A);
}
SIM int sim_flow_nll_fwdbwd(const float* packed, int C, int nb, const float* x, int M, float scale, float* nll_sum, float* dx,
                            float* /*stash: tensor-core kernel only*/) {
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.out = dx; A.nll_sum = nll_sum; A.scale = scale; A.M = M; A.n_blocks = nb;
  return sim_flow_dispatch<FLOW_NLL_FWDBWD>(C, A);
}
SIM int sim_flow_vjp(const float* packed, int C, int nb, const float* x, int M, const float* gz, const float* gld, float* dx) {
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.out = dx; A.gz = gz; A.gld = gld; A.M = M; A.n_blocks = nb;
  return sim_flow_dispatch<FLOW_NLL_FWDBWD>(C, A);
}
SIM int sim_flow_sample(const float* packed, int nb, const float* x, const float* noise, int M, float* out) {
  FlowArgs A;
  memset(&A, 0, sizeof(A));
  A.packed = packed; A.x = x; A.noise = noise; A.out = out; A.M = M; A.n_blocks = nb;
  sim_flow_launch<34, FLOW_SAMPLE>(A);
  return 0;
}

SIM int sim_occ_lift(const float* x, const float* hl, const float* ht, int M, float depth, float* pose) {
  const long long total = (long long)M * 17;
  hostsim::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [&] { occ_lift_kernel(x, hl, ht, M, depth, pose); });
  return 0;
}
SIM int sim_occ_rotate_y(const float* pose, const float* u, int M, float* out) {
  const long long total = (long long)M * 17;
  hostsim::launch(dim3((unsigned)((total + 255) / 256)), dim3(256), 0, [&] { occ_rotate_y_kernel(pose, u, M, out); });
  return 0;
}
SIM int sim_occ_mse(const float* pred, int ld_pred, const float* pose, const int* tidx, int n_out, int M, float scale,
                    float* loss_sum, void* g, void* gT, int ldT, int colT0) {
  hostsim::launch(dim3((M + 3) / 4), dim3(128), 0,
                  [&] { occ_mse_kernel(pred, ld_pred, pose, tidx, n_out, M, scale, loss_sum, (bf16*)g, (bf16*)gT, ldT, colT0); });
  return 0;
}
