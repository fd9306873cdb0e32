// C ABI, part 2: the ArcFace head (K2 forward statistics, K3 backward): argument checking, workspace
// carving and engine dispatch.  No allocation, no synchronisation.  Declared in include/b200face.h.
#include "common.cuh"
#include "head_simt.cuh"
#include "rowops.cuh"
#include "umma_api.cuh"

namespace b200f {

static inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
static inline bool dtype_ok(int dt) { return dt == B200F_F32 || dt == B200F_BF16; }
static inline size_t elem_size(int dt) { return dt == B200F_F32 ? 4 : 2; }

// ---- head workspace plan (shared by the size query and the calls) -------------------------
struct HeadPlan {
  // forward
  int n_chunks, tiles_per_chunk, m_tiles;
  size_t off_part, off_cos;
  // backward
  int64_t Cc, ldg;
  int n_splits; int64_t k_per_split;
  size_t off_G, off_r, off_dxpart;
  size_t total;
};

static HeadPlan plan_head(int64_t B, int64_t C, int D) {
  HeadPlan pl{};
  const int sms = num_sms();
  pl.m_tiles = (int)ceil_div(B, simt::BM);
  const int64_t n_tiles = ceil_div(C, simt::BN);
  int64_t want_chunks = ceil_div((int64_t)4 * sms, pl.m_tiles);
  if (want_chunks > n_tiles) want_chunks = n_tiles;
  if (want_chunks < 1) want_chunks = 1;
  pl.tiles_per_chunk = (int)ceil_div(n_tiles, want_chunks);
  pl.n_chunks = (int)ceil_div(n_tiles, pl.tiles_per_chunk);
  size_t off = 0;
  pl.off_part = off; off += align_up(sizeof(float) * pl.n_chunks * B * head_simt::PART_COLS, 256);
  pl.off_cos = off;  off += align_up(sizeof(float) * 2 * (size_t)pl.n_chunks * pl.m_tiles, 256);
  const size_t fwd_total = off;
  // backward: G chunk of at most 64 MB fp32
  const int64_t c_round = ceil_div(C, simt::BN) * simt::BN;
  int64_t cc = ((int64_t)(64u << 20) / 4 / B) / simt::BN * simt::BN;
  if (cc < simt::BN) cc = simt::BN;
  if (cc > c_round) cc = c_round;
  pl.Cc = cc; pl.ldg = cc;
  const int64_t d_tiles = ceil_div(D, simt::BN);
  int64_t splits = ceil_div((int64_t)2 * sms, pl.m_tiles * d_tiles);
  const int64_t max_splits = cc / simt::BN;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  pl.k_per_split = ceil_div(cc / simt::BN, splits) * simt::BN;
  pl.n_splits = (int)ceil_div(cc, pl.k_per_split);
  off = 0;
  pl.off_G = off;      off += align_up(sizeof(float) * (size_t)B * pl.ldg, 256);
  pl.off_r = off;      off += align_up(sizeof(float) * (size_t)cc, 256);
  pl.off_dxpart = off; off += align_up(sizeof(float) * (size_t)pl.n_splits * B * D, 256);
  pl.total = off > fwd_total ? off : fwd_total;
  return pl;
}

template <typename T>
static int head_fwd_simt(const void* x, const void* w, const float* inv_nx, const float* inv_nw,
                         const int64_t* label, int64_t B, int64_t C, int64_t class_offset, int D,
                         const b200f_head_cfg* cfg, float* row_stats, float* row_best,
                         int64_t* row_argmax, float* cos_minmax, int32_t* nan_flag, float* logits,
                         int64_t ld_logits, char* ws, const HeadPlan& pl, cudaStream_t st) {
  head_simt::FwdParams p{};
  p.x = x; p.w = w; p.inv_nx = inv_nx; p.inv_nw = inv_nw; p.label = label;
  p.B = B; p.C = C; p.class_offset = class_offset; p.D = D;
  p.hm = HeadMath{cfg->m_eff, cfg->s_eff, cfg->easy_margin};
  p.n_chunks = pl.n_chunks; p.tiles_per_chunk = pl.tiles_per_chunk;
  p.part = reinterpret_cast<float*>(ws + pl.off_part);
  p.cos_part = reinterpret_cast<float*>(ws + pl.off_cos);
  p.nan_flag = nan_flag; p.logits = logits; p.ld_logits = ld_logits;
  p.vec_x = simt::vec_friendly<T>(x, D); p.vec_w = simt::vec_friendly<T>(w, D);
  dim3 grid(pl.n_chunks, pl.m_tiles);
  head_simt::fwd_kernel<T><<<grid, simt::THREADS, 0, st>>>(p);
  B200F_LAUNCH_OK("head_simt::fwd_kernel");
  head_simt::reduce_partials_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(
      p.part, pl.n_chunks, B, p.cos_part, pl.n_chunks * pl.m_tiles, row_stats, row_best, row_argmax,
      cos_minmax);
  B200F_LAUNCH_OK("head_simt::reduce_partials_kernel");
  return B200F_OK;
}

template <typename T>
static int head_bwd_simt(const void* x, const void* w, const float* inv_nx, const float* inv_nw,
                         const int64_t* label, const float* lse, const float* grad_scale,
                         const float* dlogits, int64_t ld_dlogits, int64_t B,
                         int64_t C, int64_t class_offset, int D, const b200f_head_cfg* cfg,
                         float* dxhat, float* dw, char* ws, const HeadPlan& pl, cudaStream_t st) {
  float* G = reinterpret_cast<float*>(ws + pl.off_G);
  float* r = reinterpret_cast<float*>(ws + pl.off_r);
  float* dxpart = reinterpret_cast<float*>(ws + pl.off_dxpart);
  const bool vx = simt::vec_friendly<T>(x, D), vw = simt::vec_friendly<T>(w, D);
  const int64_t d_tiles = ceil_div(D, simt::BN);
  int chunk_no = 0;
  for (int64_t c0 = 0; c0 < C; c0 += pl.Cc, ++chunk_no) {
    const int64_t c_cnt = (C - c0 < pl.Cc) ? (C - c0) : pl.Cc;
    head_simt::BwdGParams g{};
    g.x = x; g.w = w; g.inv_nx = inv_nx; g.inv_nw = inv_nw; g.label = label; g.lse = lse;
    g.grad_scale = grad_scale; g.B = B; g.C = C; g.class_offset = class_offset; g.c0 = c0; g.Cc = pl.Cc;
    g.D = D; g.hm = HeadMath{cfg->m_eff, cfg->s_eff, cfg->easy_margin};
    g.ls_eps = cfg->label_smoothing; g.inv_Ctot = 1.0f / (float)cfg->num_classes_total;
    g.G = G; g.ldg = pl.ldg; g.r = r; g.vec_x = vx; g.vec_w = vw;
    g.dlogits = dlogits; g.ld_dlogits = ld_dlogits;
    head_simt::bwd_g_kernel<T><<<(unsigned)ceil_div(c_cnt, simt::BN), simt::THREADS, 0, st>>>(g);
    B200F_LAUNCH_OK("head_simt::bwd_g_kernel");

    head_simt::BwdDwParams dwp{};
    dwp.x = x; dwp.w = w; dwp.inv_nx = inv_nx; dwp.inv_nw = inv_nw; dwp.G = G; dwp.ldg = pl.ldg; dwp.r = r;
    dwp.B = B; dwp.C = C; dwp.c0 = c0; dwp.Cc = pl.Cc; dwp.D = D; dwp.dw = dw; dwp.vec_x = vx;
    dim3 gdw((unsigned)ceil_div(c_cnt, simt::BM), (unsigned)d_tiles);
    head_simt::bwd_dw_kernel<T><<<gdw, simt::THREADS, 0, st>>>(dwp);
    B200F_LAUNCH_OK("head_simt::bwd_dw_kernel");

    head_simt::BwdDxParams dxp{};
    dxp.w = w; dxp.inv_nw = inv_nw; dxp.G = G; dxp.ldg = pl.ldg; dxp.B = B; dxp.C = C; dxp.c0 = c0;
    dxp.Cc = pl.Cc; dxp.D = D; dxp.k_per_split = pl.k_per_split; dxp.part = dxpart; dxp.vec_w = vw;
    const int splits = (int)ceil_div(c_cnt, pl.k_per_split);
    dim3 gdx(pl.m_tiles, (unsigned)d_tiles, splits);
    head_simt::bwd_dx_kernel<T><<<gdx, simt::THREADS, 0, st>>>(dxp);
    B200F_LAUNCH_OK("head_simt::bwd_dx_kernel");
    const int64_t n = B * D;
    head_simt::reduce_splits_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(dxpart, splits, n, dxhat,
                                                                               chunk_no > 0);
    B200F_LAUNCH_OK("head_simt::reduce_splits_kernel");
  }
  return B200F_OK;
}

// sum of squares of n floats, one block, fixed order (the fp32 engine's ||dW||^2: its problems are small)
static __global__ void __launch_bounds__(1024) sumsq_kernel(const float* __restrict__ v, int64_t n, float* out) {
  __shared__ float sh[1024];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s = fmaf(v[i], v[i], s);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

}  // namespace b200f

using namespace b200f;

extern "C" {

size_t b200f_head_workspace_bytes(int64_t B, int64_t C_local, int D, int dtype, int engine) {
  (void)engine;
  if (B <= 0 || C_local <= 0 || D <= 0) return 0;
  if (dtype == B200F_F16N) return umma::head_workspace_bytes(B, C_local, D);
  return plan_head(B, C_local, D).total;
}

// dtype B200F_F16N (pre-normalised fp16 operands from K1) selects the tcgen05 engine; B200F_F32 /
// B200F_BF16 run on the fp32 CUDA-core engine.
static int check_head_args(const char* who, const void* x, const void* w, int dtype, const float* inv_nx,
                           const float* inv_nw, const int64_t* label, int64_t B, int64_t C, int D,
                           const b200f_head_cfg* cfg) {
  if (!dtype_ok(dtype) && dtype != B200F_F16N) return fail(B200F_ERR_ARG, "%s: bad dtype %d", who, dtype);
  if (B <= 0 || C <= 0 || D <= 0) return fail(B200F_ERR_ARG, "%s: bad shape B=%lld C=%lld D=%d", who, (long long)B, (long long)C, D);
  if (!x || !w || !label || !cfg) return fail(B200F_ERR_ARG, "%s: null pointer", who);
  if (dtype != B200F_F16N && (!inv_nx || !inv_nw)) return fail(B200F_ERR_ARG, "%s: inverse norms required", who);
  if (cfg->num_classes_total < C) return fail(B200F_ERR_ARG, "%s: num_classes_total < C_local", who);
  if (cfg->num_classes_total >= (int64_t)1 << 31) return fail(B200F_ERR_ARG, "%s: more than 2^31 classes", who);
  if (dtype == B200F_F16N && cfg->engine == B200F_ENGINE_SIMT)
    return fail(B200F_ERR_UNSUPPORTED, "%s: B200F_F16N operands belong to the tcgen05 engine", who);
  if (dtype != B200F_F16N && cfg->engine == B200F_ENGINE_TCGEN05)
    return fail(B200F_ERR_UNSUPPORTED, "%s: the tcgen05 engine takes B200F_F16N operands (K1 output)", who);
  return B200F_OK;
}

static int arcface_fwd_impl(const char* who, const void* x, const void* w, int dtype, const float* inv_nx, const float* inv_nw,
                            const int64_t* label, int64_t B, int64_t C_local, int64_t class_offset, int D,
                            const b200f_head_cfg* cfg, float* row_stats, float* row_best, int64_t* row_argmax,
                            float* cos_minmax, int32_t* nan_flag, float* logits_or_null, int64_t ld_logits,
                            const umma::HeadFinal* fin, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_head_args(who, x, w, dtype, inv_nx, inv_nw, label, B, C_local, D, cfg);
  if (rc) return rc;
  if (!row_stats || !nan_flag) return fail(B200F_ERR_ARG, "%s: null output", who);
  if (logits_or_null && ld_logits < C_local) return fail(B200F_ERR_ARG, "%s: ld_logits < C_local", who);
  const size_t need = b200f_head_workspace_bytes(B, C_local, D, dtype, cfg->engine);
  if (!workspace || workspace_bytes < need)
    return fail(B200F_ERR_WORKSPACE, "%s: workspace %zu < %zu", who, workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  if (dtype == B200F_F16N) {
    if (logits_or_null) return fail(B200F_ERR_UNSUPPORTED, "%s: the tcgen05 engine never stores logits", who);
    return umma::head_fwd(x, w, label, B, C_local, class_offset, D, cfg, row_stats, row_best, row_argmax, cos_minmax,
                          nan_flag, fin, static_cast<char*>(workspace), workspace_bytes, st);
  }
  const HeadPlan pl = plan_head(B, C_local, D);
  if (dtype == B200F_F32)
    rc = head_fwd_simt<float>(x, w, inv_nx, inv_nw, label, B, C_local, class_offset, D, cfg, row_stats, row_best,
                              row_argmax, cos_minmax, nan_flag, logits_or_null, ld_logits,
                              static_cast<char*>(workspace), pl, st);
  else
    rc = head_fwd_simt<__nv_bfloat16>(x, w, inv_nx, inv_nw, label, B, C_local, class_offset, D, cfg, row_stats,
                                      row_best, row_argmax, cos_minmax, nan_flag, logits_or_null, ld_logits,
                                      static_cast<char*>(workspace), pl, st);
  if (rc || fin == nullptr) return rc;
  launch_pdl(rowops::loss_kernel, dim3(1), dim3(1024), 0, st, (const float*)row_stats, B, cfg->s_eff, cfg->label_smoothing,
             (double)cfg->num_classes_total, fin->lse, fin->loss, fin->pq_norm2,
             rowops::HookCfg{fin->hook_enabled, fin->max_grad_norm, fin->phase, fin->epoch}, fin->out4);
  B200F_LAUNCH_OK("loss_kernel");
  return B200F_OK;
}

static int arcface_bwd_impl(const char* who, const void* x, const void* w, int dtype, const float* inv_nx, const float* inv_nw,
                            const int64_t* label, const float* lse, const float* grad_scale,
                            const float* dlogits_or_null, int64_t ld_dlogits, int64_t B, int64_t C_local,
                            int64_t class_offset, int D, const b200f_head_cfg* cfg, float* dxhat, float* dw,
                            const umma::HeadDx* hdx, void* workspace, size_t workspace_bytes, void* stream, int phase = 0,
                            int cluster_limit = 0) {
  int rc = check_head_args(who, x, w, dtype, inv_nx, inv_nw, label, B, C_local, D, cfg);
  if (rc) return rc;
  const bool part = phase >= umma::HEAD_BWD_PART_K3A && phase <= umma::HEAD_BWD_PART_K3C;
  if ((phase < 0 || phase > 2) && !part) return fail(B200F_ERR_ARG, "%s: phase must be 0, 1 or 2", who);
  if (part && dtype != B200F_F16N) return fail(B200F_ERR_UNSUPPORTED, "%s: parts exist on the tcgen05 engine only", who);
  if (phase == 2 && dtype != B200F_F16N) return B200F_OK;      // CUDA-core engine: phase 1 already did everything
  if (!lse || !grad_scale || !dxhat || !dw) return fail(B200F_ERR_ARG, "%s: null pointer", who);
  const size_t need = b200f_head_workspace_bytes(B, C_local, D, dtype, cfg->engine);
  if (!workspace || workspace_bytes < need)
    return fail(B200F_ERR_WORKSPACE, "%s: workspace %zu < %zu", who, workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  if (dlogits_or_null && ld_dlogits < C_local) return fail(B200F_ERR_ARG, "%s: ld_dlogits < C_local", who);
  if (dtype == B200F_F16N) {
    if (dlogits_or_null) return fail(B200F_ERR_UNSUPPORTED, "%s: the tcgen05 engine has no dlogits path", who);
    if (!inv_nw) return fail(B200F_ERR_ARG, "%s: inv_nw required", who);
    return umma::head_bwd(x, w, inv_nw, label, lse, grad_scale, B, C_local, class_offset, D, cfg, dxhat, dw, hdx,
                          static_cast<char*>(workspace), workspace_bytes, st, phase, cluster_limit);
  }
  const HeadPlan pl = plan_head(B, C_local, D);
  if (dtype == B200F_F32)
    rc = head_bwd_simt<float>(x, w, inv_nx, inv_nw, label, lse, grad_scale, dlogits_or_null, ld_dlogits, B,
                              C_local, class_offset, D, cfg, dxhat, dw, static_cast<char*>(workspace), pl, st);
  else
    rc = head_bwd_simt<__nv_bfloat16>(x, w, inv_nx, inv_nw, label, lse, grad_scale, dlogits_or_null, ld_dlogits,
                                      B, C_local, class_offset, D, cfg, dxhat, dw,
                                      static_cast<char*>(workspace), pl, st);
  if (float* sq = umma::head_dw_sqnorm_request()) {            // CUDA-core engine: one pass over its (small) dW
    umma::head_request_dw_sqnorm(nullptr);
    if (!rc) {
      sumsq_kernel<<<1, 1024, 0, st>>>(dw, C_local * (int64_t)D, sq);
      B200F_LAUNCH_OK("sumsq_kernel");
    }
  }
  if (rc || hdx == nullptr || hdx->dx == nullptr) return rc;
  // CUDA-core engine: its operands ARE the raw rows
  __nv_bfloat16* lowp = static_cast<__nv_bfloat16*>(hdx->dx_bf16);
  if (dtype == B200F_F32)
    rowops::launch_l2norm_bwd<float, false>(static_cast<const float*>(x), 1.f, inv_nx, dxhat, B, D, hdx->dx, st, lowp);
  else
    rowops::launch_l2norm_bwd<__nv_bfloat16, false>(static_cast<const __nv_bfloat16*>(x), 1.f, inv_nx, dxhat, B, D, hdx->dx, st, lowp);
  B200F_LAUNCH_OK("l2norm_bwd kernel (dx)");
  return B200F_OK;
}

int b200f_arcface_fwd(const void* x, const void* w, int dtype, const float* inv_nx, const float* inv_nw,
                      const int64_t* label, int64_t B, int64_t C_local, int64_t class_offset, int D,
                      const b200f_head_cfg* cfg, float* row_stats, float* row_best, int64_t* row_argmax,
                      float* cos_minmax, int32_t* nan_flag, float* logits_or_null, int64_t ld_logits,
                      void* workspace, size_t workspace_bytes, void* stream) {
  B200F_NVTX("b200f_arcface_fwd");
  return arcface_fwd_impl("arcface_fwd", x, w, dtype, inv_nx, inv_nw, label, B, C_local, class_offset, D, cfg, row_stats,
                          row_best, row_argmax, cos_minmax, nan_flag, logits_or_null, ld_logits, nullptr, workspace,
                          workspace_bytes, stream);
}

int b200f_arcface_fwd_loss(const void* x, const void* w, int dtype, const float* inv_nx, const float* inv_nw,
                           const int64_t* label, int64_t B, int64_t C_local, int64_t class_offset, int D,
                           const b200f_head_cfg* cfg, const b200f_hook_cfg* hook, float* row_stats, float* row_best,
                           int64_t* row_argmax, float* cos_minmax, int32_t* nan_flag, float* lse, float* loss,
                           float* pq_norm2, float* out4, void* workspace, size_t workspace_bytes, void* stream) {
  B200F_NVTX("b200f_arcface_fwd_loss");
  if (!hook || !lse || !loss || !pq_norm2 || !out4) return fail(B200F_ERR_ARG, "arcface_fwd_loss: null pointer");
  if (cfg && C_local != cfg->num_classes_total)
    return fail(B200F_ERR_ARG, "arcface_fwd_loss: a class shard needs the cross-shard sum first (b200f_arcface_fwd, "
                               "all-reduce, b200f_arcface_loss_hook)");
  const umma::HeadFinal fin{lse, loss, pq_norm2, hook->enabled, hook->max_grad_norm, hook->phase, hook->epoch, out4};
  return arcface_fwd_impl("arcface_fwd_loss", x, w, dtype, inv_nx, inv_nw, label, B, C_local, class_offset, D, cfg, row_stats,
                          row_best, row_argmax, cos_minmax, nan_flag, nullptr, 0, &fin, workspace, workspace_bytes, stream);
}

int b200f_arcface_fwd_raw(const void* x_raw_or_null, int x_dtype, void* x_f16n, float* inv_nx,
                          const void* w_raw, int w_dtype, void* w_f16n, float* inv_nw, float eps,
                          const int64_t* label, int64_t B, int64_t C_local, int64_t class_offset, int D,
                          const b200f_head_cfg* cfg, const b200f_hook_cfg* hook_or_null, float* row_stats, float* row_best,
                          int64_t* row_argmax, float* cos_minmax, int32_t* nan_flag, float* lse, float* loss,
                          float* pq_norm2, float* out4, void* workspace, size_t workspace_bytes, void* stream) {
  B200F_NVTX("b200f_arcface_fwd_raw");
  const char* who = "arcface_fwd_raw";
  int rc = check_head_args(who, x_f16n, w_f16n, B200F_F16N, inv_nx, inv_nw, label, B, C_local, D, cfg);
  if (rc) return rc;
  if (!w_raw || !inv_nx || !inv_nw || !row_stats || !nan_flag) return fail(B200F_ERR_ARG, "%s: null pointer", who);
  if (!dtype_ok(w_dtype) || (x_raw_or_null && !dtype_ok(x_dtype))) return fail(B200F_ERR_ARG, "%s: raw rows must be fp32 or bf16", who);
  if (hook_or_null) {
    if (!lse || !loss || !pq_norm2 || !out4) return fail(B200F_ERR_ARG, "%s: null loss output", who);
    if (C_local != cfg->num_classes_total)
      return fail(B200F_ERR_ARG, "%s: a class shard needs the cross-shard sum first (hook = NULL, all-reduce, b200f_arcface_loss_hook)", who);
  }
  const size_t need = b200f_head_workspace_bytes(B, C_local, D, B200F_F16N, cfg->engine);
  if (!workspace || workspace_bytes < need) return fail(B200F_ERR_WORKSPACE, "%s: workspace %zu < %zu", who, workspace_bytes, need);
  umma::HeadFinal fin{};
  if (hook_or_null) fin = umma::HeadFinal{lse, loss, pq_norm2, hook_or_null->enabled, hook_or_null->max_grad_norm, hook_or_null->phase, hook_or_null->epoch, out4};
  const umma::HeadPrep prep{x_raw_or_null, x_dtype, inv_nx, w_raw, w_dtype, inv_nw, eps};
  return umma::head_fwd(x_f16n, w_f16n, label, B, C_local, class_offset, D, cfg, row_stats, row_best, row_argmax, cos_minmax,
                        nan_flag, hook_or_null ? &fin : nullptr, static_cast<char*>(workspace), workspace_bytes, as_stream(stream), &prep);
}

int b200f_arcface_bwd(const void* x, const void* w, int dtype, const float* inv_nx, const float* inv_nw,
                      const int64_t* label, const float* lse, const float* grad_scale,
                      const float* dlogits_or_null, int64_t ld_dlogits, int64_t B,
                      int64_t C_local, int64_t class_offset, int D, const b200f_head_cfg* cfg, float* dxhat,
                      float* dw, void* workspace, size_t workspace_bytes, void* stream) {
  B200F_NVTX("b200f_arcface_bwd");
  return arcface_bwd_impl("arcface_bwd", x, w, dtype, inv_nx, inv_nw, label, lse, grad_scale, dlogits_or_null, ld_dlogits, B,
                          C_local, class_offset, D, cfg, dxhat, dw, nullptr, workspace, workspace_bytes, stream);
}

int b200f_head_request_dw_sqnorm(float* out_or_null) {
  umma::head_request_dw_sqnorm(out_or_null);
  return B200F_OK;
}

int b200f_arcface_bwd_phase(const void* x, const void* w, int dtype, const float* inv_nx, const float* inv_nw,
                            const int64_t* label, const float* lse, const float* grad_scale, int64_t B,
                            int64_t C_local, int64_t class_offset, int D, const b200f_head_cfg* cfg, float* dxhat,
                            float* dw, int phase, void* workspace, size_t workspace_bytes, void* stream) {
  B200F_NVTX("b200f_arcface_bwd_phase");
  if (phase != 1 && phase != 2) return fail(B200F_ERR_ARG, "arcface_bwd_phase: phase must be 1 or 2");
  return arcface_bwd_impl("arcface_bwd_phase", x, w, dtype, inv_nx, inv_nw, label, lse, grad_scale, nullptr, 0, B,
                          C_local, class_offset, D, cfg, dxhat, dw, nullptr, workspace, workspace_bytes, stream, phase);
}

int b200f_arcface_bwd_parts_ok(int64_t B, int64_t C_local, int D, int dtype) {
  return (dtype == B200F_F16N && umma::available()) ? umma::head_bwd_parts_ok(B, C_local, D) : 0;
}

int b200f_arcface_bwd_part(const void* x, const void* w, int dtype, const float* inv_nx, const float* inv_nw,
                           const int64_t* label, const float* lse, const float* grad_scale, int64_t B,
                           int64_t C_local, int64_t class_offset, int D, const b200f_head_cfg* cfg, float* dxhat,
                           float* dw, const void* x_raw_or_null, int x_raw_dtype, float* dx_or_null, void* dx_bf16_or_null,
                           int part, int max_clusters, void* workspace, size_t workspace_bytes, void* stream) {
  B200F_NVTX("b200f_arcface_bwd_part");
  if (part < 1 || part > 3) return fail(B200F_ERR_ARG, "arcface_bwd_part: part must be 1 (G^T), 2 (dW) or 3 (dx)");
  if (max_clusters < 0) return fail(B200F_ERR_ARG, "arcface_bwd_part: max_clusters < 0");
  if (x_raw_or_null && !dtype_ok(x_raw_dtype)) return fail(B200F_ERR_ARG, "arcface_bwd_part: bad x_raw dtype %d", x_raw_dtype);
  const umma::HeadDx hdx{x_raw_or_null, x_raw_dtype, inv_nx, dx_or_null, dx_bf16_or_null};
  const bool want_dx = part == 3 && dx_or_null != nullptr;
  if (want_dx && !inv_nx) return fail(B200F_ERR_ARG, "arcface_bwd_part: dx needs inv_nx");
  return arcface_bwd_impl("arcface_bwd_part", x, w, dtype, inv_nx, inv_nw, label, lse, grad_scale, nullptr, 0, B, C_local,
                          class_offset, D, cfg, dxhat, dw, want_dx ? &hdx : nullptr, workspace, workspace_bytes, stream,
                          umma::HEAD_BWD_PART_K3A + part - 1, max_clusters);
}

int b200f_arcface_bwd_dx(const void* x, const void* w, int dtype, const float* inv_nx, const float* inv_nw,
                         const int64_t* label, const float* lse, const float* grad_scale,
                         const float* dlogits_or_null, int64_t ld_dlogits, int64_t B,
                         int64_t C_local, int64_t class_offset, int D, const b200f_head_cfg* cfg, float* dxhat,
                         float* dw, const void* x_raw_or_null, int x_raw_dtype, float* dx, void* dx_bf16_or_null,
                         void* workspace, size_t workspace_bytes, void* stream) {
  B200F_NVTX("b200f_arcface_bwd_dx");
  if (!dx || !inv_nx) return fail(B200F_ERR_ARG, "arcface_bwd_dx: dx and inv_nx are required");
  if (x_raw_or_null && !dtype_ok(x_raw_dtype)) return fail(B200F_ERR_ARG, "arcface_bwd_dx: bad x_raw dtype %d", x_raw_dtype);
  const umma::HeadDx hdx{x_raw_or_null, x_raw_dtype, inv_nx, dx, dx_bf16_or_null};
  return arcface_bwd_impl("arcface_bwd_dx", x, w, dtype, inv_nx, inv_nw, label, lse, grad_scale, dlogits_or_null, ld_dlogits, B,
                          C_local, class_offset, D, cfg, dxhat, dw, &hdx, workspace, workspace_bytes, stream);
}

}  // extern "C"
