// C ABI, part 1: runtime (errors, version) and the HBM-bound row kernels (K1, normalise-backward,
// loss finalize, hook scalar).  Declared in include/b200face.h.
#include <cmath>
#include <stdarg.h>
#include <atomic>

#include "common.cuh"
#include "rowops.cuh"

namespace b200f {

std::string& last_error_ref() {
  static thread_local std::string err;
  return err;
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
void pdl_set(bool on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }
static std::atomic<int> g_k1_hints{0};
int k1_hints() { return g_k1_hints.load(std::memory_order_relaxed); }
int k1_hints_set(int v) { return g_k1_hints.exchange(v); }

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

static inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
static inline bool dtype_ok(int dt) { return dt == B200F_F32 || dt == B200F_BF16; }
static inline size_t elem_size(int dt) { return dt == B200F_F32 ? 4 : 2; }

}  // namespace b200f

using namespace b200f;

template <typename T>
static int tail_fwd_t(const T* z, int64_t rows, int dim, const float* gamma, const float* beta, const float* mean, const float* stat,
                      int stat_is_var, float bn_eps, const uint8_t* mask, float keep_scale, float norm_eps, float out_scale, float* y,
                      __half* yhat16, float* emb, float* inv_norm, cudaStream_t st) {
  const int nvec = dim >> 2;
  const unsigned grid = (unsigned)ceil_div(rows, rowops::WARPS_PER_BLOCK);
  const dim3 blk(rowops::WARPS_PER_BLOCK * 32);
#define TAILF(NV) launch_pdl(rowops::tail_fwd_kernel<T, NV>, dim3(grid), blk, 0, st, z, rows, dim, gamma, beta, mean, stat, stat_is_var, \
                             bn_eps, mask, keep_scale, norm_eps, out_scale, y, yhat16, emb, inv_norm)
  if (nvec <= 32) TAILF(1); else if (nvec <= 64) TAILF(2); else if (nvec <= 128) TAILF(4); else TAILF(8);
#undef TAILF
  B200F_LAUNCH_OK("tail_fwd_kernel");
  return B200F_OK;
}

extern "C" {

int b200f_version(void) { return 100; }

unsigned long long b200f_launch_count(void) { return launch_count(); }

const char* b200f_last_error(void) { return last_error_ref().c_str(); }

int b200f_has_tcgen05(void) {
  // cudaGetDeviceProperties costs milliseconds: ask once per device (this is on the per-step host path)
  static std::atomic<int> cache[64];            // 0 = unknown, 1 = no, 2 = yes
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c != 0) return c == 2;
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < 64) cache[dev].store(major == 10 ? 2 : 1, std::memory_order_relaxed);
  return major == 10 ? 1 : 0;
}

int b200f_l2norm_rows(const void* in, int in_dtype, int64_t rows, int dim, float eps, float* inv_norm,
                      void* out_or_null, int out_dtype, float out_scale, void* stream) {
  B200F_NVTX("b200f_l2norm_rows");
  if (!dtype_ok(in_dtype)) return fail(B200F_ERR_ARG, "l2norm_rows: bad input dtype %d", in_dtype);
  if (out_or_null && !(dtype_ok(out_dtype) || out_dtype == B200F_F16N))
    return fail(B200F_ERR_ARG, "l2norm_rows: bad output dtype %d", out_dtype);
  if (rows < 0 || dim <= 0) return fail(B200F_ERR_ARG, "l2norm_rows: bad shape rows=%lld dim=%d", (long long)rows, dim);
  if (rows == 0) return B200F_OK;
  if (!in || (!inv_norm && !out_or_null)) return fail(B200F_ERR_ARG, "l2norm_rows: null pointer");
  cudaStream_t st = as_stream(stream);
  const int od = out_or_null ? out_dtype : B200F_F32;
#define L2N(TI, TO) rowops::launch_l2norm_rows<TI, TO>(static_cast<const TI*>(in), rows, dim, eps, out_scale, inv_norm, \
                                                      static_cast<TO*>(out_or_null), st)
  if (in_dtype == B200F_F32) {
    if (od == B200F_F32) L2N(float, float); else if (od == B200F_BF16) L2N(float, __nv_bfloat16); else L2N(float, __half);
  } else {
    if (od == B200F_F32) L2N(__nv_bfloat16, float); else if (od == B200F_BF16) L2N(__nv_bfloat16, __nv_bfloat16);
    else L2N(__nv_bfloat16, __half);
  }
#undef L2N
  B200F_LAUNCH_OK("l2norm_rows kernel");
  return B200F_OK;
}

int b200f_l2norm_rows_pair(const void* in0, int64_t rows0, float* inv0, void* out0, const void* in1, int64_t rows1,
                           float* inv1, void* out1, int in_dtype, int dim, float eps, int out_dtype, float out_scale,
                           void* stream) {
  B200F_NVTX("b200f_l2norm_rows_pair");
  if (rows0 <= 0 || rows1 <= 0 || !in0 || !in1 || !inv0 || !inv1 || !out0 || !out1)
    return fail(B200F_ERR_ARG, "l2norm_rows_pair: both row sets need input, inverse norms and output");
  const bool al = ((reinterpret_cast<uintptr_t>(in0) | reinterpret_cast<uintptr_t>(in1) | reinterpret_cast<uintptr_t>(out0) |
                    reinterpret_cast<uintptr_t>(out1)) & 31) == 0;
  if (in_dtype == B200F_BF16 && out_dtype == B200F_F16N && dim == 512 && al) {   // the head's shape: ONE launch
    const int64_t b0 = ceil_div(rows0, rowops::ROWS_PER_BLOCK), b1 = ceil_div(rows1, rowops::ROWS_PER_BLOCK);
    launch_pdl(rowops::l2norm_rows_512x16_pair_kernel<__nv_bfloat16, __half>, dim3((unsigned)(b0 + b1)),
               dim3(rowops::WARPS_PER_BLOCK * 32), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(in0), rows0, inv0,
               static_cast<__half*>(out0), (int)b0, static_cast<const __nv_bfloat16*>(in1), rows1, inv1,
               static_cast<__half*>(out1), eps, out_scale, rows1 >= 16384 ? k1_hints() : 0);
    B200F_LAUNCH_OK("l2norm_rows pair kernel");
    return B200F_OK;
  }
  int rc = b200f_l2norm_rows(in0, in_dtype, rows0, dim, eps, inv0, out0, out_dtype, out_scale, stream);
  if (rc) return rc;
  return b200f_l2norm_rows(in1, in_dtype, rows1, dim, eps, inv1, out1, out_dtype, out_scale, stream);
}

int b200f_l2norm_bwd(const void* v, int dtype, float v_scale, const float* inv_norm, const float* dvhat, int64_t rows,
                     int dim, float* dv, void* dv_bf16_or_null, void* stream) {
  B200F_NVTX("b200f_l2norm_bwd");
  if (!dtype_ok(dtype) && dtype != B200F_F16N) return fail(B200F_ERR_ARG, "l2norm_bwd: bad dtype");
  if (rows < 0 || dim <= 0) return fail(B200F_ERR_ARG, "l2norm_bwd: bad shape");
  if (rows == 0) return B200F_OK;
  if (!v || !inv_norm || !dvhat || !dv) return fail(B200F_ERR_ARG, "l2norm_bwd: null pointer");
  if (dtype == B200F_F16N && !(v_scale > 0.f)) return fail(B200F_ERR_ARG, "l2norm_bwd: v_scale must be > 0");
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16* lowp = static_cast<__nv_bfloat16*>(dv_bf16_or_null);
  if (dtype == B200F_F32)
    rowops::launch_l2norm_bwd<float, false>(static_cast<const float*>(v), 1.f, inv_norm, dvhat, rows, dim, dv, st, lowp);
  else if (dtype == B200F_BF16)
    rowops::launch_l2norm_bwd<__nv_bfloat16, false>(static_cast<const __nv_bfloat16*>(v), 1.f, inv_norm, dvhat, rows, dim, dv, st, lowp);
  else
    rowops::launch_l2norm_bwd<__half, true>(static_cast<const __half*>(v), v_scale, inv_norm, dvhat, rows, dim, dv, st, lowp);
  B200F_LAUNCH_OK("l2norm_bwd kernel");
  return B200F_OK;
}

int b200f_arcface_loss(const float* row_stats, int64_t B, const b200f_head_cfg* cfg, float* lse, float* loss,
                       float* pq_norm2, void* stream) {
  B200F_NVTX("b200f_arcface_loss");
  if (!row_stats || !cfg || B <= 0) return fail(B200F_ERR_ARG, "arcface_loss: bad argument");
  launch_pdl(rowops::loss_kernel, dim3(1), dim3(1024), 0, as_stream(stream), row_stats, B, cfg->s_eff,
             cfg->label_smoothing, (double)cfg->num_classes_total, lse, loss, pq_norm2, rowops::HookCfg{0, 1.f, 1, 0},
             (float*)nullptr);
  B200F_LAUNCH_OK("loss_kernel");
  return B200F_OK;
}

int b200f_arcface_loss_hook(const float* row_stats, int64_t B, const b200f_head_cfg* cfg, const b200f_hook_cfg* hook,
                            float* lse, float* loss, float* pq_norm2, float* out4, void* stream) {
  B200F_NVTX("b200f_arcface_loss_hook");
  if (!row_stats || !cfg || !hook || !out4 || B <= 0) return fail(B200F_ERR_ARG, "arcface_loss_hook: bad argument");
  launch_pdl(rowops::loss_kernel, dim3(1), dim3(1024), 0, as_stream(stream), row_stats, B, cfg->s_eff,
             cfg->label_smoothing, (double)cfg->num_classes_total, lse, loss, pq_norm2,
             rowops::HookCfg{hook->enabled, hook->max_grad_norm, hook->phase, hook->epoch}, out4);
  B200F_LAUNCH_OK("loss_kernel (+ hook scalars)");
  return B200F_OK;
}

int b200f_arcface_hook_scale(const float* pq_norm2, const float* upstream, int64_t B, float s_eff,
                             int hook_enabled, float max_grad_norm, int phase, int epoch, float* out4,
                             void* stream) {
  B200F_NVTX("b200f_arcface_hook_scale");
  if (!pq_norm2 || !out4 || B <= 0) return fail(B200F_ERR_ARG, "arcface_hook_scale: bad argument");
  launch_pdl(rowops::hook_scale_kernel, dim3(1), dim3(1), 0, as_stream(stream), pq_norm2, upstream, (double)B, s_eff,
             rowops::HookCfg{hook_enabled, max_grad_norm, phase, epoch}, out4);
  B200F_LAUNCH_OK("hook_scale_kernel");
  return B200F_OK;
}

int b200f_bn_stats(const void* z, int dtype, int64_t rows, int dim, float eps, float momentum, float* running_mean,
                   float* running_var, float* mean_out, float* invstd_out, void* stream) {
  B200F_NVTX("b200f_bn_stats");
  if (!dtype_ok(dtype)) return fail(B200F_ERR_ARG, "bn_stats: bad dtype");
  if (rows <= 0 || dim <= 0 || !z || !mean_out || !invstd_out) return fail(B200F_ERR_ARG, "bn_stats: bad argument");
  const unsigned grid = (unsigned)ceil_div(dim, 32);
  if (dtype == B200F_F32)
    launch_pdl(rowops::bn_stats_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream), static_cast<const float*>(z), rows, dim, eps,
               momentum, running_mean, running_var, mean_out, invstd_out);
  else
    launch_pdl(rowops::bn_stats_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(stream), static_cast<const __nv_bfloat16*>(z),
               rows, dim, eps, momentum, running_mean, running_var, mean_out, invstd_out);
  B200F_LAUNCH_OK("bn_stats_kernel");
  return B200F_OK;
}

int b200f_tail_fwd(const void* z, int dtype, int64_t rows, int dim, const float* gamma, const float* beta, const float* mean,
                   const float* stat, int stat_is_var, float bn_eps, const uint8_t* mask_or_null, float keep_scale, float norm_eps,
                   float out_scale, float* y_or_null, void* yhat16_or_null, float* emb_or_null, float* inv_norm, void* stream) {
  B200F_NVTX("b200f_tail_fwd");
  if (!dtype_ok(dtype)) return fail(B200F_ERR_ARG, "tail_fwd: bad dtype");
  if (rows <= 0 || dim <= 0 || dim % 4 != 0 || dim > 1024) return fail(B200F_ERR_ARG, "tail_fwd: rows > 0, dim %% 4 == 0, dim <= 1024");
  if (!z || !gamma || !beta || !mean || !stat || !inv_norm) return fail(B200F_ERR_ARG, "tail_fwd: null pointer");
  cudaStream_t st = as_stream(stream);
  if (dtype == B200F_F32)
    return tail_fwd_t<float>(static_cast<const float*>(z), rows, dim, gamma, beta, mean, stat, stat_is_var, bn_eps, mask_or_null, keep_scale,
                             norm_eps, out_scale, y_or_null, static_cast<__half*>(yhat16_or_null), emb_or_null, inv_norm, st);
  return tail_fwd_t<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(z), rows, dim, gamma, beta, mean, stat, stat_is_var, bn_eps,
                                   mask_or_null, keep_scale, norm_eps, out_scale, y_or_null, static_cast<__half*>(yhat16_or_null),
                                   emb_or_null, inv_norm, st);
}

int b200f_tail_bwd(const float* dy, const uint8_t* mask_or_null, float keep_scale, const void* z, int dtype, const float* mean,
                   const float* stat, int stat_is_var, float bn_eps, const float* gamma, int batch_stats, int64_t rows, int dim,
                   float* dgamma, float* dbeta, float* dz, void* stream) {
  B200F_NVTX("b200f_tail_bwd");
  if (!dtype_ok(dtype)) return fail(B200F_ERR_ARG, "tail_bwd: bad dtype");
  if (rows <= 0 || dim <= 0 || !dy || !z || !mean || !stat || !gamma || !dgamma || !dbeta || !dz)
    return fail(B200F_ERR_ARG, "tail_bwd: bad argument");
  if (batch_stats && stat_is_var) return fail(B200F_ERR_ARG, "tail_bwd: batch statistics come as (mean, invstd)");
  cudaStream_t st = as_stream(stream);
  const unsigned gcol = (unsigned)ceil_div(dim, 32);
  const unsigned gel = (unsigned)ceil_div(rows * (int64_t)dim, 256);
  // column sums: d_gamma / d_beta (both modes; eval uses invstd from the running variance)
  if (dtype == B200F_F32) {
    if (stat_is_var) return fail(B200F_ERR_UNSUPPORTED, "tail_bwd: pass invstd (b200f_bn_stats) for the column sums");
    launch_pdl(rowops::tail_bwd_cols_kernel<float>, dim3(gcol), dim3(256), 0, st, dy, mask_or_null, keep_scale, static_cast<const float*>(z),
               mean, stat, rows, dim, dgamma, dbeta);
    B200F_LAUNCH_OK("tail_bwd_cols_kernel");
    launch_pdl(rowops::tail_bwd_apply_kernel<float>, dim3(gel), dim3(256), 0, st, dy, mask_or_null, keep_scale, static_cast<const float*>(z),
               mean, stat, stat_is_var, bn_eps, gamma, (const float*)dgamma, (const float*)dbeta, batch_stats, rows, dim, dz);
  } else {
    if (stat_is_var) return fail(B200F_ERR_UNSUPPORTED, "tail_bwd: pass invstd (b200f_bn_stats) for the column sums");
    launch_pdl(rowops::tail_bwd_cols_kernel<__nv_bfloat16>, dim3(gcol), dim3(256), 0, st, dy, mask_or_null, keep_scale,
               static_cast<const __nv_bfloat16*>(z), mean, stat, rows, dim, dgamma, dbeta);
    B200F_LAUNCH_OK("tail_bwd_cols_kernel");
    launch_pdl(rowops::tail_bwd_apply_kernel<__nv_bfloat16>, dim3(gel), dim3(256), 0, st, dy, mask_or_null, keep_scale,
               static_cast<const __nv_bfloat16*>(z), mean, stat, stat_is_var, bn_eps, gamma, (const float*)dgamma, (const float*)dbeta,
               batch_stats, rows, dim, dz);
  }
  B200F_LAUNCH_OK("tail_bwd_apply_kernel");
  return B200F_OK;
}

int b200f_head_adamw(float* w, const float* dw, float* m, float* v, float* vmax, int64_t rows, int dim, double lr,
                     double beta1, double beta2, double eps, double weight_decay, int64_t step, const float* grad_scale,
                     void* w_hat_out, float out_scale, float norm_eps, float* inv_norm, void* stream) {
  B200F_NVTX("b200f_head_adamw");
  if (rows < 0 || dim <= 0 || step < 1) return fail(B200F_ERR_ARG, "head_adamw: bad shape / step (rows=%lld dim=%d step=%lld)",
                                                    (long long)rows, dim, (long long)step);
  if (rows == 0) return B200F_OK;
  if (!w || !dw || !m || !v) return fail(B200F_ERR_ARG, "head_adamw: null pointer");
  if (dim % 4 != 0 || dim > 1024) return fail(B200F_ERR_UNSUPPORTED, "head_adamw: dim %% 4 == 0 and dim <= 1024 (got %d)", dim);
  const uintptr_t al = reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dw) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(vmax) | (reinterpret_cast<uintptr_t>(w_hat_out) << 1);
  if (al & 15) return fail(B200F_ERR_ARG, "head_adamw: buffers must be 16-byte aligned (w_hat 8-byte)");
  if (!(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0)) return fail(B200F_ERR_ARG, "head_adamw: betas must be in [0, 1)");
  // torch evaluates these scalar expressions in double on the host and hands each to its kernel as one fp32 value
  rowops::AdamWParams p{};
  p.decay = (float)(1.0 - lr * weight_decay);
  p.one_m_b1 = (float)(1.0 - beta1); p.beta2 = (float)beta2; p.one_m_b2 = (float)(1.0 - beta2);
  p.step_size = (float)(lr / (1.0 - std::pow(beta1, (double)step)));
  p.bc2_sqrt = (float)std::sqrt(1.0 - std::pow(beta2, (double)step));
  p.eps = (float)eps;
  p.grad_scale = grad_scale; p.norm_eps = norm_eps; p.out_scale = out_scale;
  cudaStream_t st = as_stream(stream);
  const bool ok = vmax ? rowops::launch_adamw_rows<true>(w, dw, m, v, vmax, rows, dim, p, static_cast<__half*>(w_hat_out), inv_norm, st)
                       : rowops::launch_adamw_rows<false>(w, dw, m, v, nullptr, rows, dim, p, static_cast<__half*>(w_hat_out), inv_norm, st);
  if (!ok) return fail(B200F_ERR_UNSUPPORTED, "head_adamw: unsupported row length %d", dim);
  B200F_LAUNCH_OK("adamw_rows kernel");
  return B200F_OK;
}

}  // extern "C"
