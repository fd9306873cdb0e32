// HBM-bound row kernels: K1 fused L2-normalise, normalise-backward, loss finalize, hook scalar.
// Reference: F.normalize(v, p=2, dim=1, eps=1e-12) at src/face_models.py:351-352,525 and its
// autograd; nn.CrossEntropyLoss(label_smoothing) src/training.py:341; hook src/face_models.py:538-567.
#pragma once
#include "common.cuh"

namespace b200f {
namespace rowops {

constexpr int ROWS_PER_BLOCK = 8;   // one warp per row

// 16-byte vector of elements -> fp32
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(r[i] << 16);
      v[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
    }
  }
};

template <typename TO> __device__ __forceinline__ void store_elem(TO* p, float v);
template <> __device__ __forceinline__ void store_elem<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_elem<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// N consecutive outputs as one vector store (N = 4 or 8; dst is N*sizeof(TO)-aligned)
template <typename TO, int N> struct StoreN;
template <> struct StoreN<float, 4> {
  static __device__ __forceinline__ void put(float* d, const float (&v)[4], float s) {
    *reinterpret_cast<float4*>(d) = make_float4(v[0] * s, v[1] * s, v[2] * s, v[3] * s);
  }
};
template <> struct StoreN<float, 8> {
  static __device__ __forceinline__ void put(float* d, const float (&v)[8], float s) {
    reinterpret_cast<float4*>(d)[0] = make_float4(v[0] * s, v[1] * s, v[2] * s, v[3] * s);
    reinterpret_cast<float4*>(d)[1] = make_float4(v[4] * s, v[5] * s, v[6] * s, v[7] * s);
  }
};
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> struct StoreN<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void put(__nv_bfloat16* d, const float (&v)[4], float s) {
    *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16x2(v[0] * s, v[1] * s), pack_bf16x2(v[2] * s, v[3] * s));
  }
};
template <> struct StoreN<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void put(__nv_bfloat16* d, const float (&v)[8], float s) {
    *reinterpret_cast<uint4*>(d) = make_uint4(pack_bf16x2(v[0] * s, v[1] * s), pack_bf16x2(v[2] * s, v[3] * s),
                                              pack_bf16x2(v[4] * s, v[5] * s), pack_bf16x2(v[6] * s, v[7] * s));
  }
};

// One warp per row.  VEC: rows are 16B-aligned and dim is a multiple of the vector width.
// MAXV: number of 16B vectors per lane kept in registers between the two passes (dim <= 32*N*MAXV);
// longer rows are re-read for the optional output pass.
template <typename TI, typename TO, bool VEC>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
l2norm_rows_kernel(const TI* __restrict__ in, int64_t rows, int dim, float eps,
                   float* __restrict__ inv_norm, TO* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* src = in + row * dim;
  constexpr int N = Vec16<TI>::N;
  constexpr int MAXV = 4;
  float keep[MAXV][N];
  float ss = 0.f;
  if (VEC) {
    const int nvec = dim / N;
#pragma unroll
    for (int it = 0; it < MAXV; ++it) {
      const int v = lane + it * 32;
      if (v < nvec) {
        Vec16<TI>::load(src + v * N, keep[it]);
#pragma unroll
        for (int e = 0; e < N; ++e) ss = fmaf(keep[it][e], keep[it][e], ss);
      }
    }
    for (int v = lane + MAXV * 32; v < nvec; v += 32) {
      float tmp[N];
      Vec16<TI>::load(src + v * N, tmp);
#pragma unroll
      for (int e = 0; e < N; ++e) ss = fmaf(tmp[e], tmp[e], ss);
    }
  } else {
    for (int d = lane; d < dim; d += 32) { const float x = to_f32<TI>(src[d]); ss = fmaf(x, x, ss); }
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
  if (lane == 0 && inv_norm != nullptr) inv_norm[row] = inv;
  if (out == nullptr) return;
  TO* dst = out + row * dim;
  if (VEC) {
    const int nvec = dim / N;
#pragma unroll
    for (int it = 0; it < MAXV; ++it) {
      const int v = lane + it * 32;
      if (v < nvec) StoreN<TO, N>::put(dst + v * N, keep[it], inv);
    }
    for (int v = lane + MAXV * 32; v < nvec; v += 32) {
      float tmp[N];
      Vec16<TI>::load(src + v * N, tmp);
      StoreN<TO, N>::put(dst + v * N, tmp, inv);
    }
  } else {
    for (int d = lane; d < dim; d += 32) store_elem<TO>(dst + d, to_f32<TI>(src[d]) * inv);
  }
}

// dv = inv * (dvhat - vhat * <vhat, dvhat>), vhat = v * inv.   One warp per row.
template <typename T>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
l2norm_bwd_kernel(const T* __restrict__ v, const float* __restrict__ inv_norm,
                  const float* dvhat, int64_t rows, int dim, float* dv) {   // dv may alias dvhat (in place)
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* src = v + row * dim;
  const float* g = dvhat + row * dim;
  const float inv = inv_norm[row];
  float dot = 0.f;
  for (int d = lane; d < dim; d += 32) dot = fmaf(to_f32<T>(src[d]) * inv, g[d], dot);
  dot = warp_sum(dot);
  float* o = dv + row * dim;
  for (int d = lane; d < dim; d += 32) o[d] = inv * (g[d] - to_f32<T>(src[d]) * inv * dot);
}

// Single block.  row_stats [B,4] -> lse[B], loss (mean), pq_norm2.  Arithmetic in double: only B
// rows, and sum_j (p-q)^2 cancels badly in fp32 once the target probability approaches 1.
static __global__ void __launch_bounds__(1024)
loss_kernel(const float* __restrict__ row_stats, int64_t B, float s_eff, float ls_eps,
            double C_total, float* __restrict__ lse_out, float* __restrict__ loss_out,
            float* __restrict__ pq_norm2_out) {
  __shared__ double sh_loss[32], sh_pq[32];
  double loss = 0.0, pq = 0.0;
  const double eps = (double)ls_eps;
  const double q_off = eps / C_total;
  const double q_sq = (1.0 - eps + q_off) * (1.0 - eps + q_off) + (C_total - 1.0) * q_off * q_off;
  for (int64_t r = threadIdx.x; r < B; r += blockDim.x) {
    const float* st = row_stats + r * B200F_STAT_COLS;
    const double se = (double)st[B200F_STAT_SUMEXP];
    const double lse = (double)s_eff + log(se);
    const double zt = (double)st[B200F_STAT_ZTARGET];
    loss += lse - (1.0 - eps) * zt - q_off * (double)st[B200F_STAT_SUMZ];
    const double pt = exp(zt - lse);
    const double p_sq = (double)st[B200F_STAT_SUMEXP2] / (se * se);
    pq += p_sq - 2.0 * ((1.0 - eps) * pt + q_off) + q_sq;
    if (lse_out != nullptr) lse_out[r] = (float)lse;
  }
  // fixed-order tree: bitwise reproducible
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_down_sync(0xffffffffu, loss, o);
    pq += __shfl_down_sync(0xffffffffu, pq, o);
  }
  if ((threadIdx.x & 31) == 0) { sh_loss[threadIdx.x >> 5] = loss; sh_pq[threadIdx.x >> 5] = pq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0, q = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { l += sh_loss[w]; q += sh_pq[w]; }
    if (loss_out != nullptr) *loss_out = (float)(l / (double)B);
    if (pq_norm2_out != nullptr) *pq_norm2_out = (float)fmax(q, 0.0);
  }
}

// src/face_models.py:538-567 on device scalars.
static __global__ void hook_scale_kernel(const float* __restrict__ pq_norm2, const float* __restrict__ upstream,
                                  double B, float s_eff, int hook_enabled, float max_grad_norm,
                                  int phase, int epoch, float* __restrict__ out3) {
  const double up = (upstream != nullptr) ? (double)*upstream : 1.0;
  const double base = (double)s_eff / B;
  const double n = fabs(up) * base * sqrt((double)*pq_norm2);
  double kappa = 1.0;
  if (hook_enabled) {
    double thr = (double)max_grad_norm;
    if (phase == 1) thr = fmin(0.5, (double)max_grad_norm);
    if (epoch < 10) thr = fmin(thr, 0.5 + 0.05 * (double)epoch);
    if (n > 3.0) thr = fmin(thr, 0.5);
    if (n > thr) kappa = thr / (n + 1e-8);
  }
  out3[0] = (float)(up * kappa * base);
  out3[1] = (float)n;
  out3[2] = (float)kappa;
}

}  // namespace rowops
}  // namespace b200f
