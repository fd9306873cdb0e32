// HBM-bound row kernels: K1 fused L2-normalise, normalise-backward, loss finalize, hook scalar.
// Reference: F.normalize(v, p=2, dim=1, eps=1e-12) at src/face_models.py:351-352,525 and its
// autograd; nn.CrossEntropyLoss(label_smoothing) src/training.py:341; hook src/face_models.py:538-567.
//
// Layout: one warp owns ROWS_PER_WARP consecutive rows and issues ALL of their 16-byte loads before the
// first reduction, so each SM keeps >= 64 KB in flight (HBM latency x bandwidth needs ~45 KB per SM).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace b200f {

template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

namespace rowops {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int ROWS_PER_WARP = 4;
constexpr int ROWS_PER_BLOCK = WARPS_PER_BLOCK * ROWS_PER_WARP;   // 32

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
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
};

template <typename TO> __device__ __forceinline__ void store_elem(TO* p, float v);
template <> __device__ __forceinline__ void store_elem<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_elem<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void store_elem<__half>(__half* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// N consecutive outputs (times s) as vector stores; dst is N*sizeof(TO)-aligned
template <typename TO, int N> struct StoreN;
template <int N> struct StoreN<float, N> {
  static __device__ __forceinline__ void put(float* d, const float (&v)[N], float s) {
#pragma unroll
    for (int i = 0; i < N; i += 4)
      *reinterpret_cast<float4*>(d + i) = make_float4(v[i] * s, v[i + 1] * s, v[i + 2] * s, v[i + 3] * s);
  }
};
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
template <> struct StoreN<__half, 4> {
  static __device__ __forceinline__ void put(__half* d, const float (&v)[4], float s) {
    *reinterpret_cast<uint2*>(d) = make_uint2(pack_f16x2(v[0] * s, v[1] * s), pack_f16x2(v[2] * s, v[3] * s));
  }
};
template <> struct StoreN<__half, 8> {
  static __device__ __forceinline__ void put(__half* d, const float (&v)[8], float s) {
    *reinterpret_cast<uint4*>(d) = make_uint4(pack_f16x2(v[0] * s, v[1] * s), pack_f16x2(v[2] * s, v[3] * s),
                                              pack_f16x2(v[4] * s, v[5] * s), pack_f16x2(v[6] * s, v[7] * s));
  }
};

// K1, vector path.  VPL = 16-byte vectors per lane per row (dim <= 32 * N * VPL, dim % N == 0).
template <typename TI, typename TO, int VPL>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_rows_vec_kernel(const TI* __restrict__ in, int64_t rows, int dim, float eps, float out_scale,
                       float* __restrict__ inv_norm, TO* __restrict__ out) {
  pdl_trigger(); pdl_wait();
  constexpr int N = Vec16<TI>::N;
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5)) * ROWS_PER_WARP;
  if (row0 >= rows) return;
  const int nvec = dim / N;
  float val[ROWS_PER_WARP][VPL][N];
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r) {
    const bool rok = row0 + r < rows;
#pragma unroll
    for (int it = 0; it < VPL; ++it) {
      const int v = lane + it * 32;
      if (rok && v < nvec) {
        Vec16<TI>::load(in + (row0 + r) * dim + v * N, val[r][it]);
      } else {
#pragma unroll
        for (int e = 0; e < N; ++e) val[r][it][e] = 0.f;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r) {
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < VPL; ++it)
#pragma unroll
      for (int e = 0; e < N; ++e) ss = fmaf(val[r][it][e], val[r][it][e], ss);
    ss = warp_sum(ss);
    if (row0 + r >= rows) continue;
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    if (lane == 0 && inv_norm != nullptr) inv_norm[row0 + r] = inv;
    if (out != nullptr) {
      const float s = inv * out_scale;
#pragma unroll
      for (int it = 0; it < VPL; ++it) {
        const int v = lane + it * 32;
        if (v < nvec) StoreN<TO, N>::put(out + (row0 + r) * dim + v * N, val[r][it], s);
      }
    }
  }
}

// K1, 512-wide 16-bit rows (the head's shape): each lane owns 32 contiguous bytes of a row -- ONE 256-bit load and
// ONE 256-bit store per lane and row (full 32 B sectors), 4 rows per warp in flight.
// hints (tunable "k1_hints", a bit mask): 1 = the source rows are read once: L2 evict_first; 2 = the operand rows are read
// again by the next kernel: L2 evict_last, so that they stay in the 126 MB L2 as dirty lines -- their write-back to HBM then
// happens under the consumer (K2 has HBM bandwidth to spare) instead of competing with this kernel's reads.
template <typename TI, typename TO>
__device__ __forceinline__ void l2norm_rows_512x16_body(const TI* __restrict__ in, int64_t rows, float eps, float out_scale,
                                                        float* __restrict__ inv_norm, TO* __restrict__ out, int64_t block,
                                                        int hints = 0) {
  static_assert(sizeof(TI) == 2 && sizeof(TO) == 2, "16-bit rows only");
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (block * WARPS_PER_BLOCK + (threadIdx.x >> 5)) * ROWS_PER_WARP;
  if (row0 >= rows) return;
  uint64_t pol_ld = 0, pol_st = 0;
  if (hints & 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_ld));
  if (hints & 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_st));
  uint32_t raw[ROWS_PER_WARP][8];
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r) {
    if (row0 + r < rows) {
      const TI* src = in + (row0 + r) * 512 + lane * 16;
      if (hints & 1)
        asm volatile("ld.global.nc.L2::cache_hint.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
                     : "=r"(raw[r][0]), "=r"(raw[r][1]), "=r"(raw[r][2]), "=r"(raw[r][3]), "=r"(raw[r][4]), "=r"(raw[r][5]),
                       "=r"(raw[r][6]), "=r"(raw[r][7])
                     : "l"(src), "l"(pol_ld));
      else
      asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(raw[r][0]), "=r"(raw[r][1]), "=r"(raw[r][2]), "=r"(raw[r][3]), "=r"(raw[r][4]), "=r"(raw[r][5]),
                     "=r"(raw[r][6]), "=r"(raw[r][7])
                   : "l"(src));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) raw[r][i] = 0u;
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS_PER_WARP; ++r) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const TI* h = reinterpret_cast<const TI*>(&raw[r][i]);
      v[2 * i] = to_f32<TI>(h[0]); v[2 * i + 1] = to_f32<TI>(h[1]);
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) ss = fmaf(v[i], v[i], ss);
    ss = warp_sum(ss);
    if (row0 + r >= rows) continue;
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    if (lane == 0 && inv_norm != nullptr) inv_norm[row0 + r] = inv;
    if (out != nullptr) {
      const float s = inv * out_scale;
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        TO pr[2];
        store_elem<TO>(&pr[0], v[2 * i] * s); store_elem<TO>(&pr[1], v[2 * i + 1] * s);
        o[i] = *reinterpret_cast<uint32_t*>(pr);
      }
      if (hints & 2)
        asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;"
                     ::"l"(out + (row0 + r) * 512 + lane * 16), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]),
                       "r"(o[6]), "r"(o[7]), "l"(pol_st)
                     : "memory");
      else
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   ::"l"(out + (row0 + r) * 512 + lane * 16), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]),
                     "r"(o[6]), "r"(o[7])
                   : "memory");
    }
  }
}
template <typename TI, typename TO>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_rows_512x16_kernel(const TI* __restrict__ in, int64_t rows, float eps, float out_scale,
                          float* __restrict__ inv_norm, TO* __restrict__ out, int hints) {
  pdl_trigger(); pdl_wait();
  l2norm_rows_512x16_body<TI, TO>(in, rows, eps, out_scale, inv_norm, out, (int64_t)blockIdx.x, hints);
}
// K1 over the batch rows + clearing a word array for the kernel behind it (the row counters of K1(W)-inside-K2)
template <typename TI, typename TO>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_rows_512x16_zero_kernel(const TI* __restrict__ in, int64_t rows, float eps, float out_scale,
                               float* __restrict__ inv_norm, TO* __restrict__ out, unsigned int* __restrict__ zero, int n_zero) {
  pdl_trigger(); pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += gridDim.x * blockDim.x) zero[i] = 0u;
  l2norm_rows_512x16_body<TI, TO>(in, rows, eps, out_scale, inv_norm, out, (int64_t)blockIdx.x);
}
// Two row sets in ONE launch (the head's K1 over the batch rows and over the class weights: the 16-block launch for
// x otherwise costs a launch latency of its own in front of the 3125-block launch for W).  Blocks [0, blocks0) take
// set 0, the rest set 1.
template <typename TI, typename TO>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_rows_512x16_pair_kernel(const TI* __restrict__ in0, int64_t rows0, float* __restrict__ inv0, TO* __restrict__ out0,
                               int blocks0, const TI* __restrict__ in1, int64_t rows1, float* __restrict__ inv1,
                               TO* __restrict__ out1, float eps, float out_scale, int hints1) {
  pdl_trigger(); pdl_wait();
  if ((int)blockIdx.x < blocks0) l2norm_rows_512x16_body<TI, TO>(in0, rows0, eps, out_scale, inv0, out0, (int64_t)blockIdx.x);
  else l2norm_rows_512x16_body<TI, TO>(in1, rows1, eps, out_scale, inv1, out1, (int64_t)blockIdx.x - blocks0, hints1);
}

// K1, generic path (any dim / alignment): one warp per row, scalar accesses.
template <typename TI, typename TO>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_rows_generic_kernel(const TI* __restrict__ in, int64_t rows, int dim, float eps, float out_scale,
                           float* __restrict__ inv_norm, TO* __restrict__ out) {
  pdl_trigger(); pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* src = in + row * dim;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) { const float x = to_f32<TI>(src[d]); ss = fmaf(x, x, ss); }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
  if (lane == 0 && inv_norm != nullptr) inv_norm[row] = inv;
  if (out == nullptr) return;
  const float s = inv * out_scale;
  for (int d = lane; d < dim; d += 32) store_elem<TO>(out + row * dim + d, to_f32<TI>(src[d]) * s);
}

// dv = inv * (dvhat - vhat * <vhat, dvhat>),  vhat = v * vmul  (vmul = inv_norm for raw rows, 1/v_scale for
// pre-normalised fp16 rows).  Vector path: one warp per row, all loads issued before the reduction.
// dv may alias dvhat (in place): every element is read and written by the same lane.
template <typename T, bool PRENORM, int VPL>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_bwd_vec_kernel(const T* __restrict__ v, float v_scale, const float* __restrict__ inv_norm, const float* dvhat,
                      int64_t rows, int dim, float* dv, __nv_bfloat16* __restrict__ dv_bf16) {
  pdl_trigger(); pdl_wait();
  constexpr int N = Vec16<T>::N;                // 8 (16-bit) or 4 (fp32) elements per 16-byte vector of v
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = dim / N;
  const float inv = inv_norm[row];
  const float vmul = PRENORM ? (1.0f / v_scale) : inv;
  float xv[VPL][N], gv[VPL][N];
#pragma unroll
  for (int it = 0; it < VPL; ++it) {
    const int k = lane + it * 32;
    if (k < nvec) {
      Vec16<T>::load(v + row * dim + k * N, xv[it]);
#pragma unroll
      for (int e = 0; e < N; e += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(dvhat + row * dim + k * N + e);
        gv[it][e] = g4.x; gv[it][e + 1] = g4.y; gv[it][e + 2] = g4.z; gv[it][e + 3] = g4.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < N; ++e) { xv[it][e] = 0.f; gv[it][e] = 0.f; }
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int it = 0; it < VPL; ++it)
#pragma unroll
    for (int e = 0; e < N; ++e) { xv[it][e] *= vmul; dot = fmaf(xv[it][e], gv[it][e], dot); }
  dot = warp_sum(dot);
#pragma unroll
  for (int it = 0; it < VPL; ++it) {
    const int k = lane + it * 32;
    if (k < nvec) {
#pragma unroll
      for (int e = 0; e < N; e += 4) {
        const float4 o = make_float4(inv * (gv[it][e] - xv[it][e] * dot), inv * (gv[it][e + 1] - xv[it][e + 1] * dot),
                                     inv * (gv[it][e + 2] - xv[it][e + 2] * dot), inv * (gv[it][e + 3] - xv[it][e + 3] * dot));
        *reinterpret_cast<float4*>(dv + row * dim + k * N + e) = o;
        if (dv_bf16 != nullptr)                               // the autograd cast of dL/dx to a bf16 input's dtype, fused
          *reinterpret_cast<uint2*>(dv_bf16 + row * dim + k * N + e) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    }
  }
}

template <typename T, bool PRENORM>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
l2norm_bwd_generic_kernel(const T* __restrict__ v, float v_scale, const float* __restrict__ inv_norm,
                          const float* dvhat, int64_t rows, int dim, float* dv, __nv_bfloat16* __restrict__ dv_bf16) {
  pdl_trigger(); pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* src = v + row * dim;
  const float* g = dvhat + row * dim;
  const float inv = inv_norm[row];
  const float vmul = PRENORM ? (1.0f / v_scale) : inv;
  float dot = 0.f;
  for (int d = lane; d < dim; d += 32) dot = fmaf(to_f32<T>(src[d]) * vmul, g[d], dot);
  dot = warp_sum(dot);
  float* o = dv + row * dim;
  for (int d = lane; d < dim; d += 32) {
    const float val = inv * (g[d] - to_f32<T>(src[d]) * vmul * dot);
    o[d] = val;
    if (dv_bf16 != nullptr) dv_bf16[row * dim + d] = __float2bfloat16_rn(val);
  }
}

// Host-side dispatch helpers (used by runtime.cu and by the tcgen05 engine).
template <typename TI, typename TO>
static inline void launch_l2norm_rows(const TI* in, int64_t rows, int dim, float eps, float out_scale, float* inv_norm,
                                      TO* out, cudaStream_t st) {
  constexpr int N = Vec16<TI>::N;
  if constexpr (sizeof(TI) == 2 && sizeof(TO) == 2) {
    if (dim == 512 && reinterpret_cast<uintptr_t>(in) % 32 == 0 && (!out || reinterpret_cast<uintptr_t>(out) % 32 == 0)) {
      // rows that fill a good part of the L2 are class weights on their way into K2: cache hints as the tunable says
      launch_pdl(l2norm_rows_512x16_kernel<TI, TO>, dim3((unsigned)ceil_div(rows, ROWS_PER_BLOCK)), dim3(WARPS_PER_BLOCK * 32), 0, st,
          in, rows, eps, out_scale, inv_norm, out, (out != nullptr && rows >= 16384) ? k1_hints() : 0);
      return;
    }
  }
  bool vec = (dim % N == 0) && (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (dim <= 32 * N * 4);
  if (out) vec = vec && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  if (vec) {
    const int nvec = dim / N;
    const unsigned grid = (unsigned)ceil_div(rows, ROWS_PER_BLOCK);
    if (nvec <= 32) launch_pdl(l2norm_rows_vec_kernel<TI, TO, 1>, dim3(grid), dim3(WARPS_PER_BLOCK * 32), 0, st, in, rows, dim, eps, out_scale, inv_norm, out);
    else if (nvec <= 64) launch_pdl(l2norm_rows_vec_kernel<TI, TO, 2>, dim3(grid), dim3(WARPS_PER_BLOCK * 32), 0, st, in, rows, dim, eps, out_scale, inv_norm, out);
    else launch_pdl(l2norm_rows_vec_kernel<TI, TO, 4>, dim3(grid), dim3(WARPS_PER_BLOCK * 32), 0, st, in, rows, dim, eps, out_scale, inv_norm, out);
  } else {
    launch_pdl(l2norm_rows_generic_kernel<TI, TO>, dim3((unsigned)ceil_div(rows, WARPS_PER_BLOCK)), dim3(WARPS_PER_BLOCK * 32), 0, st, 
        in, rows, dim, eps, out_scale, inv_norm, out);
  }
}

template <typename T, bool PRENORM>
static inline void launch_l2norm_bwd(const T* v, float v_scale, const float* inv_norm, const float* dvhat, int64_t rows,
                                     int dim, float* dv, cudaStream_t st, __nv_bfloat16* dv_bf16 = nullptr) {
  constexpr int N = Vec16<T>::N;
  const bool vec = (dim % N == 0) && (dim % 4 == 0) && (reinterpret_cast<uintptr_t>(v) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(dvhat) % 16 == 0) && (reinterpret_cast<uintptr_t>(dv) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(dv_bf16) % 8 == 0) && (dim <= 32 * N * 4);
  const unsigned grid = (unsigned)ceil_div(rows, WARPS_PER_BLOCK);
  if (vec) {
    const int nvec = dim / N;
    if (nvec <= 32) launch_pdl(l2norm_bwd_vec_kernel<T, PRENORM, 1>, dim3(grid), dim3(WARPS_PER_BLOCK * 32), 0, st, v, v_scale, inv_norm, dvhat, rows, dim, dv, dv_bf16);
    else if (nvec <= 64) launch_pdl(l2norm_bwd_vec_kernel<T, PRENORM, 2>, dim3(grid), dim3(WARPS_PER_BLOCK * 32), 0, st, v, v_scale, inv_norm, dvhat, rows, dim, dv, dv_bf16);
    else launch_pdl(l2norm_bwd_vec_kernel<T, PRENORM, 4>, dim3(grid), dim3(WARPS_PER_BLOCK * 32), 0, st, v, v_scale, inv_norm, dvhat, rows, dim, dv, dv_bf16);
  } else {
    launch_pdl(l2norm_bwd_generic_kernel<T, PRENORM>, dim3(grid), dim3(WARPS_PER_BLOCK * 32), 0, st, v, v_scale, inv_norm, dvhat, rows, dim, dv, dv_bf16);
  }
}

// K5 -- AdamW (+ AMSGrad) step of the class-weight rows fused with next step's K1 (the reference
// trains the head with torch.optim.AdamW(amsgrad=True), src/training.py:343-348, after an optional clip_grad_norm_,
// :528-533).  Per element, in torch's _single_tensor_adamw order:
//   p *= 1 - lr * wd;  m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g;  vmax = max(vmax, v)
//   p -= (lr / bc1) * m / (sqrt(vmax or v) / sqrt(bc2) + eps)            with g = dw * grad_scale (the clip coefficient)
// One warp per row with the whole updated row in registers (dim <= 1024, dim % 4 == 0), so the row norm of the NEW
// weights is at hand and w_hat16 = w * inv_norm * out_scale and inv_norm leave in the same pass: 9 fp32 streams
// + 2 B per element (38 B) against 36 B for an unfused AdamW + 6 B for K1 -- and no second kernel on the step's
// critical path.
struct AdamWParams {               // scalars rounded to fp32 from the double expressions torch evaluates on the host
  float decay, one_m_b1, beta2, one_m_b2, step_size, bc2_sqrt, eps;
  const float* grad_scale;    // device scalar multiplied into the gradient, or NULL
  float norm_eps, out_scale;
};

template <bool AMSGRAD, int NV>                               // NV float4 per lane
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
adamw_rows_kernel(float* __restrict__ w, const float* __restrict__ dw, float* __restrict__ m, float* __restrict__ v,
                  float* __restrict__ vmax, int64_t rows, int dim, AdamWParams p, __half* __restrict__ w_hat,
                  float* __restrict__ inv_norm) {
  pdl_trigger(); pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = dim >> 2;
  const float gs = p.grad_scale ? __ldg(p.grad_scale) : 1.0f;
  const float decay = p.decay;
  const float step_size = p.step_size;
  float4 wv[NV], gv[NV], mv[NV], vv[NV], xv[NV];
  const int64_t base = row * (int64_t)dim;
#pragma unroll
  for (int i = 0; i < NV; ++i) {                              // all loads in flight before the first use
    const int c = lane + 32 * i;
    if (c < nvec) {
      wv[i] = *reinterpret_cast<const float4*>(w + base + 4 * c);
      gv[i] = __ldg(reinterpret_cast<const float4*>(dw + base + 4 * c));
      mv[i] = *reinterpret_cast<const float4*>(m + base + 4 * c);
      vv[i] = *reinterpret_cast<const float4*>(v + base + 4 * c);
      if (AMSGRAD) xv[i] = *reinterpret_cast<const float4*>(vmax + base + 4 * c);
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      float* pw = reinterpret_cast<float*>(&wv[i]); float* pg = reinterpret_cast<float*>(&gv[i]);
      float* pm = reinterpret_cast<float*>(&mv[i]); float* pv = reinterpret_cast<float*>(&vv[i]);
      float* px = reinterpret_cast<float*>(&xv[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float g = pg[e] * gs;
        float wn = pw[e] * decay;
        const float mn = pm[e] + (g - pm[e]) * p.one_m_b1;
        const float vn = fmaf(pv[e], p.beta2, p.one_m_b2 * g * g);         // mul_(b2).addcmul_(g, g, 1 - b2)
        float den;
        if (AMSGRAD) { const float xn = fmaxf(px[e], vn); px[e] = xn; den = sqrtf(xn) / p.bc2_sqrt + p.eps; }
        else den = sqrtf(vn) / p.bc2_sqrt + p.eps;
        wn = wn - step_size * (mn / den);
        pw[e] = wn; pm[e] = mn; pv[e] = vn;
        ss = fmaf(wn, wn, ss);
      }
      *reinterpret_cast<float4*>(w + base + 4 * c) = wv[i];
      *reinterpret_cast<float4*>(m + base + 4 * c) = mv[i];
      *reinterpret_cast<float4*>(v + base + 4 * c) = vv[i];
      if (AMSGRAD) *reinterpret_cast<float4*>(vmax + base + 4 * c) = xv[i];
    }
  }
  if (w_hat == nullptr && inv_norm == nullptr) return;
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), p.norm_eps);
  if (lane == 0 && inv_norm != nullptr) inv_norm[row] = inv;
  if (w_hat != nullptr) {
    const float s = inv * p.out_scale;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const __half2 lo = __floats2half2_rn(wv[i].x * s, wv[i].y * s), hi = __floats2half2_rn(wv[i].z * s, wv[i].w * s);
        uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(w_hat + base + 4 * c) = o;
      }
    }
  }
}

template <bool AMSGRAD>
static inline bool launch_adamw_rows(float* w, const float* dw, float* m, float* v, float* vmax, int64_t rows, int dim,
                                     const AdamWParams& p, __half* w_hat, float* inv_norm, cudaStream_t st) {
  const int nvec = dim >> 2;
  const unsigned grid = (unsigned)ceil_div(rows, WARPS_PER_BLOCK);
  const dim3 blk(WARPS_PER_BLOCK * 32);
  if (nvec <= 32) launch_pdl(adamw_rows_kernel<AMSGRAD, 1>, dim3(grid), blk, 0, st, w, dw, m, v, vmax, rows, dim, p, w_hat, inv_norm);
  else if (nvec <= 64) launch_pdl(adamw_rows_kernel<AMSGRAD, 2>, dim3(grid), blk, 0, st, w, dw, m, v, vmax, rows, dim, p, w_hat, inv_norm);
  else if (nvec <= 128) launch_pdl(adamw_rows_kernel<AMSGRAD, 4>, dim3(grid), blk, 0, st, w, dw, m, v, vmax, rows, dim, p, w_hat, inv_norm);
  else if (nvec <= 256) launch_pdl(adamw_rows_kernel<AMSGRAD, 8>, dim3(grid), blk, 0, st, w, dw, m, v, vmax, rows, dim, p, w_hat, inv_norm);
  else return false;
  return true;
}

// src/face_models.py:538-567 on scalars: out4 = {grad_scale, n, kappa, g_scale}.
struct HookCfg { int enabled; float max_grad_norm; int phase, epoch; };
__device__ __forceinline__ void hook_scalars(double pq_norm2, double up, double B, float s_eff, const HookCfg& h,
                                             float* __restrict__ out4) {
  const double base = (double)s_eff / B;
  const double n = fabs(up) * base * sqrt(pq_norm2);
  double kappa = 1.0;
  if (h.enabled) {
    double thr = (double)h.max_grad_norm;
    if (h.phase == 1) thr = fmin(0.5, (double)h.max_grad_norm);
    if (h.epoch < 10) thr = fmin(thr, 0.5 + 0.05 * (double)h.epoch);
    if (n > 3.0) thr = fmin(thr, 0.5);
    if (n > thr) kappa = thr / (n + 1e-8);
  }
  const double gs = up * kappa * base;
  out4[0] = (float)gs;
  out4[1] = (float)n;
  out4[2] = (float)kappa;
  // power of two that puts |grad_scale| * g_scale in (512, 1024]: the fp16 range centring of the
  // tcgen05 engine's logit-gradient buffer (exact to undo)
  out4[3] = (gs != 0.0 && isfinite(gs)) ? (float)exp2(10.0 - ceil(log2(fabs(gs)))) : 1.0f;
}

// One block: row_stats [B,4] -> lse[B], loss (mean), pq_norm2 and (hook_out4 != NULL) the hook scalars for an
// upstream gradient of 1.  Arithmetic in double: only B rows, and sum_j (p-q)^2 cancels badly in fp32 once the
// target probability approaches 1.  Reads bypass L1 (__ldcg): the fused caller runs this in the LAST block of a grid
// whose other blocks wrote row_stats.  Fixed-order tree: bitwise reproducible.
__device__ __forceinline__ void loss_block(const float* row_stats, int64_t B, float s_eff, float ls_eps, double C_total,
                                           float* __restrict__ lse_out, float* __restrict__ loss_out,
                                           float* __restrict__ pq_norm2_out, const HookCfg* hook, float* __restrict__ hook_out4,
                                           double* sh_loss, double* sh_pq) {
  double loss = 0.0, pq = 0.0;
  const double eps = (double)ls_eps;
  const double q_off = eps / C_total;
  const double q_sq = (1.0 - eps + q_off) * (1.0 - eps + q_off) + (C_total - 1.0) * q_off * q_off;
  for (int64_t r = threadIdx.x; r < B; r += blockDim.x) {
    const float4 st4 = __ldcg(reinterpret_cast<const float4*>(row_stats + r * B200F_STAT_COLS));
    const double se = (double)st4.x;                          // B200F_STAT_SUMEXP, _SUMEXP2, _ZTARGET, _SUMZ
    const double lse = (double)s_eff + log(se);
    const double zt = (double)st4.z;
    loss += lse - (1.0 - eps) * zt - q_off * (double)st4.w;
    const double pt = exp(zt - lse);
    const double p_sq = (double)st4.y / (se * se);
    pq += p_sq - 2.0 * ((1.0 - eps) * pt + q_off) + q_sq;
    if (lse_out != nullptr) lse_out[r] = (float)lse;
  }
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_down_sync(0xffffffffu, loss, o);
    pq += __shfl_down_sync(0xffffffffu, pq, o);
  }
  if ((threadIdx.x & 31) == 0) { sh_loss[threadIdx.x >> 5] = loss; sh_pq[threadIdx.x >> 5] = pq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0, q = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { l += sh_loss[w]; q += sh_pq[w]; }
    q = fmax(q, 0.0);
    if (loss_out != nullptr) *loss_out = (float)(l / (double)B);
    if (pq_norm2_out != nullptr) *pq_norm2_out = (float)q;
    if (hook_out4 != nullptr) hook_scalars((double)(float)q, 1.0, (double)B, s_eff, *hook, hook_out4);
  }
}
static_assert(B200F_STAT_SUMEXP == 0 && B200F_STAT_SUMEXP2 == 1 && B200F_STAT_ZTARGET == 2 && B200F_STAT_SUMZ == 3 &&
              B200F_STAT_COLS == 4, "loss_block reads a row of statistics as one float4");

static __global__ void __launch_bounds__(1024)
loss_kernel(const float* __restrict__ row_stats, int64_t B, float s_eff, float ls_eps,
            double C_total, float* __restrict__ lse_out, float* __restrict__ loss_out,
            float* __restrict__ pq_norm2_out, HookCfg hook, float* __restrict__ hook_out4) {
  pdl_trigger(); pdl_wait();
  __shared__ double sh_loss[32], sh_pq[32];
  loss_block(row_stats, B, s_eff, ls_eps, C_total, lse_out, loss_out, pq_norm2_out, &hook, hook_out4, sh_loss, sh_pq);
}

static __global__ void hook_scale_kernel(const float* __restrict__ pq_norm2, const float* __restrict__ upstream,
                                         double B, float s_eff, HookCfg hook, float* __restrict__ out4) {
  pdl_trigger(); pdl_wait();
  const double up = (upstream != nullptr) ? (double)*upstream : 1.0;
  hook_scalars((double)*pq_norm2, up, B, s_eff, hook, out4);
}

// dx_hat[row] = scale * sum_s part[s][row]  followed at once by the normalise-backward of that row of x
// (dx = inv_nx (dx_hat - x_hat <x_hat, dx_hat>), autograd of F.normalize, src/face_models.py:351): one warp per row,
// the row never leaves registers between the two.  v = the rows the projection uses: the RAW input rows (x_hat = v *
// inv_nx, exact to fp32 -- preferred: K1's fp16 copy carries a 2^-12 rounding that the projection of a row nearly
// parallel to its class centre amplifies ~3x) or, PRENORM, K1's fp16 rows (x_hat * v_scale).  Optional outputs: dxhat
// (the un-projected sum), dx_bf16 (the cast autograd applies for a bf16 input).  dim % 8 == 0, dim <= 512.
template <typename T, bool PRENORM>
static __global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
reduce_splits_normbwd_kernel(const float* __restrict__ part, int n_splits, int64_t rows, int dim, float scale,
                             const float* __restrict__ dev_scale, const T* __restrict__ v, float v_scale,
                             const float* __restrict__ inv_nx, float* __restrict__ dxhat, float* __restrict__ dx,
                             __nv_bfloat16* __restrict__ dx_bf16) {
  pdl_trigger(); pdl_wait();
  constexpr int N = Vec16<T>::N;                              // elements per 16-byte vector of v: 4 (fp32) or 8
  constexpr int NV = 512 / (32 * N);                          // vectors per lane: 4 or 2
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float sc = (dev_scale != nullptr) ? scale / __ldg(dev_scale) : scale;
  const int nvec = dim / N;
  const int64_t n = rows * (int64_t)dim;
  const float inv = __ldg(inv_nx + row);
  const float vmul = PRENORM ? (1.0f / v_scale) : inv;
  float g[NV][N], xh[NV][N];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int e = 0; e < N; ++e) { g[i][e] = 0.f; xh[i][e] = 0.f; }
  }
  // The split partials: FOUR splits per trip, all of their 16-byte loads (4 x NV x N / 4 = 16 per lane) issued before the
  // first add.  (One split per trip left two loads in flight per lane: 512 warps x 1 KB against ~0.6 us of L2 latency is
  // 0.9 TB/s -- 17 us for the 19 MB of cfg3's 18 splits, at the very end of the step where nothing overlaps it.)  Every
  // element is still summed over k in ascending order: the same bits.
  constexpr int KU = 4;
  for (int k0 = 0; k0 < n_splits; k0 += KU) {
    float4 pv[KU][NV][N / 4];
#pragma unroll
    for (int kk = 0; kk < KU; ++kk) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
#pragma unroll
        for (int e = 0; e < N; e += 4) {
          pv[kk][i][e / 4] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < nvec && k0 + kk < n_splits)
            pv[kk][i][e / 4] = __ldcg(reinterpret_cast<const float4*>(part + (int64_t)(k0 + kk) * n + row * dim + (int64_t)c * N + e));
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < KU; ++kk) {
      if (k0 + kk < n_splits) {                                // (a skipped split must not even add +0: -0 + 0 = +0)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
#pragma unroll
          for (int e = 0; e < N; e += 4) {
            const float4 q = pv[kk][i][e / 4];
            g[i][e] += q.x; g[i][e + 1] += q.y; g[i][e + 2] += q.z; g[i][e + 3] += q.w;
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const int64_t off = row * dim + (int64_t)c * N;
      Vec16<T>::load(v + off, xh[i]);
#pragma unroll
      for (int e = 0; e < N; ++e) { g[i][e] *= sc; xh[i][e] *= vmul; dot = fmaf(xh[i][e], g[i][e], dot); }
      if (dxhat != nullptr) {
#pragma unroll
        for (int e = 0; e < N; e += 4)
          *reinterpret_cast<float4*>(dxhat + off + e) = make_float4(g[i][e], g[i][e + 1], g[i][e + 2], g[i][e + 3]);
      }
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const int64_t off = row * dim + (int64_t)c * N;
#pragma unroll
      for (int e = 0; e < N; e += 4) {
        const float4 o = make_float4(inv * (g[i][e] - xh[i][e] * dot), inv * (g[i][e + 1] - xh[i][e + 1] * dot),
                                     inv * (g[i][e + 2] - xh[i][e + 2] * dot), inv * (g[i][e + 3] - xh[i][e + 3] * dot));
        *reinterpret_cast<float4*>(dx + off + e) = o;
        if (dx_bf16 != nullptr)
          *reinterpret_cast<uint2*>(dx_bf16 + off + e) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    }
  }
}

// =====================================================================================================================
// The embedding tail in front of the head (src/face_models.py:515-525):
//     z = embedding(features)            (Linear 512 -> 512, no bias: stays a library GEMM)
//     b = BatchNorm1d(z)                 (train: batch statistics + running-stat update; eval: running statistics)
//     y = dropout(b)                     (train only; the keep mask is an input: the caller owns the generator)
//     emb = F.normalize(y)               (and the head normalises its input again, :351 -- the same values)
// fused as  [bn_stats] -> tail_fwd: ONE pass over z that applies BN and the mask, forms the row norm and emits what the
// head's K1 would: the fp16 operand rows y_hat * S, 1 / |y|, plus y itself (the rows the backward projects with).
// Backward of dropout + BatchNorm: tail_bwd_cols (column sums d_beta, d_gamma) -> tail_bwd_apply.
// =====================================================================================================================

// Column statistics of z [rows, dim] (train mode).  One block per 32 columns, 32 x 8 threads, rows strided over ty;
// double accumulators (B <= a few thousand rows: cheap, and the variance is a difference of large numbers in fp32).
// mean_out / invstd_out [dim]; running_mean / running_var updated in place with torch's rule
// (momentum; the running variance takes the UNBIASED batch variance).
template <typename T>
static __global__ void __launch_bounds__(256)
bn_stats_kernel(const T* __restrict__ z, int64_t rows, int dim, float eps, float momentum, float* running_mean,
                float* running_var, float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  pdl_trigger(); pdl_wait();
  __shared__ double sh_s[8][33], sh_q[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  double s = 0.0, q = 0.0;
  if (j < dim)
    for (int64_t r = ty; r < rows; r += 8) { const double v = (double)to_f32<T>(z[r * dim + j]); s += v; q += v * v; }
  sh_s[ty][tx] = s; sh_q[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && j < dim) {
    for (int k = 1; k < 8; ++k) { s += sh_s[k][tx]; q += sh_q[k][tx]; }
    const double n = (double)rows;
    const double mean = s / n;
    double var = q / n - mean * mean; if (var < 0.0) var = 0.0;
    mean_out[j] = (float)mean;
    invstd_out[j] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) running_mean[j] = (float)((1.0 - momentum) * (double)running_mean[j] + momentum * mean);
    if (running_var != nullptr) {
      const double unb = rows > 1 ? var * n / (n - 1.0) : var;
      running_var[j] = (float)((1.0 - momentum) * (double)running_var[j] + momentum * unb);
    }
  }
}

// mean / invstd come from bn_stats_kernel (train) or are formed from the running statistics here (eval: stat_is_var).
// mask: uint8 keep mask [rows, dim] or NULL; keep_scale = 1 / (1 - p).  Outputs (each optional but inv_norm):
// y fp32 [rows, dim], y_hat16 = fp16(y / |y| * out_scale), emb = y / |y| fp32.  One warp per row, dim % 4 == 0, dim <= 1024.
template <typename T, int NV>
static __global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
tail_fwd_kernel(const T* __restrict__ z, int64_t rows, int dim, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ mean, const float* __restrict__ stat, int stat_is_var, float bn_eps,
                const uint8_t* __restrict__ mask, float keep_scale, float norm_eps, float out_scale,
                float* __restrict__ y_out, __half* __restrict__ yhat16, float* __restrict__ emb, float* __restrict__ inv_norm) {
  pdl_trigger(); pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = dim >> 2;
  float4 y[NV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    y[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nvec) {
      const int64_t off = row * dim + 4 * c;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = to_f32<T>(z[off + e]);
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean + 4 * c));
      float4 s4 = __ldg(reinterpret_cast<const float4*>(stat + 4 * c));
      if (stat_is_var) s4 = make_float4(rsqrtf(s4.x + bn_eps), rsqrtf(s4.y + bn_eps), rsqrtf(s4.z + bn_eps), rsqrtf(s4.w + bn_eps));
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + 4 * c));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + 4 * c));
      float o[4] = {fmaf((v[0] - m4.x) * s4.x, g4.x, b4.x), fmaf((v[1] - m4.y) * s4.y, g4.y, b4.y),
                    fmaf((v[2] - m4.z) * s4.z, g4.z, b4.z), fmaf((v[3] - m4.w) * s4.w, g4.w, b4.w)};
      if (mask != nullptr) {
        const uchar4 k4 = *reinterpret_cast<const uchar4*>(mask + off);
        o[0] = k4.x ? o[0] * keep_scale : 0.f; o[1] = k4.y ? o[1] * keep_scale : 0.f;
        o[2] = k4.z ? o[2] * keep_scale : 0.f; o[3] = k4.w ? o[3] * keep_scale : 0.f;
      }
      y[i] = make_float4(o[0], o[1], o[2], o[3]);
      ss = fmaf(o[0], o[0], ss); ss = fmaf(o[1], o[1], ss); ss = fmaf(o[2], o[2], ss); ss = fmaf(o[3], o[3], ss);
      if (y_out != nullptr) *reinterpret_cast<float4*>(y_out + off) = y[i];
    }
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), norm_eps);
  if (lane == 0) inv_norm[row] = inv;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const int64_t off = row * dim + 4 * c;
      if (emb != nullptr) *reinterpret_cast<float4*>(emb + off) = make_float4(y[i].x * inv, y[i].y * inv, y[i].z * inv, y[i].w * inv);
      if (yhat16 != nullptr) {
        const float sc = inv * out_scale;
        *reinterpret_cast<uint2*>(yhat16 + off) = make_uint2(pack_f16x2(y[i].x * sc, y[i].y * sc), pack_f16x2(y[i].z * sc, y[i].w * sc));
      }
    }
  }
}

// Column sums of the BatchNorm backward: d_beta_j = sum_b d_b, d_gamma_j = sum_b d_b * xbn_bj with d = dy * mask * keep_scale
// (gradient at the BatchNorm output) and xbn = (z - mean) * invstd.  Same geometry as bn_stats_kernel.
template <typename T>
static __global__ void __launch_bounds__(256)
tail_bwd_cols_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ mask, float keep_scale, const T* __restrict__ z,
                     const float* __restrict__ mean, const float* __restrict__ invstd, int64_t rows, int dim,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_trigger(); pdl_wait();
  __shared__ double sh_b[8][33], sh_g[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  double sb = 0.0, sg = 0.0;
  if (j < dim) {
    const float mu = mean[j], is = invstd[j];
    for (int64_t r = ty; r < rows; r += 8) {
      float d = dy[r * dim + j];
      if (mask != nullptr) d = mask[r * dim + j] ? d * keep_scale : 0.f;
      sb += (double)d; sg += (double)d * (double)((to_f32<T>(z[r * dim + j]) - mu) * is);
    }
  }
  sh_b[ty][tx] = sb; sh_g[ty][tx] = sg;
  __syncthreads();
  if (ty == 0 && j < dim) {
    for (int k = 1; k < 8; ++k) { sb += sh_b[k][tx]; sg += sh_g[k][tx]; }
    dbeta[j] = (float)sb; dgamma[j] = (float)sg;
  }
}

// dz = invstd * gamma * (d - d_beta / B - xbn * d_gamma / B)  (train: batch statistics)  |  invstd * gamma * d  (eval)
template <typename T>
static __global__ void __launch_bounds__(256)
tail_bwd_apply_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ mask, float keep_scale, const T* __restrict__ z,
                      const float* __restrict__ mean, const float* __restrict__ stat, int stat_is_var, float bn_eps,
                      const float* __restrict__ gamma, const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                      int batch_stats, int64_t rows, int dim, float* __restrict__ dz) {
  pdl_trigger(); pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (int64_t)dim) return;
  const int j = (int)(i % dim);
  float d = dy[i];
  if (mask != nullptr) d = mask[i] ? d * keep_scale : 0.f;
  const float is = stat_is_var ? rsqrtf(stat[j] + bn_eps) : stat[j];
  float t = d;
  if (batch_stats) {
    const float xbn = (to_f32<T>(z[i]) - mean[j]) * is;
    const float inv_b = 1.0f / (float)rows;
    t = d - dbeta[j] * inv_b - xbn * dgamma[j] * inv_b;
  }
  dz[i] = is * gamma[j] * t;
}

}  // namespace rowops
}  // namespace b200f
