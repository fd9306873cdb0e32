// fp32 CUDA-core GEMM tile engine (128x128x16, 256 threads, 8x8 per thread).
//
// This is the "exact" engine: fp32 products and fp32 accumulation, used for fp32 inputs (the
// north star's 1e-5 bar cannot be met with bf16/tf32 tensor-core products) and for shapes the
// tcgen05 engine does not take.  Every GEMM-shaped stage of the head and of the gallery match
// is built from tile_mainloop() plus a stage-specific epilogue.
//
//   C[m,n] = sum_k op(A(m,k), B(n,k))
//   A(m,k) = A_KMAJOR ? A[m*lda + k] : A[k*lda + m]      (same for B with n)
#pragma once
#include "common.cuh"

namespace b200f {
namespace simt {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;
constexpr int THREADS = 256;
constexpr int PAD = 4;

struct __align__(16) Smem {
  float a[BK][BM + PAD];
  float b[BK][BN + PAD];
};

struct OpFma {
  static __device__ __forceinline__ float apply(float a, float b, float acc) { return fmaf(a, b, acc); }
};
// || a - b + 1e-6 ||^2 term by term, exactly as F.pairwise_distance forms it (src/app.py:59)
struct OpL2Eps {
  static __device__ __forceinline__ float apply(float a, float b, float acc) {
    float t = (a - b) + 1e-6f;
    return fmaf(t, t, acc);
  }
};

// One operand tile [128 (mn) x 16 (k)] : global -> registers -> shared.
template <typename T, bool KMAJOR>
struct TileLoader {
  const T* base;
  int64_t ld, MN, mn0;
  bool vec_ok;
  const float* kscale;   // optional scale indexed by absolute k

  __device__ __forceinline__ void fetch(int64_t k0, int64_t K_end, float (&v)[8]) const {
    const int t = threadIdx.x;
    if (KMAJOR) {
      const int mn = t & 127, kofs = (t >> 7) * 8;
      const int64_t row = mn0 + mn, k = k0 + kofs;
      int valid = 0;
      if (row < MN && k < K_end) valid = (int)min((int64_t)8, K_end - k);
      if (valid > 0) load8<T>(base + row * ld + k, valid, vec_ok, v);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      if (kscale != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (i < valid) v[i] *= __ldg(kscale + k + i);
      }
    } else {
      const int kk = t >> 4, mnofs = (t & 15) * 8;
      const int64_t k = k0 + kk, col = mn0 + mnofs;
      int valid = 0;
      if (k < K_end && col < MN) valid = (int)min((int64_t)8, MN - col);
      if (valid > 0) load8<T>(base + k * ld + col, valid, vec_ok, v);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      if (kscale != nullptr && valid > 0) {
        const float s = __ldg(kscale + k);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= s;
      }
    }
  }

  __device__ __forceinline__ void stash(float (*s)[BM + PAD], const float (&v)[8]) const {
    const int t = threadIdx.x;
    if (KMAJOR) {
      const int mn = t & 127, kofs = (t >> 7) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) s[kofs + i][mn] = v[i];
    } else {
      const int kk = t >> 4, mnofs = (t & 15) * 8;
      *reinterpret_cast<float4*>(&s[kk][mnofs])     = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&s[kk][mnofs + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
};

// local row / column of accumulator element (i, j) for this thread
__device__ __forceinline__ int acc_row(int i) {
  const int ty = threadIdx.x >> 4;
  return (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
}
__device__ __forceinline__ int acc_col(int j) {
  const int tx = threadIdx.x & 15;
  return (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4));
}

template <class Op>
__device__ __forceinline__ void tile_fma(float (&acc)[8][8], const Smem& sm, int kmax) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (kmax == BK) {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&sm.a[kk][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&sm.a[kk][64 + ty * 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&sm.b[kk][tx * 4]);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&sm.b[kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = Op::apply(a[i], b[j], acc[i][j]);
    }
  } else {
    for (int kk = 0; kk < kmax; ++kk) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&sm.a[kk][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&sm.a[kk][64 + ty * 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&sm.b[kk][tx * 4]);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&sm.b[kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = Op::apply(a[i], b[j], acc[i][j]);
    }
  }
}

// acc = sum_{k in [k_begin,k_end)} op(A(m0+.., k), B(n0+.., k)).  Ends with a __syncthreads()
// so the caller may reuse `sm` immediately.
template <class Op, typename TA, bool A_KM, typename TB, bool B_KM>
__device__ __forceinline__ void tile_mainloop(float (&acc)[8][8], Smem& sm,
                                              const TileLoader<TA, A_KM>& la,
                                              const TileLoader<TB, B_KM>& lb,
                                              int64_t k_begin, int64_t k_end) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[8], rb[8];
  la.fetch(k_begin, k_end, ra);
  lb.fetch(k_begin, k_end, rb);
  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
    la.stash(sm.a, ra);
    lb.stash(sm.b, rb);
    __syncthreads();
    if (k0 + BK < k_end) {          // prefetch the next k-slab while this one is consumed
      la.fetch(k0 + BK, k_end, ra);
      lb.fetch(k0 + BK, k_end, rb);
    }
    const int kmax = (int)min((int64_t)BK, k_end - k0);
    tile_fma<Op>(acc, sm, kmax);
    __syncthreads();
  }
}

template <typename T>
inline bool vec_friendly(const void* p, int64_t ld) {
  const int per16 = 16 / (int)sizeof(T);
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % per16 == 0) && (ld % 8 == 0 || sizeof(T) == 4);
}

}  // namespace simt
}  // namespace b200f
