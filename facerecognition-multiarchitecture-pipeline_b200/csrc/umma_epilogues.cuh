// Epilogue policies of the generic tcgen05 GEMM core (umma_gemm.cuh) and helpers shared with the
// X-stationary kernel's policies (umma_xw_epilogues.cuh).
// Each epilogue thread owns ONE accumulator row (TMEM lane); columns arrive 32 at a time.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma_gemm.cuh"

namespace b200f {
namespace umma {

constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// -------------------------------------------------------------------------------------------------
// Plain store: D (fp32) * scale -> out[split][row_offset + m][n].  Used by the self-test, the dW GEMM
// and the split-K dX GEMM.  dev_scale (optional device scalar) divides the host scale.
struct EpiStore {
  struct Params { float* out; int64_t ld; int64_t split_stride; int64_t row_offset; float scale; const float* dev_scale; };
  static __device__ __forceinline__ void run(const Params& ep, const GemmParams& p, const TileCoord& t,
                                             uint32_t tmem_acc, int quad, int lane, int epi_tid, float* scratch) {
    const int row = t.m0 + quad * 32 + lane;
    const int ncols = min(BLOCK_N, p.N - t.n0);
    const float sc = (ep.dev_scale != nullptr) ? ep.scale / __ldg(ep.dev_scale) : ep.scale;
    float* dst = ep.out + (int64_t)t.split * ep.split_stride + (ep.row_offset + row) * ep.ld + t.n0;
    for (int ch = 0; ch * 32 < ncols; ++ch) {
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + ch * 32, v);
      tmem_ld_wait();
      if (row < p.M) {
        const int cc = min(32, ncols - ch * 32);
        if (cc == 32 && (ep.ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 31) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            st_global_256(dst + ch * 32 + j, __float_as_uint(v[j] * sc), __float_as_uint(v[j + 1] * sc),
                          __float_as_uint(v[j + 2] * sc), __float_as_uint(v[j + 3] * sc), __float_as_uint(v[j + 4] * sc),
                          __float_as_uint(v[j + 5] * sc), __float_as_uint(v[j + 6] * sc), __float_as_uint(v[j + 7] * sc));
        } else if (cc == 32 && (ep.ld % 4 == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + ch * 32 + j) = make_float4(v[j] * sc, v[j + 1] * sc, v[j + 2] * sc, v[j + 3] * sc);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < cc) dst[ch * 32 + j] = v[j] * sc;
        }
      }
    }
    (void)epi_tid; (void)scratch;
  }
};

}  // namespace umma
}  // namespace b200f
