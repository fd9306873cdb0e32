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

// -------------------------------------------------------------------------------------------------
// Coefficients of the normalise-backward of a class row, formed where they are used (this replaced a reduce_r kernel
// between K3a and K3b):  coef_c = { inv_nw_c / (S g_scale),  r_c = sum_rb r_part[rb, c] }  -- K3a leaves one partial of
// r_c = sum_b G_bc cos_bc per (row group, column half); the sum runs in a fixed order (bitwise reproducible).
struct CoefSrc {
  const float* r_part; int n_rb; int64_t ldr;   // [n_rb, ldr], classes of THIS chunk
  const float* inv_nw;                          // [classes of the launch], indexed by the class id of the whole call
  const float* grad4; float S;                  // grad4[3] = g_scale
  __device__ __forceinline__ float inv_sg() const { return 1.0f / (S * __ldg(grad4 + 3)); }
  // c_chunk: class within the chunk, c_call: class within the call (= chunk offset + c_chunk)
  __device__ __forceinline__ float2 load(int64_t c_chunk, int64_t c_call, float inv_sg_v) const {
    const float inw = __ldg(inv_nw + c_call);
    float s = 0.f;
    int rb = 0;
    for (; rb + 4 <= n_rb; rb += 4) {                       // four loads in flight, summed in ascending order
      const float a = __ldg(r_part + (int64_t)rb * ldr + c_chunk), b = __ldg(r_part + (int64_t)(rb + 1) * ldr + c_chunk);
      const float c = __ldg(r_part + (int64_t)(rb + 2) * ldr + c_chunk), d = __ldg(r_part + (int64_t)(rb + 3) * ldr + c_chunk);
      s = (((s + a) + b) + c) + d;
    }
    for (; rb < n_rb; ++rb) s += __ldg(r_part + (int64_t)rb * ldr + c_chunk);
    return make_float2(inw * inv_sg_v, s);
  }
};

// -------------------------------------------------------------------------------------------------
// dW GEMM with both operands streamed (batch > 512): A = G^T rows, so an epilogue thread owns a class row and 256 of
// its features, and the normalise-backward of W is finished here like in the X-stationary K3b:
//   dW[c, d] = coef_c.x * (acc[c, d] - wh[c, d] * coef_c.y)      coef from reduce_r_kernel, wh = K1's fp16 rows.
// A tile's main loop runs K = batch >= 576 deep (>= 9 k-blocks of 256 x 256 x 64), so the 128 x 256 epilogue with its
// row-per-lane loads of wh is a small fraction of the tile: fusing it removes the separate pass over dW (read + write of
// 4 B per element plus the wh read) that used to follow.
struct EpiDwNorm {
  struct Params { float* out; int64_t ld; int64_t row_offset; CoefSrc coef; const __half* wh;
                  float* sq_part; };        // NULL, or [ceil(M / 128) * n_tiles * 4]: sum of dW^2 per (128-row block, n tile, warp)
  static __device__ __forceinline__ void run(const Params& ep, const GemmParams& p, const TileCoord& t,
                                             uint32_t tmem_acc, int quad, int lane, int epi_tid, float* scratch) {
    float sq = 0.f;
    const int row = t.m0 + quad * 32 + lane;
    const int ncols = min(BLOCK_N, p.N - t.n0);
    const bool row_ok = row < p.M;
    float2 cf = make_float2(0.f, 0.f);
    if (row_ok) cf = ep.coef.load(row, ep.row_offset + row, ep.coef.inv_sg());
    const int64_t base = (ep.row_offset + row) * ep.ld + t.n0;
    const bool vec = (ep.ld % 8 == 0);
    for (int ch = 0; ch * 32 < ncols; ++ch) {
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + ch * 32, v);
      uint4 w4[4];
      const int cc = min(32, ncols - ch * 32);
      const bool full = row_ok && cc == 32 && vec;
      if (full) {
#pragma unroll
        for (int i = 0; i < 4; ++i) w4[i] = __ldg(reinterpret_cast<const uint4*>(ep.wh + base + ch * 32) + i);
      }
      tmem_ld_wait();
      if (!row_ok) continue;
      float* dst = ep.out + base + ch * 32;
      if (full) {
        float o[32];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t q[4] = {w4[i].x, w4[i].y, w4[i].z, w4[i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&q[j]));
            o[i * 8 + 2 * j] = cf.x * fmaf(-f.x, cf.y, v[i * 8 + 2 * j]);
            o[i * 8 + 2 * j + 1] = cf.x * fmaf(-f.y, cf.y, v[i * 8 + 2 * j + 1]);
          }
        }
        if (ep.sq_part != nullptr) {
          float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) q4[u] = fmaf(o[j + u], o[j + u], q4[u]);
          }
          sq += (q4[0] + q4[1]) + (q4[2] + q4[3]);
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          st_global_256(dst + j, __float_as_uint(o[j]), __float_as_uint(o[j + 1]), __float_as_uint(o[j + 2]),
                        __float_as_uint(o[j + 3]), __float_as_uint(o[j + 4]), __float_as_uint(o[j + 5]),
                        __float_as_uint(o[j + 6]), __float_as_uint(o[j + 7]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < cc) {
            const float o1 = cf.x * fmaf(-__half2float(__ldg(ep.wh + base + ch * 32 + j)), cf.y, v[j]);
            dst[j] = o1; sq = fmaf(o1, o1, sq);
          }
      }
    }
    if (ep.sq_part != nullptr) {                              // every lane of the warp arrives here (rows beyond M hold 0)
      const float s = warp_sum(sq);
      if (lane == 0) ep.sq_part[((int64_t)(t.m0 / 128) * p.n_tiles + t.n0 / BLOCK_N) * 4 + quad] = s;
    }
    (void)epi_tid; (void)scratch;
  }
};

}  // namespace umma
}  // namespace b200f
