// Epilogue policies of the tcgen05 GEMM core (umma_gemm.cuh) for the ArcFace head.
// Each epilogue thread owns ONE accumulator row (TMEM lane); columns arrive 32 at a time.
//
// Operands are the fp16 rows K1 emits: x_hat * S and w_hat * S (already L2-normalised, S a power of
// two), so an accumulator is  acc = S^2 * cos(theta)  and the epilogues need no per-row / per-column
// scale.  Every epilogue has a fast path for whole 32-column chunks (no clamp, no target column, no
// NaN scrub -- validated after the fact from the chunk's min / max / sum) and a careful path that
// applies the reference's element-wise sequence (src/face_models.py:363-427) to the same registers.
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
        if (cc == 32 && (ep.ld % 4 == 0)) {
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
// K2: cosine logits -> margin -> scale -> softmax / cross-entropy statistics (src/face_models.py:355-427,
// training.py:515).  Nothing B x C is stored: per (n_tile, row) one PART record.
constexpr int PART_COLS = 6;   // sumexp, sumexp2, ztarget, sumz, best, bestidx (int32 bits) -- as head_simt

struct EpiFwd {
  struct Params {
    const int64_t* label;
    int64_t B, C, class_offset;
    HeadMath hm;
    float inv_scale;    // 1 / S^2 : cos = acc * inv_scale
    float* part;        // [n_tiles, B, PART_COLS]
    float* cos_part;    // [tiles * 4 warps, 2]
    int32_t* nan_flag;
  };

  static __device__ __forceinline__ void run(const Params& ep, const GemmParams& p, const TileCoord& t,
                                             uint32_t tmem_acc, int quad, int lane, int epi_tid, float* scratch) {
    const int64_t row = (int64_t)t.m0 + quad * 32 + lane;
    const bool row_ok = row < ep.B;
    int tl = -1;                                              // my target column inside this tile
    if (row_ok) {
      const int64_t tg = __ldg(ep.label + row) - ep.class_offset - t.n0;
      if (tg >= 0 && tg < BLOCK_N && t.n0 + tg < ep.C) tl = (int)tg;
    }
    const int ncols = (int)min((int64_t)BLOCK_N, ep.C - t.n0);
    const float s_eff = ep.hm.s_eff;
    const float isc = ep.inv_scale;
    const float zs = isc * s_eff;                             // z = acc * zs off the target column
    const float a = zs * LOG2E, b = -s_eff * LOG2E;           // exp(z - s_eff) = 2^(acc*a + b)
    const float lo = cos_lo(), hi = cos_hi();
    const bool fast_ok = (s_eff > 0.f);

    float sumexp = 0.f, sumexp2 = 0.f, sumz = 0.f, ztgt = 0.f, best = -INFINITY;
    int bestidx = -1;
    float cmin = INFINITY, cmax = -INFINITY;
    bool saw_nan = false;

    for (int ch = 0; ch * 32 < ncols; ++ch) {
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + ch * 32, v);
      tmem_ld_wait();
      const int cc = min(32, ncols - ch * 32);
      float ce = 0.f, ce2 = 0.f, ct = 0.f, tmn = INFINITY, tmx = -INFINITY;
      int tix = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float tt = v[j];
        const float e = ex2_approx(fmaf(tt, a, b));
        ce += e;
        ce2 = fmaf(e, e, ce2);
        ct += tt;
        tmn = fminf(tmn, tt);
        if (tt > tmx) { tmx = tt; tix = j; }
      }
      const bool has_t = (tl >= ch * 32) && (tl < ch * 32 + 32);
      bool careful = !fast_ok || has_t || (cc < 32) || !(tmx * isc <= hi) || !(tmn * isc >= lo) ||
                     !isfinite(ct) || !isfinite(ce);
      careful = __any_sync(0xffffffffu, careful);             // warp stays convergent for the next tcgen05.ld
      if (!careful) {
        sumexp += ce; sumexp2 += ce2; sumz = fmaf(ct, zs, sumz);
        cmin = fminf(cmin, tmn * isc); cmax = fmaxf(cmax, tmx * isc);
        const float bz = tmx * zs;
        if (bz > best) { best = bz; bestidx = ch * 32 + tix; }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < cc) {
            const float cosv = v[j] * isc;
            cmin = fminf(cmin, cosv); cmax = fmaxf(cmax, cosv);
            const float c = (cosv != cosv) ? cosv : fminf(fmaxf(cosv, lo), hi);
            const bool is_t = (ch * 32 + j == tl);
            const float tv = is_t ? ep.hm.phi(c) : c;
            float z = tv * s_eff;
            if (!isfinite(z)) { z = 0.f; saw_nan = true; }
            if (is_t) ztgt = z;
            const float e = exp2f((z - s_eff) * LOG2E);
            sumexp += e; sumexp2 = fmaf(e, e, sumexp2); sumz += z;
            if (z > best) { best = z; bestidx = ch * 32 + j; }
          }
        }
      }
    }
    if (row_ok) {
      const int64_t n_tile = t.n0 / BLOCK_N;
      float* dst = ep.part + (n_tile * ep.B + row) * PART_COLS;
      dst[0] = sumexp; dst[1] = sumexp2; dst[2] = ztgt; dst[3] = sumz; dst[4] = best;
      reinterpret_cast<int32_t*>(dst)[5] = (bestidx < 0) ? -1 : (int32_t)(ep.class_offset + t.n0 + bestidx);
    } else {
      cmin = INFINITY; cmax = -INFINITY;
    }
    cmin = warp_min(cmin); cmax = warp_max(cmax);
    const int w = (t.n0 / BLOCK_N) * p.m_tiles + t.m0 / BLOCK_M;
    if (lane == 0) {
      float* cp = ep.cos_part + 2 * ((int64_t)w * 4 + quad);
      cp[0] = cmin; cp[1] = cmax;
    }
    if (__any_sync(0xffffffffu, saw_nan) && lane == 0) atomicExch(ep.nan_flag, 1);
    (void)epi_tid; (void)scratch;
  }
};

// -------------------------------------------------------------------------------------------------
// K3a: recompute the logits of a class chunk and emit the logit gradient as fp16 (one L2-resident buffer
// that both consumer GEMMs read),
//   G_ij = g_scale * grad_scale * (p_ij - q_ij) * dphi/dc * 1[lo <= cos <= hi]      (SURVEY 8a closed form)
// grad4 = {grad_scale, n, kappa, g_scale} from b200f_arcface_hook_scale; g_scale is the power of two that
// puts |grad_scale| * g_scale in [512, 1024], so a target-column entry (|p-q| <= 1, dphi <~ 30) stays below
// fp16 max and entries down to p ~ 1e-7 stay normal; the consumers divide it out again.
struct EpiBwdG {
  struct Params {
    const int64_t* label; const float* lse; const float* grad4;
    int64_t B, C, class_offset, c0;      // this launch covers shard-local classes [c0, c0 + p.N)
    HeadMath hm;
    float ls_eps, inv_Ctot, inv_scale;
    uint16_t* G; int64_t ldg;
  };

  static __device__ __forceinline__ void run(const Params& ep, const GemmParams& p, const TileCoord& t,
                                             uint32_t tmem_acc, int quad, int lane, int epi_tid, float* scratch) {
    const int64_t n0 = ep.c0 + t.n0;                          // shard-local class of column 0
    const int64_t c_end = min(ep.C, ep.c0 + (int64_t)p.N);
    const int64_t row = (int64_t)t.m0 + quad * 32 + lane;
    const bool row_ok = row < ep.B;
    const float lse = row_ok ? __ldg(ep.lse + row) : 0.f;
    int tl = -1;
    if (row_ok) {
      const int64_t tg = __ldg(ep.label + row) - ep.class_offset - n0;
      if (tg >= 0 && tg < BLOCK_N && n0 + tg < c_end) tl = (int)tg;
    }
    const int ncols = (int)min((int64_t)BLOCK_N, c_end - n0);
    const float s_eff = ep.hm.s_eff;
    const float isc = ep.inv_scale;
    const float a = isc * s_eff * LOG2E, b = -lse * LOG2E;    // p = 2^(acc*a + b)
    const float lo = cos_lo(), hi = cos_hi();
    const float gs = __ldg(ep.grad4) * __ldg(ep.grad4 + 3);
    const float q_off = ep.ls_eps * ep.inv_Ctot;
    const float gq = gs * q_off;
    const bool fast_ok = (s_eff > 0.f);
    uint16_t* gdst = ep.G + row * ep.ldg + t.n0;

    for (int ch = 0; ch * 32 < ncols; ++ch) {
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + ch * 32, v);
      tmem_ld_wait();
      const int cc = min(32, ncols - ch * 32);
      float g[32];
      float tmn = INFINITY, tmx = -INFINITY, chk = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float tt = v[j];
        g[j] = fmaf(gs, ex2_approx(fmaf(tt, a, b)), -gq);
        tmn = fminf(tmn, tt); tmx = fmaxf(tmx, tt);
        chk += tt;
      }
      const bool has_t = (tl >= ch * 32) && (tl < ch * 32 + 32);
      bool careful = !fast_ok || has_t || !(tmx * isc <= hi) || !(tmn * isc >= lo) || !isfinite(chk);
      careful = __any_sync(0xffffffffu, careful);
      if (careful) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float cosv = v[j] * isc;
          const float c = (cosv != cosv) ? cosv : fminf(fmaxf(cosv, lo), hi);
          const bool is_t = (ch * 32 + j == tl);
          const float tv = is_t ? ep.hm.phi(c) : c;
          float z = tv * s_eff;
          float f = is_t ? ep.hm.dphi(c) : 1.0f;
          if (!isfinite(z)) { z = 0.f; f = 0.f; }
          if (!(cosv >= lo && cosv <= hi)) f = 0.f;
          const float pr = exp2f((z - lse) * LOG2E);
          const float q = is_t ? (1.0f - ep.ls_eps) + q_off : q_off;
          g[j] = fminf(fmaxf(gs * (pr - q) * f, -65504.f), 65504.f);
        }
      }
      if (row_ok) {
        if (cc == 32) {
          uint32_t w1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) w1[j] = pack_f16(g[2 * j], g[2 * j + 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(gdst + ch * 32 + 8 * j) = make_uint4(w1[4 * j], w1[4 * j + 1], w1[4 * j + 2], w1[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < cc) gdst[ch * 32 + j] = (uint16_t)(pack_f16(g[j], 0.f) & 0xffff);
        }
      }
    }
    (void)epi_tid; (void)scratch;
  }
};

}  // namespace umma
}  // namespace b200f
