// Epilogue policies of the tcgen05 GEMM core (umma_gemm.cuh) for the ArcFace head.
// Each epilogue thread owns ONE accumulator row (TMEM lane); columns arrive 32 at a time.
#pragma once
#include "common.cuh"
#include "umma_gemm.cuh"

namespace b200f {
namespace umma {

constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Stage the per-column scale (inv_nw of the 256 classes of this tile) in epilogue scratch.
__device__ __forceinline__ void stage_col_scale(float* scratch, const float* __restrict__ inv_nw, int64_t n0,
                                                int64_t C, int epi_tid) {
  epi_bar_sync();                                             // previous tile's readers are done
#pragma unroll
  for (int i = 0; i < BLOCK_N / EPI_THREADS; ++i) {
    const int c = epi_tid + i * EPI_THREADS;
    scratch[c] = (n0 + c < C) ? __ldg(inv_nw + n0 + c) : 0.f;
  }
  epi_bar_sync();
}

// -------------------------------------------------------------------------------------------------
// Plain store: D (fp32) -> out[split][m][n].  Used by the self-test, the dW GEMM and the split-K dX GEMM.
struct EpiStore {
  struct Params { float* out; int64_t ld; int64_t split_stride; int64_t row_offset; float scale; };
  static __device__ __forceinline__ void run(const Params& ep, const GemmParams& p, const TileCoord& t,
                                             uint32_t tmem_acc, int quad, int lane, int epi_tid, float* scratch) {
    const int row = t.m0 + quad * 32 + lane;
    const int ncols = min(BLOCK_N, p.N - t.n0);
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
            *reinterpret_cast<float4*>(dst + ch * 32 + j) =
                make_float4(v[j] * ep.scale, v[j + 1] * ep.scale, v[j + 2] * ep.scale, v[j + 3] * ep.scale);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < cc) dst[ch * 32 + j] = v[j] * ep.scale;
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
    const float* inv_nx; const float* inv_nw; const int64_t* label;
    int64_t B, C, class_offset;
    HeadMath hm;
    float* part;        // [n_tiles, B, PART_COLS]
    float* cos_part;    // [work items * 4 warps, 2]
    int32_t* nan_flag;
  };

  static __device__ __forceinline__ void run(const Params& ep, const GemmParams& p, const TileCoord& t,
                                             uint32_t tmem_acc, int quad, int lane, int epi_tid, float* scratch) {
    stage_col_scale(scratch, ep.inv_nw, t.n0, ep.C, epi_tid);
    const int64_t row = (int64_t)t.m0 + quad * 32 + lane;
    const bool row_ok = row < ep.B;
    const float inx = row_ok ? __ldg(ep.inv_nx + row) : 1.0f;
    int tl = -1;                                              // my target column inside this tile
    if (row_ok) {
      const int64_t tg = __ldg(ep.label + row) - ep.class_offset - t.n0;
      if (tg >= 0 && tg < BLOCK_N && t.n0 + tg < ep.C) tl = (int)tg;
    }
    const int ncols = (int)min((int64_t)BLOCK_N, ep.C - t.n0);
    const float s_eff = ep.hm.s_eff;
    const float zs = inx * s_eff;                             // z = t * zs for non-target columns
    const float a = zs * LOG2E, b = -s_eff * LOG2E;
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
      const float* cs = scratch + ch * 32;
      // ---- fast path: no clamp, no target, no scrub; validated after the fact -------------------
      float ce = 0.f, ce2 = 0.f, ct = 0.f, tmn = INFINITY, tmx = -INFINITY;
      int tix = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float tt = v[j] * cs[j];
        const float e = ex2_approx(fmaf(tt, a, b));
        ce += e;
        ce2 = fmaf(e, e, ce2);
        ct += tt;
        tmn = fminf(tmn, tt);
        if (tt > tmx) { tmx = tt; tix = j; }
      }
      const bool has_t = (tl >= ch * 32) && (tl < ch * 32 + 32);
      bool careful = !fast_ok || has_t || (cc < 32) || !(tmx * inx <= hi) || !(tmn * inx >= lo) ||
                     !isfinite(ct) || !isfinite(ce);
      careful = __any_sync(0xffffffffu, careful);             // keep the warp convergent for the next tcgen05.ld
      if (!careful) {
        sumexp += ce; sumexp2 += ce2; sumz = fmaf(ct, zs, sumz);
        cmin = fminf(cmin, tmn * inx); cmax = fmaxf(cmax, tmx * inx);
        const float bz = tmx * zs;
        if (bz > best) { best = bz; bestidx = ch * 32 + tix; }
      } else {
        // ---- careful path: the reference's element-wise sequence, per element -------------------
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < cc) {
            const float cosv = v[j] * cs[j] * inx;
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
  }
};

// -------------------------------------------------------------------------------------------------
// K3a: recompute the logits of a class chunk and emit the logit gradient in two 16-bit layouts,
//   G1[i, j] = G_ij * inv_nx[i]   (A operand of dW_hat = G1^T x,       raw x rows, exact bf16)
//   G2[i, j] = G_ij * inv_nw[j]   (A operand of dx_hat = G2 w,         raw w rows, exact bf16)
//   G_ij = grad_scale * (p_ij - q_ij) * dphi/dc * 1[lo <= cos <= hi]      (SURVEY 8a, H2/H3 closed form)
// both scaled by g_scale (a power of two, undone by the consumers' epilogues) so fp16 keeps its range.
template <bool G_FP16>
struct EpiBwdG {
  struct Params {
    const float* inv_nx; const float* inv_nw; const int64_t* label; const float* lse; const float* grad_scale;
    int64_t B, C, class_offset, c0;      // this launch covers shard-local classes [c0, c0 + p.N)
    HeadMath hm;
    float ls_eps, inv_Ctot, g_scale;
    uint16_t* G1; uint16_t* G2; int64_t ldg;
  };

  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    return G_FP16 ? pack_f16(lo, hi) : pack_bf16(lo, hi);
  }

  static __device__ __forceinline__ void run(const Params& ep, const GemmParams& p, const TileCoord& t,
                                             uint32_t tmem_acc, int quad, int lane, int epi_tid, float* scratch) {
    const int64_t n0 = ep.c0 + t.n0;                          // shard-local class of column 0
    const int64_t c_end = min(ep.C, ep.c0 + (int64_t)p.N);
    stage_col_scale(scratch, ep.inv_nw, n0, c_end, epi_tid);
    const int64_t row = (int64_t)t.m0 + quad * 32 + lane;
    const bool row_ok = row < ep.B;
    const float inx = row_ok ? __ldg(ep.inv_nx + row) : 1.0f;
    const float lse = row_ok ? __ldg(ep.lse + row) : 0.f;
    int tl = -1;
    if (row_ok) {
      const int64_t tg = __ldg(ep.label + row) - ep.class_offset - n0;
      if (tg >= 0 && tg < BLOCK_N && n0 + tg < c_end) tl = (int)tg;
    }
    const int ncols = (int)min((int64_t)BLOCK_N, c_end - n0);
    const float s_eff = ep.hm.s_eff;
    const float zs = inx * s_eff;
    const float a = zs * LOG2E, b = -lse * LOG2E;
    const float lo = cos_lo(), hi = cos_hi();
    const float gs = __ldg(ep.grad_scale) * ep.g_scale;
    const float q_off = ep.ls_eps * ep.inv_Ctot;
    const float gq = gs * q_off;
    const bool fast_ok = (s_eff > 0.f);
    uint16_t* g1 = ep.G1 + row * ep.ldg + t.n0;
    uint16_t* g2 = ep.G2 + row * ep.ldg + t.n0;

    for (int ch = 0; ch * 32 < ncols; ++ch) {
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + ch * 32, v);
      tmem_ld_wait();
      const int cc = min(32, ncols - ch * 32);
      const float* cs = scratch + ch * 32;
      float g[32];
      float tmn = INFINITY, tmx = -INFINITY, chk = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float tt = v[j] * cs[j];
        const float pr = ex2_approx(fmaf(tt, a, b));
        g[j] = fmaf(gs, pr, -gq);
        tmn = fminf(tmn, tt); tmx = fmaxf(tmx, tt);
        chk += tt;
      }
      const bool has_t = (tl >= ch * 32) && (tl < ch * 32 + 32);
      bool careful = !fast_ok || has_t || !(tmx * inx <= hi) || !(tmn * inx >= lo) || !isfinite(chk);
      careful = __any_sync(0xffffffffu, careful);
      if (careful) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float cosv = v[j] * cs[j] * inx;
          const float c = (cosv != cosv) ? cosv : fminf(fmaxf(cosv, lo), hi);
          const bool is_t = (ch * 32 + j == tl);
          const float tv = is_t ? ep.hm.phi(c) : c;
          float z = tv * s_eff;
          float f = is_t ? ep.hm.dphi(c) : 1.0f;
          if (!isfinite(z)) { z = 0.f; f = 0.f; }
          if (!(cosv >= lo && cosv <= hi)) f = 0.f;
          const float pr = exp2f((z - lse) * LOG2E);
          const float q = is_t ? (1.0f - ep.ls_eps) + q_off : q_off;
          g[j] = gs * (pr - q) * f;
        }
      }
      if (row_ok) {
        if (cc == 32) {
          uint32_t w1[16], w2[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            w1[j] = pack(g[2 * j] * inx, g[2 * j + 1] * inx);
            w2[j] = pack(g[2 * j] * cs[2 * j], g[2 * j + 1] * cs[2 * j + 1]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<uint4*>(g1 + ch * 32 + 8 * j) = make_uint4(w1[4 * j], w1[4 * j + 1], w1[4 * j + 2], w1[4 * j + 3]);
            *reinterpret_cast<uint4*>(g2 + ch * 32 + 8 * j) = make_uint4(w2[4 * j], w2[4 * j + 1], w2[4 * j + 2], w2[4 * j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < cc) {
              const uint32_t h1 = pack(g[j] * inx, 0.f), h2 = pack(g[j] * cs[j], 0.f);
              g1[ch * 32 + j] = (uint16_t)(h1 & 0xffff);
              g2[ch * 32 + j] = (uint16_t)(h2 & 0xffff);
            }
          }
        }
      }
    }
    (void)epi_tid;
  }
};

}  // namespace umma
}  // namespace b200f
