// ArcFace head on the fp32 CUDA-core engine: fused forward statistics (K2), backward (K3).
// Reference semantics: src/face_models.py:351-427 (forward), its autograd + the
// CrossEntropyLoss(label_smoothing) of src/training.py:341,515 (backward).
#pragma once
#include "simt_gemm.cuh"

namespace b200f {
namespace head_simt {

using simt::BM;
using simt::BN;
using simt::BK;
using simt::THREADS;

constexpr int PART_COLS = 6;   // sumexp, sumexp2, ztarget, sumz, best, bestidx(as float bits)

struct FwdParams {
  const void* x; const void* w;
  const float* inv_nx; const float* inv_nw;
  const int64_t* label;
  int64_t B, C, class_offset;
  int D;
  HeadMath hm;
  int n_chunks, tiles_per_chunk;
  float* part;          // [n_chunks, B, PART_COLS]
  float* cos_part;      // [gridDim.x*gridDim.y, 2]
  int32_t* nan_flag;
  float* logits; int64_t ld_logits;
  bool vec_x, vec_w;
};

template <typename T>
__global__ void __launch_bounds__(THREADS)
fwd_kernel(const FwdParams p) {
  __shared__ simt::Smem sm;
  __shared__ float red_min[THREADS / 32], red_max[THREADS / 32];
  const int chunk = blockIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int tx = threadIdx.x & 15;

  simt::TileLoader<T, true> la{static_cast<const T*>(p.x), p.D, p.B, m0, p.vec_x, nullptr};

  float inx[8];
  int64_t tgt[8];          // local target column of each of my rows, -1 if not in this shard
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + simt::acc_row(i);
    inx[i] = (row < p.B) ? p.inv_nx[row] : 0.f;
    int64_t t = (row < p.B) ? (p.label[row] - p.class_offset) : -1;
    tgt[i] = (t >= 0 && t < p.C) ? t : -1;
  }
  float sumexp[8], sumexp2[8], sumz[8], best[8], ztgt[8];
  int64_t bestidx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sumexp[i] = sumexp2[i] = sumz[i] = ztgt[i] = 0.f;
    best[i] = -INFINITY; bestidx[i] = INT64_MAX;
  }
  float cmin = INFINITY, cmax = -INFINITY;
  bool saw_nan = false;
  const float lo = cos_lo(), hi = cos_hi();
  const float s_eff = p.hm.s_eff;

  const int64_t n_tiles = (p.C + BN - 1) / BN;
  const int64_t t_begin = (int64_t)chunk * p.tiles_per_chunk;
  const int64_t t_end = min(n_tiles, t_begin + p.tiles_per_chunk);
  for (int64_t nt = t_begin; nt < t_end; ++nt) {
    const int64_t n0 = nt * BN;
    simt::TileLoader<T, true> lb{static_cast<const T*>(p.w), p.D, p.C, n0, p.vec_w, nullptr};
    float acc[8][8];
    simt::tile_mainloop<simt::OpFma>(acc, sm, la, lb, 0, p.D);
    float inw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t col = n0 + simt::acc_col(j);
      inw[j] = (col < p.C) ? __ldg(p.inv_nw + col) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t row = m0 + simt::acc_row(i);
      if (row >= p.B) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t col = n0 + simt::acc_col(j);
        if (col >= p.C) continue;
        const float cosv = acc[i][j] * inx[i] * inw[j];
        cmin = fminf(cmin, cosv); cmax = fmaxf(cmax, cosv);
        // torch.clamp propagates NaN; fminf/fmaxf would swallow it
        float c = (cosv != cosv) ? cosv : fminf(fmaxf(cosv, lo), hi);
        float t = (col == tgt[i]) ? p.hm.phi(c) : c;
        float z = t * s_eff;
        if (!isfinite(z)) { z = 0.f; saw_nan = true; }
        if (col == tgt[i]) ztgt[i] = z;
        const float e = expf(z - s_eff);
        sumexp[i] += e;
        sumexp2[i] = fmaf(e, e, sumexp2[i]);
        sumz[i] += z;
        const int64_t gcol = col + p.class_offset;
        if (z > best[i]) { best[i] = z; bestidx[i] = gcol; }
        if (p.logits != nullptr) p.logits[row * p.ld_logits + col] = z;
      }
    }
  }
  // reduce across the 16 threads (tx) that share my rows: they sit in one half-warp
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      sumexp[i]  += __shfl_xor_sync(0xffffffffu, sumexp[i], o);
      sumexp2[i] += __shfl_xor_sync(0xffffffffu, sumexp2[i], o);
      sumz[i]    += __shfl_xor_sync(0xffffffffu, sumz[i], o);
      ztgt[i]    += __shfl_xor_sync(0xffffffffu, ztgt[i], o);
      const float ob = __shfl_xor_sync(0xffffffffu, best[i], o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bestidx[i], o);
      if (ob > best[i] || (ob == best[i] && oi < bestidx[i])) { best[i] = ob; bestidx[i] = oi; }
    }
    const int64_t row = m0 + simt::acc_row(i);
    if (tx == 0 && row < p.B) {
      float* dst = p.part + ((int64_t)chunk * p.B + row) * PART_COLS;
      dst[0] = sumexp[i]; dst[1] = sumexp2[i]; dst[2] = ztgt[i]; dst[3] = sumz[i];
      dst[4] = best[i];
      // global class index < 2^31 in every supported configuration; keep it exact in an int slot
      reinterpret_cast<int32_t*>(dst)[5] = (bestidx[i] == INT64_MAX) ? -1 : (int32_t)bestidx[i];
    }
  }
  cmin = warp_min(cmin); cmax = warp_max(cmax);
  if ((threadIdx.x & 31) == 0) { red_min[threadIdx.x >> 5] = cmin; red_max[threadIdx.x >> 5] = cmax; }
  if (__syncthreads_or(saw_nan) && threadIdx.x == 0) atomicExch(p.nan_flag, 1);
  if (threadIdx.x == 0) {
    for (int wdx = 1; wdx < THREADS / 32; ++wdx) { cmin = fminf(cmin, red_min[wdx]); cmax = fmaxf(cmax, red_max[wdx]); }
    float* cp = p.cos_part + 2 * ((int64_t)blockIdx.y * gridDim.x + blockIdx.x);
    cp[0] = cmin; cp[1] = cmax;
  }
}

// Sum the per-chunk partials in a fixed order (bitwise reproducible), emit the shard's row stats.
__global__ void reduce_partials_kernel(const float* __restrict__ part, int n_chunks, int64_t B,
                                       const float* __restrict__ cos_part, int n_cos_part,
                                       float* __restrict__ row_stats, float* __restrict__ row_best,
                                       int64_t* __restrict__ row_argmax, float* __restrict__ cos_minmax) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row < B) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, best = -INFINITY;
    int64_t bi = -1;
    for (int c = 0; c < n_chunks; ++c) {
      const float* src = part + ((int64_t)c * B + row) * PART_COLS;
      s0 += src[0]; s1 += src[1]; s2 += src[2]; s3 += src[3];
      const int32_t idx = reinterpret_cast<const int32_t*>(src)[5];
      if (idx >= 0 && (src[4] > best || bi < 0)) { best = src[4]; bi = idx; }   // chunks ascend in class index
    }
    float* dst = row_stats + row * B200F_STAT_COLS;
    dst[B200F_STAT_SUMEXP] = s0; dst[B200F_STAT_SUMEXP2] = s1;
    dst[B200F_STAT_ZTARGET] = s2; dst[B200F_STAT_SUMZ] = s3;
    if (row_best != nullptr) row_best[row] = best;
    if (row_argmax != nullptr) row_argmax[row] = bi;
  }
  if (blockIdx.x == 0 && cos_minmax != nullptr) {
    __shared__ float smin[32], smax[32];
    float cmin = INFINITY, cmax = -INFINITY;
    for (int i = threadIdx.x; i < n_cos_part; i += blockDim.x) {
      cmin = fminf(cmin, cos_part[2 * i]); cmax = fmaxf(cmax, cos_part[2 * i + 1]);
    }
    cmin = warp_min(cmin); cmax = warp_max(cmax);
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = cmin; smax[threadIdx.x >> 5] = cmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int wdx = 1; wdx < (int)(blockDim.x >> 5); ++wdx) { cmin = fminf(cmin, smin[wdx]); cmax = fmaxf(cmax, smax[wdx]); }
      cos_minmax[0] = cmin; cos_minmax[1] = cmax;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward, per class chunk [c0, c0+Cc):
//   g_kernel : recompute S = x w^T for the chunk, write G (fp32, [B, ldg]) and
//              r_j = sum_i G_ij cos_ij  (= <w_hat_j, dw_hat_j>, the radial part removed by the
//              normalise-backward of w)
//   dw_kernel: dw[j,:] = inv_nw_j * (sum_i G_ij x_hat_i - w_hat_j r_j)
//   dx_kernel: partial dx_hat[i,:] = sum_{j in split} G_ij w_hat_j
struct BwdGParams {
  const void* x; const void* w;
  const float* inv_nx; const float* inv_nw;
  const int64_t* label; const float* lse; const float* grad_scale;
  int64_t B, C, class_offset, c0, Cc;
  int D;
  HeadMath hm;
  float ls_eps; float inv_Ctot;
  float* G; int64_t ldg;
  float* r;             // [Cc]
  const float* dlogits; int64_t ld_dlogits;   // compatibility path: upstream dL/dlogits instead of p-q
  bool vec_x, vec_w;
};

template <typename T>
__global__ void __launch_bounds__(THREADS)
bwd_g_kernel(const BwdGParams p) {
  __shared__ simt::Smem sm;
  const int64_t n0 = p.c0 + (int64_t)blockIdx.x * BN;         // first class of my tile (shard-local)
  const int64_t c_end = min(p.C, p.c0 + p.Cc);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  simt::TileLoader<T, true> lb{static_cast<const T*>(p.w), p.D, c_end, n0, p.vec_w, nullptr};
  float inw[8], rcol[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t col = n0 + simt::acc_col(j);
    inw[j] = (col < c_end) ? p.inv_nw[col] : 0.f;
    rcol[j] = 0.f;
  }
  const float gs = *p.grad_scale;
  const float lo = cos_lo(), hi = cos_hi();
  const float s_eff = p.hm.s_eff;
  const float q_off = p.ls_eps * p.inv_Ctot;
  for (int64_t m0 = 0; m0 < p.B; m0 += BM) {
    simt::TileLoader<T, true> la{static_cast<const T*>(p.x), p.D, p.B, m0, p.vec_x, nullptr};
    float acc[8][8];
    simt::tile_mainloop<simt::OpFma>(acc, sm, la, lb, 0, p.D);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t row = m0 + simt::acc_row(i);
      if (row >= p.B) continue;
      const float inx = p.inv_nx[row];
      const float lse = (p.dlogits != nullptr) ? 0.f : p.lse[row];
      const int64_t tgt = p.label[row] - p.class_offset;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t col = n0 + simt::acc_col(j);
        if (col >= c_end) continue;
        const float cosv = acc[i][j] * inx * inw[j];
        const bool is_nan = (cosv != cosv);
        const float c = is_nan ? cosv : fminf(fmaxf(cosv, lo), hi);
        const bool is_t = (col == tgt);
        float t = is_t ? p.hm.phi(c) : c;
        float z = t * s_eff;
        float f = is_t ? p.hm.dphi(c) : 1.0f;
        if (!isfinite(z)) { z = 0.f; f = 0.f; }            // scrubbed element: where() cuts the gradient
        if (!(cosv >= lo && cosv <= hi)) f = 0.f;           // clamp backward
        float g;
        if (p.dlogits != nullptr) {
          g = gs * p.dlogits[row * p.ld_dlogits + col] * f;
        } else {
          const float pr = expf(z - lse);
          const float q = is_t ? (1.0f - p.ls_eps) + q_off : q_off;
          g = gs * (pr - q) * f;
        }
        p.G[row * p.ldg + (col - p.c0)] = g;
        rcol[j] = fmaf(g, is_nan ? 0.f : cosv, rcol[j]);
      }
    }
  }
  // column sums over the 16 row-groups (ty) of the block
  float (*red)[BN + simt::PAD] = sm.a;                       // 16 x 132 floats, free after the mainloop
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][simt::acc_col(j)] = rcol[j];
  __syncthreads();
  if (threadIdx.x < BN) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) s += red[g][threadIdx.x];
    const int64_t col = n0 + threadIdx.x;
    if (col < c_end) p.r[col - p.c0] = s;
  }
  (void)tx;
}

struct BwdDwParams {
  const void* x; const void* w;
  const float* inv_nx; const float* inv_nw;
  const float* G; int64_t ldg; const float* r;
  int64_t B, C, c0, Cc;
  int D;
  float* dw;            // [C, D]
  bool vec_x;
};

template <typename T>
__global__ void __launch_bounds__(THREADS)
bwd_dw_kernel(const BwdDwParams p) {
  __shared__ simt::Smem sm;
  const int64_t cm0 = (int64_t)blockIdx.x * BM;              // class offset inside the chunk
  const int64_t d0 = (int64_t)blockIdx.y * BN;
  const int64_t c_cnt = min(p.Cc, p.C - p.c0);
  // A(m=class,k=batch) = G[k*ldg + m] (MN-major) ; B(n=d,k=batch) = x[k*D + n] * inv_nx[k]
  simt::TileLoader<float, false> la{p.G, p.ldg, c_cnt, cm0, (p.ldg % 4 == 0), nullptr};
  simt::TileLoader<T, false> lb{static_cast<const T*>(p.x), p.D, p.D, d0, p.vec_x, p.inv_nx};
  float acc[8][8];
  simt::tile_mainloop<simt::OpFma>(acc, sm, la, lb, 0, p.B);
  const T* w = static_cast<const T*>(p.w);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t cl = cm0 + simt::acc_row(i);
    if (cl >= c_cnt) continue;
    const int64_t cls = p.c0 + cl;
    const float inw = p.inv_nw[cls];
    const float rj = p.r[cl];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t d = d0 + simt::acc_col(j);
      if (d >= p.D) continue;
      const float what = to_f32<T>(w[cls * p.D + d]) * inw;
      p.dw[cls * p.D + d] = inw * (acc[i][j] - what * rj);
    }
  }
}

struct BwdDxParams {
  const void* w; const float* inv_nw;
  const float* G; int64_t ldg;
  int64_t B, C, c0, Cc;
  int D;
  int64_t k_per_split;
  float* part;          // [n_splits, B, D]
  bool vec_w;
};

template <typename T>
__global__ void __launch_bounds__(THREADS)
bwd_dx_kernel(const BwdDxParams p) {
  __shared__ simt::Smem sm;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t d0 = (int64_t)blockIdx.y * BN;
  const int64_t c_cnt = min(p.Cc, p.C - p.c0);
  const int64_t k_begin = (int64_t)blockIdx.z * p.k_per_split;
  const int64_t k_end = min(c_cnt, k_begin + p.k_per_split);
  // A(m=batch,k=class) = G[m*ldg + k] (K-major); B(n=d,k=class) = w[(c0+k)*D + n] * inv_nw[c0+k]
  simt::TileLoader<float, true> la{p.G, p.ldg, p.B, m0, (p.ldg % 4 == 0), nullptr};
  simt::TileLoader<T, false> lb{static_cast<const T*>(p.w) + p.c0 * p.D, p.D, p.D, d0, p.vec_w,
                                p.inv_nw + p.c0};
  float acc[8][8];
  simt::tile_mainloop<simt::OpFma>(acc, sm, la, lb, k_begin, k_end);
  float* out = p.part + (int64_t)blockIdx.z * p.B * p.D;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + simt::acc_row(i);
    if (row >= p.B) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t d = d0 + simt::acc_col(j);
      if (d < p.D) out[row * p.D + d] = acc[i][j];
    }
  }
}

// dst[i] (=|+=) sum_s part[s][i], fixed order
__global__ void reduce_splits_kernel(const float* __restrict__ part, int n_splits, int64_t n,
                                     float* __restrict__ dst, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = accumulate ? dst[i] : 0.f;
  for (int k = 0; k < n_splits; ++k) s += part[(int64_t)k * n + i];
  dst[i] = s;
}

}  // namespace head_simt
}  // namespace b200f
