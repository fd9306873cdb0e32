// Gallery match on the fp32 CUDA-core engine (K4, exact engine).
// Reference semantics: compare_faces, src/app.py:50-64 -- d = ||q - g + 1e-6||_2 per reference
// embedding, strict '<' so the first index wins ties, accept iff d_min <= thresh -- and the
// cosine class-centre match of src/hyperparameter_tuning.py:1039-1046,1076.
#pragma once
#include "simt_gemm.cuh"

namespace b200f {
namespace gallery {

using simt::BM;
using simt::BN;
using simt::THREADS;

// Sorted list of the K best (key ascending; ties keep the entry that arrived first).
template <int K, typename IdxT>
struct TopK {
  float key[K];
  IdxT idx[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; ++i) { key[i] = INFINITY; idx[i] = (IdxT)-1; }
  }
  // candidates arrive in ascending index order: strict '<' keeps the lowest index on ties
  __device__ __forceinline__ void insert_ordered(float k, IdxT id) {
    if (!(k < key[K - 1])) return;          // also drops NaN and +inf, like `dist < min_dist`
    key[K - 1] = k; idx[K - 1] = id;
#pragma unroll
    for (int s = K - 1; s > 0; --s) {
      const bool sw = (key[s] < key[s - 1]);
      if (sw) {
        float tk = key[s]; key[s] = key[s - 1]; key[s - 1] = tk;
        IdxT ti = idx[s]; idx[s] = idx[s - 1]; idx[s - 1] = ti;
      }
    }
  }
  // arbitrary arrival order: (key, idx) lexicographic
  __device__ __forceinline__ void insert_lex(float k, IdxT id) {
    if (id < 0 || k != k) return;
    const bool better_than_last = (idx[K - 1] < 0) || (k < key[K - 1]) || (k == key[K - 1] && id < idx[K - 1]);
    if (!better_than_last) return;
    key[K - 1] = k; idx[K - 1] = id;
#pragma unroll
    for (int s = K - 1; s > 0; --s) {
      const bool sw = (idx[s - 1] < 0) || (key[s] < key[s - 1]) || (key[s] == key[s - 1] && idx[s] < idx[s - 1]);
      if (sw) {
        float tk = key[s]; key[s] = key[s - 1]; key[s - 1] = tk;
        IdxT ti = idx[s]; idx[s] = idx[s - 1]; idx[s - 1] = ti;
      }
    }
  }
};

struct Params {
  const void* q; const void* g;
  const float* q_inv; const float* g_inv;
  int64_t Q, N, index_offset;
  int D, k, metric;
  int tiles_per_chunk;
  int64_t* cand_idx;    // [n_chunks, Q, k]
  float* cand_score;    // [n_chunks, Q, k]
  const uint8_t* only_rows;   // optional [Q]: compute only the queries flagged 1 (re-run of the tensor engine's
                              // unverified queries); blocks without a flagged query leave at once
  bool vec_q, vec_g;
};

constexpr int SCORE_LD = BM + 1;
inline size_t dyn_smem_bytes() { return sizeof(float) * BN * SCORE_LD; }

template <typename T, class Op, int K>
__global__ void __launch_bounds__(THREADS)
topk_kernel(const Params p) {
  pdl_trigger(); pdl_wait();                                  // no-ops unless launched with programmatic serialization (the gated fallback)
  __shared__ simt::Smem sm;
  extern __shared__ float score_tile[];                       // [BN cols][SCORE_LD rows]
  const int chunk = blockIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  simt::TileLoader<T, true> la{static_cast<const T*>(p.q), p.D, p.Q, m0, p.vec_q, nullptr};
  const bool cosine = (p.metric == B200F_METRIC_COS);

  const int my_row = threadIdx.x & (BM - 1), my_half = threadIdx.x >> 7;
  if (p.only_rows != nullptr) {
    const bool mine = (m0 + my_row < p.Q) && p.only_rows[m0 + my_row] != 0;
    if (!__syncthreads_or(mine)) return;
  }
  TopK<K, int32_t> top;
  top.init();

  float qi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + simt::acc_row(i);
    qi[i] = (cosine && p.q_inv != nullptr && row < p.Q) ? p.q_inv[row] : 1.0f;
  }
  const int64_t n_tiles = (p.N + BN - 1) / BN;
  const int64_t t_begin = (int64_t)chunk * p.tiles_per_chunk;
  const int64_t t_end = min(n_tiles, t_begin + p.tiles_per_chunk);
  for (int64_t nt = t_begin; nt < t_end; ++nt) {
    const int64_t n0 = nt * BN;
    simt::TileLoader<T, true> lb{static_cast<const T*>(p.g), p.D, p.N, n0, p.vec_g, nullptr};
    float acc[8][8];
    simt::tile_mainloop<Op>(acc, sm, la, lb, 0, p.D);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = simt::acc_col(j);
      const int64_t gcol = n0 + col;
      const float gi = (cosine && p.g_inv != nullptr && gcol < p.N) ? __ldg(p.g_inv + gcol) : 1.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // key: smaller is better.  l2eps: the distance itself; cos: minus the cosine
        const float key = cosine ? -(acc[i][j] * qi[i] * gi) : sqrtf(acc[i][j]);
        score_tile[col * SCORE_LD + simt::acc_row(i)] = key;
      }
    }
    __syncthreads();
    const int c_lo = my_half * (BN / 2);
    const int c_cnt = (int)min((int64_t)(BN / 2), p.N - n0 - c_lo);
    for (int c = 0; c < c_cnt; ++c)
      top.insert_ordered(score_tile[(c_lo + c) * SCORE_LD + my_row], (int32_t)(n0 + c_lo + c));
    __syncthreads();
  }
  // merge the two half-row lists: upper half parks its list in shared memory
  float* park_key = score_tile;
  int32_t* park_idx = reinterpret_cast<int32_t*>(score_tile + BM * K);
  if (my_half == 1) {
#pragma unroll
    for (int s = 0; s < K; ++s) { park_key[s * BM + my_row] = top.key[s]; park_idx[s * BM + my_row] = top.idx[s]; }
  }
  __syncthreads();
  if (my_half == 0) {
#pragma unroll
    for (int s = 0; s < K; ++s) top.insert_lex(park_key[s * BM + my_row], park_idx[s * BM + my_row]);
    const int64_t row = m0 + my_row;
    if (row < p.Q) {
      int64_t* oi = p.cand_idx + ((int64_t)chunk * p.Q + row) * p.k;
      float* os = p.cand_score + ((int64_t)chunk * p.Q + row) * p.k;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        if (s < p.k) {
          const bool ok = top.idx[s] >= 0;
          oi[s] = ok ? (p.index_offset + top.idx[s]) : -1;
          os[s] = ok ? (cosine ? -top.key[s] : top.key[s]) : (cosine ? -INFINITY : INFINITY);
        }
      }
    }
  }
}

// One thread per query: merge P lists of k candidates (score form, global indices).
template <int K>
__global__ void merge_kernel(const int64_t* __restrict__ idx_all, const float* __restrict__ score_all,
                             int P, int64_t Q, int k, int metric, float thresh,
                             int64_t* __restrict__ idx, float* __restrict__ score,
                             uint8_t* __restrict__ accept, const uint8_t* __restrict__ only_rows) {
  pdl_trigger(); pdl_wait();
  const int64_t qid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (qid >= Q) return;
  if (only_rows != nullptr && only_rows[qid] == 0) return;
  const bool cosine = (metric == B200F_METRIC_COS);
  TopK<K, int64_t> top;
  top.init();
  for (int pth = 0; pth < P; ++pth) {
    const int64_t* si = idx_all + ((int64_t)pth * Q + qid) * k;
    const float* ss = score_all + ((int64_t)pth * Q + qid) * k;
    for (int s = 0; s < k; ++s) top.insert_lex(cosine ? -ss[s] : ss[s], si[s]);
  }
#pragma unroll
  for (int s = 0; s < K; ++s) {
    if (s < k) {
      const bool ok = top.idx[s] >= 0;
      idx[qid * k + s] = ok ? top.idx[s] : -1;
      score[qid * k + s] = ok ? (cosine ? -top.key[s] : top.key[s]) : (cosine ? -INFINITY : INFINITY);
    }
  }
  if (accept != nullptr) {
    const bool ok = top.idx[0] >= 0;
    const float best = cosine ? -top.key[0] : top.key[0];
    accept[qid] = ok && (cosine ? (best >= thresh) : (best <= thresh));   // src/app.py:64
  }
}

}  // namespace gallery
}  // namespace b200f
