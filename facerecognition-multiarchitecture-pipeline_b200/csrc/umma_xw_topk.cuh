// K4 on tensor cores: epilogue policy of the X-stationary kernel (umma_xw.cuh) that turns the query x gallery
// product into per-query candidate lists, plus the prepare / select kernels around it.
//
//   scan   : xw_kernel<PAIR, K-major, XwTopK<KT>>  -- queries (bf16) resident in shared memory, gallery rows (bf16)
//            streamed once from HBM; an epilogue thread owns ONE query and keeps the KT smallest approximate keys
//            of the gallery columns it sees in registers (sorted insertion, first index wins ties)
//              L2EPS: a_ij = (|g_j|^2 - 2e-6 sum g_j) - 2 <q_i, g_j>      (= d_ij^2 minus a per-query constant)
//              COS  : a_ij = - <q_i, g_j / |g_j|>
//   select : per query, merge the per-(chunk, half) lists to the KT best, RE-SCORE them exactly in fp32 against the
//            original gallery rows with the reference formula (src/app.py:59: ||q - g + 1e-6||), order by (score,
//            index), and VERIFY that no row the scan excluded can belong to the top-k: every excluded row has an
//            approximate key >= the worst kept one, and |approx - exact| <= delta (bf16 operand rounding bound), so
//            exact_kth < worst_kept - delta proves exactness.  Queries that fail the test are flagged and recomputed
//            by the exact fp32 CUDA-core engine in the same stream (no host round trip).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma_xw.cuh"

namespace b200f {
namespace umma {

constexpr float GALLERY_EPS = 1e-6f;            // F.pairwise_distance default eps, added to the difference
// |<q,g> - <bf16(q), bf16(g)>| <= (2u + u^2) |q||g| with u = 2^-9 (round to nearest), plus fp32 accumulation
constexpr float GALLERY_DOT_ERR = 3.97e-3f;          // includes the 2^-18 relative index packing of the keys
// fp16 operands: u = 2^-12 in the normal range, absolute 2^-25 per element below 2^-14
constexpr float GALLERY_DOT_ERR_FP16 = 4.96e-4f;
constexpr float GALLERY_ABS_ERR_FP16 = 3.0e-8f;      // per element; enters as abs * sqrt(D) * (|q| + |g|max)

template <int KT>
struct XwTopK {
  struct Params {
    const float* bias;          // [N] per gallery row (L2EPS) or NULL (COS)
    float mult;                 // -2 (L2EPS) / -1 (COS)
    float* cand_key;            // [Q, n_lists, KT]
    int32_t* cand_idx;          // [Q, n_lists, KT]   gallery row (shard-local), -1 = empty
    int n_lists;                // n_chunks * 2
    const float* tau0;          // [Q] upper bound of each query's KT-th best key from a scanned SAMPLE of the
                                // gallery (gallery_tau_kernel), or NULL
    // compact output (cnt != NULL; needs tau0): a list appends only its non-empty entries to its query's candidate
    // array -- with the sample bound in place a query ends up with a few dozen candidates instead of n_lists * KT
    // mostly empty slots, and gallery_select_block_kernel ranks them and re-scores the winners in one round.  cnt[q] counts what was
    // offered (the tau kernel zeroes it); entries beyond `cap` are dropped and the query goes to the exact engine.
    int32_t* cnt; int cap;
  };
  struct State { float key[KT]; int32_t idx[KT]; float lim0; bool row_ok; };

  static __device__ __forceinline__ void item_begin(State& st, const Params& ep, const XwParams& p, const XwItem& it) {
#pragma unroll
    for (int s = 0; s < KT; ++s) { st.key[s] = INFINITY; st.idx[s] = -1; }
    st.row_ok = it.row < p.B;
    // Nothing above the sample's KT-th best key can be among the KT best of the whole gallery.  Without this bound each
    // of a query's ~300 lists (one per CTA and column half) warms up alone -- ~100 insertions each, serialised over the
    // warp: 87 % of the scan's instructions; with it a list only ever takes the few globally competitive elements.
    // A few ulps of slack keep the sample's own KT best inside, so the merged candidates still number >= KT and the
    // proof's bound (every dropped key >= merged KT-th key) holds unchanged.
    st.lim0 = 3.0e38f;
    if (ep.tau0 != nullptr && st.row_ok) {
      const float t0 = __ldg(ep.tau0 + it.row);
      st.lim0 = t0 + fabsf(t0) * 1.6e-5f + 1e-30f;
    }
  }
  static __device__ __forceinline__ void tile_begin(State&, const Params&, const XwParams&, const XwItem&, int, int, int) {}

  static __device__ __forceinline__ void insert(State& st, float k, int32_t id) {
    st.key[KT - 1] = k; st.idx[KT - 1] = id;
#pragma unroll
    for (int s = KT - 1; s > 0; --s) {
      if (st.key[s] < st.key[s - 1]) {                        // strict: an equal key stays behind the earlier index
        const float tk = st.key[s]; st.key[s] = st.key[s - 1]; st.key[s - 1] = tk;
        const int32_t ti = st.idx[s]; st.idx[s] = st.idx[s - 1]; st.idx[s - 1] = ti;
      }
    }
  }

  // The slice position j is packed into the 5 low mantissa bits of the key (2^-18 relative: far inside the bf16
  // error budget of the proof, and it makes every key of a slice distinct), so a min tree yields value AND index.
  // Qualifying elements are then extracted one per trip of a warp-uniform loop -- smallest first -- instead of
  // running the 15-step insertion at each of the 32 positions whenever ANY lane of the warp needs it there
  // (which made the first version 5x slower than the scan's HBM time).
  static __device__ __forceinline__ void slice(State& st, const Params& ep, const XwParams& p, const XwItem&,
                                               float (&v)[32], int cls0) {
    const int cc = min(32, p.C - cls0);
    constexpr float PAD = 3.0e38f;                            // beyond the gallery: finite, never selected
    float a[32];
    if (ep.bias != nullptr) {
      if (cc == 32 && (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) {   // the slice's 32 biases as eight 16-byte loads (warp-uniform address)
        const float4* b4 = reinterpret_cast<const float4*>(ep.bias + cls0);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = __ldg(b4 + (j >> 2));
          a[j] = fmaf(v[j], ep.mult, bb.x); a[j + 1] = fmaf(v[j + 1], ep.mult, bb.y);
          a[j + 2] = fmaf(v[j + 2], ep.mult, bb.z); a[j + 3] = fmaf(v[j + 3], ep.mult, bb.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = (j < cc) ? fmaf(v[j], ep.mult, __ldg(ep.bias + cls0 + j)) : PAD;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = (j < cc) ? v[j] * ep.mult : PAD;
    }
    // (NaN keys never win an fminf and never enter a list -- the exact engine drops NaN distances too.  Operands that
    //  turn non-finite only in 16 bits are caught when they are prepared, see gallery_prepare_kernel.)
    // Reject test before any key is packed: a packed key differs from its element by < 2^-18 |element|, so an element
    // whose packed key would beat the bound is itself below  bound + 2^-17 |bound|.  With the sample bound in place
    // nearly every slice ends here -- one FFMA and one FMNMX per element (above 128 queries the scan is bound by this
    // epilogue, not by HBM).
    {
      float u4[4] = {PAD, PAD, PAD, PAD};
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) u4[u] = fminf(u4[u], a[j + u]);
      }
      const float umin = fminf(fminf(u4[0], u4[1]), fminf(u4[2], u4[3]));
      const float bound = fminf(st.key[KT - 1], st.lim0);
      const float bound_hi = bound + fabsf(bound) * 7.7e-6f + 1e-37f;       // 2^-17 = 7.63e-6
      if (!__any_sync(0xffffffffu, umin < bound_hi)) return;
    }
    float m4[4] = {PAD, PAD, PAD, PAD};
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[j + u] = __uint_as_float((__float_as_uint(a[j + u]) & ~31u) | (uint32_t)(j + u));
        m4[u] = fminf(m4[u], a[j + u]);
      }
    }
    float smin = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
    while (__any_sync(0xffffffffu, smin < fminf(st.key[KT - 1], st.lim0))) {
      const bool mine = smin < fminf(st.key[KT - 1], st.lim0);
      const float taken = smin;
      if (mine) insert(st, smin, cls0 + (int)(__float_as_uint(smin) & 31u));
      // next smallest element after the one just taken (lanes that did not insert are done with this slice)
      float n4[4] = {PAD, PAD, PAD, PAD};
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) n4[u] = fminf(n4[u], (a[j + u] > taken) ? a[j + u] : PAD);
      }
      smin = mine ? fminf(fminf(n4[0], n4[1]), fminf(n4[2], n4[3])) : PAD;
    }
  }

  static __device__ __forceinline__ void item_end(State& st, const Params& ep, const XwParams&, const XwItem& it, float*) {
    if (!st.row_ok) return;
    if (ep.cnt != nullptr) {
      int nv = 0;
#pragma unroll
      for (int s = 0; s < KT; ++s) nv += (st.idx[s] >= 0) ? 1 : 0;       // sorted: the valid entries come first
      if (nv == 0) return;
      const int pos = atomicAdd(ep.cnt + it.row, nv);
      float* dk = ep.cand_key + (int64_t)it.row * ep.cap;
      int32_t* di = ep.cand_idx + (int64_t)it.row * ep.cap;
#pragma unroll
      for (int s = 0; s < KT; ++s)
        if (s < nv && pos + s < ep.cap) { dk[pos + s] = st.key[s]; di[pos + s] = st.idx[s]; }
      return;
    }
    const int64_t base = ((int64_t)it.row * ep.n_lists + it.chunk * 2 + it.half) * KT;
#pragma unroll
    for (int s = 0; s < KT; s += 4) {
      *reinterpret_cast<float4*>(ep.cand_key + base + s) = make_float4(st.key[s], st.key[s + 1], st.key[s + 2], st.key[s + 3]);
      *reinterpret_cast<int4*>(ep.cand_idx + base + s) = make_int4(st.idx[s], st.idx[s + 1], st.idx[s + 2], st.idx[s + 3]);
    }
  }
};

// Sample pre-pass, min-only: a thread keeps GALLERY_MIN_SUB minima, one per residue class of its slice counter (disjoint
// columns, so the minima belong to distinct gallery rows) -- the bound below needs nothing else, and the full KT-deep sorted
// insertion of XwTopK was most of the pre-pass (24 us for one tile per CTA at Q = 128, 185 us at Q = 8192 x 125 k).
constexpr int GALLERY_MIN_SUB = 4;
struct XwMinKey {
  struct Params { const float* bias; float mult; float* out; int n_lists; };   // out [Q, n_lists, GALLERY_MIN_SUB]
  struct State { float m[GALLERY_MIN_SUB]; int s; bool row_ok; };
  static __device__ __forceinline__ void item_begin(State& st, const Params&, const XwParams& p, const XwItem& it) {
#pragma unroll
    for (int b = 0; b < GALLERY_MIN_SUB; ++b) st.m[b] = 3.0e38f;
    st.s = 0; st.row_ok = it.row < p.B;
  }
  static __device__ __forceinline__ void tile_begin(State&, const Params&, const XwParams&, const XwItem&, int, int, int) {}
  static __device__ __forceinline__ void slice(State& st, const Params& ep, const XwParams& p, const XwItem&,
                                               float (&v)[32], int cls0) {
    const int cc = min(32, p.C - cls0);
    float m4[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
    if (ep.bias != nullptr && cc == 32 && (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) {
      const float4* b4 = reinterpret_cast<const float4*>(ep.bias + cls0);   // eight 16-byte loads, not 32 scalar ones
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bb = __ldg(b4 + (j >> 2));
        m4[0] = fminf(m4[0], fmaf(v[j], ep.mult, bb.x)); m4[1] = fminf(m4[1], fmaf(v[j + 1], ep.mult, bb.y));
        m4[2] = fminf(m4[2], fmaf(v[j + 2], ep.mult, bb.z)); m4[3] = fminf(m4[3], fmaf(v[j + 3], ep.mult, bb.w));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float a = (ep.bias != nullptr) ? fmaf(v[j + u], ep.mult, (j + u < cc) ? __ldg(ep.bias + cls0 + j + u) : 0.f)
                                               : v[j + u] * ep.mult;
          m4[u] = fminf(m4[u], (j + u < cc) ? a : 3.0e38f);    // NaN keys never win an fminf
        }
      }
    }
    const float sm = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
    const int b = st.s & (GALLERY_MIN_SUB - 1);
    ++st.s;
#pragma unroll
    for (int u = 0; u < GALLERY_MIN_SUB; ++u) if (u == b) st.m[u] = fminf(st.m[u], sm);
  }
  static __device__ __forceinline__ void item_end(State& st, const Params& ep, const XwParams&, const XwItem& it, float*) {
    static_assert(GALLERY_MIN_SUB == 4, "one float4 per thread");
    if (st.row_ok)
      *reinterpret_cast<float4*>(ep.out + ((int64_t)it.row * ep.n_lists + it.chunk * 2 + it.half) * GALLERY_MIN_SUB) =
          make_float4(st.m[0], st.m[1], st.m[2], st.m[3]);
  }
};

// tau0[q] = KT-th smallest of the query's n_lists (<= 1024) minima: KT DISTINCT gallery rows have a key <= it
// (the lists cover disjoint rows), so it bounds the KT-th best key of the whole gallery.  One warp per query, the
// values in registers, KT rounds of a shuffle arg-min.  Also zeroes the query's candidate counter of the compact scan.
constexpr int GALLERY_TAU_MIN_PER_LANE = 32;                    // minima a lane holds at most: 1024 per query
template <int KT, int L>
__global__ void __launch_bounds__(128)
gallery_tau_min_kernel(const float* __restrict__ key_min, int n_lists, int64_t Q, float* __restrict__ tau0,
                       int32_t* __restrict__ cnt) {
  pdl_trigger(); pdl_wait();
  static_assert(L <= GALLERY_TAU_MIN_PER_LANE, "minima per lane");
  const int lane = threadIdx.x & 31;
  const int64_t qi = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (qi >= Q) return;
  float v[L];
#pragma unroll
  for (int u = 0; u < L; ++u) { const int l = lane + 32 * u; v[u] = (l < n_lists) ? __ldg(key_min + qi * n_lists + l) : INFINITY; }
  float last = INFINITY;
  for (int r = 0; r < KT; ++r) {
    float k = INFINITY; int who = 1 << 20;
#pragma unroll
    for (int u = 0; u < L; ++u) if (v[u] < k) { k = v[u]; who = lane + 32 * u; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ok = __shfl_xor_sync(0xffffffffu, k, o);
      const int ow = __shfl_xor_sync(0xffffffffu, who, o);
      if (ok < k || (ok == k && ow < who)) { k = ok; who = ow; }
    }
    if (who >= (1 << 20)) { last = INFINITY; break; }        // fewer than KT ranked lists: no bound
#pragma unroll
    for (int u = 0; u < L; ++u) if (who == lane + 32 * u) v[u] = INFINITY;
    last = k;
  }
  if (lane == 0) { tau0[qi] = last; if (cnt != nullptr) cnt[qi] = 0; }
}

// ---- select, compact form ------------------------------------------------------------------------------------------
// Input: the query's compact candidate array (cnt[q] entries of (key, gallery row): with the sample bound at the 1e-3
// quantile, ~1000 of a 1 M-row gallery).  One block of 128 threads per query, every step parallel over the block:
//   candidates -> registers (16 per thread); the KT-th smallest of the 128 per-thread minima bounds the KT-th best key
//   (rank by counting, no arg-min rounds) -> the few survivors under it -> shared memory -> rank by counting: the KT best
//   approximate keys by (key, row) -> ALL KT winners re-scored exactly in ONE round (thread t owns elements t, t + 128, ..:
//   4 x KT independent loads in flight per thread; the padded-list kernel below takes four rounds of four rows with a
//   barrier pair each) -> ordered by (exact key, row) by counting -> thread 0 proves the top-k and writes the outputs.
// The re-score adds in the SAME order as gallery_select_kernel (per-thread chain over d = t, t + 128, ..; warp butterfly;
// (w0 + w1) + (w2 + w3)), so a gallery gives the same score bits whichever of the two kernels its size selects (sharded
// and unsharded results are compared bit for bit).  History: one WARP per query 57 us for 128 queries (32 blocks, every
// phase latency-bound); this block form with arg-min rounds over all ~1000 candidates 54 us; padded lists 41 us.
constexpr int GALLERY_SEL_PER_THREAD = 16;                     // candidates a thread holds: 2048 per query, else the exact engine
constexpr int GALLERY_SEL_MAX_SURV = 1024;
template <int KT>
__global__ void __launch_bounds__(128)
gallery_select_block_kernel(const float* __restrict__ cand_key, const int32_t* __restrict__ cand_idx,
                            const int32_t* __restrict__ cnt, int cap, const float* __restrict__ q, const float* __restrict__ g,
                            const float* __restrict__ q_inv, const float* __restrict__ g_inv,
                            const float* __restrict__ gmax_ptr, const uint8_t* __restrict__ q_bad, int64_t Q, int D, int k,
                            int metric, int fmt, float thresh, int64_t index_offset, int64_t* __restrict__ idx_out,
                            float* __restrict__ score_out, uint8_t* __restrict__ accept, uint8_t* __restrict__ redo,
                            int32_t* __restrict__ redo_count) {
  pdl_trigger(); pdl_wait();
  constexpr int PER = GALLERY_SEL_PER_THREAD;
  __shared__ float skey[GALLERY_SEL_MAX_SURV]; __shared__ int sidx[GALLERY_SEL_MAX_SURV];
  __shared__ float mins[128];
  __shared__ float red_k[4]; __shared__ float red_s[4];
  __shared__ float win_key[KT]; __shared__ int win_idx[KT]; __shared__ float ex_key[KT];
  __shared__ float srt_key[KT]; __shared__ int srt_idx[KT]; __shared__ float srt_win[KT];
  __shared__ float qstat[2]; __shared__ float bound_s; __shared__ int n_surv;
  __shared__ float red_acc[4][KT];
  const int64_t qi = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool cosine = (metric == B200F_METRIC_COS);
  const int n_all = __ldcg(cnt + qi);
  int ns = n_all < cap ? n_all : cap;
  bool overflow = n_all > cap || ns > 128 * PER;             // more than fits: the exact engine takes the query
  if (ns > 128 * PER) ns = 128 * PER;
  // ---- candidates -> registers, per-thread minimum
  float ck[PER]; int ci[PER];
  float lmin = INFINITY;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int i = tid + 128 * u;
    ck[u] = (i < ns) ? __ldcg(cand_key + qi * cap + i) : INFINITY;
    ci[u] = (i < ns) ? __ldcg(cand_idx + qi * cap + i) : -1;
    if (ci[u] >= 0) lmin = fminf(lmin, ck[u]);
  }
  mins[tid] = lmin;
  if (tid == 0) { n_surv = 0; bound_s = INFINITY; }
  if (tid < KT) { win_key[tid] = INFINITY; win_idx[tid] = -1; }
  const bool any_marked = (q_bad != nullptr && q_bad[qi] != 0) ||
                          (gmax_ptr != nullptr && reinterpret_cast<const int*>(gmax_ptr)[1] != 0);
  // the query row: thread t owns elements t, t + 128, t + 256, t + 384 (D <= 512)
  float xq[4];
  float nq = 0.f, sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = tid + 128 * i;
    xq[i] = (d < D) ? q[qi * D + d] : 0.f;
    if (d < D) { nq = fmaf(xq[i], xq[i], nq); sq += xq[i]; }
  }
  nq = warp_sum(nq); sq = warp_sum(sq);
  if (lane == 0) { red_k[wid] = nq; red_s[wid] = sq; }
  __syncthreads();
  if (tid == 0) { qstat[0] = red_k[0] + red_k[1] + red_k[2] + red_k[3]; qstat[1] = red_s[0] + red_s[1] + red_s[2] + red_s[3]; }
  // ---- bound: the KT-th smallest of the 128 per-thread minima (128 distinct candidates are <= their minima's maximum;
  // +inf when fewer than KT threads hold anything: keep all)
  {
    int before = 0;
    for (int j = 0; j < 128; ++j) { const float mj = mins[j]; before += (mj < lmin || (mj == lmin && j < tid)) ? 1 : 0; }
    if (before == KT - 1) bound_s = lmin;
  }
  __syncthreads();
  const float bound = bound_s;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    if (ci[u] >= 0 && ck[u] <= bound) {
      const int pos = atomicAdd(&n_surv, 1);
      if (pos < GALLERY_SEL_MAX_SURV) { skey[pos] = ck[u]; sidx[pos] = ci[u]; }
    }
  }
  __syncthreads();
  int nsv = n_surv;
  if (nsv > GALLERY_SEL_MAX_SURV) { nsv = GALLERY_SEL_MAX_SURV; overflow = true; }   // a crowd of equal keys
  // ---- the KT best approximate keys by (key, row): rank by counting (rows are distinct: a total order)
  for (int t = tid; t < nsv; t += 128) {
    const float kt = skey[t]; const int it = sidx[t];
    int rank = 0;
    for (int j = 0; j < nsv; ++j) { const float kj = skey[j]; rank += (kj < kt || (kj == kt && sidx[j] < it)) ? 1 : 0; }
    if (rank < KT) { win_key[rank] = kt; win_idx[rank] = it; }
  }
  __syncthreads();
  // ---- exact re-score of all KT winners at once (reference formula, fp32)
  {
    int id[KT];
    float y[4][KT];
#pragma unroll
    for (int u = 0; u < KT; ++u) id[u] = win_idx[u];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = tid + 128 * i;
#pragma unroll
      for (int u = 0; u < KT; ++u) y[i][u] = (d < D && id[u] >= 0) ? __ldg(g + (int64_t)id[u] * D + d) : 0.f;
    }
    float acc[KT];
#pragma unroll
    for (int u = 0; u < KT; ++u) acc[u] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (tid + 128 * i < D) {
#pragma unroll
        for (int u = 0; u < KT; ++u) {
          if (cosine) acc[u] = fmaf(xq[i], y[i][u], acc[u]);
          else { const float df = xq[i] - y[i][u] + GALLERY_EPS; acc[u] = fmaf(df, df, acc[u]); }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < KT; ++u) {
      const float a = warp_sum(acc[u]);
      if (lane == 0) red_acc[wid][u] = a;
    }
  }
  __syncthreads();
  if (tid < KT) {
    const int idt = win_idx[tid];
    const float tot = (red_acc[0][tid] + red_acc[1][tid]) + (red_acc[2][tid] + red_acc[3][tid]);
    float e;                                                   // ordering key, smaller is better (as the exact engine)
    if (idt < 0) e = INFINITY;
    else if (cosine) e = -(tot * (q_inv ? q_inv[qi] : 1.0f) * (g_inv ? g_inv[idt] : 1.0f));
    else e = sqrtf(tot);
    ex_key[tid] = e;
  }
  __syncthreads();
  // ---- order the KT winners by (exact key, row), NaN keys and empty slots last: the stable insertion sort of the
  // padded-list kernel, as a rank by counting (a precedes b iff first(a, b), or neither precedes and a came earlier)
  if (tid < KT) {
    const float ka = ex_key[tid]; const int ia = win_idx[tid];
    int pos = 0;
    for (int j = 0; j < KT; ++j) {
      if (j == tid) continue;
      const float kb = ex_key[j]; const int ib = win_idx[j];
      const bool j_first = (ib >= 0) && (ia < 0 || (kb == kb && (ka != ka || kb < ka || (kb == ka && ib < ia))));
      const bool me_first = (ia >= 0) && (ib < 0 || (ka == ka && (kb != kb || ka < kb || (ka == kb && ia < ib))));
      pos += (j_first || (!me_first && j < tid)) ? 1 : 0;
    }
    srt_key[pos] = ka; srt_idx[pos] = ia; srt_win[pos] = win_key[tid];
  }
  __syncthreads();
  if (tid == 0) {
    // worst kept approximate key = lower bound of every excluded row's approximate key
    float a_excl = -INFINITY; int n_valid = 0;
    for (int r = 0; r < KT; ++r) if (srt_idx[r] >= 0) { a_excl = fmaxf(a_excl, srt_win[r]); ++n_valid; }
    bool verified = !any_marked && !overflow;
    if (verified && n_valid == KT) {                           // something may have been excluded
      const int kk = min(k, n_valid);
      const float ek = srt_key[kk - 1];                        // exact k-th best ordering key
      const float qn = sqrtf(qstat[0]);
      const float gmax = gmax_ptr ? __int_as_float(*reinterpret_cast<const int*>(gmax_ptr)) : 1.0f;
      float exact_in_approx_units, delta;
      const float rel = (fmt == B200F_OPERAND_FP16) ? GALLERY_DOT_ERR_FP16 : GALLERY_DOT_ERR;
      const float abs_e = (fmt == B200F_OPERAND_FP16) ? GALLERY_ABS_ERR_FP16 * sqrtf((float)D) : 0.f;
      if (cosine) {
        const float qv = q_inv ? q_inv[qi] : 1.0f;
        exact_in_approx_units = (qv > 0.f) ? ek / qv : -INFINITY;
        delta = rel * qn * 1.01f + abs_e * (qn + 1.0f);
      } else {
        exact_in_approx_units = ek * ek - (qstat[0] + 2.0f * GALLERY_EPS * qstat[1] + (float)D * GALLERY_EPS * GALLERY_EPS);
        delta = 2.0f * (rel * qn * gmax + abs_e * (qn + gmax)) + 1e-6f * (qstat[0] + gmax * gmax + 1.0f);
      }
      verified = (ek == ek) && (exact_in_approx_units < a_excl - delta);
    }
    if (accept != nullptr) {
      const bool ok = srt_idx[0] >= 0 && srt_key[0] == srt_key[0];
      const float best = cosine ? -srt_key[0] : srt_key[0];
      accept[qi] = ok && (cosine ? (best >= thresh) : (best <= thresh));
    }
    redo[qi] = verified ? 0 : 1;
    if (!verified && redo_count != nullptr) atomicAdd(redo_count, 1);
  }
  if (tid < k) {
    const bool ok = (tid < KT) && srt_idx[tid] >= 0 && srt_key[tid] == srt_key[tid];
    idx_out[qi * k + tid] = ok ? (index_offset + srt_idx[tid]) : -1;
    score_out[qi * k + tid] = ok ? (cosine ? -srt_key[tid] : srt_key[tid]) : (cosine ? -INFINITY : INFINITY);
  }
}

// ---- prepare: rows -> 16-bit scan operand (+ bias, + max row norm) -------------------------------------
// One warp per row (D <= 512).  This scalar form takes any D; rows of whole 8-element chunks at 16-byte aligned
// addresses -- everything the tensor engine scans -- go through gallery_prepare_vec8_kernel below.  metric COS: out = g / max(|g|, 1e-12); L2EPS: out = g; stored as bf16
// (any range) or fp16 (8x tighter error bound; |values| must stay well inside +-65504 -- an overflow is safe, it only
// sends the affected queries to the exact engine).
// bias[r] = |g|^2 - 2 eps sum(g) (L2EPS) / 0 (COS); two extra slots: bias[rows] = max row norm (atomicMax),
// bias[rows + 1] = number of rows the 16-bit operand cannot represent (int).  row_bad (optional): the per-row flag.
// Largest row norm through an atomicMax on the int view (norms are >= 0), attempted only when the value beats what the
// slot already holds: one unconditional atomic per row on ONE address serialised in the L2 and made the kernel take
// 1.65 ms on a 1 M x 512 gallery, against 0.5 ms of HBM time (profiles/r01_rowops_full.md).  A stale read only costs
// an extra atomic.
__device__ __forceinline__ void raise_max_norm(float* slot, float eff) {
  if (eff == eff && __float_as_int(eff) > __ldcg(reinterpret_cast<const int*>(slot)))
    atomicMax(reinterpret_cast<int*>(slot), __float_as_int(eff));
}

template <typename TI>
__global__ void __launch_bounds__(256)
gallery_prepare_kernel(const TI* __restrict__ in, int64_t rows, int dim, int metric, int fmt, uint16_t* __restrict__ out,
                       float* __restrict__ bias, uint8_t* __restrict__ row_bad) {
  pdl_trigger(); pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[16];                                                // dim <= 512: 16 elements per lane
  float ss = 0.f, sm = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int d = lane + 32 * i;
    v[i] = (d < dim) ? to_f32<TI>(in[row * dim + d]) : 0.f;
    ss = fmaf(v[i], v[i], ss); sm += v[i];
  }
  ss = warp_sum(ss); sm = warp_sum(sm);
  const float nrm = sqrtf(ss);
  const float sc = (metric == B200F_METRIC_COS) ? 1.0f / fmaxf(nrm, 1e-12f) : 1.0f;
  // a row that is finite in fp32 but not in its 16-bit form (fp16 overflow), or that holds an infinity, cannot be
  // ranked by the scan the way the exact engine ranks it: flag it (all-NaN-distance rows are dropped by both)
  bool lost = false;
  if (out != nullptr) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int d = lane + 32 * i;
      if (d < dim) {
        const float f = v[i] * sc;
        const uint16_t h = (fmt == B200F_OPERAND_FP16) ? __half_as_ushort(__float2half_rn(f)) : __bfloat16_as_ushort(__float2bfloat16_rn(f));
        const float back = (fmt == B200F_OPERAND_FP16) ? __half2float(__ushort_as_half(h)) : __bfloat162float(__ushort_as_bfloat16(h));
        lost = lost || ((back - back != 0.f) && (f == f));    // non-finite result from a non-NaN input
        out[row * dim + d] = h;
      }
    }
  }
  lost = __any_sync(0xffffffffu, lost);
  if (lane == 0 && row_bad != nullptr) row_bad[row] = lost ? 1 : 0;
  if (bias != nullptr && lane == 0) {
    bias[row] = (metric == B200F_METRIC_COS) ? 0.f : (ss - 2.0f * GALLERY_EPS * sm);
    const float eff = (metric == B200F_METRIC_COS) ? 1.0f : nrm;
    raise_max_norm(bias + rows, eff);
    if (lost) atomicAdd(reinterpret_cast<int*>(bias + rows + 1), 1);                       // unrankable rows
  }
}

// The same for dim % 8 == 0 with 16-byte aligned rows (what the tensor engine takes anyway): a lane owns chunks of 8
// consecutive elements (chunks lane and lane + 32), loaded as 32-byte (fp32) / 16-byte (bf16) vectors and stored as
// one 16-byte vector of 16-bit values each.  The scalar kernel above issues 2-byte stores and ran at 1.9 TB/s on a
// 1 M x 512 gallery (profiles/r01_rowops_full.md).
template <typename TI>
__global__ void __launch_bounds__(256)
gallery_prepare_vec8_kernel(const TI* __restrict__ in, int64_t rows, int dim, int metric, int fmt, uint16_t* __restrict__ out,
                            float* __restrict__ bias, uint8_t* __restrict__ row_bad) {
  pdl_trigger(); pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int chunks = dim >> 3;                                // <= 64
  float v[2][8];
  float ss = 0.f, sm = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = lane + 32 * i;
    if (c < chunks) {
      load8<TI>(in + row * dim + 8 * c, 8, true, v[i]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { ss = fmaf(v[i][e], v[i][e], ss); sm += v[i][e]; }
  }
  ss = warp_sum(ss); sm = warp_sum(sm);
  const float nrm = sqrtf(ss);
  const float sc = (metric == B200F_METRIC_COS) ? 1.0f / fmaxf(nrm, 1e-12f) : 1.0f;
  bool lost = false;
  if (out != nullptr) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      if (c < chunks) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          uint16_t h2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float f = v[i][e + u] * sc;
            h2[u] = (fmt == B200F_OPERAND_FP16) ? __half_as_ushort(__float2half_rn(f)) : __bfloat16_as_ushort(__float2bfloat16_rn(f));
            const float back = (fmt == B200F_OPERAND_FP16) ? __half2float(__ushort_as_half(h2[u])) : __bfloat162float(__ushort_as_bfloat16(h2[u]));
            lost = lost || ((back - back != 0.f) && (f == f));  // non-finite result from a non-NaN input
          }
          pk[e >> 1] = (uint32_t)h2[0] | ((uint32_t)h2[1] << 16);
        }
        *reinterpret_cast<uint4*>(out + row * dim + 8 * c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  lost = __any_sync(0xffffffffu, lost);
  if (lane == 0 && row_bad != nullptr) row_bad[row] = lost ? 1 : 0;
  if (bias != nullptr && lane == 0) {
    bias[row] = (metric == B200F_METRIC_COS) ? 0.f : (ss - 2.0f * GALLERY_EPS * sm);
    const float eff = (metric == B200F_METRIC_COS) ? 1.0f : nrm;
    raise_max_norm(bias + rows, eff);
    if (lost) atomicAdd(reinterpret_cast<int*>(bias + rows + 1), 1);                       // unrankable rows
  }
}

// ---- sample bound: KT-th smallest key over a query's (<= 256) sorted sample lists ------------------------------
// One warp per query, lane l walks lists l, l + 32, ..: KT rounds of a shuffle arg-min pop the global order.
// tau0 = +inf when the sample holds fewer than KT ranked elements (no bound).
constexpr int GALLERY_TAU_LISTS_PER_LANE = 8;
template <int KT>
__global__ void __launch_bounds__(128)
gallery_tau_kernel(const float* __restrict__ cand_key, const int32_t* __restrict__ cand_idx, int n_lists, int64_t Q,
                   float* __restrict__ tau0) {
  pdl_trigger(); pdl_wait();
  constexpr int L = GALLERY_TAU_LISTS_PER_LANE;
  const int lane = threadIdx.x & 31;
  const int64_t qi = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (qi >= Q) return;
  const float* lk = cand_key + qi * n_lists * KT;
  const int32_t* li = cand_idx + qi * n_lists * KT;
  int pos[L];
  float head[L];                                             // current head of each of my lists (+inf: exhausted)
  // both loads of every head are issued before any is used: guarded by the index (`id >= 0 ? key : inf`) the key load
  // depended on the index load and the eight heads of a lane cost sixteen serial global-memory latencies (11 of 14 us)
  int hid[L];
#pragma unroll
  for (int u = 0; u < L; ++u) {
    pos[u] = 0;
    const int l = lane + 32 * u;
    const bool in = l < n_lists;
    hid[u] = in ? __ldg(li + (in ? l : 0) * KT) : -1;
    head[u] = in ? __ldg(lk + (in ? l : 0) * KT) : INFINITY;
  }
#pragma unroll
  for (int u = 0; u < L; ++u) head[u] = (hid[u] >= 0) ? head[u] : INFINITY;
  float last = INFINITY;
  bool short_of = false;
  // With many lists (>= 4 KT: the one-row-group pre-pass has 256) the KT-th smallest list HEAD is used: KT distinct
  // elements are <= it, so it bounds the KT-th best key as well, it is nearly as tight (the best KT elements of 16 k
  // rows rarely share one of 256 lists), and no round waits for a dependent global load (14 us -> ~3 us).
  const bool heads_only = n_lists >= 4 * KT;
  for (int r = 0; r < KT; ++r) {
    float k = INFINITY; int who = 1 << 20;
#pragma unroll
    for (int u = 0; u < L; ++u)
      if (head[u] < k) { k = head[u]; who = lane + 32 * u; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ok = __shfl_xor_sync(0xffffffffu, k, o);
      const int ow = __shfl_xor_sync(0xffffffffu, who, o);
      if (ok < k || (ok == k && ow < who)) { k = ok; who = ow; }
    }
    if (who >= (1 << 20)) { short_of = true; break; }
#pragma unroll
    for (int u = 0; u < L; ++u) {
      if (who == lane + 32 * u) {                             // pop my list
        ++pos[u];
        head[u] = (!heads_only && pos[u] < KT && li[who * KT + pos[u]] >= 0) ? lk[who * KT + pos[u]] : INFINITY;
      }
    }
    last = k;
  }
  if (lane == 0) tau0[qi] = short_of ? INFINITY : last;
}

// ---- select: merge lists, exact re-rank, verification ----------------------------------------------------
// One block of 128 threads per query.  Candidates in shared memory; KT rounds of block-wide lexicographic
// arg-min pick the KT best approximate keys; each is re-scored exactly (thread t owns elements t, t+128, ..).
template <typename TG, int KT>
__global__ void __launch_bounds__(128)
gallery_select_kernel(const float* __restrict__ cand_key, const int32_t* __restrict__ cand_idx, int n_cand,
                      const float* __restrict__ q, const TG* __restrict__ g, const float* __restrict__ q_inv,
                      const float* __restrict__ g_inv, const float* __restrict__ gmax_ptr,
                      const uint8_t* __restrict__ q_bad, int64_t Q, int D, int k,
                      int metric, int fmt, float thresh, int64_t index_offset, int64_t* __restrict__ idx_out,
                      float* __restrict__ score_out, uint8_t* __restrict__ accept, uint8_t* __restrict__ redo,
                      int32_t* __restrict__ redo_count) {
  pdl_trigger(); pdl_wait();
  extern __shared__ uint8_t sel_smem[];
  float* ckey = reinterpret_cast<float*>(sel_smem);
  int32_t* cidx = reinterpret_cast<int32_t*>(ckey + n_cand);
  __shared__ float red_k[4]; __shared__ int red_i[4]; __shared__ int red_pos[4];
  __shared__ float win_key[KT]; __shared__ int win_idx[KT]; __shared__ float ex_key[KT];
  __shared__ float qstat[2];
  const int64_t qi = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool cosine = (metric == B200F_METRIC_COS);
  // The list heads are also kept as a dense array: read from ckey[h * KT] they sit KT words apart, a 16-way bank
  // conflict on every read of the KT selection rounds below.
  constexpr int MAX_HEADS = 512;
  __shared__ float head_k[MAX_HEADS]; __shared__ int head_i[MAX_HEADS];
  const bool dense_heads = (n_cand / KT) <= MAX_HEADS && (n_cand & 3) == 0;
  if ((n_cand & 3) == 0) {                                   // n_cand = n_lists * KT: 16-byte copies
    const float4* gk = reinterpret_cast<const float4*>(cand_key + qi * n_cand);
    const int4* gi = reinterpret_cast<const int4*>(cand_idx + qi * n_cand);
    for (int i = tid; i < n_cand / 4; i += 128) {
      const float4 kv = __ldg(gk + i); const int4 iv = __ldg(gi + i);
      reinterpret_cast<float4*>(ckey)[i] = kv;
      reinterpret_cast<int4*>(cidx)[i] = iv;
      if (dense_heads && (i % (KT / 4)) == 0) { head_k[i / (KT / 4)] = kv.x; head_i[i / (KT / 4)] = iv.x; }
    }
  } else {
    for (int i = tid; i < n_cand; i += 128) { ckey[i] = cand_key[qi * n_cand + i]; cidx[i] = cand_idx[qi * n_cand + i]; }
  }
  // no proof is possible for a query or a gallery whose 16-bit operand lost values (overflow, infinities)
  const bool any_marked = (q_bad != nullptr && q_bad[qi] != 0) ||
                          (gmax_ptr != nullptr && reinterpret_cast<const int*>(gmax_ptr)[1] != 0);
  // per-query constants: |q|^2, sum q
  float nq = 0.f, sq = 0.f;
  for (int d = tid; d < D; d += 128) { const float x = q[qi * D + d]; nq = fmaf(x, x, nq); sq += x; }
  nq = warp_sum(nq); sq = warp_sum(sq);
  if (lane == 0) { red_k[wid] = nq; ex_key[wid] = sq; }
  __syncthreads();
  if (tid == 0) { qstat[0] = red_k[0] + red_k[1] + red_k[2] + red_k[3]; qstat[1] = ex_key[0] + ex_key[1] + ex_key[2] + ex_key[3]; }
  __syncthreads();
  // ---- the KT best approximate keys.  Every list is sorted, so the KT-th smallest list HEAD bounds the answer:
  // the KT smallest heads are already KT candidates <= it.  Warp 0 finds that bound over the heads (KT rounds of
  // shuffle arg-min, no block barrier), all threads compact the few candidates under it, warp 0 ranks the survivors.
  __shared__ float bound_s; __shared__ int n_surv;
  const int n_heads = n_cand / KT;
  if (tid == 0) n_surv = 0;
  if (wid == 0) {
    float taken_k = -INFINITY; int taken_i = -1;
    float hb = INFINITY;
    for (int r = 0; r < KT; ++r) {
      float bk = INFINITY; int bi = INT32_MAX;
      for (int h = lane; h < n_heads; h += 32) {
        const int id = dense_heads ? head_i[h] : cidx[h * KT]; const float kk = dense_heads ? head_k[h] : ckey[h * KT];
        const bool after = (kk > taken_k) || (kk == taken_k && id > taken_i);
        if (id >= 0 && after && (kk < bk || (kk == bk && id < bi))) { bk = kk; bi = id; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
      }
      if (bi == INT32_MAX) { hb = INFINITY; break; }          // fewer than KT non-empty lists: keep everything
      taken_k = bk; taken_i = bi; hb = bk;
    }
    if (lane == 0) bound_s = hb;
  }
  __syncthreads();
  const float bound = bound_s;
  float* skey = reinterpret_cast<float*>(cidx + n_cand);       // survivors (capacity n_cand)
  int32_t* sidx = reinterpret_cast<int32_t*>(skey + n_cand);
  for (int i = tid; i < n_cand; i += 128) {
    const int id = cidx[i]; const float kk = ckey[i];
    if (id >= 0 && kk <= bound) { const int pos = atomicAdd(&n_surv, 1); skey[pos] = kk; sidx[pos] = id; }
  }
  __syncthreads();
  if (wid == 0) {
    const int ns = n_surv;
    float taken_k = -INFINITY; int taken_i = -1;
    for (int r = 0; r < KT; ++r) {
      float bk = INFINITY; int bi = INT32_MAX;
      for (int i = lane; i < ns; i += 32) {
        const float kk = skey[i]; const int id = sidx[i];
        const bool after = (kk > taken_k) || (kk == taken_k && id > taken_i);
        if (after && (kk < bk || (kk == bk && id < bi))) { bk = kk; bi = id; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
      }
      if (lane == 0) { win_key[r] = bk; win_idx[r] = (bi == INT32_MAX) ? -1 : bi; }
      if (bi != INT32_MAX) { taken_k = bk; taken_i = bi; }
    }
  }
  __syncthreads();
  // ---- exact re-score of the winners (reference formula, fp32), four rows at a time: the rows are random 2 KB reads
  // from HBM, and one row per barrier round their latencies added up to half of this kernel
  __shared__ float red_acc[4][4];
  for (int r0 = 0; r0 < KT; r0 += 4) {
    int id[4];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) id[u] = win_idx[r0 + u];
    for (int d = tid; d < D; d += 128) {
      const float x = q[qi * D + d];
      float y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) y[u] = (id[u] >= 0) ? to_f32<TG>(g[(int64_t)id[u] * D + d]) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (cosine) acc[u] = fmaf(x, y[u], acc[u]);
        else { const float df = x - y[u] + GALLERY_EPS; acc[u] = fmaf(df, df, acc[u]); }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float a = warp_sum(acc[u]);
      if (lane == 0) red_acc[wid][u] = a;
    }
    __syncthreads();
    if (tid < 4) {
      const int idt = win_idx[r0 + tid];
      const float tot = (red_acc[0][tid] + red_acc[1][tid]) + (red_acc[2][tid] + red_acc[3][tid]);
      // ordering key, smaller is better: the distance / minus the cosine (as the exact engine)
      float e;
      if (idt < 0) e = INFINITY;
      else if (cosine) e = -(tot * (q_inv ? q_inv[qi] : 1.0f) * (g_inv ? g_inv[idt] : 1.0f));
      else e = sqrtf(tot);
      ex_key[r0 + tid] = e;
    }
    __syncthreads();
  }
  if (tid == 0) {
    // insertion sort of the KT winners by (exact key, index); NaN keys go last
    for (int a = 1; a < KT; ++a) {
      const float ka = ex_key[a]; const int ia = win_idx[a]; const float wa = win_key[a];
      int b = a - 1;
      while (b >= 0) {
        const float kb = ex_key[b]; const int ib = win_idx[b];
        const bool a_first = (ia >= 0) && (ib < 0 || (ka == ka && (kb != kb || ka < kb || (ka == kb && ia < ib))));
        if (!a_first) break;
        ex_key[b + 1] = kb; win_idx[b + 1] = ib; win_key[b + 1] = win_key[b];
        --b;
      }
      ex_key[b + 1] = ka; win_idx[b + 1] = ia; win_key[b + 1] = wa;
    }
    // worst kept approximate key = lower bound of every excluded row's approximate key
    float a_excl = -INFINITY; int n_valid = 0;
    for (int r = 0; r < KT; ++r) if (win_idx[r] >= 0) { a_excl = fmaxf(a_excl, win_key[r]); ++n_valid; }
    bool verified = !any_marked;
    if (verified && n_valid == KT && n_cand > 0) {                         // something may have been excluded
      const int kk = min(k, n_valid);
      const float ek = ex_key[kk - 1];                         // exact k-th best ordering key
      const float qn = sqrtf(qstat[0]);
      const float gmax = gmax_ptr ? __int_as_float(*reinterpret_cast<const int*>(gmax_ptr)) : 1.0f;
      float exact_in_approx_units, delta;
      const float rel = (fmt == B200F_OPERAND_FP16) ? GALLERY_DOT_ERR_FP16 : GALLERY_DOT_ERR;
      const float abs_e = (fmt == B200F_OPERAND_FP16) ? GALLERY_ABS_ERR_FP16 * sqrtf((float)D) : 0.f;
      if (cosine) {
        // approx a = -<q, g_hat>; exact ordering key e = -cos * ... = -(<q,g> g_inv) q_inv  ->  -<q,g_hat> = e / q_inv
        const float qv = q_inv ? q_inv[qi] : 1.0f;
        exact_in_approx_units = (qv > 0.f) ? ek / qv : -INFINITY;
        delta = rel * qn * 1.01f + abs_e * (qn + 1.0f);
      } else {
        // approx a = d^2 - (|q|^2 + 2 eps sum q + D eps^2)
        exact_in_approx_units = ek * ek - (qstat[0] + 2.0f * GALLERY_EPS * qstat[1] + (float)D * GALLERY_EPS * GALLERY_EPS);
        delta = 2.0f * (rel * qn * gmax + abs_e * (qn + gmax)) + 1e-6f * (qstat[0] + gmax * gmax + 1.0f);
      }
      verified = (ek == ek) && (exact_in_approx_units < a_excl - delta);
    }
    for (int s = 0; s < k; ++s) {
      const bool ok = (s < KT) && win_idx[s] >= 0 && ex_key[s] == ex_key[s];
      idx_out[qi * k + s] = ok ? (index_offset + win_idx[s]) : -1;
      score_out[qi * k + s] = ok ? (cosine ? -ex_key[s] : ex_key[s]) : (cosine ? -INFINITY : INFINITY);
    }
    if (accept != nullptr) {
      const bool ok = win_idx[0] >= 0 && ex_key[0] == ex_key[0];
      const float best = cosine ? -ex_key[0] : ex_key[0];
      accept[qi] = ok && (cosine ? (best >= thresh) : (best <= thresh));
    }
    redo[qi] = verified ? 0 : 1;
    if (!verified && redo_count != nullptr) atomicAdd(redo_count, 1);
  }
}

}  // namespace umma
}  // namespace b200f
