// Persistent, warp-specialised tcgen05 GEMM core (1 CTA per SM), both operands streamed:
//
//   D[m, n] = sum_k A(m, k) * B(n, k)      bf16 (or fp16) operands, fp32 accumulation in TMEM
//
// PAIR = 1: cta_group::1, one CTA per 128 x 256 tile, 4 stages of (16 KB A | 32 KB B).
// PAIR = 2: the two CTAs of a cluster form a cta_group::2 pair on one 256 x 256 tile: each loads ITS 128 rows of A
//           and ITS 128 of the tile's 256 B rows (6 stages of 16 KB | 16 KB), the leader issues 256 x 256 x 16 MMAs that
//           read both CTAs' shared memory.  Per k-block a CTA now ingests 32 KB for the work it took 48 KB for:
//           the single-CTA core is L2->SM bound (615 MB for cfg3's dX GEMM, 13.7 TB/s), the pair moves 2/3 of that.
//
//   warp 0      : TMA producer   (global -> 128B-swizzled shared tiles, 4-stage mbarrier ring)
//   warp 1      : MMA issuer     (one elected lane issues tcgen05.mma 128x256x16; accumulators are
//                                 double-buffered in TMEM: 2 x 256 columns)
//   warps 2..5  : epilogue       (tcgen05.ld 32 lanes x 32 columns at a time; thread t of the warp
//                                 owns accumulator row 32*(warp%4)+t, so per-row reductions are
//                                 thread-local -- no shuffles)
//
// Operand layouts (both selectable per operand):
//   K-major : element (r, k) at base[r*ld + k]   (TMA box 64 k x ROWS rows)
//   MN-major: element (r, k) at base[k*ld + r]   (TMA boxes of 64 r x 64 k, one per 64-wide block)
// Work items are (m_tile, n_tile, k_split) with m fastest so CTAs that run concurrently share the same
// B tiles through L2.  The stage-specific work lives in the Epilogue policy.
#pragma once
#include "common.cuh"
#include "umma_core.cuh"

namespace b200f {
namespace umma {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;                                // PAIR = 1; see gemm_stages()
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KB
constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;      // 32 KB
constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARP0 = 2;
constexpr int EPI_THREADS = 128;
constexpr int MN_BLOCK_BYTES = 64 * BLOCK_K * 2;          // one 64-wide MN block of 64 k-rows: 8 KB
constexpr int EPI_SMEM_FLOATS = 1024;                     // scratch for the epilogue policy
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + EPI_SMEM_FLOATS * 4 + 256;
constexpr int MAX_STAGES = 6;
constexpr int FOLLOW_TILE_K = 256;                         // follow mode: k extent of one tile of the neighbour kernel (its 256-class tiles)
__host__ __device__ constexpr int gemm_b_rows(int pair) { return BLOCK_N / pair; }                  // B rows a CTA loads
__host__ __device__ constexpr int gemm_stage_bytes(int pair) { return A_TILE_BYTES + gemm_b_rows(pair) * BLOCK_K * 2; }
__host__ __device__ constexpr int gemm_stages(int pair) { return STAGES * STAGE_BYTES / gemm_stage_bytes(pair); }   // 4 | 6

struct GemmParams {
  int M, N, K;              // extents of the m, n and k index spaces
  int m_tiles, n_tiles, k_splits;
  int k_per_split;          // multiple of BLOCK_K
  // descriptor knobs (bytes); defaults in default_params(); the self-test can override them
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep;
  uint32_t idesc;
  int n_fastest;            // 1: work items run (n_tile, m_tile, k_split) with n fastest: the clusters that run side by side
                            // share their A rows through L2 (dW at batch > 512: A = G^T, 2 MB per 256 class rows, would
                            // otherwise come from HBM once per n tile -- 2.24 GB read per launch at 4096 x 125 k, ncu)
  // "follow" mode (K3c beside K3b, b200f_arcface_bwd_part): the k range is walked in the order in which the dW kernel that
  // runs next to this one reads the same class rows -- its follow_chunks class chunks of the follow_tiles 256-class tiles,
  // each last tile first (follow_rev) or first tile first, all chunks in step -- dealt round-robin to the k_splits splits.
  // Both kernels read G^T and w_hat (204 MB at cfg3) and nothing else of size; walked at their own orders one kernel's
  // lines are gone from the L2 before the other asks for them.  0 = off (a split is a contiguous k range).
  int follow_chunks, follow_tiles, follow_rev;
  // stream-K: the (tile, k-block) space -- tiles in decode order, k_units k-blocks each -- is cut into one contiguous,
  // equally long range per cluster; a range that crosses a tile boundary gives its cluster two (rarely three) pieces.
  // A piece's output slot (TileCoord::split) is its cluster's rank among the clusters that touch the tile; the reduction
  // sums stream_k_pieces(tile) slots.  Why: with 32 output tiles and 74 clusters (dx at batch 4096) split-K by 2 leaves
  // 10 clusters idle and split-K by 3 runs two waves; 0 = off.
  int stream_k, k_units;
  // A operand in [32 classes x 64 batch rows] blocks of 4 KB (a_blocked = 1; G^T at batch > 512, umma_head.cu "gt_blocked"):
  // tm_a is a rank-4 map {batch row within block, class within block, batch block, class block}
  int a_blocked;
  int early;                // 1: neither operand nor the output is touched by the predecessor grid (K3c behind K3b: both only READ
                            // G^T): nobody waits for it up front -- CTAs start on SMs the predecessor has left -- and the epilogue
                            // warps wait at their END, so that this grid still completes after its predecessor (stream order holds
                            // for everything behind it)
};

struct TileCoord { int m0, n0, split, k_begin, k_end, nb0; };   // m0: first row of THIS CTA; nb0: first B row it loads

// m_tiles counts tiles of 128 * PAIR rows
template <int PAIR>
__device__ __forceinline__ TileCoord decode_work(const GemmParams& p, int w, int rank) {
  TileCoord t;
  const int m = p.n_fastest ? (w / p.n_tiles) % p.m_tiles : w % p.m_tiles;
  const int n = p.n_fastest ? w % p.n_tiles : (w / p.m_tiles) % p.n_tiles;
  t.split = w / (p.m_tiles * p.n_tiles);
  t.m0 = (m * PAIR + rank) * BLOCK_M; t.n0 = n * BLOCK_N;
  t.nb0 = t.n0 + rank * gemm_b_rows(PAIR);
  if (p.follow_chunks > 0) {                                    // LOGICAL k range: this split's share of the tile sequence
    t.k_begin = 0;
    t.k_end = ((p.follow_tiles - t.split + p.k_splits - 1) / p.k_splits) * FOLLOW_TILE_K;
  } else {
    t.k_begin = t.split * p.k_per_split;
    t.k_end = min(p.K, t.k_begin + p.k_per_split);
  }
  return t;
}

// ---- work items of a cluster: split-K (item w = cluster, cluster + n_clusters, ...) or stream-K ------------------------
struct WorkIter { long long pos, hi; };
__device__ __forceinline__ void work_begin(const GemmParams& p, int cluster_id, int n_clusters, WorkIter& it) {
  if (p.stream_k) {
    const long long total = (long long)p.m_tiles * p.n_tiles * p.k_units;
    it.pos = total * cluster_id / n_clusters; it.hi = total * (cluster_id + 1) / n_clusters;
  } else {
    it.pos = cluster_id; it.hi = (long long)p.m_tiles * p.n_tiles * p.k_splits;
  }
}
// the cluster whose range holds position pos of the (tile, k-block) space (ranges: [total c / n, total (c + 1) / n))
__host__ __device__ __forceinline__ int stream_k_owner(long long pos, long long total, int n_clusters) {
  return (int)(((pos + 1) * n_clusters - 1) / total);
}
// slots the reduction has to sum for a tile
__host__ __device__ __forceinline__ int stream_k_pieces(int tile, int k_units, long long total, int n_clusters) {
  return stream_k_owner((long long)(tile + 1) * k_units - 1, total, n_clusters) - stream_k_owner((long long)tile * k_units, total, n_clusters) + 1;
}
template <int PAIR>
__device__ __forceinline__ bool work_next(const GemmParams& p, int cluster_id, int n_clusters, int rank, WorkIter& it, TileCoord& t) {
  if (it.pos >= it.hi) return false;
  if (!p.stream_k) { t = decode_work<PAIR>(p, (int)it.pos, rank); it.pos += n_clusters; return true; }
  const int U = p.k_units;
  const int tile = (int)(it.pos / U);
  const int kb = (int)(it.pos - (long long)tile * U);
  long long len = it.hi - it.pos;
  if (len > U - kb) len = U - kb;
  t = decode_work<PAIR>(p, tile, rank);                         // m0 / n0 / nb0 of the tile
  t.k_begin = kb * BLOCK_K;
  t.k_end = min(p.K, (kb + (int)len) * BLOCK_K);
  t.split = cluster_id - stream_k_owner((long long)tile * U, (long long)p.m_tiles * p.n_tiles * U, n_clusters);
  it.pos += len;
  return true;
}

// follow mode: first k of the tile that stands at position n of the neighbour's reading order.  Its chunk c covers tiles
// [c T / nc, (c + 1) T / nc) -- base or base + 1 of them; step i of all chunks comes before step i + 1 of any.
__device__ __forceinline__ int follow_tile_k(const GemmParams& p, int n) {
  const int T = p.follow_tiles, nc = p.follow_chunks;
  const int base = T / nc;
  int i, c;
  if (n < base * nc) { i = n / nc; c = n - i * nc; }
  else {                                                       // the (n - base nc)-th of the chunks that have base + 1 tiles
    i = base; c = 0;
    int left = n - base * nc;
    for (; c < nc; ++c) {
      const int len = (int)(((int64_t)(c + 1) * T) / nc) - (int)(((int64_t)c * T) / nc);
      if (len > base && left-- == 0) break;
    }
  }
  const int tb = (int)(((int64_t)c * T) / nc), te = (int)(((int64_t)(c + 1) * T) / nc);
  return (p.follow_rev ? te - 1 - i : tb + i) * FOLLOW_TILE_K;
}

// Epilogue policy interface:
//   struct Epi { struct Params; static __device__ void run(const Params&, const GemmParams&, const TileCoord&,
//                uint32_t tmem_acc /*lane 0, first column of this accumulator stage*/, int quad /*warp%4*/,
//                int lane, int epi_tid /*0..127*/, float* scratch /*EPI_SMEM_FLOATS, epilogue-only smem*/); }
// run() must finish all its tcgen05.ld (tmem_ld_wait) before returning.
template <int PAIR, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
            const GemmParams p, const typename Epi::Params ep) {
  constexpr int NST = gemm_stages(PAIR);
  constexpr int STG_BYTES = gemm_stage_bytes(PAIR);
  constexpr int B_ROWS = gemm_b_rows(PAIR);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;                                        // NST x (A | B), 1024 B aligned
  float* scratch = reinterpret_cast<float*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + EPI_SMEM_FLOATS);
  uint64_t* full_bar = bars;                                    // [NST]   TMA -> MMA (leader CTA)
  uint64_t* empty_bar = bars + MAX_STAGES;                      // [NST]   MMA -> TMA (both CTAs)
  uint64_t* acc_full = bars + 2 * MAX_STAGES;                   // [ACC_STAGES] MMA -> epilogue (both CTAs)
  uint64_t* acc_empty = bars + 2 * MAX_STAGES + ACC_STAGES;     // [ACC_STAGES] epilogue (both CTAs) -> MMA (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 2 * ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (PAIR == 2) ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / PAIR;
  const int n_clusters = gridDim.x / PAIR;

  // An early kernel (K3c) lets its dependents be scheduled right away: they wait for its completion themselves.  Every other
  // use triggers only AFTER its own wait (below): a dependent that starts early (K3c behind the streamed K3b) relies on
  // everything older than its predecessor being complete, and that holds only if the predecessor could not release it
  // before having waited itself.  (Triggering at the top let K3c read G^T while K3a was still writing it whenever the
  // grids were small enough to be co-resident: NaN in dx at B = 640 x 24 k classes in three chunks.)
  if (p.early) pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], PAIR); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], PAIR * EPI_THREADS / 32); }
    fence_barrier_init();
  }
  if (warp == 1) xw_tmem_alloc<PAIR>(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();                           // the peer's barriers exist before anyone signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (!p.early) { pdl_wait(); pdl_trigger(); }

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    {
      const bool leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      bool ok = true;
      WorkIter wi; work_begin(p, cluster_id, n_clusters, wi);
      TileCoord t;
      while (ok && work_next<PAIR>(p, cluster_id, n_clusters, rank, wi, t)) {
        int k_tile = 0;                                         // follow mode: physical k of the current logical tile
        for (int kl = t.k_begin; kl < t.k_end && ok; kl += BLOCK_K) {
          int k0 = kl;
          if (p.follow_chunks > 0) {
            if ((kl & (FOLLOW_TILE_K - 1)) == 0) k_tile = follow_tile_k(p, (kl / FOLLOW_TILE_K) * p.k_splits + t.split);
            k0 = k_tile + (kl & (FOLLOW_TILE_K - 1));
          }
          ok = mbar_wait(&empty_bar[stage], phase ^ 1);
          if (!ok) break;
          if (leader) {
            uint8_t* sa = tiles + (size_t)stage * STG_BYTES;
            uint8_t* sb = sa + A_TILE_BYTES;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], PAIR * STG_BYTES);
            else mbar_arrive_cluster(&full_bar[stage], 0);
            if (p.a_blocked) {                                  // class rows and batch rows are multiples of 64 here
              if (!A_MN) {                                      // rows = classes (t.m0), k = batch
                xw_tma_load4<PAIR>(sa, &tm_a, &full_bar[stage], 0, 0, k0 >> 6, t.m0 >> 5);
              } else {                                          // rows = batch (t.m0), k = classes
#pragma unroll
                for (int j = 0; j < BLOCK_M / 64; ++j)
                  xw_tma_load4<PAIR>(sa + j * MN_BLOCK_BYTES, &tm_a, &full_bar[stage], 0, 0, (t.m0 + 64 * j) >> 6, k0 >> 5);
              }
            } else
            if (!A_MN) {
              xw_tma_load<PAIR>(sa, &tm_a, &full_bar[stage], k0, t.m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                xw_tma_load<PAIR>(sa + j * MN_BLOCK_BYTES, &tm_a, &full_bar[stage], t.m0 + 64 * j, k0);
            }
            if (!B_MN) {
              xw_tma_load<PAIR>(sb, &tm_b, &full_bar[stage], k0, t.nb0);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j)
                xw_tma_load<PAIR>(sb + j * MN_BLOCK_BYTES, &tm_b, &full_bar[stage], t.nb0 + 64 * j, k0);
            }
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA of a pair) =================
    // convergent warp, one elected lane issues; descriptors = base descriptor + 14-bit address offset
    if (rank == 0) {
      const bool leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      bool ok = true;
      const uint64_t desc_a0 = make_smem_desc(smem_u32(tiles), p.a_lbo, p.a_sbo);
      const uint64_t desc_b0 = make_smem_desc(smem_u32(tiles) + A_TILE_BYTES, p.b_lbo, p.b_sbo);
      const uint32_t a_kstep = p.a_kstep >> 4, b_kstep = p.b_kstep >> 4;
      WorkIter wi; work_begin(p, cluster_id, n_clusters, wi);
      TileCoord t;
      while (ok && work_next<PAIR>(p, cluster_id, n_clusters, rank, wi, t)) {
        ok = mbar_wait(&acc_empty[acc], acc_phase ^ 1);        // epilogue has drained this accumulator
        if (!ok) break;
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        uint32_t accumulate = 0;
        for (int k0 = t.k_begin; k0 < t.k_end && ok; k0 += BLOCK_K) {
          ok = mbar_wait(&full_bar[stage], phase);
          if (!ok) break;
          tc_fence_after_sync();
          if (leader) {
            const uint64_t soff = (uint64_t)((uint32_t)stage * (STG_BYTES >> 4));
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
              xw_mma<PAIR>(d_tmem, desc_a0 + soff + (uint64_t)(kk * a_kstep), desc_b0 + soff + (uint64_t)(kk * b_kstep),
                           p.idesc, accumulate);
              accumulate = 1;
            }
            xw_commit<PAIR>(&empty_bar[stage]);                // smem slot reusable (in both CTAs) once these MMAs retire
          }
          accumulate = 1;
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        if (leader) xw_commit<PAIR>(&acc_full[acc]);           // accumulator ready for the epilogues
        __syncwarp();
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (both CTAs: own 128 rows x 256 columns) =================
    const int quad = warp & 3;                                 // TMEM lanes [32*quad, 32*quad+32)
    const int epi_tid = (warp - EPI_WARP0) * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    bool ok = true;
    WorkIter wi; work_begin(p, cluster_id, n_clusters, wi);
    TileCoord t;
    while (ok && work_next<PAIR>(p, cluster_id, n_clusters, rank, wi, t)) {
      ok = mbar_wait(&acc_full[acc], acc_phase);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after_sync();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BLOCK_N);
      Epi::run(ep, p, t, tmem_acc, quad, lane, epi_tid, scratch);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (PAIR == 2) mbar_arrive_cluster(&acc_empty[acc], 0);
        else mbar_arrive(&acc_empty[acc]);
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (p.early) pdl_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();                           // nobody leaves while the leader may still signal it
  if (warp == 1) { tc_fence_after_sync(); xw_tmem_dealloc<PAIR>(tmem_base, TMEM_COLS); }
}

}  // namespace umma
}  // namespace b200f
