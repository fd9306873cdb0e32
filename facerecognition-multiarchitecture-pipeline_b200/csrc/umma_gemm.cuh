// Persistent, warp-specialised tcgen05 GEMM core (1 CTA per SM, cta_group::1):
//
//   D[m, n] = sum_k A(m, k) * B(n, k)      bf16 (or fp16) operands, fp32 accumulation in TMEM
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
constexpr int STAGES = 4;
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

struct GemmParams {
  int M, N, K;              // extents of the m, n and k index spaces
  int m_tiles, n_tiles, k_splits;
  int k_per_split;          // multiple of BLOCK_K
  // descriptor knobs (bytes); defaults in default_params(); the self-test can override them
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep;
  uint32_t idesc;
};

struct TileCoord { int m0, n0, split, k_begin, k_end; };

__device__ __forceinline__ TileCoord decode_work(const GemmParams& p, int w) {
  TileCoord t;
  const int m = w % p.m_tiles;
  const int n = (w / p.m_tiles) % p.n_tiles;
  t.split = w / (p.m_tiles * p.n_tiles);
  t.m0 = m * BLOCK_M; t.n0 = n * BLOCK_N;
  t.k_begin = t.split * p.k_per_split;
  t.k_end = min(p.K, t.k_begin + p.k_per_split);
  return t;
}

// Epilogue policy interface:
//   struct Epi { struct Params; static __device__ void run(const Params&, const GemmParams&, const TileCoord&,
//                uint32_t tmem_acc /*lane 0, first column of this accumulator stage*/, int quad /*warp%4*/,
//                int lane, int epi_tid /*0..127*/, float* scratch /*EPI_SMEM_FLOATS, epilogue-only smem*/); }
// run() must finish all its tcgen05.ld (tmem_ld_wait) before returning.
template <bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
            const GemmParams p, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles = smem;                                        // STAGES x (A | B), 1024 B aligned
  float* scratch = reinterpret_cast<float*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + EPI_SMEM_FLOATS);
  uint64_t* full_bar = bars;                                    // [STAGES]   TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;                          // [STAGES]   MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;                       // [ACC_STAGES] MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + ACC_STAGES;         // [ACC_STAGES] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_work = p.m_tiles * p.n_tiles * p.k_splits;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_THREADS / 32); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    {
      const bool leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      bool ok = true;
      for (int w = blockIdx.x; w < total_work && ok; w += gridDim.x) {
        const TileCoord t = decode_work(p, w);
        for (int k0 = t.k_begin; k0 < t.k_end && ok; k0 += BLOCK_K) {
          ok = mbar_wait(&empty_bar[stage], phase ^ 1);
          if (!ok) break;
          if (leader) {
            uint8_t* sa = tiles + (size_t)stage * STAGE_BYTES;
            uint8_t* sb = sa + A_TILE_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d(sa, &tm_a, &full_bar[stage], k0, t.m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_2d(sa + j * MN_BLOCK_BYTES, &tm_a, &full_bar[stage], t.m0 + 64 * j, k0);
            }
            if (!B_MN) {
              tma_load_2d(sb, &tm_b, &full_bar[stage], k0, t.n0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_2d(sb + j * MN_BLOCK_BYTES, &tm_b, &full_bar[stage], t.n0 + 64 * j, k0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // convergent warp, one elected lane issues; descriptors = base descriptor + 14-bit address offset
    {
      const bool leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      bool ok = true;
      const uint64_t desc_a0 = make_smem_desc(smem_u32(tiles), p.a_lbo, p.a_sbo);
      const uint64_t desc_b0 = make_smem_desc(smem_u32(tiles) + A_TILE_BYTES, p.b_lbo, p.b_sbo);
      const uint32_t a_kstep = p.a_kstep >> 4, b_kstep = p.b_kstep >> 4;
      for (int w = blockIdx.x; w < total_work && ok; w += gridDim.x) {
        const TileCoord t = decode_work(p, w);
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
            const uint64_t soff = (uint64_t)((uint32_t)stage * (STAGE_BYTES >> 4));
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
              mma_f16_ss(d_tmem, desc_a0 + soff + (uint64_t)(kk * a_kstep), desc_b0 + soff + (uint64_t)(kk * b_kstep),
                         p.idesc, accumulate);
              accumulate = 1;
            }
            mma_commit(&empty_bar[stage]);                     // smem slot reusable once these MMAs retire
          }
          accumulate = 1;
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (leader) mma_commit(&acc_full[acc]);                // accumulator ready for the epilogue
        __syncwarp();
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue =================
    const int quad = warp & 3;                                 // TMEM lanes [32*quad, 32*quad+32)
    const int epi_tid = (warp - EPI_WARP0) * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    bool ok = true;
    for (int w = blockIdx.x; w < total_work && ok; w += gridDim.x) {
      const TileCoord t = decode_work(p, w);
      ok = mbar_wait(&acc_full[acc], acc_phase);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after_sync();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BLOCK_N);
      Epi::run(ep, p, t, tmem_acc, quad, lane, epi_tid, scratch);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) { tc_fence_after_sync(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

}  // namespace umma
}  // namespace b200f
