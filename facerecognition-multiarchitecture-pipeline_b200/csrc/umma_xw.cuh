// X-stationary tcgen05 kernel for the two cosine-logit stages of the head (K2 forward statistics, K3a logit
// gradient):   S[b, c] = sum_k x_hat(b, k) * w_hat(c, k)      fp16 operands, fp32 accumulation in TMEM.
//
// Why not the generic GEMM core (umma_gemm.cuh): with 128 x 256 output tiles every k-block re-reads 48 KB of
// operands from L2 per 512 tensor cycles -- 94 B/clk/SM against an L2 that sustains ~42 B/clk/SM chip-wide --
// so the generic core is L2-bound at ~1/3 of the tensor peak.  Here
//   * a CTA keeps its 128 rows of x_hat (<= 128 KB, D <= 512) resident in shared memory for a whole work
//     item and streams only class-weight tiles through a 5-stage TMA ring, and
//   * (PAIR = 2) two CTAs of a cluster form one tcgen05 cta_group::2 pair: one 256 x 256 x 16 MMA per k-step
//     spans both SMs, each CTA loads only ITS 128 classes of the weight tile, so the per-SM L2 ingest drops
//     to 16 KB per 512 tensor cycles (32 B/clk/SM) and W is read from L2 once per 256 rows of the batch.
//
// Roles per CTA (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA of the pair only),
// warps 2..9 = epilogue (two warps per TMEM lane quadrant; each takes half of the tile's columns; the next
// 32-column tcgen05.ld is in flight while the current one is processed).  Accumulators are double-buffered
// in TMEM (2 x 256 columns).
//
// Work: item = (row group g of 128*PAIR rows, class chunk); a cluster walks items cluster_id, +n_clusters, ..
// The Epilogue policy sees every 32-column slice of the accumulators plus item begin / end hooks, so per-row
// statistics stay in registers across all class tiles of an item (one partial record per row and chunk).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma_core.cuh"

namespace b200f {
namespace umma {

constexpr int XW_M = 128;                       // rows of x per CTA
constexpr int XW_WROWS = 128;                   // class rows per CTA per stage (tile width = 128 * PAIR)
constexpr int XW_K = 64;                        // k-block (one 128 B swizzle row of fp16)
constexpr int XW_MAX_KB = 8;                    // D <= 512 stays resident
constexpr int XW_STAGES = 5;

constexpr int XW_TILE_BYTES = XW_M * XW_K * 2;  // 16 KB: one k-block of x, or one stage of w
constexpr int XW_EPI_WARPS = 8;                   // epilogue warps of ONE group (two per TMEM lane quadrant)
constexpr int XW_THREADS = 64 + 32 * XW_EPI_WARPS;   // 320: producer + MMA issuer + one epilogue group
constexpr int XW_MAX_EPI_GROUPS = 2;
constexpr int XW_SCRATCH_FLOATS = 2048;
constexpr int XW_MAX_ACC = 4;                     // accumulator stages: 2 x 256 columns (CTA pair) or 4 x 128 (single CTA)
constexpr int XW_NUM_BARS = 1 + XW_MAX_KB + 2 * XW_STAGES + 2 * XW_MAX_ACC;   // x_empty, x_full[kb], ring full / empty, accumulators
constexpr size_t XW_SMEM_BYTES = 1024 + (size_t)(XW_MAX_KB + XW_STAGES) * XW_TILE_BYTES + XW_SCRATCH_FLOATS * 4 + 256;
constexpr uint32_t XW_ACC_STRIDE = 256;         // TMEM columns between accumulator stages of a CTA pair (single CTA: 128)

// tcgen05.ld without the wait (software pipelining), and a wait that carries the registers as operands so
// the compiler cannot schedule their uses above it.
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep(float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
        "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
        "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :
      : "memory");
}
// 16-column forms (policies with kSliceCols = 16: half the live accumulator registers per slice)
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep(float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
      :
      : "memory");
}
// one column (32 lanes x 1 fp32), synchronous: a single accumulator element per thread, for rarely taken paths
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(r) : "r"(taddr) : "memory");
  return __uint_as_float(r);
}
// barrier over the 256 threads of ONE epilogue group (ids 1, 2: the groups run at their own pace)
__device__ __forceinline__ void epi_bar_sync(int grp = 0) {
  asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
}
// ... over a group of NT threads (policies with kEpiSplit = 4: 512)
template <int NT> __device__ __forceinline__ void epi_bar_sync_n(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(NT) : "memory");
}

// X_MN / W_MN: layout of the resident ("x") and the streamed ("w") operand: false = K-major (rows x k, row-major),
//   true = MN-major (stored [k, rows] row-major).  The dW stage runs as dW^T[d, c] = sum_b x_hat[b, d] * G^T[c, b]:
//   x_hat (MN-major view of x_hat^T) stays resident, logit-gradient rows stream, and an epilogue thread owns one
//   feature d, so its 32 lanes write 128 contiguous bytes of a dW row.  "B" is then the extent of m (D), "C" of n
//   (classes), "D" of k (the batch).
// SWAP: the streamed operand feeds the A side of the MMA, so accumulator LANES are streamed rows (classes) and COLUMNS
//   are the resident rows (batch).  K3a runs this way: a thread owns one class, so the column sums over the batch that
//   the normalise-backward of W needs (r_j = sum_i G_ij cos_ij) are a private running sum instead of a 31-shuffle
//   butterfly per 32 classes, and G is emitted class-major.
struct XwParams {
  int B, C, D;                          // extents of m (rows of x), n (classes of this launch) and k
  int kb_count;                         // ceil(D / 64) <= XW_MAX_KB
  int m_groups, n_tiles, n_chunks;      // row groups of 128*PAIR rows, class tiles of 128*PAIR, class chunks
  int tn;                               // 128 * PAIR: tile width = rows of a resident group
  int reverse;                          // walk the tiles of a chunk last-to-first: a consumer of what the previous
                                        // kernel just wrote finds the freshest part still in L2
  int prefetch;                         // TILES of the streamed operand pulled into L2 ahead of the ring (0 = off) with ONE
                                        // contiguous cp.async.bulk.prefetch.L2 per CTA and tile (K-major rows are contiguous:
                                        // a CTA's 128 rows x D are one block of memory).  Why: the ring holds 5 of a tile's 8
                                        // k-blocks (80 KB in flight per SM); against HBM latency under load (~2 us) that
                                        // sustains ~40 GB/s per SM where a tensor-bound K2 at cfg3 needs 60.
  const void* w_base;                   // streamed operand: first row of this launch, row pitch in bytes (prefetch only)
  int64_t w_row_bytes;
  int w_hint;                           // L2 policy of the streamed operand's loads (l2_policy(): 0 none, 1 evict_first, 2 evict_last)
  int early_operands;                   // 1: BOTH operands were complete before the predecessor grid started (K3a: x_hat / w_hat
                                        // come from K1, the predecessor only supplies lse / grad4 to the epilogue), so the TMA and
                                        // MMA warps do not wait for the predecessor: loads and MMAs of the first tiles overlap its
                                        // tail.  The epilogue warps wait (griddepcontrol.wait) before they touch anything.
                                        // 2: only the RESIDENT operand is that old (K3b: x_hat from K1; its streamed G^T is the
                                        // predecessor's output): the producer loads the first item's resident rows, THEN waits, then
                                        // streams.  It does not release dependents itself (a kernel in front of an early starter
                                        // must have waited before it triggers: the epilogue warps do both, in that order).
  uint32_t idesc;
  int x_whole;                          // 1: the MMA issuer waits for the whole resident operand before the item's first MMA
  // ---- operand preparation fused into the kernel (policies with kPrepWarps > 0: K2 of the head) ------------------------
  // The streamed operand does not exist yet when the kernel starts: prep warps of ALL CTAs produce it (K1 of the class
  // weights: row L2-normalise -> fp16 rows * out_scale + 1/||w||) in the order the tiles are consumed, and count finished
  // rows per 128-row block in prep_ready[]; a TMA producer loads a tile only after its block is complete.  The 41 us
  // K1(W) pass of the cfg3 step (102 MB read + 102 MB written, HBM-bound, tensor pipe idle) then runs under K2's MMAs,
  // and K2 reads the fresh rows from L2.  prep_ready must be zero when the kernel starts (the predecessor kernel clears it).
  const void* prep_src;                 // raw rows of the streamed operand [C, D], bf16 (prep_f32 = 0) or fp32 (1); null = off
  int prep_f32;
  uint16_t* prep_dst;                   // fp16 rows [C, D] = what tm_w reads
  float* prep_inv;                      // [C] 1 / max(||row||, eps)
  unsigned int* prep_ready;             // [ceil(C / 128)] rows finished per 128-row block
  float prep_eps, prep_scale;
  int prep_cw;                          // chunks whose tiles are consumed at the same time (clusters / row groups)
#ifdef B200F_TIMELINE
  unsigned long long* tl;               // -DB200F_TIMELINE builds (tools/timeline_probe.py): SM clock stamps, see XW_TL
#endif
};

// Timeline stamps (-DB200F_TIMELINE builds only; the shipped library has neither the field nor the stores): the first two
// clusters record clock64() per (CTA, role, tile of the CTA, event) -- roles: 0 MMA issuer {wait for the accumulator stage,
// stage granted, first ring stage of the tile full, last MMA issued}, 1 / 2 first / last epilogue warp {wait for the
// accumulator, accumulator full, slices done}, 3 TMA producer {first slot wait of the tile, last load issued}.
#ifdef B200F_TIMELINE
#define XW_TL(cond, role, tile, k)                                                                                   \
  do {                                                                                                               \
    if ((cond) && p.tl != nullptr && blockIdx.x < 4 && (tile) < 32)                                                   \
      p.tl[((((int)blockIdx.x * 4 + (role)) * 32 + (int)(tile)) << 2) + (k)] = (unsigned long long)clock64();         \
  } while (0)
#else
#define XW_TL(cond, role, tile, k) do {} while (0)
#endif

struct XwItem;
struct XwParams;
// policies may take the epilogue scratch (warp-private staging) in slice(): detected at compile time
template <class Epi, class = void> struct xw_slice_wants_scratch { static constexpr bool value = false; };
template <class Epi> struct xw_slice_wants_scratch<Epi, decltype((void)Epi::kSliceScratch)> { static constexpr bool value = true; };

// A policy may trade ring stages for warp-private staging ("aux") space: with kRingStages = S < XW_STAGES the
// streamed operand gets S stages and every epilogue warp owns (XW_STAGES - S) * 16 KB / 8 bytes at XwItem::aux plus two
// mbarriers (count 1) at XwItem::aux_bar -- e.g. for epilogue operands it fetches with its own TMA loads.
#ifdef B200F_TL_RING     // instrumented builds only (tools/build_timeline.sh): a shorter ring, to tell latency from throughput
template <class Epi, class = void> struct xw_ring_stages { static constexpr int value = B200F_TL_RING; };
#else
template <class Epi, class = void> struct xw_ring_stages { static constexpr int value = XW_STAGES; };
#endif
template <class Epi> struct xw_ring_stages<Epi, decltype((void)Epi::kRingStages)> { static constexpr int value = Epi::kRingStages; };

// A policy may ask for shared memory beyond XW_SMEM_BYTES (kExtraSmem bytes behind the barriers) and lay out its own
// warp-private areas (aux_layout): the dW kernel with TMA stores keeps boxes AND store staging per warp.
template <class Epi, class = void> struct xw_extra_smem { static constexpr int value = 0; };
template <class Epi> struct xw_extra_smem<Epi, decltype((void)Epi::kExtraSmem)> { static constexpr int value = Epi::kExtraSmem; };
template <class Epi> constexpr size_t xw_smem_bytes() { return XW_SMEM_BYTES + (size_t)xw_extra_smem<Epi>::value; }

// Epilogue groups.  A policy with kEpiGroups = 2 runs TWO groups of 8 epilogue warps (576 threads, <= 112 registers):
// group g takes the tiles whose running number n has n % 2 == g, i.e. with two accumulator stages each group owns one
// stage.  Why: with one group every tile costs E (epilogue) + hand-over latency in series with the next-but-one tile's
// MMAs (measured at cfg3: E = 4400 busy cycles, M = 4100 tensor cycles, period 6140); with two groups Epi(t) and
// Epi(t + 1) overlap and the period is max(M, (E + M) / 2).  Per-row state is per group: item_end emits one partial
// record per (row, chunk, group).
template <class Epi, class = void> struct xw_epi_groups { static constexpr int value = 1; };
template <class Epi> struct xw_epi_groups<Epi, decltype((void)Epi::kEpiGroups)> { static constexpr int value = Epi::kEpiGroups; };
// Column split of a tile inside ONE group.  Default 2: eight warps, two per TMEM lane quadrant, each half of the columns.
// kEpiSplit = 4: SIXTEEN warps on the same tile, four per quadrant, each a quarter of the columns.  Against two groups on
// alternating tiles (same warp count): a stage is held for M + hand-over + E, and there are only two stages, so the period
// per tile is (M + h + E) / 2 either way -- but with all warps on one tile E halves (the epilogue is latency-bound, not
// issue-bound: 43 % issue slots used), while with alternating groups it does not.
template <class Epi, class = void> struct xw_epi_split { static constexpr int value = 2; };
template <class Epi> struct xw_epi_split<Epi, decltype((void)Epi::kEpiSplit)> { static constexpr int value = Epi::kEpiSplit; };
template <class Epi> __host__ __device__ constexpr int xw_group_warps() { return 4 * xw_epi_split<Epi>::value; }
// columns per slice() call: 32 (default) or 16 (kSliceCols = 16)
template <class Epi, class = void> struct xw_slice_cols { static constexpr int value = 32; };
template <class Epi> struct xw_slice_cols<Epi, decltype((void)Epi::kSliceCols)> { static constexpr int value = Epi::kSliceCols; };
// warps that prepare the streamed operand inside the kernel (XwParams::prep_*): 0 (default) or kPrepWarps
template <class Epi, class = void> struct xw_prep_warps { static constexpr int value = 0; };
template <class Epi> struct xw_prep_warps<Epi, decltype((void)Epi::kPrepWarps)> { static constexpr int value = Epi::kPrepWarps; };
template <class Epi> __host__ __device__ constexpr int xw_threads() {
  return 64 + 32 * xw_group_warps<Epi>() * xw_epi_groups<Epi>::value + 32 * xw_prep_warps<Epi>::value;
}
// registers per thread: 1 CTA per SM either way (shared memory).  The register file is handed out per FOUR warps:
// 10 warps count as 12 (168 registers), 18 warps as 20 (96 registers; 112 is refused at launch: "too many resources").
// 12 warps (10 + 2 spare): 168; up to 16 warps: 128; 20 (two groups, or one group of 16-column slices + 10 prep warps): 96.
template <class Epi> __host__ __device__ constexpr int xw_maxnreg() {
  return xw_threads<Epi>() <= 384 ? 168 : (xw_threads<Epi>() <= 512 ? 128 : 96);
}


// ---- operand preparation inside the kernel (XwParams::prep_*) ---------------------------------------------------------
// Row L2-normalise of 512-wide rows: the arithmetic of rowops::l2norm_rows_512x16_body (same per-lane element order, same
// reduction), so the rows are bit-identical to the stand-alone K1's for bf16 sources.  A warp keeps ROWS rows (ROWS x 1 KB
// of bf16, ROWS x 2 KB of fp32) in flight; rows are visited in the order the TMA producers consume them: rounds of "the
// k-th tile of every chunk that is being worked on", row blocks of a round spread over all prep warps of the grid.
__device__ __forceinline__ bool xw_wait_ready(const unsigned int* flag, unsigned int need) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= need) return true;
    __nanosleep(40);
  }
  atomicExch(&g_umma_timeout_flag, 1u);
  __threadfence_system();
  __trap();
  return false;
}

// Publishing finished rows: the count of a 128-row block is raised with a release at gpu scope (fence.acq_rel + red; not
// __threadfence(), which is fence.sc).  The release waits for the warp's outstanding stores (1.5-2.5 us under this kernel's
// memory traffic), so it is paid once per RUN of rows in the same block -- a warp's rows of one round -- not once per trip:
// with one release per 2-3 rows the ten prep warps of an SM spent most of their time in it (K2 115-123 us).
struct XwPrepPending { int blk; unsigned int cnt; };
__device__ __forceinline__ void xw_prep_flush(const XwParams& p, XwPrepPending& pd, int lane) {
  if (pd.cnt == 0) return;                                      // warp-uniform
  __syncwarp();
  if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.prep_ready + pd.blk), "r"(pd.cnt) : "memory");
  pd.cnt = 0;
}
template <int ROWS>
__device__ __forceinline__ void xw_prep_note(const XwParams& p, XwPrepPending& pd, const int (&row)[ROWS], int lane) {
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    if (row[j] < 0) continue;                                   // warp-uniform
    const int b = row[j] >> 7;
    if (b != pd.blk) { xw_prep_flush(p, pd, lane); pd.blk = b; }
    ++pd.cnt;
  }
}

// Lane -> element mapping (it fixes the order of the sum of squares, hence the bits of 1/||w||):
//   bf16 source: lane owns elements [16 lane, 16 lane + 16): one 32-byte load and one 32-byte store per row, as
//                rowops::l2norm_rows_512x16_body;
//   fp32 source: lane owns the float4 vectors lane, lane + 32, lane + 64, lane + 96, as rowops::l2norm_rows_vec_kernel
//                <float, __half, 4> (and K5, which emits the same operands after an optimizer step).
// A warp walks ITS rows (rpw per round, all rounds back to back) ROWS at a time through two register buffers: the loads
// of trip t + 1 are in flight while trip t is reduced and stored, and the rows of trip t - 1 are published (fence +
// counter) right after those loads were issued.  The arithmetic of a trip is written phase by phase over its rows
// (squares, butterfly, inverse norms, stores) so that the rows' dependent chains interleave: six warps per SM have to
// turn 680 rows around in ~50 us.
template <int ROWS, bool F32>
struct XwPrepBuf {
  static constexpr int RW = F32 ? 16 : 8;                      // 32-bit words a lane holds per row (16 elements)
  int row[ROWS];                                               // C < 2^30 (check_shape)
  uint32_t raw[ROWS][RW];
};

// The warp's position in its row sequence: entry (k, slot) = slot-th of its rpw rows of round k.  Lane l holds what is
// fixed per slot l for the current wave group -- r0 = first-round row, cnt = tiles of that slot's chunk -- so a row is
// r0 + k * TN while k < cnt: no division in the loop.
struct XwPrepPos { int k, slot; };

template <int TN, int ROWS, bool F32>
__device__ __forceinline__ void xw_prep_load(const XwParams& p, XwPrepBuf<ROWS, F32>& b, XwPrepPos& pos, int rpw, int maxT,
                                             int r0_mine, int cnt_mine, int lane) {
  constexpr int RW = XwPrepBuf<ROWS, F32>::RW;
  // (Measured and dropped: bulk L2 prefetches of the rows of the next three rounds, one per slot -- K2 104.7 us against
  //  98.4 without: the prep warps are not waiting for their loads.)
  int k = pos.k, s = pos.slot + lane;                           // lane j < ROWS resolves the j-th entry from pos
  while (s >= rpw) { s -= rpw; ++k; }
  const int r0 = __shfl_sync(0xffffffffu, r0_mine, s & 31);
  const int cnt = __shfl_sync(0xffffffffu, cnt_mine, s & 31);
  int mine = -1;
  if (lane < ROWS && k < maxT && k < cnt) { const int64_t r = (int64_t)r0 + (int64_t)k * TN; if (r < p.C) mine = (int)r; }
  pos.slot += ROWS;
  while (pos.slot >= rpw) { pos.slot -= rpw; ++pos.k; }
  const char* src = static_cast<const char*>(p.prep_src);
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    b.row[j] = __shfl_sync(0xffffffffu, mine, j);
    if (b.row[j] >= 0) {
      if (F32) {
        const char* a = src + (int64_t)b.row[j] * 2048 + lane * 16;
#pragma unroll
        for (int it = 0; it < 4; ++it)
          asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(b.raw[j][4 * it]), "=r"(b.raw[j][4 * it + 1]), "=r"(b.raw[j][4 * it + 2]), "=r"(b.raw[j][4 * it + 3])
                       : "l"(a + 512 * it));
      } else {
        const char* a = src + (int64_t)b.row[j] * 1024 + lane * 32;
        asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(b.raw[j][0]), "=r"(b.raw[j][1]), "=r"(b.raw[j][2]), "=r"(b.raw[j][3]),
                       "=r"(b.raw[j][4]), "=r"(b.raw[j][5]), "=r"(b.raw[j][6]), "=r"(b.raw[j][7])
                     : "l"(a));
      }
    } else {
#pragma unroll
      for (int i = 0; i < RW; ++i) b.raw[j][i] = 0u;
    }
  }
}

// packed fp32x2 arithmetic (sm_100): one instruction per element PAIR
__device__ __forceinline__ uint64_t f32x2_pack(float lo, float hi) {
  uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ uint64_t f32x2_mul(uint64_t a, uint64_t b) {
  uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
// element pair i (elements 2i, 2i + 1) of row j as a packed fp32x2
template <int ROWS, bool F32>
__device__ __forceinline__ uint64_t xw_prep_pair(const XwPrepBuf<ROWS, F32>& b, int j, int i) {
  if (F32) return f32x2_pack(__uint_as_float(b.raw[j][2 * i]), __uint_as_float(b.raw[j][2 * i + 1]));
  return f32x2_pack(__uint_as_float(b.raw[j][i] << 16), __uint_as_float(b.raw[j][i] & 0xffff0000u));
}

// The sum of squares runs as two chains per lane (even and odd elements, packed FFMA2), added at the end: not the
// stand-alone K1's single chain, so 1/||w|| can differ from its value in the last bit (and with it, rarely, an operand's).
template <int ROWS, bool F32>
__device__ __forceinline__ void xw_prep_finish(const XwParams& p, const XwPrepBuf<ROWS, F32>& b, int lane) {
  uint64_t ss2[ROWS];
#pragma unroll
  for (int j = 0; j < ROWS; ++j) ss2[j] = 0ull;
#pragma unroll
  for (int i = 0; i < 8; ++i)                                   // pair-major: ROWS independent chains
#pragma unroll
    for (int j = 0; j < ROWS; ++j) { const uint64_t v = xw_prep_pair<ROWS, F32>(b, j, i); ss2[j] = f32x2_fma(v, v, ss2[j]); }
  float ss[ROWS];
#pragma unroll
  for (int j = 0; j < ROWS; ++j) { float lo, hi; f32x2_unpack(ss2[j], lo, hi); ss[j] = lo + hi; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)                        // warp_sum's butterfly, the rows side by side
#pragma unroll
    for (int j = 0; j < ROWS; ++j) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], off);
  // the square root and the division once per ROW (lane j takes row j), not once per row and lane
  float m = ss[0];
#pragma unroll
  for (int j = 1; j < ROWS; ++j) if (lane == j) m = ss[j];
  const float inv_m = 1.0f / fmaxf(sqrtf(m), p.prep_eps);
  if (lane < ROWS) {
    int r = b.row[0];
#pragma unroll
    for (int j = 1; j < ROWS; ++j) if (lane == j) r = b.row[j];
    if (r >= 0) p.prep_inv[r] = inv_m;
  }
#pragma unroll
  for (int j = 0; j < ROWS; ++j) {
    const float sc = __shfl_sync(0xffffffffu, inv_m, j) * p.prep_scale;
    if (b.row[j] < 0) continue;                                 // warp-uniform
    const uint64_t sc2 = f32x2_pack(sc, sc);
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float lo, hi;
      f32x2_unpack(f32x2_mul(xw_prep_pair<ROWS, F32>(b, j, i), sc2), lo, hi);
      const __half2 h = __floats2half2_rn(lo, hi);
      o[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    uint16_t* dst = p.prep_dst + (int64_t)b.row[j] * 512;
    if (F32) {
#pragma unroll
      for (int it = 0; it < 4; ++it)
        asm volatile("st.global.v2.b32 [%0], {%1, %2};" ::"l"(dst + (lane + 32 * it) * 4), "r"(o[2 * it]), "r"(o[2 * it + 1]) : "memory");
    } else {
      st_global_256(dst + lane * 16, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
    }
  }
}

// rpw (rows of a round per prep warp) <= 32 so that a lane can hold a slot: the host enables the fused path only then.
template <int TN, int ROWS, bool F32>
__device__ __forceinline__ void xw_prep_run(const XwParams& p, int gw, int nW) {
  const int lane = threadIdx.x & 31;
  const int cw = p.prep_cw;
  const int n_groups = (p.n_chunks + cw - 1) / cw;
  const int maxT = (p.n_tiles + p.n_chunks - 1) / p.n_chunks;  // a chunk has floor or ceil(n_tiles / n_chunks) tiles
  const int round_rows = cw * TN;
  const int rpw = (round_rows + nW - 1) / nW;
  const int q_lo = gw * rpw;
  const int q_hi = (q_lo + rpw < round_rows) ? q_lo + rpw : round_rows;
  if (q_lo >= round_rows) return;
  XwPrepBuf<ROWS, F32> a, b;
  XwPrepPending pd{-1, 0u};
#pragma unroll 1
  for (int W = 0; W < n_groups; ++W) {
    int r0_mine = 0, cnt_mine = 0;                              // slot `lane` of this warp in wave group W
    if (lane < rpw && q_lo + lane < q_hi) {
      const int q = q_lo + lane;
      const int c = W * cw + q / TN;
      if (c < p.n_chunks) {
        const int tb = (int)((int64_t)c * p.n_tiles / p.n_chunks);
        const int te = (int)((int64_t)(c + 1) * p.n_tiles / p.n_chunks);
        r0_mine = tb * TN + (q % TN);
        cnt_mine = te - tb;
      }
    }
    XwPrepPos pos{0, 0};
    const int n_total = maxT * rpw;
    xw_prep_load<TN, ROWS, F32>(p, a, pos, rpw, maxT, r0_mine, cnt_mine, lane);
#pragma unroll 1
    for (int n0 = 0; n0 < n_total; n0 += 2 * ROWS) {
      xw_prep_load<TN, ROWS, F32>(p, b, pos, rpw, maxT, r0_mine, cnt_mine, lane);
      xw_prep_finish<ROWS, F32>(p, a, lane);
      xw_prep_note<ROWS>(p, pd, a.row, lane);
      xw_prep_load<TN, ROWS, F32>(p, a, pos, rpw, maxT, r0_mine, cnt_mine, lane);
      xw_prep_finish<ROWS, F32>(p, b, lane);
      xw_prep_note<ROWS>(p, pd, b.row, lane);
    }
  }
  xw_prep_flush(p, pd, lane);
}

struct XwItem {                         // what an epilogue thread knows about its work item
  int item, chunk, group;
  int rank, ew, quad, half, lane;       // CTA rank in the pair, epilogue warp 0..7 of its group, TMEM quadrant, column half, lane
  int grp;                              // epilogue group 0..kEpiGroups-1
  int first_tile;                       // position (in the item's walk order) of the first tile this group takes
  int64_t row;                          // global row of x owned by this thread
  uint32_t taddr0;                      // TMEM address of this thread's lane, column 0 of the current accumulator stage (SWAP mode)
  uint8_t* aux;                         // this warp's staging bytes (nullptr unless the policy reserves them)
  uint8_t* stage;                       // this warp's store staging (policies with kExtraSmem: set by their aux_layout)
  uint64_t* aux_bar;                    // its two mbarriers
  mutable uint32_t aux_phase;           // their parity bits; lives across the items of the kernel
};

template <class Epi, class State, class Params, int SC>
__device__ __forceinline__ void xw_call_slice(State& st, const Params& ep, const XwParams& p, const XwItem& it,
                                              float (&v)[SC], int cls0, float* scratch) {
  if constexpr (xw_slice_wants_scratch<Epi>::value) Epi::slice(st, ep, p, it, v, cls0, scratch);
  else Epi::slice(st, ep, p, it, v, cls0);
}

// Epilogue policy interface:
//   struct Epi { struct Params; struct State;
//     static __device__ void item_begin(State&, const Params&, const XwParams&, const XwItem&);
//     static __device__ void tile_begin(State&, const Params&, const XwParams&, const XwItem&, int cls0, int next_cls0, int ncols);
//         once per class tile before its slices: this warp owns classes [cls0, cls0 + ncols) now and
//         [next_cls0, next_cls0 + ncols) in its next tile (-1: none) -- the place for L2 prefetches
//     static __device__ void slice(State&, const Params&, const XwParams&, const XwItem&, float (&v)[32], int cls0);
//         v = accumulators of row it.row for classes [cls0, cls0 + 32) of this launch (warp-uniform cls0 < C)
//     static __device__ void item_end(State&, const Params&, const XwParams&, const XwItem&, float* scratch); }
template <int PAIR, bool X_MN, bool W_MN, bool SWAP, class Epi>
__global__ void __maxnreg__(xw_maxnreg<Epi>())
xw_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const XwParams p,
          const __grid_constant__ typename Epi::Params ep) {
  constexpr int TN = XW_WROWS * PAIR;                        // class-tile width of the cluster
  constexpr int STAGES = xw_ring_stages<Epi>::value;         // ring stages of the streamed operand
  static_assert(STAGES >= 2 && STAGES <= XW_STAGES, "ring stages");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xres = smem;                                      // XW_MAX_KB x 16 KB
  uint8_t* ring = smem + (size_t)XW_MAX_KB * XW_TILE_BYTES;  // XW_STAGES x 16 KB
  float* scratch = reinterpret_cast<float*>(ring + (size_t)XW_STAGES * XW_TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + XW_SCRATCH_FLOATS);
  // one barrier per k-block of the resident operand: the first tile's MMAs of k-block kb start when ITS 16 KB have landed,
  // not when all 128 KB have (148 CTAs pull 19 MB of x_hat out of L2 at the start of every kernel: ~3 us)
  uint64_t* x_empty = bars;                                  // MMA -> TMA (both CTAs)
  uint64_t* x_full = bars + 1;                               // [XW_MAX_KB] TMA -> MMA (leader)
  uint64_t* full_bar = bars + 1 + XW_MAX_KB;                 // [STAGES] TMA -> MMA (leader)
  uint64_t* empty_bar = full_bar + XW_STAGES;                // [STAGES] MMA -> TMA (both CTAs)
  // A single CTA's tile is 128 columns wide: its 512 TMEM columns hold FOUR accumulator stages, so the MMA issuer can run
  // two tiles ahead of the slowest epilogue warp (with two stages every hand-off latency sits on the critical path).
  constexpr int ACC = (PAIR == 1) ? 4 : 2;
  constexpr uint32_t ACC_STRIDE = 512 / ACC;
  uint64_t* acc_full = empty_bar + XW_STAGES;                // [ACC] MMA -> epilogue (both CTAs)
  uint64_t* acc_empty = acc_full + XW_MAX_ACC;               // [ACC] epilogue (both CTAs) -> MMA (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + XW_NUM_BARS);
  uint8_t* aux = ring + (size_t)STAGES * XW_TILE_BYTES;       // (XW_STAGES - STAGES) x 16 KB of warp-private staging
  constexpr int SPLIT = xw_epi_split<Epi>::value;             // column split of a tile inside a group
  constexpr int GW = xw_group_warps<Epi>();                  // epilogue warps of one group
  static_assert(SPLIT == 2 || SPLIT == 4, "column split");
  constexpr int EPI_WARPS_ALL = GW * xw_epi_groups<Epi>::value;
  constexpr int EXTRA = xw_extra_smem<Epi>::value;
  uint8_t* extra = reinterpret_cast<uint8_t*>(bars) + 256;    // EXTRA bytes behind the barrier block
  // 2 mbarriers per epilogue warp: at the end of the scratch area, or (policies that use all of it) at the end of `extra`
  uint64_t* aux_bar = (EXTRA > 0) ? reinterpret_cast<uint64_t*>(extra + EXTRA - 256)
                                  : reinterpret_cast<uint64_t*>(scratch + XW_SCRATCH_FLOATS) - 2 * EPI_WARPS_ALL;
  constexpr int AUX_WARP_BYTES = (XW_STAGES - STAGES) * XW_TILE_BYTES / EPI_WARPS_ALL;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (PAIR == 2) ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / PAIR;
  const int n_clusters = gridDim.x / PAIR;
  const int items = p.m_groups * p.n_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int kb = 0; kb < XW_MAX_KB; ++kb) mbar_init(&x_full[kb], PAIR);
    mbar_init(x_empty, 1);
    for (int s = 0; s < XW_STAGES; ++s) { mbar_init(&full_bar[s], PAIR); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], PAIR * GW); }
    if (STAGES < XW_STAGES) for (int s = 0; s < 2 * EPI_WARPS_ALL; ++s) mbar_init(&aux_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) xw_tmem_alloc<PAIR>(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();                         // the peer's barriers exist before anyone signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // prologue done; predecessor grids complete from here on -- for everyone, or (early_operands) for the epilogue warps only:
  // every CTA has epilogue warps, so no CTA (hence not the grid) completes before its predecessor has.
  const bool waits_now = !p.early_operands || warp >= 2;
  if (waits_now) pdl_wait();
  // Dependents may be scheduled from here on -- AFTER the wait, not at the top of the kernel: a dependent that starts
  // "early" (above) relies on everything older than its predecessor being complete, and that holds only if the
  // predecessor could not release it before having waited itself (K1 -> K2 -> statistics -> K3a: K3a reads K1's output).
  if (waits_now || p.early_operands == 1) pdl_trigger();

  constexpr int PREP_WARPS = xw_prep_warps<Epi>::value;
  if (PREP_WARPS > 0 && warp >= 2 + EPI_WARPS_ALL) {
    // ================= operand preparation (K1 of the streamed rows, all CTAs) =================
    if (p.prep_src != nullptr) {
      const int gw = (int)blockIdx.x * PREP_WARPS + (warp - 2 - EPI_WARPS_ALL);
      const int nW = (int)gridDim.x * PREP_WARPS;
      if (p.prep_f32) xw_prep_run<TN, 1, true>(p, gw, nW);
      else xw_prep_run<TN, 2, false>(p, gw, nW);
    }
  } else if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    // The whole warp walks the loops convergently (loop state stays in uniform registers); one elected lane
    // issues the TMA traffic.
    {
      const bool leader = elect_one();
      const uint64_t w_pol = l2_policy(p.w_hint);
      int stage = 0; uint32_t phase = 0;
      int item_no = 0;
      bool ok = true;
      [[maybe_unused]] int tl_tile = 0;
      for (int item = cluster_id; item < items && ok; item += n_clusters, ++item_no) {
        const int g = item % p.m_groups, chunk = item / p.m_groups;
        const int t_begin = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
        const int t_end = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
        if (item_no > 0) { ok = mbar_wait(x_empty, (uint32_t)((item_no - 1) & 1)); if (!ok) break; }
        const int row0 = (g * PAIR + rank) * XW_M;
        if (leader) {
          for (int kb = 0; kb < p.kb_count; ++kb) {
            if (rank == 0) mbar_arrive_expect_tx(&x_full[kb], (uint32_t)(PAIR * XW_TILE_BYTES));
            else mbar_arrive_cluster(&x_full[kb], 0);
            uint8_t* dst = xres + (size_t)kb * XW_TILE_BYTES;
            if (!X_MN) {
              xw_tma_load<PAIR>(dst, &tm_x, &x_full[kb], kb * XW_K, row0);
            } else {                                              // two 64-wide row blocks of 64 k-rows each
              xw_tma_load<PAIR>(dst, &tm_x, &x_full[kb], row0, kb * XW_K);
              xw_tma_load<PAIR>(dst + XW_TILE_BYTES / 2, &tm_x, &x_full[kb], row0 + 64, kb * XW_K);
            }
          }
        }
        if (p.early_operands == 2 && item_no == 0) pdl_wait();      // the streamed operand is the predecessor's output
        // L2 prefetch of whole tiles ahead of the ring (K-major streamed operand only)
        auto prefetch_tile = [&](int tj) {                      // tj: position in the walk order of this item
          if (W_MN || p.prefetch <= 0 || tj >= t_end) return;
          const int tt = p.reverse ? (t_end - 1 - (tj - t_begin)) : tj;
          const int64_t r0 = (int64_t)tt * TN + rank * XW_WROWS;
          int64_t nrows = (int64_t)p.C - r0; if (nrows > XW_WROWS) nrows = XW_WROWS;
          if (nrows <= 0) return;
          const char* a = static_cast<const char*>(p.w_base) + r0 * p.w_row_bytes;
          const uint32_t nbytes = (uint32_t)(nrows * p.w_row_bytes) & ~15u;
          if (nbytes >= 16 && (reinterpret_cast<uintptr_t>(a) & 15) == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(nbytes) : "memory");
        };
        if (leader) for (int i = 1; i < p.prefetch; ++i) prefetch_tile(t_begin + i);
        for (int ti = t_begin; ti < t_end && ok; ++ti) {
          const int t = p.reverse ? (t_end - 1 - (ti - t_begin)) : ti;
          const int n0 = t * TN + rank * XW_WROWS;
          if (PREP_WARPS > 0 && p.prep_src != nullptr && n0 < p.C) {
            // this CTA's 128 rows of the tile are being written by prep warps somewhere on the chip: wait for the block's
            // row count, then order the (generic-proxy) rows before this thread's async-proxy reads of them
            const int left = p.C - n0;
            xw_wait_ready(p.prep_ready + (n0 >> 7), (unsigned int)(left < XW_WROWS ? left : XW_WROWS));
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          XW_TL(leader, 3, tl_tile, 0);
          for (int kb = 0; kb < p.kb_count; ++kb) {
            XW_TL(leader && tl_tile == 3, 3, 16 + kb, 0);
            ok = mbar_wait(&empty_bar[stage], phase ^ 1);
            if (!ok) break;
            XW_TL(leader && tl_tile == 3, 3, 16 + kb, 1);
            if (leader) {
              if (kb == 0) prefetch_tile(ti + p.prefetch);
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(PAIR * XW_TILE_BYTES));
              else mbar_arrive_cluster(&full_bar[stage], 0);
              uint8_t* dst = ring + (size_t)stage * XW_TILE_BYTES;
              if (!W_MN) {
                if (p.w_hint) xw_tma_load_hint<PAIR>(dst, &tm_w, &full_bar[stage], kb * XW_K, n0, w_pol);
                else xw_tma_load<PAIR>(dst, &tm_w, &full_bar[stage], kb * XW_K, n0);
              } else {
                xw_tma_load<PAIR>(dst, &tm_w, &full_bar[stage], n0, kb * XW_K);
                xw_tma_load<PAIR>(dst + XW_TILE_BYTES / 2, &tm_w, &full_bar[stage], n0 + 64, kb * XW_K);
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          XW_TL(leader, 3, tl_tile, 1);
          ++tl_tile;
        }
      }
      // do not leave while the leader's last commit may still signal this CTA's barriers
      if (ok && item_no > 0) mbar_wait(x_empty, (uint32_t)((item_no - 1) & 1));
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA of the pair) =================
    // Convergent warp, one elected lane issues: stage / k-block counters live in uniform registers and a
    // descriptor is the base descriptor plus a 14-bit address offset, so the issue loop stays far shorter than
    // the 128 tensor cycles each MMA takes (the first version rebuilt both descriptors per MMA from a divergent
    // lane and was issue-bound: tensor pipe 49 % active with every queue full).
    if (rank == 0) {
      const bool leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int item_no = 0;
      bool ok = true;
      [[maybe_unused]] int tl_tile = 0;
      // K-major: +32 B per 16 k; MN-major: +16 k-rows of 128 B, LBO = next 64-wide row block (8 KB)
      const uint32_t x_kstep = X_MN ? 2048u >> 4 : 32u >> 4;     // descriptor address units (16 B)
      const uint32_t w_kstep = W_MN ? 2048u >> 4 : 32u >> 4;
      const uint64_t desc_x0 = make_smem_desc(smem_u32(xres), X_MN ? (uint32_t)(XW_TILE_BYTES / 2) : 0u, 1024);
      const uint64_t desc_w0 = make_smem_desc(smem_u32(ring), W_MN ? (uint32_t)(XW_TILE_BYTES / 2) : 0u, 1024);
      for (int item = cluster_id; item < items && ok; item += n_clusters, ++item_no) {
        const int chunk = item / p.m_groups;
        const int t_begin = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
        const int t_end = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
        for (int t = t_begin; t < t_end && ok; ++t) {
          XW_TL(leader, 0, tl_tile, 0);
          ok = mbar_wait(&acc_empty[acc], acc_phase ^ 1);
          if (!ok) break;
          tc_fence_after_sync();
          XW_TL(leader, 0, tl_tile, 1);
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_STRIDE;
          for (int kb = 0; kb < p.kb_count; ++kb) {
            if (t == t_begin) {                                   // the item's resident operand, k-block by k-block
              if (p.x_whole) {                                    // (tunable "x_whole" = 1: all of it before the first MMA, as round 1 did)
                if (kb == 0) for (int k2 = 0; k2 < p.kb_count; ++k2) mbar_wait(&x_full[k2], (uint32_t)(item_no & 1));
              } else {
                ok = mbar_wait(&x_full[kb], (uint32_t)(item_no & 1));
                if (!ok) break;
              }
            }
            ok = mbar_wait(&full_bar[stage], phase);
            if (!ok) break;
            tc_fence_after_sync();
            XW_TL(leader && kb == 0, 0, tl_tile, 2);
            XW_TL(leader && tl_tile == 3, 0, 16 + kb, 0);
            if (leader) {
              const uint64_t dx = desc_x0 + (uint64_t)((uint32_t)kb * (XW_TILE_BYTES >> 4));
              const uint64_t dw = desc_w0 + (uint64_t)((uint32_t)stage * (XW_TILE_BYTES >> 4));
#pragma unroll
              for (int kk = 0; kk < XW_K / 16; ++kk) {
                // SWAP: the streamed operand is A (accumulator lanes = streamed rows), the resident one is B
                if (!SWAP) xw_mma<PAIR>(d_tmem, dx + (uint64_t)(kk * x_kstep), dw + (uint64_t)(kk * w_kstep), p.idesc,
                                        (uint32_t)((kb | kk) != 0));
                else xw_mma<PAIR>(d_tmem, dw + (uint64_t)(kk * w_kstep), dx + (uint64_t)(kk * x_kstep), p.idesc,
                                  (uint32_t)((kb | kk) != 0));
              }
              xw_commit<PAIR>(&empty_bar[stage]);
            }
            XW_TL(leader && tl_tile == 3, 0, 16 + kb, 1);
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (!ok) break;
          XW_TL(leader, 0, tl_tile, 3);
          ++tl_tile;
          if (leader) xw_commit<PAIR>(&acc_full[acc]);
          __syncwarp();
          if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
        }
        if (ok && leader) xw_commit<PAIR>(x_empty);
        __syncwarp();
      }
    }
  } else if (warp < 2 + EPI_WARPS_ALL) {
    // ================= epilogue (8 warps per group, both CTAs) =================
    constexpr int EG = xw_epi_groups<Epi>::value;
    constexpr int SC = xw_slice_cols<Epi>::value;
    static_assert(EG >= 1 && EG <= XW_MAX_EPI_GROUPS && (SC == 16 || SC == 32), "epilogue geometry");
    XwItem it;
    it.rank = rank; it.grp = (warp - 2) / GW; it.ew = (warp - 2) % GW;
    it.quad = warp & 3; it.half = it.ew >> 2; it.lane = lane;
    it.aux = (STAGES < XW_STAGES) ? aux + (size_t)(it.grp * GW + it.ew) * AUX_WARP_BYTES : nullptr;
    it.stage = nullptr;
    if constexpr (EXTRA > 0) Epi::aux_layout(it, ring, reinterpret_cast<uint8_t*>(scratch), extra);
    it.aux_bar = aux_bar + 2 * (it.grp * GW + it.ew); it.aux_phase = 0; it.first_tile = 0;
    float* const gscratch = scratch + it.grp * (XW_SCRATCH_FLOATS / XW_MAX_EPI_GROUPS);   // this group's half of the scratch
    uint32_t tile_no = 0;                                      // running tile number of this CTA (all groups count alike)
    bool ok = true;
    constexpr int SLICES = TN / SPLIT / SC;                    // slices per warp per tile
    static_assert(SLICES >= 2 && SLICES % 2 == 0, "two slices per trip");
    if constexpr (SWAP) {
      // Accumulator lanes = streamed rows (classes), columns = the resident operand's rows (the batch rows of the
      // group): the thread owns ONE class per tile and sees half of the group's batch rows.
      for (int item = cluster_id; item < items && ok; item += n_clusters) {
        it.item = item; it.group = item % p.m_groups; it.chunk = item / p.m_groups;
        const int t_begin = (int)((int64_t)it.chunk * p.n_tiles / p.n_chunks);
        const int t_end = (int)((int64_t)(it.chunk + 1) * p.n_tiles / p.n_chunks);
        typename Epi::State stt;
        it.first_tile = (EG > 1) ? (int)((it.grp + EG - (tile_no % EG)) % EG) : 0;
        Epi::item_begin(stt, ep, p, it, gscratch, TN);         // per-column tables of the group -> shared memory
        const int col_base = it.half * (TN / SPLIT);
        for (int ti = t_begin; ti < t_end; ++ti) {
          const uint32_t n = tile_no++;
          if (EG > 1 && (int)(n % EG) != it.grp) continue;     // the other group's tile
          const int acc = (int)(n % ACC); const uint32_t acc_phase = (n / ACC) & 1u;
          const int t = p.reverse ? (t_end - 1 - (ti - t_begin)) : ti;
          XW_TL(lane == 0 && (it.ew == 0 || it.ew == GW - 1), it.ew == 0 ? 1 : 2, n, 0);
          ok = mbar_wait(&acc_full[acc], acc_phase);
          ok = __all_sync(0xffffffffu, ok);
          if (!ok) break;
          tc_fence_after_sync();
          XW_TL(lane == 0 && (it.ew == 0 || it.ew == GW - 1), it.ew == 0 ? 1 : 2, n, 1);
          it.row = (int64_t)t * TN + rank * XW_WROWS + it.quad * 32 + lane;   // the class this thread owns in tile t
          const uint32_t taddr = tmem_base + (uint32_t)acc * ACC_STRIDE + ((uint32_t)(it.quad * 32) << 16) + (uint32_t)col_base;
          it.taddr0 = taddr - (uint32_t)col_base;
          Epi::tile_begin(stt, ep, p, it);
          float va[SC], vb[SC];
          tmem_ld32_async(taddr, va);
#pragma unroll 1
          for (int s = 0; s < SLICES; s += 2) {
            // __syncwarp() behind each prefetch: ptxas otherwise sinks the tcgen05.ld to the END of the slice's first basic
            // block (behind all of its math: it gives the next slice's accumulators the registers of this slice's
            // temporaries), and the load's latency then sits at the top of the next slice
            tmem_ld_wait_dep(va);
            tmem_ld32_async(taddr + (uint32_t)(s + 1) * SC, vb);
            __syncwarp();
            Epi::slice(stt, ep, p, it, va, col_base + s * SC, gscratch);
            tmem_ld_wait_dep(vb);
            if (s + 2 < SLICES) tmem_ld32_async(taddr + (uint32_t)(s + 2) * SC, va);
            __syncwarp();
            Epi::slice(stt, ep, p, it, vb, col_base + (s + 1) * SC, gscratch);
          }
          XW_TL(lane == 0 && (it.ew == 0 || it.ew == GW - 1), it.ew == 0 ? 1 : 2, n, 2);
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) {
            if (PAIR == 2) mbar_arrive_cluster(&acc_empty[acc], 0);
            else mbar_arrive(&acc_empty[acc]);
          }
          Epi::tile_end(stt, ep, p, it);
        }
        if (!ok) break;
        Epi::item_end_swap(stt, ep, p, it, PAIR);
      }
    } else
    for (int item = cluster_id; item < items && ok; item += n_clusters) {
      it.item = item; it.group = item % p.m_groups; it.chunk = item / p.m_groups;
      it.row = (int64_t)(it.group * PAIR + rank) * XW_M + it.quad * 32 + lane;
      const int t_begin = (int)((int64_t)it.chunk * p.n_tiles / p.n_chunks);
      const int t_end = (int)((int64_t)(it.chunk + 1) * p.n_tiles / p.n_chunks);
      typename Epi::State stt;
      Epi::item_begin(stt, ep, p, it);
      for (int ti = t_begin; ti < t_end; ++ti) {
        const uint32_t n = tile_no++;
        if (EG > 1 && (int)(n % EG) != it.grp) continue;       // the other group's tile
        const int acc = (int)(n % ACC); const uint32_t acc_phase = (n / ACC) & 1u;
        const int t = p.reverse ? (t_end - 1 - (ti - t_begin)) : ti;
        XW_TL(lane == 0 && (it.ew == 0 || it.ew == GW - 1), it.ew == 0 ? 1 : 2, n, 0);
        ok = mbar_wait(&acc_full[acc], acc_phase);
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        tc_fence_after_sync();
        XW_TL(lane == 0 && (it.ew == 0 || it.ew == GW - 1), it.ew == 0 ? 1 : 2, n, 1);
        const int col_base = it.half * (TN / SPLIT);
        const uint32_t taddr = tmem_base + (uint32_t)acc * ACC_STRIDE + ((uint32_t)(it.quad * 32) << 16) + (uint32_t)col_base;
        const int cls_base = t * TN + col_base;
        Epi::tile_begin(stt, ep, p, it, cls_base, (ti + 1 < t_end) ? cls_base + (p.reverse ? -TN : TN) : -1, TN / SPLIT);
        float va[SC], vb[SC];
        tmem_ld32_async(taddr, va);
        // two slices per trip, NOT fully unrolled: the policy code exists twice, not 2 * SLICES times (i-cache)
#pragma unroll 1
        for (int s = 0; s < SLICES; s += 2) {
          tmem_ld_wait_dep(va);
          tmem_ld32_async(taddr + (uint32_t)(s + 1) * SC, vb);
          __syncwarp();                                          // keeps the prefetch where it is written (see the SWAP loop)
          if (cls_base + s * SC < p.C) xw_call_slice<Epi>(stt, ep, p, it, va, cls_base + s * SC, gscratch);
          tmem_ld_wait_dep(vb);
          if (s + 2 < SLICES) tmem_ld32_async(taddr + (uint32_t)(s + 2) * SC, va);
          __syncwarp();
          if (cls_base + (s + 1) * SC < p.C) xw_call_slice<Epi>(stt, ep, p, it, vb, cls_base + (s + 1) * SC, gscratch);
        }
        XW_TL(lane == 0 && (it.ew == 0 || it.ew == GW - 1), it.ew == 0 ? 1 : 2, n, 2);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) {
          if (PAIR == 2) mbar_arrive_cluster(&acc_empty[acc], 0);
          else mbar_arrive(&acc_empty[acc]);
        }
      }
      if (!ok) break;
      Epi::item_end(stt, ep, p, it, gscratch);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();
  if (warp == 1) { tc_fence_after_sync(); xw_tmem_dealloc<PAIR>(tmem_base, 512); }
}

}  // namespace umma
}  // namespace b200f
