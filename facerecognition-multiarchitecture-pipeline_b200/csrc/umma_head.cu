// tcgen05 / TMEM / TMA engine: host side (ArcFace head K2 / K3, gallery scan K4t, probes).
// Operands of the head are the fp16, L2-normalised, power-of-two-scaled rows K1 emits (dtype B200F_F16N).
//   K2  forward statistics : xw_kernel<PAIR, XW_KK,   XwFwd>    x_hat[B,D] . w_hat[C,D]^T, x_hat resident in smem
//   K3a logit gradient     : xw_kernel<PAIR, XW_SWAP, XwBwdGT>  recompute, class-major G^T (fp16) + r column sums
//   K3b dW                 : xw_kernel<PAIR, XW_MK,   XwDw>     dW^T = x_hat^T . G, normalise-backward fused (B <= 512)
//                            gemm_kernel<K,MN,EpiStore> + l2norm_bwd (B > 512: both operands streamed)
//   K3c dx_hat = G w_hat   : gemm_kernel<MN,MN,EpiStore> split-K, then a fixed-order reduction
//   K4t gallery scan       : xw_kernel<PAIR, XW_KK,   XwTopK<KT>> + gallery_select_kernel
// PAIR = 2 runs the xw kernels on tcgen05 cta_group::2 CTA pairs (clusters of two), PAIR = 1 on single CTAs.
// ALL tcgen05 kernels of the library live in this one translation unit (g_umma_timeout_flag).
#include "umma_api.cuh"
#include "umma_epilogues.cuh"
#include "umma_xw_epilogues.cuh"
#include "umma_xw_topk.cuh"
#include "rowops.cuh"

#include <cudaTypedefs.h>
#include <atomic>
#include <cstdlib>
#include <mutex>

namespace b200f {
namespace umma {

// ---- TMA descriptor encode (driver entry point fetched through the runtime: no libcuda link) -----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D row-major 16-bit tensor [outer, inner] (inner contiguous, row stride ld elements), 128B swizzle.
static int make_tmap(CUtensorMap* m, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                     int box_outer, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(B200F_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8))
    return fail(B200F_ERR_ARG, "TMA operand must be 16B aligned with a row stride that is a multiple of 8 elements");
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200F_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return B200F_OK;
}
// rank-4 map of G^T stored in [32 classes x 64 batch rows] blocks of 4 KB: element (c, b) at
// base[(((c >> 5) * nb + (b >> 6)) << 11) + ((c & 31) << 6) + (b & 63)], nb = batch blocks per class block.  Dimensions
// {b & 63, c & 31, b >> 6, c >> 5}; a box {64, 32, 1, cb} lands in shared memory as [32 cb classes][64 batch rows = 128 bytes]:
// K-major rows for the dW GEMM, and the same bytes are the MN-major [class][64 batch rows] blocks the dx GEMM wants.
// 128-byte swizzle over full 128-byte inner rows (with 64-byte inner rows -- 2 KB blocks -- the boxes did not arrive in the
// layout the MMA descriptors describe).
static int make_tmap_gt_blocked(CUtensorMap* m, const void* base, int64_t classes, int64_t nb, int box_class_blocks) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(B200F_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail(B200F_ERR_ARG, "TMA operand must be 16B aligned");
  cuuint64_t gdim[4] = {64, 32, (cuuint64_t)nb, (cuuint64_t)ceil_div(classes, (int64_t)32)};
  cuuint64_t gstr[3] = {128, 4096, (cuuint64_t)nb * 4096};  // bytes: next class of the block, next batch block, next class block
  cuuint32_t box[4] = {64, 32, 1, (cuuint32_t)box_class_blocks};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200F_ERR_CUDA, "cuTensorMapEncodeTiled (rank 4) failed (%d)", (int)r);
  return B200F_OK;
}
// fp32 matrix [outer, inner] (row stride ld floats) for TMA tensor STORES of [box_outer x box_inner] blocks
static int make_tmap_f32(CUtensorMap* m, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer,
                         CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(B200F_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4))
    return fail(B200F_ERR_ARG, "TMA store target must be 16B aligned with a row stride that is a multiple of 4 floats");
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200F_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed (%d)", (int)r);
  return B200F_OK;
}
// operand whose rows are the m/n index and whose contiguous axis is k
static int tmap_kmajor(CUtensorMap* m, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  return make_tmap(m, base, K, rows, ld, BLOCK_K, box_rows);
}
// operand whose rows are the k index and whose contiguous axis is m/n
static int tmap_mnmajor(CUtensorMap* m, const void* base, int64_t MN, int64_t K, int64_t ld) {
  return make_tmap(m, base, MN, K, ld, 64, BLOCK_K);
}

static GemmParams gemm_params(int M, int N, int K, int k_splits, bool a_mn, bool b_mn, uint32_t a_fmt, uint32_t b_fmt,
                              int pair = 1) {
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (int)ceil_div(M, BLOCK_M * pair);           // tiles of 128 * pair rows
  p.n_tiles = (int)ceil_div(N, BLOCK_N);
  int kchunks = (int)ceil_div(K, BLOCK_K);
  if (k_splits > kchunks) k_splits = kchunks;
  if (k_splits < 1) k_splits = 1;
  p.k_per_split = (int)ceil_div(kchunks, k_splits) * BLOCK_K;
  p.k_splits = (int)ceil_div(K, p.k_per_split);
  // K-major: SBO = 1024 (8 rows x 128 B), k-step 32 B.  MN-major: LBO = 8 KB (next 64-wide block),
  // SBO = 1024 (next 8 k-rows), k-step = 16 rows x 128 B.   (validated on B200 by tools/umma_probe.py)
  p.a_lbo = a_mn ? MN_BLOCK_BYTES : 0; p.a_sbo = 1024; p.a_kstep = a_mn ? 2048 : 32;
  p.b_lbo = b_mn ? MN_BLOCK_BYTES : 0; p.b_sbo = 1024; p.b_kstep = b_mn ? 2048 : 32;
  p.idesc = make_idesc(a_fmt, b_fmt, a_mn, b_mn, BLOCK_M * pair, BLOCK_N);
  return p;
}

static int xw_max_clusters(int pair);

static int gemm_clusters(int pair, int cluster_limit) {
  int clusters = (pair == 2) ? xw_max_clusters(2) : num_sms();
  if (cluster_limit > 0 && cluster_limit < clusters) clusters = cluster_limit;
  return clusters;
}
// tunable "stream_k": 1 = the dx GEMM may cut its (tile, k) space into equal ranges per cluster.  Measured at 4096 x 125 k (32 tiles,
// 74 clusters): 405 -> 497 us -- with split-K the clusters of a wave walk the SAME k range, so the 16 that share a G row block and
// the 2 x 16 that share a w_hat column block find each other's operand tiles in L2; staggered ranges read every tile's operands
// from HBM (4.1 GB instead of 1.15 GB).  Correct, tested, off.
static std::atomic<int> g_stream_k{0};
// Turn p into a stream-K launch when plain split-K would leave more than 7 % of the clusters' time idle and the partial
// buffers (slots of them) suffice.  Returns the slots the reduction sums at most (p.k_splits otherwise).
static int gemm_stream_k(GemmParams& p, int pair, int cluster_limit, int slots) {
  if (!g_stream_k.load(std::memory_order_relaxed) || p.follow_chunks > 0) return p.k_splits;
  const int clusters = gemm_clusters(pair, cluster_limit);
  const int tiles = p.m_tiles * p.n_tiles;
  const int work = tiles * p.k_splits;
  const int rounds = (work + clusters - 1) / clusters;
  if ((double)work >= 0.93 * (double)rounds * clusters) return p.k_splits;
  const int U = (int)ceil_div(p.K, BLOCK_K);
  const long long total = (long long)tiles * U;
  if (total < 4LL * clusters) return p.k_splits;              // too little work to cut
  int most = 0;
  for (int t = 0; t < tiles; ++t) { const int n = stream_k_pieces(t, U, total, clusters); if (n > most) most = n; }
  if (most > slots) return p.k_splits;
  p.stream_k = 1; p.k_units = U;
  return most;
}

// PAIR = 2: clusters of two CTAs on 256 x 256 tiles (p from gemm_params(..., pair = 2); a K-major B map needs
// 128-row boxes: gemm_b_rows(2))
template <int PAIR, bool A_MN, bool B_MN, class Epi>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       const typename Epi::Params& ep, cudaStream_t st, const char* what, int cluster_limit = 0) {
  auto kern = gemm_kernel<PAIR, A_MN, B_MN, Epi>;
  B200F_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  const int work = p.m_tiles * p.n_tiles * p.k_splits;
  int clusters = gemm_clusters(PAIR, cluster_limit);
  if (!p.stream_k && clusters > work) clusters = work;      // stream-K: every cluster has a range (gemm_stream_k checked that)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * PAIR)); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p, ep);
  if (e != cudaSuccess) return fail(B200F_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  B200F_LAUNCH_OK(what);
  return B200F_OK;
}

// What the LAST block of reduce_row_partials_kernel does when the caller wants the loss in the same launch (single
// shard: no all-reduce sits between the statistics and the loss): loss, lse, ||p - q||^2 and the hook scalars for an
// upstream gradient of 1 (rowops::loss_block).  counter: a zeroed word of the workspace (K2 zeroes it; the last
// block leaves it zero again).
struct FwdFinal {
  unsigned int* counter;      // NULL: statistics only
  float s_eff, ls_eps; double C_total;
  float* lse; float* loss; float* pq_norm2;
  rowops::HookCfg hook; float* out4;
};

// One warp per row: sum the per-tile partial records in a fixed order (bitwise reproducible).
__global__ void __launch_bounds__(256)
reduce_row_partials_kernel(const float* __restrict__ part, int n_parts, int64_t B, const float* __restrict__ cos_part,
                           int n_cos, float* row_stats, float* __restrict__ row_best,
                           int64_t* __restrict__ row_argmax, float* __restrict__ cos_minmax, FwdFinal fin) {
  // wait, THEN let the dependent be scheduled: K3a follows and starts its operand loads without waiting for this grid
  // (early_operands); what it loads must be complete, and with K1(W) fused into K2 the operand rows are K2's own output
  pdl_wait(); pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row < B) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, best = -INFINITY;
    int bi = INT32_MAX;
    for (int c = lane; c < n_parts; c += 32) {
      const float* src = part + (row * n_parts + c) * PART_COLS;       // [B, n_parts, PART_COLS]: coalesced
      s0 += src[0]; s1 += src[1]; s2 += src[2]; s3 += src[3];
      const int idx = reinterpret_cast<const int32_t*>(src)[5];
      if (idx >= 0 && (src[4] > best || (src[4] == best && idx < bi))) { best = src[4]; bi = idx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      float* dst = row_stats + row * B200F_STAT_COLS;
      dst[B200F_STAT_SUMEXP] = s0; dst[B200F_STAT_SUMEXP2] = s1; dst[B200F_STAT_ZTARGET] = s2; dst[B200F_STAT_SUMZ] = s3;
      if (row_best) row_best[row] = best;
      if (row_argmax) row_argmax[row] = (bi == INT32_MAX) ? -1 : bi;
    }
  }
  if (blockIdx.x == 0 && cos_minmax != nullptr) {
    __shared__ float smin[8], smax[8];
    float cmin = INFINITY, cmax = -INFINITY;
    for (int i = threadIdx.x; i < n_cos; i += blockDim.x) { cmin = fminf(cmin, cos_part[2 * i]); cmax = fmaxf(cmax, cos_part[2 * i + 1]); }
    cmin = warp_min(cmin); cmax = warp_max(cmax);
    if (lane == 0) { smin[threadIdx.x >> 5] = cmin; smax[threadIdx.x >> 5] = cmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) { cmin = fminf(cmin, smin[w]); cmax = fmaxf(cmax, smax[w]); }
      cos_minmax[0] = cmin; cos_minmax[1] = cmax;
    }
  }
  if (fin.counter != nullptr) {                               // uniform over the grid
    __shared__ bool is_last;
    __shared__ double sh_loss[8], sh_pq[8];
    __threadfence();                                          // this block's row_stats are visible device-wide ...
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(fin.counter, 1u) == gridDim.x - 1);   // ... before it is counted
    __syncthreads();
    if (is_last) {
      __threadfence();
      if (threadIdx.x == 0) *fin.counter = 0u;
      rowops::loss_block(row_stats, B, fin.s_eff, fin.ls_eps, fin.C_total, fin.lse, fin.loss, fin.pq_norm2, &fin.hook,
                         fin.out4, sh_loss, sh_pq);
    }
  }
}

// dst (=|+=) scale / *dev_scale * sum_s part[s], fixed order
// stream-K partials (GemmParams::stream_k): how many slots a tile of the [rows, D] output has is a function of the tile
struct StreamGeom { int on, D, tile_m, tile_n, m_tiles, k_units, n_clusters; long long total; };
__global__ void reduce_splits_kernel(const float* __restrict__ part, int n_splits, int64_t n, float* __restrict__ dst,
                                     int accumulate, float scale, const float* __restrict__ dev_scale, const StreamGeom sg) {
  pdl_trigger(); pdl_wait();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (sg.on) {                                                // a float4 lies inside one tile (tile_n % 4 == 0)
    const int64_t row = i / sg.D;
    const int col = (int)(i - row * sg.D);
    const int tile = (int)(row / sg.tile_m) + sg.m_tiles * (col / sg.tile_n);
    n_splits = stream_k_pieces(tile, sg.k_units, sg.total, sg.n_clusters);
  }
  const float sc = (dev_scale != nullptr) ? scale / __ldg(dev_scale) : scale;
  float4 s = accumulate ? *reinterpret_cast<const float4*>(dst + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k0 = 0; k0 < n_splits; k0 += 4) {                  // four loads in flight per thread, summed in ascending order
    float4 v[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      if (k0 + kk < n_splits) v[kk] = __ldcg(reinterpret_cast<const float4*>(part + (int64_t)(k0 + kk) * n + i));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      if (k0 + kk < n_splits) { acc.x += v[kk].x; acc.y += v[kk].y; acc.z += v[kk].z; acc.w += v[kk].w; }
  }
  s.x = fmaf(acc.x, sc, s.x); s.y = fmaf(acc.y, sc, s.y); s.z = fmaf(acc.z, sc, s.z); s.w = fmaf(acc.w, sc, s.w);
  *reinterpret_cast<float4*>(dst + i) = s;
}

// ---- X-stationary kernel: launch geometry ---------------------------------------------------------
static std::atomic<int> g_pair{2};                 // 2 = tcgen05 cta_group::2 CTA pairs, 1 = single CTAs
// L2 policy hints, a bit mask: 1 K3a streams w_hat evict_first; 2 K3a stores G^T evict_last; 4 K3b stores dW evict_first;
// 8 K3b fetches its w_hat boxes evict_first; 16 K2 streams w_hat evict_last (K3a reads it again); 32 K3b streams G^T evict_last
static std::atomic<int> g_l2_hints{6};
// Epilogue geometry per kernel, as measured at cfg3 on one box (tools/run_r02_*.sh, gpurun_out/r02[a-f]_*): K2 59.8 us with
// one group / 61-65 with two (it is bound by the TMA -> MMA pipeline, 3.0 of its 3.5 us per tile); K3a 75 / 70.5 (epilogue-
// bound: two groups overlap Epi(t) with Epi(t+1)); K3b 91.6 / 99.4 (store-bound: 64-byte pieces per lane and 1 KB w_hat
// boxes move the same bytes less efficiently than 128-byte pieces and 2 KB boxes).
static std::atomic<int> g_k2_groups{1};             // K2: epilogue groups
static std::atomic<int> g_k3b_groups{1};            // K3b: 2 = two epilogue groups on 16-feature slices, 1 = one group on 32
static std::atomic<int> g_early{2};                 // >= 1: K3a / K3c start their loads and MMAs without waiting for the predecessor grid,
                                                    // 2: K3b also loads its resident x_hat^T before it waits
static std::atomic<int> g_epi_groups{1};            // K3a: 2 = two epilogue groups of 8 warps on alternating tiles (16-column slices), 1 = one group of 8,
                                                    //      4 = one group of 16 warps on every tile (column quarters)
#ifdef B200F_PROBES                                 // tools/ builds only; the shipped library has neither the branch nor the switch
static std::atomic<int> g_k3a_ablate{0};            // probe: 2 = no G^T stores (WRONG results)
static std::atomic<int> g_k3b_ablate{0};            // probe: 1 = no w_hat loads, 2 = no dW stores (WRONG results)
#endif
static std::atomic<int> g_k3b_reverse{1};
// tunable "gt_blocked": G^T of a batch > 512 (a multiple of 256 rows, CTA pairs) is stored in [32 classes x 64 batch rows]
// blocks of 4 KB, so that what a K3a warp writes in two consecutive slices is 4 KB contiguous (128 bytes per lane) instead of
// 32 row pieces 8 KB apart.  Probe builds that only moved K3a's stores (tools/build_probe.sh -DB200F_GT_BLOCKED_PROBE): 412
// instead of 465 us at 4096 x 125 k with 2 KB blocks; per 256-row groups ([group][class][256]: pieces 512 bytes apart) it
// stayed at 458; at batch <= 512 (row pitch 1 KB) the probe gains 1.4 us of 60.6, so that path keeps the plain layout.
static std::atomic<int> g_gt_blocked{1};
static std::atomic<int> g_k3a_tma_store{0};         // tunable "k3a_tma_store": 1 = G^T through shared-memory staging + TMA tensor stores (XwBwdGTS)
static std::atomic<int> g_k3a_reverse{0};           // tunable "k3a_reverse": K3a walks each chunk last tile first (K2 read those w_hat rows last)
static std::atomic<int> g_k3c_follow{0};            // tunable "k3c_follow": 1 = the dx part beside the dW part reads the class rows in the dW kernel's order (measured at cfg3: step 263.6 -> 262.3 us, e2e 1.763 -> 1.743 M samples/s: within noise, off)
static std::atomic<int> g_dw_n_fastest{1};          // tunable "dw_n_fastest": streamed dW GEMM (batch > 512) runs the n tiles of a class block side by side
static std::atomic<int> g_k3b_tma_store{0};         // tunable "k3b_tma_store": 1 = dW through shared-memory staging + TMA tensor stores (XwDwTS)
static std::atomic<int> g_target_patch{2};          // tunable "target_patch": 0 = a slice with a target goes element-wise (round 1), 1 = K2 / K3a exchange the element
                                                    // in place, 2 = ... and K3a queues its patches to the end of the item
static std::atomic<int> g_x_whole{0};               // tunable "x_whole": 1 = first MMA of an item waits for the whole resident operand
// "stage_events" tunable: record a CUDA event pair around each GEMM kernel of the head on the launching stream
// (bench.py's per-kernel durations; eager launches only -- never inside a graph capture).
static std::atomic<int> g_stage_events{0};
enum { EV_K2 = 0, EV_K3A, EV_K3B, EV_K3C, EV_COUNT };
static const char* const kEvNames[EV_COUNT] = {"k2", "k3a", "k3b", "k3c"};
constexpr int EV_MAX_CHUNKS = 32;                       // class chunks of one backward call that get their own event pair
struct StageEvents { cudaEvent_t beg[EV_COUNT][EV_MAX_CHUNKS], end[EV_COUNT][EV_MAX_CHUNKS]; int n[EV_COUNT] = {}; bool made = false; };
static thread_local StageEvents g_ev;
static void stage_reset(int which) { g_ev.n[which] = 0; }
static void stage_event(int which, bool is_end, cudaStream_t st) {
  if (!g_stage_events.load(std::memory_order_relaxed)) return;
  if (!g_ev.made) {
    for (int i = 0; i < EV_COUNT; ++i)
      for (int j = 0; j < EV_MAX_CHUNKS; ++j) { cudaEventCreate(&g_ev.beg[i][j]); cudaEventCreate(&g_ev.end[i][j]); }
    g_ev.made = true;
  }
  const int j = g_ev.n[which];
  if (j >= EV_MAX_CHUNKS) return;
  cudaEventRecord(is_end ? g_ev.end[which][j] : g_ev.beg[which][j], st);
  if (is_end) g_ev.n[which] = j + 1;
}           // K3b walks each chunk last tile first
static std::atomic<int> g_prefetch{0};             // L2 prefetch distance of the xw producer (tiles of the streamed operand)
static std::atomic<int> g_chunk_mb{112};           // budget of the fp16 logit-gradient buffer G per class chunk

// MODE of the X-stationary kernel: 0 = both operands K-major (K2, gallery scan, probes); 1 = resident operand
// MN-major, streamed operand K-major (K3b: x_hat^T resident, class-major G rows streamed); 2 = K-major operands with
// the streamed one on the A side of the MMA (K3a: accumulator lanes = classes).
enum { XW_KK = 0, XW_MK = 1, XW_SWAP = 2, XW_SWAP_MK = 3 };   // _MK: resident operand MN-major; SWAP: streamed rows on the A side
template <int PAIR, int MODE, class Epi>
struct XwKernel {
  static constexpr auto fn = xw_kernel<PAIR, MODE == XW_MK || MODE == XW_SWAP_MK, false, MODE == XW_SWAP || MODE == XW_SWAP_MK, Epi>;
};

template <int PAIR, int MODE, class Epi>
static int xw_set_smem() {
  static thread_local int done_dev = -1;
  int dev = 0;
  B200F_CUDA_OK(cudaGetDevice(&dev));
  if (done_dev != dev) {
    B200F_CUDA_OK(cudaFuncSetAttribute(XwKernel<PAIR, MODE, Epi>::fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xw_smem_bytes<Epi>()));
    done_dev = dev;
  }
  return B200F_OK;
}

// clusters of `pair` CTAs that can be co-resident (1 CTA per SM: 227 KB of shared memory each)
static int xw_max_clusters(int pair) {
#ifdef B200F_TIMELINE     // instrumented builds only: fewer clusters, to tell a per-SM limit from a chip-wide one
  if (const char* e = getenv("B200F_TL_CLUSTERS")) { const int n = atoi(e); if (n > 0) return n; }
#endif
  if (pair == 1) return num_sms();
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return num_sms() / 2;
  if (dev != cached_dev) {
    int n = 0;
    if (xw_set_smem<2, XW_KK, XwFwd>() == B200F_OK) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)num_sms() / 2 * 2); cfg.blockDim = dim3(XW_THREADS); cfg.dynamicSmemBytes = XW_SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, XwKernel<2, XW_KK, XwFwd>::fn, &cfg) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
    }
    if (n <= 0 || n > num_sms() / 2) n = num_sms() / 2;
    cached = n; cached_dev = dev;
  }
  return cached;
}

static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

struct XwPlan { int pair, m_groups, n_tiles, n_clusters, n_chunks, items, grid; };

static XwPlan xw_plan(int64_t B, int64_t C, int pair, int max_chunks = 0, int cluster_limit = 0) {
  XwPlan q{};
  q.pair = pair;
  q.m_groups = (int)ceil_div(B, (int64_t)XW_M * pair);
  q.n_tiles = (int)ceil_div(C, (int64_t)XW_WROWS * pair);
  q.n_clusters = xw_max_clusters(pair);
  if (cluster_limit > 0 && cluster_limit < q.n_clusters) q.n_clusters = cluster_limit;   // part of the chip (head_bwd parts)
  if (max_chunks > 0) {                                   // few, long chunks (the gallery's sample pre-pass)
    q.n_chunks = max_chunks < q.n_tiles ? max_chunks : q.n_tiles;
    q.items = q.m_groups * q.n_chunks;
    q.grid = pair * (q.items < q.n_clusters ? q.items : q.n_clusters);
    return q;
  }
  // items = m_groups * n_chunks is a multiple of the cluster count whenever the class range allows it
  int nc = q.n_clusters / gcd_int(q.n_clusters, q.m_groups);
  if (nc > q.n_tiles) nc = q.n_tiles;
  if (nc < 1) nc = 1;
  q.n_chunks = nc;
  q.items = q.m_groups * q.n_chunks;
  q.grid = pair * (q.items < q.n_clusters ? q.items : q.n_clusters);
  return q;
}

// K1 of the streamed rows inside the kernel (XwParams::prep_*; policies with kPrepWarps > 0)
struct XwPrepArgs { const void* src; int f32; uint16_t* dst; float* inv; unsigned int* ready; float eps, scale; };

#ifdef B200F_TIMELINE
// tools/timeline_probe.py: the (which)-th X-stationary launch after this call stamps its timeline into buf (XW_TL)
static unsigned long long* g_tl_buf = nullptr;
static int g_tl_which = -1, g_tl_count = 0;
extern "C" int b200f_xw_timeline(void* buf, int which) {
  g_tl_buf = static_cast<unsigned long long*>(buf); g_tl_which = which; g_tl_count = 0;
  return B200F_OK;
}
#endif

template <int PAIR, int MODE, class Epi>
static int launch_xw(const CUtensorMap& tx, const CUtensorMap& tw, const XwPlan& q, int64_t B, int64_t C, int D,
                     const typename Epi::Params& ep, cudaStream_t st, const char* what, uint32_t fmt = FMT_F16,
                     bool reverse = false, const void* w_base = nullptr, int64_t w_row_bytes = 0, int early = 0,
                     int w_hint = 0, const XwPrepArgs* prep = nullptr, int prefetch_tiles = -1) {
  int rc = xw_set_smem<PAIR, MODE, Epi>(); if (rc) return rc;
  XwParams p{};
  p.B = (int)B; p.C = (int)C; p.D = D;
  p.kb_count = (int)ceil_div(D, XW_K);
  p.m_groups = q.m_groups; p.n_tiles = q.n_tiles; p.n_chunks = q.n_chunks;
  p.prefetch = (w_base != nullptr) ? (prefetch_tiles >= 0 ? prefetch_tiles : g_prefetch.load(std::memory_order_relaxed)) : 0;
  p.w_base = w_base; p.w_row_bytes = w_row_bytes;
  p.early_operands = (early && g_early.load(std::memory_order_relaxed)) ? early : 0;   // 1: both operands older than the predecessor, 2: the resident one
  p.w_hint = w_hint;
#ifdef B200F_TIMELINE
  p.tl = (g_tl_buf != nullptr && g_tl_count++ == g_tl_which) ? g_tl_buf : nullptr;
#endif
  p.tn = XW_WROWS * PAIR;
  p.reverse = reverse ? 1 : 0;
  p.x_whole = g_x_whole.load(std::memory_order_relaxed);
  if (prep != nullptr) {
    p.prep_src = prep->src; p.prep_f32 = prep->f32; p.prep_dst = prep->dst; p.prep_inv = prep->inv; p.prep_ready = prep->ready;
    p.prep_eps = prep->eps; p.prep_scale = prep->scale;
    int cw = q.n_clusters / (q.m_groups > 0 ? q.m_groups : 1);       // chunks in flight at once: one per m_groups clusters
    if (cw < 1) cw = 1;
    if (cw > q.n_chunks) cw = q.n_chunks;
    p.prep_cw = cw;
    p.prefetch = 0;                                                   // nothing to pull into L2: the rows are being written
  }
  p.idesc = make_idesc(fmt, fmt, MODE == XW_MK, MODE == XW_SWAP_MK, XW_M * PAIR, XW_WROWS * PAIR);   // A = resident, except SWAP modes
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)q.grid); cfg.blockDim = dim3((unsigned)xw_threads<Epi>()); cfg.dynamicSmemBytes = xw_smem_bytes<Epi>(); cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, XwKernel<PAIR, MODE, Epi>::fn, tx, tw, p, ep);
  if (e != cudaSuccess) return fail(B200F_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  B200F_LAUNCH_OK(what);
  return B200F_OK;
}

// ---- plan -------------------------------------------------------------------------------------------
struct Plan {
  XwPlan fwd;
  size_t off_part, off_cos, off_counter, off_ready;
  int64_t Cc, ldg; int n_chunks, dx_splits, dx_slots;   // dx_slots: partial buffers of the dx GEMM the workspace holds
  size_t off_G, off_dxpart, off_rpart, off_sq, n_sq;
  int n_rb;                 // 32-row blocks of the batch that K3a emits r partials for
  bool fused_dw;            // B <= 512: dW GEMM on the MN-major X-stationary kernel (x_hat^T resident); above that
                            // on the generic core, transposed the same way; the normalise-backward is fused in both
  size_t total;
};

static Plan make_plan(int64_t B, int64_t C, int D) {
  Plan pl{};
  const int pair = g_pair.load(std::memory_order_relaxed);
  pl.fwd = xw_plan(B, C, pair);
  const int m_tiles = (int)ceil_div(B, BLOCK_M);
  size_t off = 0;
  // sized for either pairing so that a workspace stays valid when the mode is switched
  const XwPlan p1 = xw_plan(B, C, 1), p2 = xw_plan(B, C, 2);
  const size_t part_max = (size_t)(p1.n_chunks > p2.n_chunks ? p1.n_chunks : p2.n_chunks);
  const size_t cos_max = (size_t)(p1.items > 2 * p2.items ? p1.items : 2 * p2.items) * XW_EPI_WARPS;
  pl.off_part = off; off += align_up(sizeof(float) * part_max * XW_MAX_EPI_GROUPS * B * PART_COLS, 256);   // one record per epilogue group
  pl.off_cos = off;  off += align_up(sizeof(float) * 2 * cos_max * XW_MAX_EPI_GROUPS, 256);
  pl.off_counter = off; off += 256;                       // the fused loss finalize counts its blocks here
  pl.off_ready = off; off += align_up(sizeof(unsigned int) * (size_t)ceil_div(C, (int64_t)XW_WROWS), 256);   // K1(W)-in-K2 row counts
  const size_t fwd_total = off;
  // backward: classes are processed in chunks whose fp16 logit gradient G fits the budget.  G is written once
  // and read twice, all while the tensor pipe (not HBM) is the bound, so it need not stay L2-resident: cfg3's
  // 102 MB is ONE chunk (fewer launches, long per-CTA streams); B = 4096 shards use ~12 k-class chunks.
  const int64_t ldgt = ceil_div(B, 64) * 64;               // G^T[class][batch row], row stride in elements
  // batches above 512 rows get 12x the budget (1.3 GB: one rank's share of cfg4 at 8 GPUs is ONE chunk): long chunks
  // keep the per-launch fill / drain and the wave quantisation small (memory is not the constraint on a 180 GB part;
  // measured at B = 4096 x 125 k classes: 1 chunk 2147 us, 3 chunks 2193 us, 6 chunks 2322 us)
  const int64_t budget_mb = (int64_t)g_chunk_mb.load(std::memory_order_relaxed) * (B > (int64_t)XW_MAX_KB * XW_K ? 12 : 1);
  int64_t cc_max = (budget_mb * (1 << 20) / 2 / ldgt) / BLOCK_N * BLOCK_N;
  if (cc_max < BLOCK_N) cc_max = BLOCK_N;
  pl.n_chunks = (int)ceil_div(C, cc_max);
  pl.Cc = ceil_div(ceil_div(C, pl.n_chunks), BLOCK_N) * BLOCK_N;
  pl.n_chunks = (int)ceil_div(C, pl.Cc);
  pl.ldg = ldgt;
  // split-K of the dX GEMM: one wave of CTAs (pair = 1) or of clusters (pair = 2) over m-tiles x n-tiles x splits
  const int out_tiles = (int)ceil_div(B, BLOCK_M * pair) * (int)ceil_div(D, BLOCK_N);
  int splits = (pair == 2 ? xw_max_clusters(2) : num_sms()) / out_tiles;
  if (splits < 1) splits = 1;
  const int kchunks = (int)(pl.Cc / BLOCK_K);
  if (splits > kchunks) splits = kchunks;
  pl.dx_splits = splits;
  off = 0;
  pl.off_G = off;      off += align_up(2 * (size_t)pl.Cc * pl.ldg, 1024);
  int splits_max = num_sms() / ((int)ceil_div(B, BLOCK_M) * (int)ceil_div(D, BLOCK_N));   // either pairing fits
  if (splits_max < splits) splits_max = splits;
  if (splits_max < 1) splits_max = 1;
  {  // stream-K of the dx GEMM: a tile is touched by at most ceil(clusters / tiles) + 1 clusters (GemmParams::stream_k)
    const int cl = (pair == 2) ? xw_max_clusters(2) : num_sms();
    const int sk = (cl + out_tiles - 1) / out_tiles + 1;
    if (splits_max < sk) splits_max = sk;
  }
  pl.dx_slots = splits_max;
  pl.off_dxpart = off; off += align_up(sizeof(float) * (size_t)splits_max * B * D, 256);
  pl.fused_dw = (B <= (int64_t)XW_MAX_KB * XW_K);         // x_hat^T resident (else: both operands streamed)
  pl.n_rb = (int)ceil_div(B, 2 * XW_M) * 2 * 4;          // covers either pairing
  pl.off_rpart = off; off += align_up(sizeof(float) * (size_t)pl.n_rb * pl.Cc, 256);
  // sum-of-squares partials of the dW epilogues (b200f_head_request_dw_sqnorm): per (item, CTA, epilogue warp) on the
  // X-stationary kernel, per (128-row block, n tile, warp) on the streamed one
  const XwPlan qw1 = xw_plan(D, pl.Cc, 1), qw2 = xw_plan(D, pl.Cc, 2);
  size_t n_sq = (size_t)(qw1.items > 2 * qw2.items ? qw1.items : 2 * qw2.items) * XW_MAX_EPI_GROUPS * XW_EPI_WARPS;
  const size_t n_sq_stream = (size_t)ceil_div(pl.Cc, (int64_t)128) * (size_t)ceil_div((int64_t)D, (int64_t)BLOCK_N) * 4;
  if (n_sq_stream > n_sq) n_sq = n_sq_stream;
  pl.n_sq = n_sq;
  pl.off_sq = off; off += align_up(sizeof(float) * n_sq, 256);
  pl.total = (off > fwd_total ? off : fwd_total) + 1024;
  return pl;
}

static bool device_is_sm100() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev != cached_dev) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
    cached = (major == 10) ? 1 : 0;
    cached_dev = dev;
  }
  return cached == 1;
}

bool available() { return device_is_sm100(); }

size_t head_workspace_bytes(int64_t B, int64_t C, int D) { return make_plan(B, C, D).total; }

static char* ws_align(char* ws) { return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023); }

static int check_shape(int64_t B, int64_t C, int D, const b200f_head_cfg* cfg) {
  if (!device_is_sm100()) return fail(B200F_ERR_UNSUPPORTED, "tcgen05 engine needs an sm_100 device");
  if (D % 8 != 0 || D > XW_MAX_KB * XW_K)
    return fail(B200F_ERR_UNSUPPORTED, "tcgen05 engine needs D %% 8 == 0 and D <= %d (got %d)", XW_MAX_KB * XW_K, D);
  if (B >= ((int64_t)1 << 30) || C >= ((int64_t)1 << 30)) return fail(B200F_ERR_UNSUPPORTED, "tcgen05 engine: B, C < 2^30");
  if (!(cfg->operand_scale > 0.f)) return fail(B200F_ERR_ARG, "tcgen05 engine: cfg.operand_scale must be > 0");
  return B200F_OK;
}

int head_dx_finish(const void* xh, float S, const HeadDx* hdx, const float* dxhat, int64_t B, int D, cudaStream_t st) {
  __nv_bfloat16* lowp = static_cast<__nv_bfloat16*>(hdx->dx_bf16);
  if (hdx->x_raw != nullptr && hdx->x_raw_dtype == B200F_BF16)
    rowops::launch_l2norm_bwd<__nv_bfloat16, false>(static_cast<const __nv_bfloat16*>(hdx->x_raw), 1.f, hdx->inv_nx, dxhat, B, D, hdx->dx, st, lowp);
  else if (hdx->x_raw != nullptr && hdx->x_raw_dtype == B200F_F32)
    rowops::launch_l2norm_bwd<float, false>(static_cast<const float*>(hdx->x_raw), 1.f, hdx->inv_nx, dxhat, B, D, hdx->dx, st, lowp);
  else
    rowops::launch_l2norm_bwd<__half, true>(static_cast<const __half*>(xh), S, hdx->inv_nx, dxhat, B, D, hdx->dx, st, lowp);
  B200F_LAUNCH_OK("l2norm_bwd kernel (dx)");
  return B200F_OK;
}

// tunable "k2_prep": 1 = K1 of the class weights runs inside K2 (HeadPrep: prep warps + per-128-row hand-over counters),
// 0 = as its own pass in front of K2 (default), 2 = probe: the 20-warp kernel with idle prep warps.  Measured at cfg3 on one
// box each (gpurun_out/r02k_*, r02m_*): K1(W) 38.5-40.5 us + K2 60 us apart; fused 227 us with 2 prep warps (168
// registers), 140 us with 6 (128), 115 us with 10 (96 registers, 16-column epilogue -- which alone still runs K2 in
// 60.6 us), 123 us with packed fp32x2 math on shorter trips, 98.4 us once the gpu-scope release of the row counters was
// paid per run of rows instead of per trip (it waits for the warp's stores: ~2 us), 104.7 us with L2 prefetches on top.
// So the hand-over works and the rows agree to the last bit or one, but the fused kernel only breaks even: normalising a
// row is latency-bound scalar work, ten warps per SM turn over ~7 rows/us where the stand-alone K1 (16 warps, the SM to
// itself) does 16.4, and K2 finishes ~20 us after its last row arrives.  What K2 has to spare is HBM bandwidth, not warps.
static std::atomic<int> g_k2_prep{0};

__global__ void zero_words_kernel(unsigned int* p, int n) {
  pdl_trigger(); pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0u;
}

bool head_prep_fused(int64_t B, int64_t C, int D, const HeadPrep* prep, const void* wh) {
  // a prep warp keeps one slot per lane: rows of a round per warp <= 32
  const XwPlan q = xw_plan(B, C, g_pair.load(std::memory_order_relaxed));
  int cw = q.n_clusters / (q.m_groups > 0 ? q.m_groups : 1);
  if (cw < 1) cw = 1;
  if (cw > q.n_chunks) cw = q.n_chunks;
  const int64_t n_warps = (int64_t)q.grid * XwFwdP::kPrepWarps;
  if (ceil_div((int64_t)cw * XW_WROWS * q.pair, n_warps) > 32) return false;
  return prep != nullptr && prep->w_raw != nullptr && g_k2_prep.load(std::memory_order_relaxed) == 1 &&
         g_k2_groups.load(std::memory_order_relaxed) == 1 && D == 512 &&
         ((reinterpret_cast<uintptr_t>(prep->w_raw) | reinterpret_cast<uintptr_t>(wh)) & 31) == 0;
}

int head_fwd(const void* xh, const void* wh, const int64_t* label, int64_t B, int64_t C, int64_t class_offset, int D,
             const b200f_head_cfg* cfg, float* row_stats, float* row_best, int64_t* row_argmax, float* cos_minmax,
             int32_t* nan_flag, const HeadFinal* fin, char* ws, size_t ws_bytes, cudaStream_t st, const HeadPrep* prep) {
  int rc = check_shape(B, C, D, cfg); if (rc) return rc;
  const Plan pl = make_plan(B, C, D);
  if (ws_bytes < pl.total) return fail(B200F_ERR_WORKSPACE, "umma head_fwd: workspace too small");
  ws = ws_align(ws);
  // K1: the batch rows now (they are the resident operand), the class weights inside K2 (prep warps) or as a pass of their own
  const bool fused_prep = head_prep_fused(B, C, D, prep, wh);
  unsigned int* ready = reinterpret_cast<unsigned int*>(ws + pl.off_ready);
  const int n_ready = (int)ceil_div(C, (int64_t)XW_WROWS);
  if (prep != nullptr) {
    const float S = cfg->operand_scale;
    bool zeroed = false;
    if (prep->x_raw != nullptr) {
      __half* xo = static_cast<__half*>(const_cast<void*>(xh));
      if (fused_prep && prep->x_dtype == B200F_BF16 && D == 512 &&
          ((reinterpret_cast<uintptr_t>(prep->x_raw) | reinterpret_cast<uintptr_t>(xh)) & 31) == 0) {
        launch_pdl(rowops::l2norm_rows_512x16_zero_kernel<__nv_bfloat16, __half>, dim3((unsigned)ceil_div(B, (int64_t)rowops::ROWS_PER_BLOCK)),
                   dim3(rowops::WARPS_PER_BLOCK * 32), 0, st, static_cast<const __nv_bfloat16*>(prep->x_raw), B, prep->eps, S,
                   prep->inv_nx, xo, ready, n_ready);
        zeroed = true;
      } else if (prep->x_dtype == B200F_BF16) {
        rowops::launch_l2norm_rows<__nv_bfloat16, __half>(static_cast<const __nv_bfloat16*>(prep->x_raw), B, D, prep->eps, S, prep->inv_nx, xo, st);
      } else {
        rowops::launch_l2norm_rows<float, __half>(static_cast<const float*>(prep->x_raw), B, D, prep->eps, S, prep->inv_nx, xo, st);
      }
      B200F_LAUNCH_OK("K1 l2norm_rows (x)");
    }
    if (fused_prep && !zeroed) {
      launch_pdl(zero_words_kernel, dim3((unsigned)ceil_div((int64_t)n_ready, (int64_t)256)), dim3(256), 0, st, ready, n_ready);
      B200F_LAUNCH_OK("zero_words_kernel");
    }
    if (!fused_prep && prep->w_raw != nullptr) {
      __half* wo = static_cast<__half*>(const_cast<void*>(wh));
      if (prep->w_dtype == B200F_BF16)
        rowops::launch_l2norm_rows<__nv_bfloat16, __half>(static_cast<const __nv_bfloat16*>(prep->w_raw), C, D, prep->eps, S, prep->inv_nw, wo, st);
      else
        rowops::launch_l2norm_rows<float, __half>(static_cast<const float*>(prep->w_raw), C, D, prep->eps, S, prep->inv_nw, wo, st);
      B200F_LAUNCH_OK("K1 l2norm_rows (W)");
    }
  }
  CUtensorMap tx, tw;
  rc = tmap_kmajor(&tx, xh, B, D, D, XW_M); if (rc) return rc;
  rc = tmap_kmajor(&tw, wh, C, D, D, XW_WROWS); if (rc) return rc;
  const XwPlan& q = pl.fwd;
  XwFwd::Params ep{};
  ep.label = label; ep.class_offset = class_offset;
  ep.hm = HeadMath{cfg->m_eff, cfg->s_eff, cfg->easy_margin};
  ep.inv_scale = 1.0f / (cfg->operand_scale * cfg->operand_scale);
  ep.part = reinterpret_cast<float*>(ws + pl.off_part);
  ep.cos_part = reinterpret_cast<float*>(ws + pl.off_cos);
  ep.nan_flag = nan_flag; ep.pair = q.pair;
  ep.zero_word = reinterpret_cast<unsigned int*>(ws + pl.off_counter);
  ep.whole_slice_targets = g_target_patch.load(std::memory_order_relaxed) ? 0 : 1;
  const int eg = g_k2_groups.load(std::memory_order_relaxed);
  const int64_t wrb = (int64_t)D * 2;
  const int k2_hint = (g_l2_hints.load(std::memory_order_relaxed) & 16) ? 2 : 0;
  B200F_NVTX("K2 cosine GEMM + margin + softmax statistics");
  stage_reset(EV_K2);
  stage_event(EV_K2, false, st);
  const bool p_kernel_only = !fused_prep && g_k2_prep.load(std::memory_order_relaxed) == 2;   // probe: the 20-warp kernel, prep warps idle
  if (fused_prep || p_kernel_only) {
    XwFwdP::Params epp{};
    epp.label = ep.label; epp.class_offset = ep.class_offset; epp.hm = ep.hm; epp.inv_scale = ep.inv_scale; epp.part = ep.part;
    epp.cos_part = ep.cos_part; epp.nan_flag = ep.nan_flag; epp.pair = ep.pair; epp.zero_word = ep.zero_word;
    epp.whole_slice_targets = ep.whole_slice_targets;
    const XwPrepArgs pa{prep->w_raw, prep->w_dtype == B200F_F32 ? 1 : 0, static_cast<uint16_t*>(const_cast<void*>(wh)), prep->inv_nw,
                        ready, prep->eps, cfg->operand_scale};
    rc = (q.pair == 2) ? launch_xw<2, XW_KK, XwFwdP>(tx, tw, q, B, C, D, epp, st, "umma K2 arcface_fwd + K1(W) (cta pair)", FMT_F16, false, wh, wrb, false, k2_hint, fused_prep ? &pa : nullptr)
                       : launch_xw<1, XW_KK, XwFwdP>(tx, tw, q, B, C, D, epp, st, "umma K2 arcface_fwd + K1(W)", FMT_F16, false, wh, wrb, false, k2_hint, fused_prep ? &pa : nullptr);
  } else if (eg == 2) {
    XwFwd2::Params ep2{};
    ep2.label = ep.label; ep2.class_offset = ep.class_offset; ep2.hm = ep.hm; ep2.inv_scale = ep.inv_scale; ep2.part = ep.part;
    ep2.cos_part = ep.cos_part; ep2.nan_flag = ep.nan_flag; ep2.pair = ep.pair; ep2.zero_word = ep.zero_word;
    ep2.whole_slice_targets = ep.whole_slice_targets;
    rc = (q.pair == 2) ? launch_xw<2, XW_KK, XwFwd2>(tx, tw, q, B, C, D, ep2, st, "umma K2 arcface_fwd (cta pair, 2 epilogue groups)", FMT_F16, false, wh, wrb, false, k2_hint)
                       : launch_xw<1, XW_KK, XwFwd2>(tx, tw, q, B, C, D, ep2, st, "umma K2 arcface_fwd (2 epilogue groups)", FMT_F16, false, wh, wrb, false, k2_hint);
  } else {
    rc = (q.pair == 2) ? launch_xw<2, XW_KK, XwFwd>(tx, tw, q, B, C, D, ep, st, "umma K2 arcface_fwd (cta pair)", FMT_F16, false, wh, wrb, false, k2_hint)
                       : launch_xw<1, XW_KK, XwFwd>(tx, tw, q, B, C, D, ep, st, "umma K2 arcface_fwd", FMT_F16, false, wh, wrb, false, k2_hint);
  }
  stage_event(EV_K2, true, st);
  if (rc) return rc;
  FwdFinal ff{};
  if (fin != nullptr) {                                     // single shard: the loss comes out of the same launch
    ff.counter = ep.zero_word; ff.s_eff = cfg->s_eff; ff.ls_eps = cfg->label_smoothing; ff.C_total = (double)cfg->num_classes_total;
    ff.lse = fin->lse; ff.loss = fin->loss; ff.pq_norm2 = fin->pq_norm2;
    ff.hook = rowops::HookCfg{fin->hook_enabled, fin->max_grad_norm, fin->phase, fin->epoch}; ff.out4 = fin->out4;
  }
  launch_pdl(reduce_row_partials_kernel, dim3((unsigned)ceil_div(B, 8)), dim3(256), 0, st, ep.part, q.n_chunks * eg, B, ep.cos_part,
                                                                      q.items * q.pair * XW_EPI_WARPS * eg, row_stats, row_best,
                                                                      row_argmax, cos_minmax, ff);
  B200F_LAUNCH_OK("reduce_row_partials_kernel");
  return B200F_OK;
}

// ---- ||dW||^2 as a side output of the dW epilogues ---------------------------------------------------------------
// b200f_head_request_dw_sqnorm(out): the calling thread's NEXT backward (b200f_arcface_bwd / _bwd_dx, or the phase 1 +
// phase 2 pair) also leaves sum(dW^2) of its class rows in out[0].  The epilogues add up what they store (fixed order:
// bitwise reproducible), one small kernel folds the per-warp partials.
static thread_local float* g_dw_sq_out = nullptr;
void head_request_dw_sqnorm(float* out) { g_dw_sq_out = out; }
float* head_dw_sqnorm_request() { return g_dw_sq_out; }

__global__ void __launch_bounds__(256) fold_partials_kernel(const float* __restrict__ part, int n, float* out, int accumulate) {
  pdl_trigger(); pdl_wait();
  __shared__ float sh[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += part[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + sh[0] : sh[0];
}

int head_bwd_parts_ok(int64_t B, int64_t C, int D) {
  if (B <= 0 || C <= 0 || D <= 0 || D % 8 != 0 || D > XW_MAX_KB * XW_K) return 0;
  const Plan pl = make_plan(B, C, D);
  return (pl.n_chunks == 1 && pl.fused_dw) ? 1 : 0;
}

int head_bwd(const void* xh, const void* wh, const float* inv_nw, const int64_t* label, const float* lse,
             const float* grad4, int64_t B, int64_t C, int64_t class_offset, int D, const b200f_head_cfg* cfg,
             float* dxhat, float* dw, const HeadDx* hdx, char* ws, size_t ws_bytes, cudaStream_t st, int phase,
             int cluster_limit) {
  int rc = check_shape(B, C, D, cfg); if (rc) return rc;
  const Plan pl = make_plan(B, C, D);
  if (ws_bytes < pl.total) return fail(B200F_ERR_WORKSPACE, "umma head_bwd: workspace too small");
  const bool part = phase >= HEAD_BWD_PART_K3A && phase <= HEAD_BWD_PART_K3C;
  if ((phase < 0 || phase > 2) && !part) return fail(B200F_ERR_ARG, "umma head_bwd: phase must be 0, 1 or 2 (or a part)");
  if (phase != 0 && phase != HEAD_BWD_PART_K3C && hdx != nullptr)
    return fail(B200F_ERR_ARG, "umma head_bwd: the fused dL/dx tail belongs to the unsplit call or to the dx part");
  // Parts (b200f_arcface_bwd_part): the three GEMM stages as separate calls, the dW and the dx stage each on a bounded
  // number of clusters, so that the caller can run them SIDE BY SIDE on two streams behind K3a.  One class chunk only.
  if (part && (pl.n_chunks != 1 || !pl.fused_dw))
    return fail(B200F_ERR_UNSUPPORTED, "umma head_bwd: parts need a single class chunk and batch <= 512");
  // phase 0: per class chunk K3a -> K3b -> K3c -> split reduction (one call does everything).
  // phases 1 / 2 (class shards): the dx_hat half first -- per chunk K3a -> K3c -> split reduction -> K3b, WITHOUT the last
  // chunk's K3b (phase 1) -- so that the caller can start the cross-rank all-reduce of dx_hat on another stream, and then
  // that last dW GEMM (phase 2: it needs the last chunk's G^T and r partials, which stay in the workspace) runs beside it.
  ws = ws_align(ws);
  uint16_t* G = reinterpret_cast<uint16_t*>(ws + pl.off_G);
  float* dxpart = reinterpret_cast<float*>(ws + pl.off_dxpart);
  float* const sq_out = g_dw_sq_out;                        // consumed by the call that runs the last dW GEMM
  if (phase == 0 || phase == 2 || phase == HEAD_BWD_PART_K3B) g_dw_sq_out = nullptr;
  float* const sq_part = sq_out ? reinterpret_cast<float*>(ws + pl.off_sq) : nullptr;
  const float S = cfg->operand_scale;
  CUtensorMap tx_k, tx_mn;
  rc = tmap_kmajor(&tx_k, xh, B, D, D, XW_M); if (rc) return rc;
  rc = tmap_mnmajor(&tx_mn, xh, D, B, D); if (rc) return rc;
  stage_reset(EV_K3A); stage_reset(EV_K3B); stage_reset(EV_K3C);
  // K3a's epilogue geometry decides how many r partial rows K3b sums: read the tunable ONCE per backward -- and in phase 2
  // take what this thread's phase 1 used (another thread may have moved the tunable between the two calls)
  static thread_local int t_k3a_mode = 1;
  if (phase == 0 || phase == 1 || phase == HEAD_BWD_PART_K3A) t_k3a_mode = g_epi_groups.load(std::memory_order_relaxed);
  const int k3a_mode = t_k3a_mode;
  const int gpair = pl.fwd.pair;                            // generic core: single CTAs or cta_group::2 pairs, like K2 / K3a
  const int64_t wrb = (int64_t)D * 2;
  int chunk_no = 0;
  for (int64_t c0 = 0; c0 < C; c0 += pl.Cc, ++chunk_no) {
    const int64_t cnt = (C - c0 < pl.Cc) ? (C - c0) : pl.Cc;
    const bool last_chunk = c0 + pl.Cc >= C;
    const uint16_t* wc = static_cast<const uint16_t*>(wh) + c0 * D;
    if (phase == 2 && !last_chunk) continue;                // everything but the last chunk's dW ran in phase 1
    const XwPlan qg = xw_plan(B, cnt, pl.fwd.pair);
    // G^T in 2 KB blocks (see g_gt_blocked): the streamed path only, whole 256-row groups only; read once per backward call --
    // phase 2 (the held-back dW GEMM) must read the layout phase 1 wrote
    static thread_local int t_blocked = 0;
    if (phase == 0 || phase == 1 || phase == HEAD_BWD_PART_K3A)
      t_blocked = g_gt_blocked.load(std::memory_order_relaxed) != 0 ? 1 : 0;
    const bool blocked = t_blocked != 0 && !pl.fused_dw && pl.fwd.pair == 2 && B % (2 * XW_M) == 0;
    const int64_t NB = B / 64;                                  // batch blocks per class block
    float* r_part = reinterpret_cast<float*>(ws + pl.off_rpart);
    const int hints = g_l2_hints.load(std::memory_order_relaxed);
    // --- K3a: logit gradient of the chunk, class-major: G^T[c, b] (x_hat resident, w_hat rows [c0, c0 + cnt) streamed
    //     on the A side, so a thread owns a class and r_c = sum_b G cos is a private sum)
    auto run_k3a = [&]() -> int {
    CUtensorMap tw_k;
    int rc = tmap_kmajor(&tw_k, wc, cnt, D, D, XW_WROWS); if (rc) return rc;
    const int hints_k3a = hints;
    auto fill = [&](auto& e) {
      e.label = label; e.lse = lse; e.grad4 = grad4; e.class_offset = class_offset + c0;
      e.hm = HeadMath{cfg->m_eff, cfg->s_eff, cfg->easy_margin};
      e.ls_eps = cfg->label_smoothing; e.inv_Ctot = 1.0f / (float)cfg->num_classes_total; e.inv_scale = 1.0f / (S * S);
      e.GT = G; e.ldgt = pl.ldg; e.gstride = (int64_t)XW_M * qg.pair; e.blocked_nb = blocked ? (int)NB : 0;
      e.r_part = r_part; e.ldr = pl.Cc;
      e.gt_hint = (hints_k3a & 2) ? 2 : 0;
      e.whole_slice_targets = g_target_patch.load(std::memory_order_relaxed) ? 0 : 1;
      e.defer_targets = g_target_patch.load(std::memory_order_relaxed) == 2 ? 1 : 0;
#ifdef B200F_PROBES
      e.ablate = g_k3a_ablate.load(std::memory_order_relaxed);
#endif
    };
    const int k3a_whint = (hints & 1) ? 1 : 0;
    const bool k3a_rev = g_k3a_reverse.load(std::memory_order_relaxed) != 0;
    { B200F_NVTX("K3a logit gradient (recompute + G^T)");
    stage_event(EV_K3A, false, st);
    if (g_k3a_tma_store.load(std::memory_order_relaxed) != 0 && k3a_mode == 1 && (pl.ldg % 8) == 0 && !blocked) {
      XwBwdGTS::Params es{}; fill(es);
      rc = make_tmap(&es.tm_gt, G, B, cnt, pl.ldg, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B); if (rc) return rc;
      rc = (qg.pair == 2) ? launch_xw<2, XW_SWAP, XwBwdGTS>(tx_k, tw_k, qg, B, cnt, D, es, st, "umma K3a logit-grad (cta pair, TMA stores)", FMT_F16, k3a_rev, wc, wrb, true, k3a_whint)
                          : launch_xw<1, XW_SWAP, XwBwdGTS>(tx_k, tw_k, qg, B, cnt, D, es, st, "umma K3a logit-grad (TMA stores)", FMT_F16, k3a_rev, wc, wrb, true, k3a_whint);
    } else if (k3a_mode == 4) {
      XwBwdGT4::Params e4{}; fill(e4);
      rc = (qg.pair == 2) ? launch_xw<2, XW_SWAP, XwBwdGT4>(tx_k, tw_k, qg, B, cnt, D, e4, st, "umma K3a logit-grad (cta pair, 16 warps per tile)", FMT_F16, k3a_rev, wc, wrb, true, k3a_whint)
                          : launch_xw<1, XW_SWAP, XwBwdGT4>(tx_k, tw_k, qg, B, cnt, D, e4, st, "umma K3a logit-grad (16 warps per tile)", FMT_F16, k3a_rev, wc, wrb, true, k3a_whint);
    } else if (k3a_mode == 2) {
      XwBwdGT2::Params e2{}; fill(e2);
      rc = (qg.pair == 2) ? launch_xw<2, XW_SWAP, XwBwdGT2>(tx_k, tw_k, qg, B, cnt, D, e2, st, "umma K3a logit-grad (cta pair, 2 epilogue groups)", FMT_F16, k3a_rev, wc, wrb, true, k3a_whint)
                          : launch_xw<1, XW_SWAP, XwBwdGT2>(tx_k, tw_k, qg, B, cnt, D, e2, st, "umma K3a logit-grad (2 epilogue groups)", FMT_F16, k3a_rev, wc, wrb, true, k3a_whint);
    } else {
      XwBwdGT::Params eg{}; fill(eg);
      rc = (qg.pair == 2) ? launch_xw<2, XW_SWAP, XwBwdGT>(tx_k, tw_k, qg, B, cnt, D, eg, st, "umma K3a logit-grad (cta pair)", FMT_F16, k3a_rev, wc, wrb, true, k3a_whint)
                          : launch_xw<1, XW_SWAP, XwBwdGT>(tx_k, tw_k, qg, B, cnt, D, eg, st, "umma K3a logit-grad", FMT_F16, k3a_rev, wc, wrb, true, k3a_whint);
    }
    stage_event(EV_K3A, true, st); }
    return rc; };
    // --- K3b: dW[c0 + c, d] = inv_nw_c (sum_b G^T[c, b] x_hat[b, d] - w_hat[c, d] r_c), class-major (the thread owns a
    //     class row), normalise-backward fused; its coefficients { inv_nw_c / (S g_scale), r_c } are formed in the epilogue
    //     from K3a's partials (CoefSrc)
    const CoefSrc coef{r_part, qg.m_groups * (k3a_mode == 4 ? 4 : 2), pl.Cc, inv_nw, grad4, S};
    auto run_k3b = [&]() -> int {
    int rc = B200F_OK;
    int n_sq_used = 0;
    { B200F_NVTX("K3b dW = G^T x_hat (+ normalise-backward of W)");
    stage_event(EV_K3B, false, st);
    if (pl.fused_dw) {                                      // x_hat^T resident, G^T rows streamed
      CUtensorMap tg_k;
      rc = tmap_kmajor(&tg_k, G, cnt, B, pl.ldg, XW_WROWS); if (rc) return rc;
      const XwPlan qw = xw_plan(D, cnt, qg.pair, 0, phase == HEAD_BWD_PART_K3B ? cluster_limit : 0);
      const bool k3b_rev = g_k3b_reverse.load(std::memory_order_relaxed) != 0;   // read K3a's freshest G^T rows first
      const int gt_lhint = (hints & 32) ? 2 : 0;
      // x_hat^T (K1's output) is older than K3a: the producers load it while K3a drains ("early" = 2: then wait, then G^T)
      const int k3b_early = (g_early.load(std::memory_order_relaxed) >= 2) ? 2 : 0;
      if (g_k3b_tma_store.load(std::memory_order_relaxed) != 0 && (D % 4) == 0) {
        XwDwTS::Params ew{};
        rc = make_tmap(&ew.tm_wh, wc, D, cnt, D, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B); if (rc) return rc;
        // the map covers THIS launch's class rows only: a warp's rows beyond the chunk are clipped, not written as zeros
        rc = make_tmap_f32(&ew.tm_dw, dw + c0 * (int64_t)D, D, cnt, D, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B); if (rc) return rc;
        ew.coef = coef; ew.dw = dw; ew.c0 = c0; ew.ld = D; ew.sq_part = sq_part;
        ew.dw_hint = (hints & 4) ? 1 : 0; ew.wh_hint = (hints & 8) ? 1 : 0;
        n_sq_used = qw.items * qw.pair * 1 * XW_EPI_WARPS;
#ifdef B200F_PROBES
        ew.ablate = g_k3b_ablate.load(std::memory_order_relaxed);
#endif
        // two ring stages for G^T: its tiles are pulled into L2 two ahead so that the ring's loads are L2 hits
        rc = (qw.pair == 2) ? launch_xw<2, XW_SWAP_MK, XwDwTS>(tx_mn, tg_k, qw, D, cnt, (int)B, ew, st, "umma K3b dW class-major, TMA stores (cta pair)", FMT_F16, k3b_rev, G, pl.ldg * 2, false, gt_lhint, nullptr, 2)
                            : launch_xw<1, XW_SWAP_MK, XwDwTS>(tx_mn, tg_k, qw, D, cnt, (int)B, ew, st, "umma K3b dW class-major, TMA stores", FMT_F16, k3b_rev, G, pl.ldg * 2, false, gt_lhint, nullptr, 2);
      } else if (g_k3b_groups.load(std::memory_order_relaxed) == 2) {
        XwDwT2::Params ew{};
        rc = make_tmap(&ew.tm_wh, wc, D, cnt, D, 16, 32, CU_TENSOR_MAP_SWIZZLE_NONE); if (rc) return rc;
        ew.coef = coef; ew.dw = dw; ew.c0 = c0; ew.ld = D; ew.sq_part = sq_part;
        ew.dw_hint = (hints & 4) ? 1 : 0; ew.wh_hint = (hints & 8) ? 1 : 0;
        n_sq_used = qw.items * qw.pair * 2 * XW_EPI_WARPS;
#ifdef B200F_PROBES
        ew.ablate = g_k3b_ablate.load(std::memory_order_relaxed);
#endif
        rc = (qw.pair == 2) ? launch_xw<2, XW_SWAP_MK, XwDwT2>(tx_mn, tg_k, qw, D, cnt, (int)B, ew, st, "umma K3b dW class-major (cta pair, 2 epilogue groups)", FMT_F16, k3b_rev, G, pl.ldg * 2, false, gt_lhint)
                            : launch_xw<1, XW_SWAP_MK, XwDwT2>(tx_mn, tg_k, qw, D, cnt, (int)B, ew, st, "umma K3b dW class-major (2 epilogue groups)", FMT_F16, k3b_rev, G, pl.ldg * 2, false, gt_lhint);
      } else {
        XwDwT::Params ew{};
        rc = make_tmap(&ew.tm_wh, wc, D, cnt, D, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B); if (rc) return rc;
        ew.coef = coef; ew.dw = dw; ew.c0 = c0; ew.ld = D; ew.sq_part = sq_part;
        ew.dw_hint = (hints & 4) ? 1 : 0; ew.wh_hint = (hints & 8) ? 1 : 0;
        n_sq_used = qw.items * qw.pair * 1 * XW_EPI_WARPS;
#ifdef B200F_PROBES
        ew.ablate = g_k3b_ablate.load(std::memory_order_relaxed);
#endif
        rc = (qw.pair == 2) ? launch_xw<2, XW_SWAP_MK, XwDwT>(tx_mn, tg_k, qw, D, cnt, (int)B, ew, st, "umma K3b dW class-major (cta pair)", FMT_F16, k3b_rev, G, pl.ldg * 2, k3b_early, gt_lhint)
                            : launch_xw<1, XW_SWAP_MK, XwDwT>(tx_mn, tg_k, qw, D, cnt, (int)B, ew, st, "umma K3b dW class-major", FMT_F16, k3b_rev, G, pl.ldg * 2, k3b_early, gt_lhint);
      }
      if (rc) return rc;
    } else {
      // batch > 512: x_hat^T cannot stay resident, both operands stream through the generic core:
      // acc[m, :] = sum_b G^T[m, b] x_hat[b, :] with the class rows on the M side, so the epilogue thread owns a class
      // row and finishes the normalise-backward in place (EpiDwNorm) -- no separate pass over dW.
      // (An earlier TRANSPOSED fused epilogue on this core measured 2.3x slower at B = 4096: per-element global loads.)
      CUtensorMap tg_km;
      if (blocked) rc = make_tmap_gt_blocked(&tg_km, G, cnt, NB, BLOCK_M / 32);
      else rc = tmap_kmajor(&tg_km, G, cnt, B, pl.ldg, BLOCK_M);
      if (rc) return rc;
      GemmParams pw = gemm_params((int)cnt, D, (int)B, 1, false, true, FMT_F16, FMT_F16, gpair);
      pw.a_blocked = blocked ? 1 : 0;
      pw.n_fastest = g_dw_n_fastest.load(std::memory_order_relaxed) ? 1 : 0;
      EpiDwNorm::Params ew{dw, (int64_t)D, c0, coef, static_cast<const __half*>(wh), sq_part};
      n_sq_used = (int)(ceil_div(cnt, (int64_t)128) * pw.n_tiles * 4);
      rc = (gpair == 2) ? launch_gemm<2, false, true, EpiDwNorm>(tg_km, tx_mn, pw, ew, st, "umma K3b dW (streamed, cta pair)")
                        : launch_gemm<1, false, true, EpiDwNorm>(tg_km, tx_mn, pw, ew, st, "umma K3b dW (streamed)");
      if (rc) return rc;
    }
    stage_event(EV_K3B, true, st); }
    if (sq_part != nullptr) {                               // fold this chunk's partials into the caller's word
      if ((size_t)n_sq_used > pl.n_sq) return fail(B200F_ERR_WORKSPACE, "umma head_bwd: dW^2 partials exceed the plan");
      launch_pdl(fold_partials_kernel, dim3(1), dim3(256), 0, st, (const float*)sq_part, n_sq_used, sq_out, chunk_no > 0 ? 1 : 0);
      B200F_LAUNCH_OK("fold_partials_kernel");
    }
    return B200F_OK; };
    // --- K3c: dx_hat partials = G[:, k-range] w_hat[c0 + k-range, :]   (A = G read MN-major from G^T)
    auto run_k3c = [&]() -> int {
    CUtensorMap tg_mn, tw_mn;
    int rc = blocked ? make_tmap_gt_blocked(&tg_mn, G, cnt, NB, BLOCK_K / 32) : tmap_mnmajor(&tg_mn, G, B, cnt, pl.ldg);
    if (rc) return rc;
    rc = tmap_mnmajor(&tw_mn, wc, D, cnt, D); if (rc) return rc;
    // the dx part on a bounded number of clusters: as many K splits as those clusters carry output tiles
    const int k3c_limit = (phase == HEAD_BWD_PART_K3C) ? cluster_limit : 0;
    int dx_splits = pl.dx_splits;
    if (k3c_limit > 0) {
      const int out_tiles = (int)ceil_div(B, (int64_t)BLOCK_M * gpair) * (int)ceil_div((int64_t)D, (int64_t)BLOCK_N);
      dx_splits = k3c_limit / out_tiles;
      if (dx_splits < 1) dx_splits = 1;
      if (dx_splits > pl.dx_splits) dx_splits = pl.dx_splits;
    }
    GemmParams px = gemm_params((int)B, D, (int)cnt, dx_splits, true, true, FMT_F16, FMT_F16, gpair);
    // K3c reads G^T (K3a) and w_hat, writes dxpart: nothing of K3b's -- behind K3b it need not wait for it; directly behind
    // K3a (phases 1 / 2) it does
    px.early = (phase == 0) ? g_early.load(std::memory_order_relaxed) : 0;
    px.a_blocked = blocked ? 1 : 0;
    if (k3c_limit > 0 && gpair == 2 && pl.fused_dw && g_k3c_follow.load(std::memory_order_relaxed) != 0) {
      // beside the dW part, which runs on the rest of the chip: walk the class rows in ITS order (GemmParams::follow_*)
      const int rest = xw_max_clusters(2) - k3c_limit;
      if (rest > 0) {
        const XwPlan qw = xw_plan(D, cnt, 2, 0, rest);
        if (qw.n_tiles >= px.k_splits && qw.n_chunks >= 1) { px.follow_chunks = qw.n_chunks; px.follow_tiles = qw.n_tiles; }
        px.follow_rev = g_k3b_reverse.load(std::memory_order_relaxed) != 0 ? 1 : 0;
      }
    }
    // flat split reduction behind it (not the fused dL/dx tail of the single-chunk, single-shard step): stream-K allowed
    const bool fused_tail = hdx != nullptr && hdx->dx != nullptr && pl.n_chunks == 1;
    const int k3c_climit = (gpair == 2) ? k3c_limit : k3c_limit * 2;
    const int n_slots = fused_tail ? px.k_splits : gemm_stream_k(px, gpair, k3c_climit, pl.dx_slots);
    StreamGeom sg{};
    if (px.stream_k) {
      sg.on = 1; sg.D = D; sg.tile_m = BLOCK_M * gpair; sg.tile_n = BLOCK_N; sg.m_tiles = px.m_tiles; sg.k_units = px.k_units;
      sg.n_clusters = gemm_clusters(gpair, k3c_climit);
      sg.total = (long long)px.m_tiles * px.n_tiles * px.k_units;
    }
    EpiStore::Params ex{dxpart, (int64_t)D, B * (int64_t)D, 0, 1.0f, nullptr};
    stage_event(EV_K3C, false, st);
    rc = (gpair == 2) ? launch_gemm<2, true, true, EpiStore>(tg_mn, tw_mn, px, ex, st, "umma K3c dX (cta pair)", k3c_limit)
                      : launch_gemm<1, true, true, EpiStore>(tg_mn, tw_mn, px, ex, st, "umma K3c dX", k3c_limit * 2);
    stage_event(EV_K3C, true, st);
    if (rc) return rc;
    const int64_t n = B * (int64_t)D;
    if (hdx != nullptr && hdx->dx != nullptr && pl.n_chunks == 1) {
      // single chunk, single shard: the split reduction finishes dL/dx itself (normalise-backward of the row, and the
      // bf16 copy autograd would make), no round trip of dx_hat through HBM and two launches fewer
      const unsigned grid = (unsigned)ceil_div(B, rowops::WARPS_PER_BLOCK);
      const dim3 blk(rowops::WARPS_PER_BLOCK * 32);
      __nv_bfloat16* lowp = static_cast<__nv_bfloat16*>(hdx->dx_bf16);
      if (hdx->x_raw != nullptr && hdx->x_raw_dtype == B200F_BF16)
        launch_pdl(rowops::reduce_splits_normbwd_kernel<__nv_bfloat16, false>, dim3(grid), blk, 0, st, dxpart, px.k_splits, B, D, 1.0f / S,
                   grad4 + 3, static_cast<const __nv_bfloat16*>(hdx->x_raw), 1.0f, hdx->inv_nx, dxhat, hdx->dx, lowp);
      else if (hdx->x_raw != nullptr && hdx->x_raw_dtype == B200F_F32)
        launch_pdl(rowops::reduce_splits_normbwd_kernel<float, false>, dim3(grid), blk, 0, st, dxpart, px.k_splits, B, D, 1.0f / S,
                   grad4 + 3, static_cast<const float*>(hdx->x_raw), 1.0f, hdx->inv_nx, dxhat, hdx->dx, lowp);
      else
        launch_pdl(rowops::reduce_splits_normbwd_kernel<__half, true>, dim3(grid), blk, 0, st, dxpart, px.k_splits, B, D, 1.0f / S,
                   grad4 + 3, static_cast<const __half*>(xh), S, hdx->inv_nx, dxhat, hdx->dx, lowp);
      B200F_LAUNCH_OK("umma reduce_splits_normbwd_kernel");
    } else {
      launch_pdl(reduce_splits_kernel, dim3((unsigned)ceil_div(n / 4, 256)), dim3(256), 0, st, dxpart, n_slots, n, dxhat, chunk_no > 0,
                                                                          1.0f / S, grad4 + 3, sg);
      B200F_LAUNCH_OK("umma reduce_splits_kernel");
      if (last_chunk && hdx != nullptr && hdx->dx != nullptr) {
        rc = head_dx_finish(xh, S, hdx, dxhat, B, D, st); if (rc) return rc;
      }
    }
    return B200F_OK; };
    if (phase == 0) {
      rc = run_k3a(); if (rc) return rc;
      rc = run_k3b(); if (rc) return rc;
      rc = run_k3c(); if (rc) return rc;
    } else if (phase == 1) {
      rc = run_k3a(); if (rc) return rc;
      rc = run_k3c(); if (rc) return rc;
      if (!last_chunk) { rc = run_k3b(); if (rc) return rc; }
    } else if (phase == HEAD_BWD_PART_K3A) {
      rc = run_k3a(); if (rc) return rc;
    } else if (phase == HEAD_BWD_PART_K3C) {
      rc = run_k3c(); if (rc) return rc;
    } else {                                                  // phase 2, or the dW part
      rc = run_k3b(); if (rc) return rc;
    }
  }
  return B200F_OK;
}

// ---- K4 on tensor cores: host side ------------------------------------------------------------------
// Sample pre-pass of the gallery scan.  One row group (Q <= 128 * PAIR): 128 chunks of one tile -- every CTA scans one
// tile of the first 16 k (32 k) rows.  Several row groups: 4 chunks of 8 tiles, so that a work item amortises its
// x load over 8 tiles (one-tile items made the pre-pass cost a quarter of the scan at Q = 8192).
constexpr int GALLERY_SAMPLE_MAX_LISTS = 32 * GALLERY_TAU_LISTS_PER_LANE;
// tunable "gallery_compact": 1 (default) = min-only sample pre-pass, compact candidate arrays, block select that re-scores all
// winners in one round; 0 = the padded-list pre-pass / select of round 1.  Per 128-query call against 1 M x 512 (ncu launch
// lists, gpurun_out/r02u / r02v): pre-pass 24.2 + 7.8 -> 13.4 + 6.0 us, main scan 162.6 -> 158.0 us, select 40.5 -> see
// profiles/.  Both selects add the exact re-score in the same order: same score bits.
static std::atomic<int> g_gallery_compact{1};
struct GalleryScanPlan {
  XwPlan q, qs;                 // main scan / sample pre-pass
  int KT, n_lists;
  int64_t n_sample;             // gallery rows of the sample (0: no pre-pass)
  bool compact;                 // one row group with a sample bound: min-only pre-pass, compact candidates, warp select
  size_t off_q16, off_ckey, off_cidx, off_skey, off_sidx, off_tau, off_qbad, off_cnt, total;
};

static GalleryScanPlan gallery_scan_plan(int64_t Q, int64_t N, int D, int k) {
  GalleryScanPlan g{};
  g.q = xw_plan(Q, N, Q > XW_M ? 2 : 1);                 // one CTA holds up to 128 queries; pairs above that
  g.KT = k <= 1 ? 8 : (k <= 6 ? 16 : 32);
  g.n_lists = g.q.n_chunks * 2;
  size_t off = 0;
  const size_t per_q = (size_t)g.n_lists * g.KT;            // candidate slots per query: padded lists, or the compact array
  g.off_q16 = off;  off += align_up(2 * (size_t)Q * D, 1024);
  g.off_ckey = off; off += align_up(sizeof(float) * (size_t)Q * per_q, 256);
  g.off_cidx = off; off += align_up(sizeof(int32_t) * (size_t)Q * per_q, 256);
  // sample pre-pass over the first rows of the gallery (skipped when that is a quarter of it or more)
  const int s_chunks = (g.q.m_groups == 1) ? GALLERY_SAMPLE_MAX_LISTS / 2 : 4;
  const int s_tiles = (g.q.m_groups == 1) ? 1 : 8;
  const int64_t ns = (int64_t)s_chunks * s_tiles * XW_WROWS * g.q.pair;
  g.n_sample = (N >= 4 * ns) ? ns : 0;
  g.qs = xw_plan(Q, g.n_sample > 0 ? g.n_sample : 1, g.q.pair, s_chunks);
  const size_t sl = (size_t)Q * GALLERY_SAMPLE_MAX_LISTS * g.KT;
  g.off_skey = off; off += align_up(sizeof(float) * sl, 256);
  g.off_sidx = off; off += align_up(sizeof(int32_t) * sl, 256);
  g.off_tau = off;  off += align_up(sizeof(float) * (size_t)Q, 256);
  g.off_qbad = off; off += align_up((size_t)Q, 256);
  g.off_cnt = off;  off += align_up(sizeof(int32_t) * (size_t)Q, 256);
  // compact mode: the sample yields n_chunks * 2 * GALLERY_MIN_SUB minima per query (1024 for one row group: one tile per
  // CTA; 32 for several: 4 chunks of 8 tiles) -- at least 2 KT of them for the KT-th smallest to be a bound worth having
  const int n_min = g.qs.n_chunks * 2 * GALLERY_MIN_SUB;
  g.compact = g_gallery_compact.load(std::memory_order_relaxed) != 0 && g.n_sample > 0 &&
              n_min >= 2 * g.KT && n_min <= 32 * GALLERY_TAU_MIN_PER_LANE;
  g.total = off + 1024;
  return g;
}

bool gallery_tc_supported(int D) { return device_is_sm100() && D % 8 == 0 && D <= XW_MAX_KB * XW_K; }

size_t gallery_scan_workspace(int64_t Q, int64_t N, int D, int k) { return gallery_scan_plan(Q, N, D, k).total; }

// rows of 8-element chunks at 16-byte aligned addresses: the vector form of the prepare kernel applies
static bool prepare_vec_ok(const void* in, const void* out, int D) {
  return D % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}

int gallery_prepare(const void* g, int dtype, int64_t N, int D, int metric, int fmt, void* g16, float* bias, cudaStream_t st) {
  if (D > 512) return fail(B200F_ERR_UNSUPPORTED, "gallery_prepare: D <= 512");
  if (bias) B200F_CUDA_OK(cudaMemsetAsync(bias + N, 0, 2 * sizeof(float), st));
  const unsigned grid = (unsigned)ceil_div(N, 8);
  const bool vec = prepare_vec_ok(g, g16, D);
  if (dtype == B200F_F32)
    launch_pdl(vec ? gallery_prepare_vec8_kernel<float> : gallery_prepare_kernel<float>, dim3(grid), dim3(256), 0, st,
               static_cast<const float*>(g), N, D, metric, fmt, static_cast<uint16_t*>(g16), bias, (uint8_t*)nullptr);
  else
    launch_pdl(vec ? gallery_prepare_vec8_kernel<__nv_bfloat16> : gallery_prepare_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st,
               static_cast<const __nv_bfloat16*>(g), N, D, metric, fmt, static_cast<uint16_t*>(g16), bias, (uint8_t*)nullptr);
  B200F_LAUNCH_OK("gallery_prepare_kernel");
  return B200F_OK;
}

template <int KT>
static int gallery_scan_kt(const GalleryScanPlan& gp, const void* g16, const CUtensorMap& tx, const CUtensorMap& tw, const float* q,
                           const float* g, const float* bias, const float* q_inv, const float* g_inv, int64_t Q, int64_t N,
                           int64_t index_offset, int D, int k, int metric, int fmt, float thresh, int64_t* idx, float* score,
                           uint8_t* accept, uint8_t* redo, int32_t* redo_count, float* ckey, int32_t* cidx, float* skey,
                           int32_t* sidx, float* tau0, const uint8_t* qbad, int32_t* cnt, cudaStream_t st) {
  const uint32_t mfmt = fmt == B200F_OPERAND_FP16 ? FMT_F16 : FMT_BF16;
  typename XwTopK<KT>::Params ep{};
  ep.bias = (metric == B200F_METRIC_COS) ? nullptr : bias;
  ep.mult = (metric == B200F_METRIC_COS) ? -1.0f : -2.0f;
  if (gp.compact) {
    // sample pre-pass, min-only: every CTA scans one tile of the first n_sample rows and keeps the smallest key per
    // (CTA, column half); tau0[q] = KT-th smallest of those minima; the same kernel zeroes the candidate counters
    XwMinKey::Params es{ep.bias, ep.mult, skey, gp.qs.n_chunks * 2};
    const int n_min = gp.qs.n_chunks * 2 * GALLERY_MIN_SUB;
    int rcs = (gp.qs.pair == 2) ? launch_xw<2, XW_KK, XwMinKey>(tx, tw, gp.qs, Q, gp.n_sample, D, es, st, "umma K4 gallery sample minima (cta pair)", mfmt)
                                : launch_xw<1, XW_KK, XwMinKey>(tx, tw, gp.qs, Q, gp.n_sample, D, es, st, "umma K4 gallery sample minima", mfmt);
    if (rcs) return rcs;
    const dim3 tgrid((unsigned)ceil_div(Q, 4));
    if (n_min <= 32) launch_pdl(gallery_tau_min_kernel<KT, 1>, tgrid, dim3(128), 0, st, (const float*)skey, n_min, Q, tau0, cnt);
    else if (n_min <= 256) launch_pdl(gallery_tau_min_kernel<KT, 8>, tgrid, dim3(128), 0, st, (const float*)skey, n_min, Q, tau0, cnt);
    else launch_pdl(gallery_tau_min_kernel<KT, 32>, tgrid, dim3(128), 0, st, (const float*)skey, n_min, Q, tau0, cnt);
    B200F_LAUNCH_OK("gallery_tau_min_kernel");
    ep.tau0 = tau0; ep.cnt = cnt; ep.cap = gp.n_lists * KT;
  } else if (gp.n_sample > 0) {
    // sample pre-pass: the same scan over the first n_sample rows with few long chunks, then tau0[q] = KT-th best key
    typename XwTopK<KT>::Params es = ep;
    es.cand_key = skey; es.cand_idx = sidx; es.n_lists = gp.qs.n_chunks * 2; es.tau0 = nullptr; es.cnt = nullptr;
    int rcs = (gp.qs.pair == 2) ? launch_xw<2, XW_KK, XwTopK<KT>>(tx, tw, gp.qs, Q, gp.n_sample, D, es, st, "umma K4 gallery sample scan (cta pair)", mfmt)
                                : launch_xw<1, XW_KK, XwTopK<KT>>(tx, tw, gp.qs, Q, gp.n_sample, D, es, st, "umma K4 gallery sample scan", mfmt);
    if (rcs) return rcs;
    launch_pdl(gallery_tau_kernel<KT>, dim3((unsigned)ceil_div(Q, 4)), dim3(128), 0, st, skey, sidx, es.n_lists, Q, tau0);
    B200F_LAUNCH_OK("gallery_tau_kernel");
    ep.tau0 = tau0;
  }
  ep.cand_key = ckey; ep.cand_idx = cidx; ep.n_lists = gp.n_lists;
  // Both operands of the main scan are older than its predecessor (q16: the query prepare, two or three kernels back; g16:
  // the gallery): with a sample bound in front its TMA and MMA warps start without waiting for the tau kernel -- the 128 KB
  // of queries per CTA and the first tiles load while that kernel runs -- and only the epilogue (tau0, the counters) waits.
  const bool scan_early = gp.n_sample > 0;
  int rc = (gp.q.pair == 2) ? launch_xw<2, XW_KK, XwTopK<KT>>(tx, tw, gp.q, Q, N, D, ep, st, "umma K4 gallery scan (cta pair)", mfmt, false, g16, (int64_t)D * 2, scan_early)
                            : launch_xw<1, XW_KK, XwTopK<KT>>(tx, tw, gp.q, Q, N, D, ep, st, "umma K4 gallery scan", mfmt, false, g16, (int64_t)D * 2, scan_early);
  if (rc) return rc;
  if (gp.compact) {
    const int cap = gp.n_lists * KT;
    launch_pdl(gallery_select_block_kernel<KT>, dim3((unsigned)Q), dim3(128), 0, st, (const float*)ckey, (const int32_t*)cidx,
               (const int32_t*)cnt, cap, q, g, q_inv, g_inv, (const float*)(bias ? bias + N : nullptr), qbad, Q, D, k, metric, fmt, thresh,
               index_offset, idx, score, accept, redo, redo_count);
    B200F_LAUNCH_OK("gallery_select_block_kernel");
    return B200F_OK;
  }
  const int n_cand = gp.n_lists * KT;
  const size_t smem = (size_t)n_cand * 16;                 // candidates + survivors, (key, idx) each
  auto kern = gallery_select_kernel<float, KT>;
  B200F_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 160 * 1024 ? 160 * 1024 : (smem < 1024 ? 1024 : smem))));
  if (smem > 160 * 1024) return fail(B200F_ERR_UNSUPPORTED, "gallery scan: candidate lists do not fit shared memory");
  launch_pdl(kern, dim3((unsigned)Q), dim3(128), smem, st, ckey, cidx, n_cand, q, g, q_inv, g_inv,
             (const float*)(bias ? bias + N : nullptr), qbad, Q, D, k, metric, fmt, thresh, index_offset, idx, score, accept, redo,
             redo_count);
  B200F_LAUNCH_OK("gallery_select_kernel");
  return B200F_OK;
}

// q, g: fp32 [Q,D], [N,D]; g16 / bias from gallery_prepare (same metric).  Fills idx / score / accept for every
// query and redo[Q] = 1 where exactness could not be proven (the caller re-runs those on the exact engine).
int gallery_scan_select(const float* q, const float* g, const void* g16, const float* bias, const float* q_inv,
                        const float* g_inv, int64_t Q, int64_t N, int64_t index_offset, int D, int k, int metric, int fmt,
                        float thresh, int64_t* idx, float* score, uint8_t* accept, uint8_t* redo, int32_t* redo_count,
                        char* ws, size_t ws_bytes, cudaStream_t st) {
  if (!gallery_tc_supported(D)) return fail(B200F_ERR_UNSUPPORTED, "gallery scan: needs sm_100, D %% 8 == 0, D <= 512");
  if (N >= ((int64_t)1 << 31) - 512 || Q >= ((int64_t)1 << 30)) return fail(B200F_ERR_UNSUPPORTED, "gallery scan: N < 2^31, Q < 2^30");
  const GalleryScanPlan gp = gallery_scan_plan(Q, N, D, k);
  if (ws_bytes < gp.total) return fail(B200F_ERR_WORKSPACE, "gallery scan: workspace too small");
  ws = ws_align(ws);
  uint16_t* q16 = reinterpret_cast<uint16_t*>(ws + gp.off_q16);
  float* ckey = reinterpret_cast<float*>(ws + gp.off_ckey);
  int32_t* cidx = reinterpret_cast<int32_t*>(ws + gp.off_cidx);
  float* skey = reinterpret_cast<float*>(ws + gp.off_skey);
  int32_t* sidx = reinterpret_cast<int32_t*>(ws + gp.off_sidx);
  float* tau0 = reinterpret_cast<float*>(ws + gp.off_tau);
  uint8_t* qbad = reinterpret_cast<uint8_t*>(ws + gp.off_qbad);
  int32_t* cnt = reinterpret_cast<int32_t*>(ws + gp.off_cnt);
  launch_pdl(prepare_vec_ok(q, q16, D) ? gallery_prepare_vec8_kernel<float> : gallery_prepare_kernel<float>, dim3((unsigned)ceil_div(Q, 8)),
             dim3(256), 0, st, q, Q, D, B200F_METRIC_L2EPS, fmt, q16, (float*)nullptr, qbad);
  B200F_LAUNCH_OK("gallery_prepare_kernel (queries)");
  CUtensorMap tx, tw;
  int rc = tmap_kmajor(&tx, q16, Q, D, D, XW_M); if (rc) return rc;
  rc = tmap_kmajor(&tw, g16, N, D, D, XW_WROWS); if (rc) return rc;
#define B200F_SCAN(KT) gallery_scan_kt<KT>(gp, g16, tx, tw, q, g, bias, q_inv, g_inv, Q, N, index_offset, D, k, metric, fmt, thresh, idx, \
                                           score, accept, redo, redo_count, ckey, cidx, skey, sidx, tau0, qbad, cnt, st)
  if (gp.KT == 8) return B200F_SCAN(8);
  if (gp.KT == 16) return B200F_SCAN(16);
  return B200F_SCAN(32);
#undef B200F_SCAN
}

}  // namespace umma
}  // namespace b200f

using namespace b200f;
using namespace b200f::umma;

extern "C" {

// Self-test / probe of the GEMM core: out[M,N] (fp32) = sum_k A(m,k) B(n,k) for every operand layout.
//   a_mn / b_mn: 0 = K-major (a is [M,K] row-major), 1 = MN-major (a is [K,M] row-major)
//   fmt: 0 = bf16 x bf16, 1 = fp16 (A) x bf16 (B) [unsupported by the hardware: faults], 2 = fp16 x fp16
int b200f_umma_selftest(const void* a, const void* b, float* out, int M, int N, int K, int a_mn, int b_mn,
                        int fmt, int k_splits, int a_lbo, int a_sbo, int a_kstep, int b_lbo, int b_sbo,
                        int b_kstep, void* stream) {
  if (!a || !b || !out || M <= 0 || N <= 0 || K <= 0) return fail(B200F_ERR_ARG, "umma_selftest: bad argument");
  if (!device_is_sm100()) return fail(B200F_ERR_UNSUPPORTED, "umma_selftest: device is not sm_100");
  const int pair = g_pair.load(std::memory_order_relaxed);  // the "pair" tunable selects cta_group::1 or ::2 here too
  CUtensorMap ta, tb;
  int rc = a_mn ? tmap_mnmajor(&ta, a, M, K, M) : tmap_kmajor(&ta, a, M, K, K, BLOCK_M);
  if (rc) return rc;
  rc = b_mn ? tmap_mnmajor(&tb, b, N, K, N) : tmap_kmajor(&tb, b, N, K, K, gemm_b_rows(pair));
  if (rc) return rc;
  GemmParams p = gemm_params(M, N, K, k_splits, a_mn != 0, b_mn != 0, fmt >= 1 ? FMT_F16 : FMT_BF16,
                             fmt == 2 ? FMT_F16 : FMT_BF16, pair);
  if (a_lbo >= 0) p.a_lbo = a_lbo;
  if (a_sbo >= 0) p.a_sbo = a_sbo;
  if (a_kstep >= 0) p.a_kstep = a_kstep;
  if (b_lbo >= 0) p.b_lbo = b_lbo;
  if (b_sbo >= 0) p.b_sbo = b_sbo;
  if (b_kstep >= 0) p.b_kstep = b_kstep;
  EpiStore::Params ep{out, (int64_t)N, (int64_t)M * N, 0, 1.0f, nullptr};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pair == 2) {
    if (!a_mn && !b_mn) return launch_gemm<2, false, false, EpiStore>(ta, tb, p, ep, st, "umma selftest KK (cta pair)");
    if (a_mn && b_mn) return launch_gemm<2, true, true, EpiStore>(ta, tb, p, ep, st, "umma selftest MM (cta pair)");
    if (!a_mn && b_mn) return launch_gemm<2, false, true, EpiStore>(ta, tb, p, ep, st, "umma selftest KM (cta pair)");
    return launch_gemm<2, true, false, EpiStore>(ta, tb, p, ep, st, "umma selftest MK (cta pair)");
  }
  if (!a_mn && !b_mn) return launch_gemm<1, false, false, EpiStore>(ta, tb, p, ep, st, "umma selftest KK");
  if (a_mn && b_mn) return launch_gemm<1, true, true, EpiStore>(ta, tb, p, ep, st, "umma selftest MM");
  if (!a_mn && b_mn) return launch_gemm<1, false, true, EpiStore>(ta, tb, p, ep, st, "umma selftest KM");
  return launch_gemm<1, true, false, EpiStore>(ta, tb, p, ep, st, "umma selftest MK");
}

// Self-test of the X-stationary kernel: out[B,C] (fp32) = x[B,D] . w[C,D]^T with fp16 operands, on single CTAs
// (pair = 1) or tcgen05 cta_group::2 CTA pairs (pair = 2).
int b200f_umma_xw_selftest(const void* x, const void* w, float* out, int B, int C, int D, int pair, void* stream) {
  if (!x || !w || !out || B <= 0 || C <= 0 || D <= 0 || (pair != 1 && pair != 2))
    return fail(B200F_ERR_ARG, "umma_xw_selftest: bad argument");
  if (!device_is_sm100()) return fail(B200F_ERR_UNSUPPORTED, "umma_xw_selftest: device is not sm_100");
  if (D % 8 != 0 || D > XW_MAX_KB * XW_K) return fail(B200F_ERR_UNSUPPORTED, "umma_xw_selftest: D %% 8 == 0 and D <= 512");
  CUtensorMap tx, tw;
  int rc = tmap_kmajor(&tx, x, B, D, D, XW_M); if (rc) return rc;
  rc = tmap_kmajor(&tw, w, C, D, D, XW_WROWS); if (rc) return rc;
  const XwPlan q = xw_plan(B, C, pair);
  XwStore::Params ep{out, (int64_t)C};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return (pair == 2) ? launch_xw<2, XW_KK, XwStore>(tx, tw, q, B, C, D, ep, st, "umma xw selftest (cta pair)")
                     : launch_xw<1, XW_KK, XwStore>(tx, tw, q, B, C, D, ep, st, "umma xw selftest");
}

// Pipeline probe (tools/xw_probe.py): same GEMM, but the epilogue only adds every accumulator of a row into
// rowsum[B] (must be zeroed by the caller): measures the TMA / tcgen05 / TMEM pipeline without epilogue work.
int b200f_umma_xw_probe(const void* x, const void* w, float* rowsum, int B, int C, int D, int pair, void* stream) {
  if (!x || !w || !rowsum || B <= 0 || C <= 0 || D <= 0 || (pair != 1 && pair != 2))
    return fail(B200F_ERR_ARG, "umma_xw_probe: bad argument");
  if (!device_is_sm100()) return fail(B200F_ERR_UNSUPPORTED, "umma_xw_probe: device is not sm_100");
  if (D % 8 != 0 || D > XW_MAX_KB * XW_K) return fail(B200F_ERR_UNSUPPORTED, "umma_xw_probe: D %% 8 == 0 and D <= 512");
  CUtensorMap tx, tw;
  int rc = tmap_kmajor(&tx, x, B, D, D, XW_M); if (rc) return rc;
  rc = tmap_kmajor(&tw, w, C, D, D, XW_WROWS); if (rc) return rc;
  const XwPlan q = xw_plan(B, C, pair);
  XwNull::Params ep{rowsum};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return (pair == 2) ? launch_xw<2, XW_KK, XwNull>(tx, tw, q, B, C, D, ep, st, "umma xw probe (cta pair)")
                     : launch_xw<1, XW_KK, XwNull>(tx, tw, q, B, C, D, ep, st, "umma xw probe");
}

// Selects single-CTA (1) or CTA-pair (2, default) execution of K2 / K3a; returns the previous setting.
int b200f_umma_set_pair(int pair) {
  if (pair != 1 && pair != 2) return g_pair.load();
  return g_pair.exchange(pair);
}

// Tunables (tests / bench sweeps): "pair" (1 | 2) and "g_chunk_mb" (MB of logit-gradient buffer per class chunk).
// Returns the previous value, or -1 for an unknown name.  Workspaces must be re-queried after a change.
int b200f_set_tunable(const char* name, int value) {
  if (!name) return -1;
  const std::string n(name);
  if (n == "pair") return b200f_umma_set_pair(value);
  if (n == "pdl") { const int old_v = pdl_enabled() ? 1 : 0; if (value == 0 || value == 1) pdl_set(value != 0); return old_v; }
  if (n == "stage_events") { if (value != 0 && value != 1) return g_stage_events.load(); return g_stage_events.exchange(value); }
  if (n == "k2_prep") { if (value < 0 || value > 2) return g_k2_prep.load(); return g_k2_prep.exchange(value); }
  if (n == "target_patch") { if (value < 0 || value > 2) return g_target_patch.load(); return g_target_patch.exchange(value); }
  if (n == "x_whole") { if (value != 0 && value != 1) return g_x_whole.load(); return g_x_whole.exchange(value); }
  if (n == "k1_hints") { if (value < 0 || value > 3) return k1_hints(); return k1_hints_set(value); }
  if (n == "l2_hints") { if (value < 0) return g_l2_hints.load(); return g_l2_hints.exchange(value); }
  if (n == "k2_groups") { if (value != 1 && value != 2) return g_k2_groups.load(); return g_k2_groups.exchange(value); }
  if (n == "k3b_groups") { if (value != 1 && value != 2) return g_k3b_groups.load(); return g_k3b_groups.exchange(value); }
  if (n == "gallery_compact") { if (value != 0 && value != 1) return g_gallery_compact.load(); return g_gallery_compact.exchange(value); }
  if (n == "early") { if (value < 0 || value > 2) return g_early.load(); return g_early.exchange(value); }
  if (n == "epi_groups") { if (value != 1 && value != 2 && value != 4) return g_epi_groups.load(); return g_epi_groups.exchange(value); }
#ifdef B200F_PROBES
  // probes that make the backward skip memory traffic (WRONG results): -DB200F_PROBES builds only (tools/)
  if (n == "k3a_ablate") { if (value < 0) return g_k3a_ablate.load(); return g_k3a_ablate.exchange(value); }
  if (n == "k3b_ablate") { if (value < 0) return g_k3b_ablate.load(); return g_k3b_ablate.exchange(value); }
#endif
  if (n == "k3b_tma_store") { if (value != 0 && value != 1) return g_k3b_tma_store.load(); return g_k3b_tma_store.exchange(value); }
  if (n == "stream_k") { if (value != 0 && value != 1) return g_stream_k.load(); return g_stream_k.exchange(value); }
  if (n == "gt_blocked") { if (value != 0 && value != 1) return g_gt_blocked.load(); return g_gt_blocked.exchange(value); }
  if (n == "k3a_tma_store") { if (value != 0 && value != 1) return g_k3a_tma_store.load(); return g_k3a_tma_store.exchange(value); }
  if (n == "k3a_reverse") { if (value != 0 && value != 1) return g_k3a_reverse.load(); return g_k3a_reverse.exchange(value); }
  if (n == "k3c_follow") { if (value != 0 && value != 1) return g_k3c_follow.load(); return g_k3c_follow.exchange(value); }
  if (n == "dw_n_fastest") { if (value != 0 && value != 1) return g_dw_n_fastest.load(); return g_dw_n_fastest.exchange(value); }
  if (n == "k3b_reverse") { if (value != 0 && value != 1) return g_k3b_reverse.load(); return g_k3b_reverse.exchange(value); }
  if (n == "xw_prefetch") { if (value < 0) return g_prefetch.load(); return g_prefetch.exchange(value); }
  if (n == "g_chunk_mb") { if (value < 1) return g_chunk_mb.load(); return g_chunk_mb.exchange(value); }
  return -1;
}

// Duration (ms) of the "k2" | "k3a" | "k3b" | "k3c" kernel(s) of the last head call this thread made while the
// "stage_events" tunable was 1 (summed over the call's class chunks).  Synchronises on the kernels' end events.
int b200f_stage_ms(const char* name, float* ms) {
  if (!name || !ms) return fail(B200F_ERR_ARG, "stage_ms: null argument");
  for (int i = 0; i < EV_COUNT; ++i) {
    if (std::string(name) != kEvNames[i]) continue;
    if (!g_ev.made || g_ev.n[i] == 0) return fail(B200F_ERR_ARG, "stage_ms: no '%s' kernel was recorded on this thread", name);
    float total = 0.f;
    for (int j = 0; j < g_ev.n[i]; ++j) {                     // one launch per class chunk of the call: summed
      float t = 0.f;
      B200F_CUDA_OK(cudaEventSynchronize(g_ev.end[i][j]));
      B200F_CUDA_OK(cudaEventElapsedTime(&t, g_ev.beg[i][j], g_ev.end[i][j]));
      total += t;
    }
    *ms = total;
    return B200F_OK;
  }
  return fail(B200F_ERR_ARG, "stage_ms: unknown stage '%s'", name);
}

// Reads (and optionally clears) the pipeline-timeout flag.  Synchronises: test / bench use only.
int b200f_umma_timeout_flag(int reset) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_umma_timeout_flag, sizeof(v)) != cudaSuccess) return -1;
  if (reset) { unsigned int z = 0; cudaMemcpyToSymbol(g_umma_timeout_flag, &z, sizeof(z)); }
  return (int)v;
}

}  // extern "C"
