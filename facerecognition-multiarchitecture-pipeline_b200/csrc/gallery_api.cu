// C ABI, part 3: the gallery match (K4) and the top-k merge.  Declared in include/b200face.h.
#include "common.cuh"
#include "gallery_simt.cuh"
#include "umma_api.cuh"

using namespace b200f;

static inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
static inline bool dtype_ok(int dt) { return dt == B200F_F32 || dt == B200F_BF16; }
static inline size_t elem_size(int dt) { return dt == B200F_F32 ? 4 : 2; }

// ---- gallery --------------------------------------------------------------------------------
struct GalleryPlan { int q_tiles, n_chunks, tiles_per_chunk, K; size_t off_idx, off_score, total; };

static GalleryPlan plan_gallery(int64_t Q, int64_t N, int k) {
  GalleryPlan pl{};
  pl.q_tiles = (int)ceil_div(Q, simt::BM);
  const int64_t n_tiles = ceil_div(N, simt::BN);
  int64_t want = ceil_div((int64_t)2 * num_sms(), pl.q_tiles);
  if (want > n_tiles) want = n_tiles;
  if (want < 1) want = 1;
  pl.tiles_per_chunk = (int)ceil_div(n_tiles, want);
  pl.n_chunks = (int)ceil_div(n_tiles, pl.tiles_per_chunk);
  pl.K = k <= 1 ? 1 : (k <= 4 ? 4 : (k <= 8 ? 8 : 16));
  size_t off = 0;
  pl.off_idx = off;   off += align_up(sizeof(int64_t) * (size_t)pl.n_chunks * Q * k, 256);
  pl.off_score = off; off += align_up(sizeof(float) * (size_t)pl.n_chunks * Q * k, 256);
  pl.total = off;
  return pl;
}


template <typename T, class Op, int K>
static int launch_gallery(const gallery::Params& p, const GalleryPlan& pl, cudaStream_t st) {
  auto kern = gallery::topk_kernel<T, Op, K>;
  const size_t dyn = gallery::dyn_smem_bytes();
  B200F_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  dim3 grid(pl.n_chunks, pl.q_tiles);
  // the gated fallback behind the tensor scan (only_rows) sits in a chain of programmatically serialised launches: its
  // launch latency then overlaps the select kernel instead of following it
  if (p.only_rows != nullptr) launch_pdl(kern, grid, dim3(simt::THREADS), dyn, st, p);
  else kern<<<grid, simt::THREADS, dyn, st>>>(p);
  B200F_LAUNCH_OK("gallery::topk_kernel");
  return B200F_OK;
}

template <typename T, class Op>
static int launch_gallery_k(const gallery::Params& p, const GalleryPlan& pl, cudaStream_t st) {
  switch (pl.K) {
    case 1: return launch_gallery<T, Op, 1>(p, pl, st);
    case 4: return launch_gallery<T, Op, 4>(p, pl, st);
    case 8: return launch_gallery<T, Op, 8>(p, pl, st);
    default: return launch_gallery<T, Op, 16>(p, pl, st);
  }
}

static int launch_merge(const int64_t* idx_all, const float* score_all, int P, int64_t Q, int k, int metric,
                        float thresh, int64_t* idx, float* score, uint8_t* accept, cudaStream_t st,
                        const uint8_t* only_rows = nullptr) {
  const unsigned grid = (unsigned)ceil_div(Q, 128);
  auto launch = [&](auto kern) {
    if (only_rows != nullptr) launch_pdl(kern, dim3(grid), dim3(128), 0, st, idx_all, score_all, P, Q, k, metric, thresh, idx, score, accept, only_rows);
    else kern<<<grid, 128, 0, st>>>(idx_all, score_all, P, Q, k, metric, thresh, idx, score, accept, only_rows);
  };
  if (k <= 1) launch(gallery::merge_kernel<1>);
  else if (k <= 4) launch(gallery::merge_kernel<4>);
  else if (k <= 8) launch(gallery::merge_kernel<8>);
  else launch(gallery::merge_kernel<16>);
  B200F_LAUNCH_OK("gallery::merge_kernel");
  return B200F_OK;
}

// exact engine: per-chunk top-k on the fp32 CUDA cores + merge (optionally only the rows flagged in only_rows)
static int run_exact(const void* q, const void* g, int dtype, const float* q_inv, const float* g_inv, int64_t Q, int64_t N,
                     int64_t index_offset, int D, int k, int metric, float thresh, int64_t* idx, float* score,
                     uint8_t* accept, char* ws, const GalleryPlan& pl, const uint8_t* only_rows, cudaStream_t st) {
  gallery::Params p{};
  p.q = q; p.g = g; p.q_inv = q_inv; p.g_inv = g_inv; p.Q = Q; p.N = N; p.index_offset = index_offset;
  p.D = D; p.k = k; p.metric = metric; p.tiles_per_chunk = pl.tiles_per_chunk;
  p.cand_idx = reinterpret_cast<int64_t*>(ws + pl.off_idx);
  p.cand_score = reinterpret_cast<float*>(ws + pl.off_score);
  p.only_rows = only_rows;
  int rc;
  if (dtype == B200F_F32) {
    p.vec_q = simt::vec_friendly<float>(q, D); p.vec_g = simt::vec_friendly<float>(g, D);
    rc = (metric == B200F_METRIC_COS) ? launch_gallery_k<float, simt::OpFma>(p, pl, st)
                                      : launch_gallery_k<float, simt::OpL2Eps>(p, pl, st);
  } else {
    p.vec_q = simt::vec_friendly<__nv_bfloat16>(q, D); p.vec_g = simt::vec_friendly<__nv_bfloat16>(g, D);
    rc = (metric == B200F_METRIC_COS) ? launch_gallery_k<__nv_bfloat16, simt::OpFma>(p, pl, st)
                                      : launch_gallery_k<__nv_bfloat16, simt::OpL2Eps>(p, pl, st);
  }
  if (rc) return rc;
  return launch_merge(p.cand_idx, p.cand_score, pl.n_chunks, Q, k, metric, thresh, idx, score, accept, st, only_rows);
}

extern "C" {

size_t b200f_gallery_workspace_bytes(int64_t Q, int64_t N_local, int D, int k, int dtype, int engine) {
  (void)D; (void)dtype; (void)engine;
  if (Q <= 0 || N_local <= 0 || k <= 0) return 256;
  return plan_gallery(Q, N_local, k).total;
}

int b200f_gallery_topk(const void* q, const void* g, int dtype, const float* q_inv, const float* g_inv,
                       int64_t Q, int64_t N_local, int64_t index_offset, int D, int k, int metric, float thresh,
                       int engine, int64_t* idx, float* score, uint8_t* accept, void* workspace,
                       size_t workspace_bytes, void* stream) {
  B200F_NVTX("b200f_gallery_topk");
  if (!dtype_ok(dtype)) return fail(B200F_ERR_ARG, "gallery_topk: bad dtype");
  if (Q < 0 || N_local < 0 || D <= 0) return fail(B200F_ERR_ARG, "gallery_topk: bad shape");
  if (k < 1 || k > 16) return fail(B200F_ERR_ARG, "gallery_topk: k=%d outside [1,16]", k);
  if (metric != B200F_METRIC_L2EPS && metric != B200F_METRIC_COS) return fail(B200F_ERR_ARG, "gallery_topk: bad metric");
  if (Q == 0) return B200F_OK;
  if (!q || !idx || !score) return fail(B200F_ERR_ARG, "gallery_topk: null pointer");
  cudaStream_t st = as_stream(stream);
  if (N_local == 0) {   // empty gallery: compare_faces returns ("Unknown", inf, None), src/app.py:51
    return launch_merge(nullptr, nullptr, 0, Q, k, metric, thresh, idx, score, accept, st);
  }
  if (!g) return fail(B200F_ERR_ARG, "gallery_topk: null gallery");
  (void)engine;   // this entry is the exact fp32 CUDA-core engine; the tensor engine is b200f_gallery_topk_tc
  const GalleryPlan pl = plan_gallery(Q, N_local, k);
  if (!workspace || workspace_bytes < pl.total)
    return fail(B200F_ERR_WORKSPACE, "gallery_topk: workspace %zu < %zu", workspace_bytes, pl.total);
  return run_exact(q, g, dtype, q_inv, g_inv, Q, N_local, index_offset, D, k, metric, thresh, idx, score, accept,
                   static_cast<char*>(workspace), pl, nullptr, st);
}

// ---- K4 on tensor cores ---------------------------------------------------------------------------------
int b200f_gallery_has_tc(int D) { return umma::gallery_tc_supported(D) ? 1 : 0; }

int b200f_gallery_prepare(const void* g, int dtype, int64_t N, int D, int metric, int operand_fmt, void* g16, float* bias,
                          void* stream) {
  B200F_NVTX("b200f_gallery_prepare");
  if (!dtype_ok(dtype)) return fail(B200F_ERR_ARG, "gallery_prepare: bad dtype");
  if (N < 0 || D <= 0) return fail(B200F_ERR_ARG, "gallery_prepare: bad shape");
  if (metric != B200F_METRIC_L2EPS && metric != B200F_METRIC_COS) return fail(B200F_ERR_ARG, "gallery_prepare: bad metric");
  if (N == 0) return B200F_OK;
  if (!g || !g16 || !bias) return fail(B200F_ERR_ARG, "gallery_prepare: null pointer");
  if (operand_fmt != B200F_OPERAND_BF16 && operand_fmt != B200F_OPERAND_FP16) return fail(B200F_ERR_ARG, "gallery_prepare: bad operand format");
  return umma::gallery_prepare(g, dtype, N, D, metric, operand_fmt, g16, bias, as_stream(stream));
}

size_t b200f_gallery_tc_workspace_bytes(int64_t Q, int64_t N_local, int D, int k) {
  if (Q <= 0 || N_local <= 0 || k <= 0) return 256;
  return align_up(umma::gallery_scan_workspace(Q, N_local, D, k), 256) + align_up((size_t)Q, 256) +
         plan_gallery(Q, N_local, k).total + 256;
}

int b200f_gallery_topk_tc(const void* q, const void* g, const void* g16, const float* bias, const float* q_inv,
                          const float* g_inv, int64_t Q, int64_t N_local, int64_t index_offset, int D, int k, int metric,
                          int operand_fmt, float thresh, int64_t* idx, float* score, uint8_t* accept, int32_t* redo_count,
                          void* workspace, size_t workspace_bytes, void* stream) {
  B200F_NVTX("b200f_gallery_topk_tc");
  if (Q < 0 || N_local <= 0 || D <= 0) return fail(B200F_ERR_ARG, "gallery_topk_tc: bad shape");
  if (k < 1 || k > 16) return fail(B200F_ERR_ARG, "gallery_topk_tc: k=%d outside [1,16]", k);
  if (metric != B200F_METRIC_L2EPS && metric != B200F_METRIC_COS) return fail(B200F_ERR_ARG, "gallery_topk_tc: bad metric");
  if (Q == 0) return B200F_OK;
  if (!q || !g || !g16 || !bias || !idx || !score) return fail(B200F_ERR_ARG, "gallery_topk_tc: null pointer");
  const size_t need = b200f_gallery_tc_workspace_bytes(Q, N_local, D, k);
  if (!workspace || workspace_bytes < need) return fail(B200F_ERR_WORKSPACE, "gallery_topk_tc: workspace %zu < %zu", workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  const size_t scan_bytes = align_up(umma::gallery_scan_workspace(Q, N_local, D, k), 256);
  uint8_t* redo = reinterpret_cast<uint8_t*>(ws + scan_bytes);
  char* exact_ws = ws + scan_bytes + align_up((size_t)Q, 256);
  int rc = umma::gallery_scan_select(static_cast<const float*>(q), static_cast<const float*>(g), g16, bias, q_inv, g_inv, Q,
                                     N_local, index_offset, D, k, metric, operand_fmt, thresh, idx, score, accept, redo, redo_count,
                                     ws, scan_bytes, st);
  if (rc) return rc;
  // queries whose top-k could not be proven exact: recomputed by the exact engine, gated on the device-side flags
  const GalleryPlan pl = plan_gallery(Q, N_local, k);
  return run_exact(q, g, B200F_F32, q_inv, g_inv, Q, N_local, index_offset, D, k, metric, thresh, idx, score, accept,
                   exact_ws, pl, redo, st);
}

int b200f_gallery_merge(const int64_t* idx_all, const float* score_all, int P, int64_t Q, int k, int metric,
                        float thresh, int64_t* idx, float* score, uint8_t* accept, void* stream) {
  B200F_NVTX("b200f_gallery_merge");
  if (P < 0 || Q < 0 || k < 1 || k > 16) return fail(B200F_ERR_ARG, "gallery_merge: bad shape");
  if (Q == 0) return B200F_OK;
  if ((P > 0 && (!idx_all || !score_all)) || !idx || !score) return fail(B200F_ERR_ARG, "gallery_merge: null pointer");
  return launch_merge(idx_all, score_all, P, Q, k, metric, thresh, idx, score, accept, as_stream(stream));
}

}  // extern "C"
