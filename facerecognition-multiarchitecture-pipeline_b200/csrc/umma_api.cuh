// tcgen05 / TMEM / TMA engine entry points (operand dtype B200F_F16N).  Defined in umma_head.cu.
#pragma once
#include "common.cuh"

namespace b200f {
namespace umma {

bool available();                                    // current device is sm_100
size_t head_workspace_bytes(int64_t B, int64_t C, int D);

// Optional tail of the forward (single shard): loss, lse, ||p-q||^2 and the hook scalars for an upstream gradient of 1
// come out of the statistics reduction's last block.
struct HeadFinal { float* lse; float* loss; float* pq_norm2; int hook_enabled; float max_grad_norm; int phase, epoch; float* out4; };
// Optional tail of the backward (single shard): dL/dx = normalise-backward of dx_hat, against the RAW input rows when
// given (x_hat = x * inv_nx exactly; else against K1's fp16 rows), plus its bf16 copy.
struct HeadDx { const void* x_raw; int x_raw_dtype; const float* inv_nx; float* dx; void* dx_bf16; };

// Optional head of the forward: K1 from the RAW rows in the same call.  x_raw (may be null: xh is already prepared) -> xh,
// inv_nx; w_raw -> wh, inv_nw -- inside K2 itself when the shape allows it (D = 512, 32-byte aligned rows; tunable
// "k2_prep"), else as a pass of its own in front of K2.  xh / wh are then OUTPUTS of head_fwd.
struct HeadPrep { const void* x_raw; int x_dtype; float* inv_nx; const void* w_raw; int w_dtype; float* inv_nw; float eps; };

int head_fwd(const void* xh, const void* wh, const int64_t* label, int64_t B, int64_t C, int64_t class_offset, int D,
             const b200f_head_cfg* cfg, float* row_stats, float* row_best, int64_t* row_argmax, float* cos_minmax,
             int32_t* nan_flag, const HeadFinal* fin, char* ws, size_t ws_bytes, cudaStream_t st,
             const HeadPrep* prep = nullptr);

int head_bwd(const void* xh, const void* wh, const float* inv_nw, const int64_t* label, const float* lse,
             const float* grad4, int64_t B, int64_t C, int64_t class_offset, int D, const b200f_head_cfg* cfg,
             float* dxhat, float* dw, const HeadDx* hdx, char* ws, size_t ws_bytes, cudaStream_t st, int phase = 0,
             int cluster_limit = 0);
// phase values beyond 0 / 1 / 2: ONE of the three GEMM stages (b200f_arcface_bwd_part), single class chunk, batch <= 512
constexpr int HEAD_BWD_PART_K3A = 10;   // logit gradient G^T + r partials
constexpr int HEAD_BWD_PART_K3B = 11;   // dW from them (cluster_limit: clusters it may occupy, 0 = all)
constexpr int HEAD_BWD_PART_K3C = 12;   // dx_hat (+ the fused dL/dx tail) from them (cluster_limit likewise)
// 1 if head_bwd can run in parts for this shape
int head_bwd_parts_ok(int64_t B, int64_t C, int D);

// ||dW||^2 side output: the calling thread's next head_bwd (phase 0, or phases 1 + 2) leaves sum(dW^2) in out[0]
void head_request_dw_sqnorm(float* out);
float* head_dw_sqnorm_request();

// dx = normalise-backward(dxhat) with the rows HeadDx names (one launch of rowops::l2norm_bwd)
int head_dx_finish(const void* xh, float S, const HeadDx* hdx, const float* dxhat, int64_t B, int D, cudaStream_t st);

// K4 on tensor cores (umma_xw_topk.cuh): bf16 tcgen05 scan + exact fp32 re-rank + verification
bool gallery_tc_supported(int D);
size_t gallery_scan_workspace(int64_t Q, int64_t N, int D, int k);
int gallery_prepare(const void* g, int dtype, int64_t N, int D, int metric, int fmt, void* g16, float* bias, cudaStream_t st);
int gallery_scan_select(const float* q, const float* g, const void* g16, const float* bias, const float* q_inv,
                        const float* g_inv, int64_t Q, int64_t N, int64_t index_offset, int D, int k, int metric, int fmt,
                        float thresh, int64_t* idx, float* score, uint8_t* accept, uint8_t* redo, int32_t* redo_count,
                        char* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace umma
}  // namespace b200f
