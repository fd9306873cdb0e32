// tcgen05 / TMEM / TMA engine entry points (bf16 inputs).  Defined in umma_head.cu.
#pragma once
#include "common.cuh"

namespace b200f {
namespace umma {

// true when this call is served by the tcgen05 engine (bf16, supported shape, no logits output)
bool head_engine_selected(int64_t B, int64_t C, int D, int dtype, int engine, bool wants_logits);
size_t head_workspace_bytes(int64_t B, int64_t C, int D, int dtype, int engine);

int head_fwd(const void* x, const void* w, const float* inv_nx, const float* inv_nw, const int64_t* label,
             int64_t B, int64_t C, int64_t class_offset, int D, const b200f_head_cfg* cfg, float* row_stats,
             float* row_best, int64_t* row_argmax, float* cos_minmax, int32_t* nan_flag, char* ws,
             size_t ws_bytes, cudaStream_t st);

int head_bwd(const void* x, const void* w, const float* inv_nx, const float* inv_nw, const int64_t* label,
             const float* lse, const float* grad_scale, int64_t B, int64_t C, int64_t class_offset, int D,
             const b200f_head_cfg* cfg, float* dxhat, float* dw, char* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace umma
}  // namespace b200f
