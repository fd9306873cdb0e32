// tcgen05 / TMEM / TMA engine entry points (operand dtype B200F_F16N).  Defined in umma_head.cu.
#pragma once
#include "common.cuh"

namespace b200f {
namespace umma {

bool available();                                    // current device is sm_100
size_t head_workspace_bytes(int64_t B, int64_t C, int D);

int head_fwd(const void* xh, const void* wh, const int64_t* label, int64_t B, int64_t C, int64_t class_offset, int D,
             const b200f_head_cfg* cfg, float* row_stats, float* row_best, int64_t* row_argmax, float* cos_minmax,
             int32_t* nan_flag, char* ws, size_t ws_bytes, cudaStream_t st);

int head_bwd(const void* xh, const void* wh, const float* inv_nw, const int64_t* label, const float* lse,
             const float* grad4, int64_t B, int64_t C, int64_t class_offset, int D, const b200f_head_cfg* cfg,
             float* dxhat, float* dw, char* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace umma
}  // namespace b200f
