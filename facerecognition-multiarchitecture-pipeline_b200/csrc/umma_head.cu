// tcgen05 engine -- placeholder until the UMMA kernels land (next commit).
#include "umma_api.cuh"

namespace b200f {
namespace umma {

bool head_engine_selected(int64_t, int64_t, int, int, int, bool) { return false; }
size_t head_workspace_bytes(int64_t, int64_t, int, int, int) { return 0; }
int head_fwd(const void*, const void*, const float*, const float*, const int64_t*, int64_t, int64_t, int64_t, int,
             const b200f_head_cfg*, float*, float*, int64_t*, float*, int32_t*, char*, size_t, cudaStream_t) {
  return fail(B200F_ERR_UNSUPPORTED, "tcgen05 engine not built");
}
int head_bwd(const void*, const void*, const float*, const float*, const int64_t*, const float*, const float*,
             int64_t, int64_t, int64_t, int, const b200f_head_cfg*, float*, float*, char*, size_t, cudaStream_t) {
  return fail(B200F_ERR_UNSUPPORTED, "tcgen05 engine not built");
}

}  // namespace umma
}  // namespace b200f
