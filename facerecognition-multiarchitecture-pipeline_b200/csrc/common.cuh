// Shared helpers for libb200face.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost a null-pointer test unless a profiler injected itself

#include "../../include/b200face.h"

namespace b200f {

// ---- error state (thread-local, INTEGRATION.md: the Streamlit UI thread and the webcam thread
// may both call in, src/app.py:331-335,639) --------------------------------------------------
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);
void count_launch();   // diagnostic counter behind b200f_launch_count()

#define B200F_CUDA_OK(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess)                                                           \
      return ::b200f::fail(B200F_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,           \
                           cudaGetErrorString(_e), __FILE__, __LINE__);              \
  } while (0)

#define B200F_LAUNCH_OK(what)                                                        \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess)                                                           \
      return ::b200f::fail(B200F_ERR_CUDA, "launch of %s failed: %s", what,          \
                           cudaGetErrorString(_e));                                  \
    ::b200f::count_launch();                                                         \
  } while (0)

// ---- NVTX ranges around every C-ABI stage (nsys / ncu timelines name the stages) ----------------------
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define B200F_NVTX_CAT2(a, b) a##b
#define B200F_NVTX_CAT(a, b) B200F_NVTX_CAT2(a, b)
#define B200F_NVTX(name) ::b200f::NvtxRange B200F_NVTX_CAT(b200f_nvtx_range_, __LINE__)(name)

// ---- programmatic dependent launch -----------------------------------------------------------------------
// Every kernel of the head step is short (3-100 us) and the step is a chain of a dozen of them: with plain stream
// order each boundary costs the launch latency plus the next kernel's prologue (barrier init, TMEM allocation,
// descriptor prefetch).  Kernels therefore (1) allow their successor to be scheduled right away -- pdl_trigger() at
// their top -- and (2) run their prologue, then pdl_wait(), which returns once the predecessor grid has completed and
// its memory is visible, before the first access to anything a predecessor wrote.  Launches carry the
// programmatic-stream-serialization attribute (launch_pdl); without it both calls are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();            // tunable "pdl" (default on)
int k1_hints();                // tunable "k1_hints": L2 cache-policy bits of K1 over the class weights (rowops.cuh)
int k1_hints_set(int v);
void pdl_set(bool on);

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();   // cached per device

// ---- constants of the reference head (src/face_models.py:363,388) ---------------------------
__device__ __forceinline__ float cos_lo() { return -1.0f + 1e-7f; }   // rounds to -(1-2^-23) in fp32
__device__ __forceinline__ float cos_hi() { return 1.0f - 1e-7f; }
#define B200F_PI_CLAMP 3.14149265358979323846f   /* pi - 1e-4 */

// ---- element access ---------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}

// Load 8 consecutive elements as fp32; `valid` (0..8) leading elements are in bounds, the rest
// are zero-filled.  `vec_ok`: pointer is 16B(bf16)/32B(f32)-vector friendly (aligned rows).
template <typename T>
__device__ __forceinline__ void load8(const T* __restrict__ p, int valid, bool vec_ok, float (&out)[8]);

template <>
__device__ __forceinline__ void load8<float>(const float* __restrict__ p, int valid, bool vec_ok,
                                             float (&out)[8]) {
  if (valid >= 8 && vec_ok) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
    out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = (i < valid) ? __ldg(p + i) : 0.0f;
  }
}

template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* __restrict__ p, int valid,
                                                     bool vec_ok, float (&out)[8]) {
  if (valid >= 8 && vec_ok) {
    uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      out[2 * i]     = __uint_as_float(r[i] << 16);
      out[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = (i < valid) ? __bfloat162float(p[i]) : 0.0f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- the head's element-wise math, shared by every engine so all of them agree bit-for-bit on
// the epilogue (src/face_models.py:363-427 and its autograd) -------------------------------
struct HeadMath {
  float m_eff, s_eff;
  int easy;

  // pre-scale target logit phi(c) for a CLAMPED cosine c           (:366-397)
  __device__ __forceinline__ float phi(float c) const {
    float theta = acosf(c);
    if (easy) return (c > 0.0f) ? cosf(theta + m_eff) : c;
    const float tm = theta + m_eff;
    // torch.minimum propagates NaN (fminf would return the clamp and hide it from the :423 scrub)
    return cosf((tm != tm) ? tm : fminf(B200F_PI_CLAMP, tm));
  }
  // d phi / d c  (autograd of acos -> +m -> minimum -> cos)
  __device__ __forceinline__ float dphi(float c) const {
    float theta = acosf(c);
    float tm = theta + m_eff;
    float inv_sin = 1.0f / sqrtf(1.0f - c * c);
    if (easy) return (c > 0.0f) ? sinf(tm) * inv_sin : 1.0f;
    return (tm < B200F_PI_CLAMP) ? sinf(tm) * inv_sin : 0.0f;
  }
};

}  // namespace b200f
