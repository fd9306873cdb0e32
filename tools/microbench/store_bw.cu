// Store-path microbenchmark (development aid, not part of the library): how fast can ONE CTA per SM with 8 / 16 / 32
// warps write a 205 MB fp32 matrix [C, 512] when
//   v0: a warp writes 512 contiguous bytes per instruction (st.v4 per lane)
//   v1: a lane owns a row and writes its 128 B segment as 4 x 32 B (st.v8) -- the class-major dW epilogue pattern
//   v2: v1's tile staged in shared memory, each lane issues one 128 B cp.async.bulk shared -> global
//   v3: staged, one 4 KB bulk copy per warp (contiguous destination: the TMA store ceiling)
//   v4: read-only, a lane reads 64 B of its row (2 x ld.v8 / 4 x ld.v4), the w_hat pattern, 1 slice in flight
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/store_bw tools/microbench/store_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int V>
__global__ void __launch_bounds__(1024, 1) store_kernel(float* __restrict__ out, const uint16_t* __restrict__ in, int64_t rows, float seed) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // work unit: a [32 rows x 32 floats] slice; slices of a 128-row x 128-col tile belong together like in the epilogue
  const int64_t n_units = rows / 32 * 16;                    // 16 column slices of 32 floats per 32-row block
  float acc = seed;
  uint8_t* stage = smem + (size_t)warp * 4608;                // 32 rows x 144 B (padded) or 4 KB dense
  for (int64_t u = (int64_t)blockIdx.x * nw + warp; u < n_units; u += (int64_t)gridDim.x * nw) {
    const int64_t rb = u / 16; const int cs = (int)(u % 16);
    const int64_t row = rb * 32 + lane;
    float* dst = out + row * 512 + cs * 32;
    if (V == 0) {
      // 32 rows x 128 B written as 8 instructions of 512 contiguous bytes (4 rows each)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float* d = out + (rb * 32 + i * 4 + (lane >> 3)) * 512 + cs * 32 + (lane & 7) * 4;
        *reinterpret_cast<float4*>(d) = make_float4(acc, acc, acc, acc);
      }
    } else if (V == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t a = __float_as_uint(acc + i);
        asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(dst + i * 8), "r"(a) : "memory");
      }
    } else if (V == 2) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      float4* s = reinterpret_cast<float4*>(stage + lane * 144);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = make_float4(acc, acc, acc, acc);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(dst), "r"(smem_u32(s)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    } else if (V == 3) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      float4* s = reinterpret_cast<float4*>(stage) + lane;
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i * 32] = make_float4(acc, acc, acc, acc);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(out + u * 1024), "r"(smem_u32(stage)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else if (V == 4) {
      const uint16_t* src = in + row * 512 + cs * 32;
      uint32_t r[16];
      asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(src));
      asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "l"(src + 16));
      uint32_t x = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) x ^= r[i];
      acc += __uint_as_float(x & 0x3f800000u);
    }
  }
  if (V == 2 || V == 3) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 123.456f) out[0] = acc;
}

template <int V>
static int run(const char* name, float* out, uint16_t* in, int64_t rows, int threads, float* flush, size_t flush_n) {
  CK(cudaFuncSetAttribute(store_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaMemsetAsync(flush, 0, flush_n));
    CK(cudaEventRecord(e0));
    store_kernel<V><<<148, threads, 200 * 1024>>>(out, in, rows, 1.0f);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it > 0 && ms < best) best = ms;
  }
  const double bytes = (V == 4) ? (double)rows * 1024 : (double)rows * 2048;
  printf("%-34s warps/SM=%2d  %7.1f us  %6.2f TB/s\n", name, threads / 32, best * 1e3, bytes / best / 1e9);
  return 0;
}

// G^T pattern of K3a: fp16 matrix [rows, ld bytes]; a warp owns 32 rows and walks `span / 2` bytes of each in pieces of
// PIECE bytes per lane (64 = one 32-column slice, 128 = two); `span` bytes of every row per work item (the 512 B a CTA
// pair covers per tile), items of the same rows on consecutive CTAs like the kernel's row groups.
template <int PIECE>
__global__ void __launch_bounds__(256, 1) gt_kernel(uint16_t* __restrict__ out, int64_t rows, int64_t ld_bytes, int span) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t groups = ld_bytes / span;
  const int64_t n_items = (rows / 128) * groups;             // item = (128-row block, group), group fastest
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int64_t rb = item / groups, g = item % groups;
    const int quad = warp & 3, half = warp >> 2;
    const int64_t row = rb * 128 + quad * 32 + lane;
    uint8_t* base = reinterpret_cast<uint8_t*>(out) + row * ld_bytes + g * span + half * (span / 2);
    for (int off = 0; off < span / 2; off += PIECE) {
#pragma unroll
      for (int i = 0; i < PIECE / 32; ++i) {
        const uint32_t a = (uint32_t)(off + i);
        asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(base + off + i * 32), "r"(a) : "memory");
      }
    }
  }
}

template <int PIECE>
static int run_gt(const char* name, uint16_t* out, int64_t rows, int64_t ld_bytes, int span, float* flush, size_t flush_n) {
  CK(cudaFuncSetAttribute(gt_kernel<PIECE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaMemsetAsync(flush, 0, flush_n));
    CK(cudaEventRecord(e0));
    gt_kernel<PIECE><<<148, 256, 200 * 1024>>>(out, rows, ld_bytes, span);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it > 0 && ms < best) best = ms;
  }
  printf("%-26s rows=%6lld ld=%5lld B span=%4d  %7.1f us  %6.2f TB/s\n", name, (long long)rows, (long long)ld_bytes, span,
         best * 1e3, (double)rows * ld_bytes / best / 1e9);
  return 0;
}

int main() {
  const int64_t rows = 100000 / 32 * 32;
  float* out; uint16_t* in; float* flush; const size_t flush_n = 256u << 20;
  CK(cudaMalloc(&out, rows * 2048)); CK(cudaMalloc(&in, rows * 1024)); CK(cudaMalloc(&flush, flush_n));
  CK(cudaMemset(in, 0, rows * 1024));
  for (int threads : {256, 512, 1024}) {
    if (run<0>("v0 coalesced st.v4", out, in, rows, threads, flush, flush_n)) return 1;
    if (run<1>("v1 row-per-lane 4 x st.v8", out, in, rows, threads, flush, flush_n)) return 1;
    if (run<2>("v2 staged, 128 B bulk per lane", out, in, rows, threads, flush, flush_n)) return 1;
    if (run<3>("v3 staged, 4 KB bulk per warp", out, in, rows, threads, flush, flush_n)) return 1;
    if (run<4>("v4 row-per-lane loads 64 B", out, in, rows, threads, flush, flush_n)) return 1;
  }
  {
    uint16_t* gt; CK(cudaMalloc(&gt, (size_t)99968 * 8192));
    for (int span : {512, 1024, 8192}) {
      if (run_gt<64>("gt 64 B pieces", gt, 41728, 8192, span, flush, flush_n)) return 1;
      if (run_gt<128>("gt 128 B pieces", gt, 41728, 8192, span, flush, flush_n)) return 1;
    }
    if (run_gt<64>("gt 64 B pieces (cfg3)", gt, 99968, 1024, 512, flush, flush_n)) return 1;
    if (run_gt<128>("gt 128 B pieces (cfg3)", gt, 99968, 1024, 512, flush, flush_n)) return 1;
  }
  return 0;
}
