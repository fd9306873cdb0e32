// TMA ingest microbenchmark (development aid, not part of the library): how many bytes per second can ONE CTA per SM pull
// into shared memory through cp.async.bulk.tensor when nothing consumes them -- by box shape, ring depth, and whether the
// source is L2-resident.  Why: the per-tile timeline of the X-stationary kernel (tools/timeline_probe.py) shows the MMA
// phase of a 256 x 256 x 512 tile at 3.1 us however few clusters run (2.1 us is the tensor floor): each CTA ingests its
// 128 KB of the streamed operand at ~42 GB/s, with 8 clusters as with 74.  Is that the TMA path's ceiling for
// [128 rows x 128 B] boxes of a K-major operand, or the ring's (80 KB in flight)?
//   v0: box [64 k x 128 rows] (16 KB, 128-byte swizzle) -- what the kernel does: 8 boxes walk one 128-row block
//   v1: box [64 k x 256 rows] (32 KB)
//   v2: two boxes [64 k x 64 rows] per stage
//   v3: box [256 k x 32 rows], no swizzle: 512 B contiguous per row (not an MMA layout: the per-row-segment cost)
//   v4: 3-D map [64 k][8 k-blocks][rows], box (64, 2, 128): two k-blocks per instruction (32 KB)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/tma_bw tools/microbench/tma_bw.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// A CTA walks row blocks b = blockIdx.x, + gridDim.x, ... of `rows_per_block` rows; per block it loads all 512 k.
template <int V>
__global__ void __launch_bounds__(64, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, int stages, int stage_bytes, int n_blocks, int repeats) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty = full + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  constexpr int ROWS = (V == 1) ? 256 : (V == 3 ? 32 : 128);       // rows per stage-load group
  constexpr int LOADS_PER_BLOCK = (V == 3) ? 2 : (V == 4 ? 4 : 8);   // stage loads that cover the block's 512 k
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int r = 0; r < repeats; ++r)
      for (int b = blockIdx.x; b < n_blocks; b += gridDim.x)
        for (int i = 0; i < LOADS_PER_BLOCK; ++i) {
          mbar_wait(&empty[s], ph ^ 1);
          if (lane == 0) {
            mbar_expect(&full[s], (uint32_t)stage_bytes);
            uint8_t* dst = smem + (size_t)s * stage_bytes;
            if (V == 0 || V == 1) tma2d(dst, &tm, &full[s], i * 64, b * ROWS);
            else if (V == 2) { tma2d(dst, &tm, &full[s], i * 64, b * ROWS); tma2d(dst + stage_bytes / 2, &tm, &full[s], i * 64, b * ROWS + 64); }
            else if (V == 3) tma2d(dst, &tm, &full[s], i * 256, b * ROWS);
            else tma3d(dst, &tm, &full[s], 0, i * 2, b * ROWS);
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1; }
        }
  } else {
    int s = 0; uint32_t ph = 0;
    for (int r = 0; r < repeats; ++r)
      for (int b = blockIdx.x; b < n_blocks; b += gridDim.x)
        for (int i = 0; i < LOADS_PER_BLOCK; ++i) {
          mbar_wait(&full[s], ph);
          if (lane == 0) mbar_arrive(&empty[s]);
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1; }
        }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int V>
static int run(EncodeFn enc, uint16_t* w, int64_t rows, int grid, int stages, int repeats, const char* what) {
  CUtensorMap tm;
  cuuint32_t estr[3] = {1, 1, 1};
  int stage_bytes;
  CUresult r;
  if (V == 4) {
    cuuint64_t gdim[3] = {64, 8, (cuuint64_t)rows}; cuuint64_t gstr[2] = {128, 1024}; cuuint32_t box[3] = {64, 2, 128};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, w, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    stage_bytes = 32768;
  } else {
    cuuint64_t gdim[2] = {512, (cuuint64_t)rows}; cuuint64_t gstr[1] = {1024};
    cuuint32_t box[2] = {64, 128};
    if (V == 1) box[1] = 256;
    if (V == 2) box[1] = 64;
    if (V == 3) { box[0] = 256; box[1] = 32; }
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, w, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            V == 3 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    stage_bytes = (V == 1) ? 32768 : 16384;
  }
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int rows_per_block = (V == 1) ? 256 : (V == 3 ? 32 : 128);
  const int n_blocks = (int)(rows / rows_per_block);
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 512;
  CK(cudaFuncSetAttribute(tma_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  tma_kernel<V><<<grid, 64, smem>>>(tm, stages, stage_bytes, n_blocks, 1);       // warm (and fill the L2 when it fits)
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  tma_kernel<V><<<grid, 64, smem>>>(tm, stages, stage_bytes, n_blocks, repeats);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double bytes = (double)rows * 1024 * repeats;
  printf("%-44s rows %8lld grid %3d stages %2d x %2d KB: %7.1f us  %6.2f TB/s  %6.1f GB/s per SM\n", what, (long long)rows, grid, stages,
         stage_bytes / 1024, ms * 1e3, bytes / ms / 1e9, bytes / ms / 1e6 / grid);
  return 0;
}

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeFn enc = reinterpret_cast<EncodeFn>(p);
  const int64_t big = 1 << 20;                               // 1 M rows x 1 KB = 1 GB: streams from HBM
  uint16_t* w; CK(cudaMalloc(&w, big * 1024)); CK(cudaMemset(w, 1, big * 1024));
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  for (int64_t rows : {(int64_t)32768, big}) {               // 32 MB: L2-resident on the second pass; 1 GB: HBM
    const int rep = rows == big ? 1 : 16;
    printf("---- source %s\n", rows == big ? "1 GB (HBM)" : "32 MB (L2-resident)");
    for (int grid : {nsm, 16}) {
      for (int st : {2, 3, 5, 8, 12}) if (run<0>(enc, w, rows, grid, st, rep, "v0 box 64k x 128 rows (16 KB)")) return 1;
      for (int st : {2, 3, 6}) if (run<1>(enc, w, rows, grid, st, rep, "v1 box 64k x 256 rows (32 KB)")) return 1;
      for (int st : {5, 12}) if (run<2>(enc, w, rows, grid, st, rep, "v2 2 boxes 64k x 64 rows per stage")) return 1;
      for (int st : {5, 12}) if (run<3>(enc, w, rows, grid, st, rep, "v3 box 256k x 32 rows, 512 B rows, no swizzle")) return 1;
      for (int st : {3, 6}) if (run<4>(enc, w, rows, grid, st, rep, "v4 3-D box (64, 2 k-blocks, 128 rows) 32 KB")) return 1;
    }
  }
  return 0;
}
