// Microbenchmarks behind the epilogue design of tc_rows.cu / tc_chain.cu (sm_100a):
//   A. TMEM read throughput: tcgen05.ld.32x32b.x32 per SM for 4 / 8 / 16 reader warps, one or two
//      loads in flight per warp before tcgen05.wait::ld.
//   B. TMA store throughput for small boxes: 64 x R element (128 B x R rows, 128B swizzle) stores of
//      fp16 from shared memory, R = 32 (4 KB: what every epilogue warp issues per 64-column step)
//      against R = 128 (16 KB), streaming over a buffer much larger than L2.
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_tma tmem_tma.cu && ./tmem_tma
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---------------------------------------------------------------- A: TMEM reads
template <int kInFlight>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  __shared__ unsigned long long worst;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) worst = 0ull;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base_s + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[kInFlight][32];
#pragma unroll
    for (int j = 0; j < kInFlight; ++j) {
      const uint32_t col = static_cast<uint32_t>(((i * kInFlight + j) * 32) & 511);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[j][0]), "=r"(r[j][1]), "=r"(r[j][2]), "=r"(r[j][3]), "=r"(r[j][4]), "=r"(r[j][5]), "=r"(r[j][6]),
            "=r"(r[j][7]), "=r"(r[j][8]), "=r"(r[j][9]), "=r"(r[j][10]), "=r"(r[j][11]), "=r"(r[j][12]),
            "=r"(r[j][13]), "=r"(r[j][14]), "=r"(r[j][15]), "=r"(r[j][16]), "=r"(r[j][17]), "=r"(r[j][18]),
            "=r"(r[j][19]), "=r"(r[j][20]), "=r"(r[j][21]), "=r"(r[j][22]), "=r"(r[j][23]), "=r"(r[j][24]),
            "=r"(r[j][25]), "=r"(r[j][26]), "=r"(r[j][27]), "=r"(r[j][28]), "=r"(r[j][29]), "=r"(r[j][30]),
            "=r"(r[j][31])
          : "r"(taddr + col));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < kInFlight; ++j)
#pragma unroll
      for (int e = 0; e < 32; ++e) acc ^= r[j][e];
  }
  const long long t1 = clock64();
  atomicMax(&worst, static_cast<unsigned long long>(t1 - t0));
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = worst;
  if (acc == 0x12345u) sink[0] = acc;
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s) : "memory");
  }
}

template <int kInFlight>
static void run_tmem(int warps, int iters) {
  unsigned long long* cyc;
  uint32_t* sink;
  CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaMalloc(&sink, 4));
  tmem_read_kernel<kInFlight><<<148, warps * 32>>>(iters, cyc, sink);   // warm-up
  tmem_read_kernel<kInFlight><<<148, warps * 32>>>(iters, cyc, sink);
  CK(cudaDeviceSynchronize());
  unsigned long long h[148];
  CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += static_cast<double>(h[i]);
  mean /= 148;
  const double bytes = static_cast<double>(warps) * iters * kInFlight * 4096.0;
  printf("TMEM read: %2d warps, %d x32 loads in flight per warp: %7.1f B/cycle/SM  (%.0f cycles per warp-load)\n",
         warps, kInFlight, bytes / mean, mean / (iters * kInFlight));
  cudaFree(cyc); cudaFree(sink);
}

// ---------------------------------------------------------------- B: TMA stores
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(512, 1) tma_store_kernel(const __grid_constant__ CUtensorMap map, int box_rows,
                                                           int iters, int64_t rows_total) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int slab_bytes = box_rows * 128;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  uint8_t* mine = base + static_cast<size_t>(warp) * 2 * slab_bytes;       // two slabs per warp
  for (int e = lane; e < 2 * slab_bytes / 4; e += 32) reinterpret_cast<uint32_t*>(mine)[e] = e + warp;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    const int64_t stream_id = static_cast<int64_t>(blockIdx.x) * warps + warp;
    const int64_t streams = static_cast<int64_t>(gridDim.x) * warps;
    for (int i = 0; i < iters; ++i) {
      const int64_t row = ((static_cast<int64_t>(i) * streams + stream_id) * box_rows) % rows_total;
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(mine + (i & 1) * slab_bytes)), "r"(0),
                     "r"(static_cast<int32_t>(row))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

static void run_tma(EncodeFn enc, void* buf, int64_t rows_total, int box_rows, int warps, double total_bytes) {
  CUtensorMap map;
  cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(rows_total)};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  const int slab = box_rows * 128;
  const size_t smem = 1024 + static_cast<size_t>(warps) * 2 * slab;
  CK(cudaFuncSetAttribute(tma_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = static_cast<int>(total_bytes / (148.0 * warps * slab));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  tma_store_kernel<<<148, warps * 32, smem>>>(map, box_rows, iters, rows_total);
  CK(cudaEventRecord(a));
  tma_store_kernel<<<148, warps * 32, smem>>>(map, box_rows, iters, rows_total);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  const double bytes = 148.0 * warps * iters * slab;
  printf("TMA store: %2d warps/SM x %3d-row boxes (%5d B): %7.1f GB/s  (%.2f us per store per SM-warp)\n", warps,
         box_rows, slab, bytes / ms / 1e6, ms * 1e3 / iters);
}

int main() {
  printf("== A. tcgen05.ld.32x32b.x32 throughput per SM (4096 B per warp-load)\n");
  for (int w : {4, 8, 16}) { run_tmem<1>(w, 4000); run_tmem<2>(w, 2000); }
  printf("== B. TMA 2-D stores from shared memory (fp16, 128-byte rows, streaming over 4 GiB)\n");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  const int64_t rows_total = (4ll << 30) / 128;
  void* buf;
  CK(cudaMalloc(&buf, 4ll << 30));
  for (int w : {2, 4, 8, 16}) run_tma(reinterpret_cast<EncodeFn>(fn), buf, rows_total, 32, w, 8e9);
  for (int w : {1, 2, 4, 6}) run_tma(reinterpret_cast<EncodeFn>(fn), buf, rows_total, 128, w, 8e9);
  cudaFree(buf);
  return 0;
}
