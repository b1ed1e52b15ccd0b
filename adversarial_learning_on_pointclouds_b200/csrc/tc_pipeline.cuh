// Shared pieces of the tensor-core GEMM kernels: tile constants, the shared-memory
// ring, the TMA producer and the tcgen05.mma issuer loops.
#pragma once
#include "tc_common.cuh"

namespace pcadv {
namespace tc {

constexpr int kStages = 3;                        // ring depth of kernels that also stage an epilogue
constexpr int kMaxStages = 4;                     // ring depth of kernels without epilogue staging
constexpr int kBlockK = 64;                       // elements per K chunk = one 128-byte swizzle row
constexpr int kTileM = 128;
constexpr int kMaxTileN = 256;
constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KB
constexpr int kBBytes = kMaxTileN * kBlockK * 2;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSlabBytes = kTileM * 128;          // epilogue staging slab: 128 rows x 128 B, swizzled
constexpr int kEpiBytes = 4 * kSlabBytes;         // 4 slabs (2 output + 2 mask, or 4 output)
constexpr int kBiasFloats = 3 * kMaxTileN;        // bias + two per-cloud bias rows of the tile
constexpr int kSmemBytes =
    kStages * kStageBytes + kEpiBytes + kBiasFloats * 4 + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kTmemCols = 512;

struct TensorMaps {
  CUtensorMap act[PCADV_MAX_SEG];   // activation segments [rows, k_i]
  CUtensorMap w;                    // weights [n, ktot]  (wgrad: dz [rows, n])
  CUtensorMap out;                  // output [rows, n]: box 128 rows x 128 B (TMA store)
  CUtensorMap mask;                 // saved activation [rows, n]: box 128 rows x 64 cols (TMA load)
  CUtensorMap w_half;               // weights, box of bn / 2 rows: one CTA's share of a multicast weight tile
};

struct SharedTail {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t mask_full[4];
  uint32_t tmem_base;
};

struct SmemLayout {
  uint8_t* stages;
  uint8_t* epi;
  float* bias;
  SharedTail* tail;
};

__device__ __forceinline__ SmemLayout carve_smem(uint8_t* raw) {
  SmemLayout L;
  L.stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  L.epi = L.stages + kStages * kStageBytes;
  L.bias = reinterpret_cast<float*>(L.epi + kEpiBytes);
  L.tail = reinterpret_cast<SharedTail*>(L.epi + kEpiBytes + kBiasFloats * 4);
  return L;
}

// one thread: barrier init; warp 1: TMEM allocation; everyone: publish
__device__ __forceinline__ uint32_t pipeline_setup(const SmemLayout& L, int warp, int lane,
                                                   uint32_t tmem_empty_count,
                                                   uint32_t empty_count = 1) {
  SharedTail* st = L.tail;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&st->full[i], 1);
      mbar_init(&st->empty[i], empty_count);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st->tmem_full[i], 1);
      mbar_init(&st->tmem_empty[i], tmem_empty_count);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&st->mask_full[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&st->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return st->tmem_base;
}

__device__ __forceinline__ void pipeline_teardown(int warp, uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// K-major operands (forward / dgrad): for every 64-wide K chunk issue 4 MMAs (K = 16 each).
// 128B swizzle, 8-row groups 1024 B apart; +32 B of start address per K = 16 step.
__device__ __forceinline__ void mma_chunk_kmajor(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr,
                                                 uint32_t idesc, bool first_chunk) {
#pragma unroll
  for (int k = 0; k < kBlockK / 16; ++k) {
    const uint64_t adesc = make_smem_desc(a_addr + k * 32, 16, 1024);
    const uint64_t bdesc = make_smem_desc(b_addr + k * 32, 16, 1024);
    umma_f16(d_tmem, adesc, bdesc, idesc, (first_chunk && k == 0) ? 0u : 1u);
  }
}

// MN-major operands (wgrad): tiles are [64 K rows][64 channels] boxes; 64-channel blocks are
// 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO); +16 rows * 128 B per K = 16 step.
__device__ __forceinline__ void mma_chunk_mnmajor(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr,
                                                  uint32_t idesc, bool first_chunk) {
#pragma unroll
  for (int k = 0; k < kBlockK / 16; ++k) {
    const uint64_t adesc = make_smem_desc(a_addr + k * 2048, 8192, 1024);
    const uint64_t bdesc = make_smem_desc(b_addr + k * 2048, 8192, 1024);
    umma_f16(d_tmem, adesc, bdesc, idesc, (first_chunk && k == 0) ? 0u : 1u);
  }
}

int num_sms();
int ensure_smem(const void* kernel);

}  // namespace tc
}  // namespace pcadv
