// Tensor-core engine, CHAINED pointwise layers: y_l = act_l(y_{l-1} W_l^T + b_l) for up to four
// consecutive layers of width <= 256 in one persistent kernel (conv2 -> conv3 -> conv4 of the
// trunk; the discriminator's conv1 -> .. -> conv4 + max over channels).
//
// The output slab an epilogue warp packs for the TMA store of layer l -- 32 rows x 128 bytes,
// 128B-swizzled, four warps stacked -- is exactly the K-major A tile [128 rows x 64 K] the next
// layer's MMA wants, so layer l + 1 multiplies straight out of shared memory and no intermediate
// activation is read back from HBM (they are still written once: the backward needs them).
//
// 64 + 128 G threads, G = 2 or 3 epilogue groups (three where every accumulator is <= 128 columns wide:
// TMEM then holds a third tile, and one layer of one tile being a single dependency chain of ~5 600
// cycles, a third tile in flight is worth 19 % on the trunk and 26 % on the discriminator's chain).
// As in tc_rows.cu: warp 0 TMA producer (layer 0's A chunks and every layer's weight
// chunks, in the order the MMA thread consumes them), warp 1 MMA issuer, warps 2..9 two epilogue
// halves.  Half h owns TMEM columns [256 h, 256 h + 256) and the shared-memory tile H[h]; the CTA's
// tiles alternate between the halves, and for every (tile, layer) the MMA thread and the half
// ping-pong on two barriers: acc_full[h] (MMA done -> epilogue) and in_ready[h] (H[h] holds the
// layer's output and the accumulator is free -> next MMA).
#include <stdlib.h>
#include "tc_pipeline.cuh"

namespace pcadv {
namespace tc {

constexpr int kChainMaxGroups = 3;                // epilogue groups = tiles in flight
__host__ __device__ constexpr int chain_threads(int groups) { return 64 + 128 * groups; }
constexpr int kChainMaxLayers = 4;
constexpr int kChainMaxStages = 6;
constexpr int kChainSmemMax = 232448;
constexpr int kChainBitsWords = kMaxTileN / 32;

struct ChainMaps {
  CUtensorMap x;                              // layer 0 input [rows, k0]: box 64 x 128 rows
  CUtensorMap w[kChainMaxLayers];             // weights [n_l, k_l]: box 64 x n_l rows
  CUtensorMap out[kChainMaxLayers];           // outputs [rows, n_l]: box 64 x 32 rows (TMA store)
};

struct ChainTail {
  uint64_t full[kChainMaxStages];
  uint64_t empty[kChainMaxStages];
  uint64_t acc_full[kChainMaxGroups];
  uint64_t in_ready[kChainMaxGroups];
  uint32_t tmem_base;
};

struct ChainParams {
  int64_t rows;
  int64_t tiles;
  int num_layers;
  int k[kChainMaxLayers];                     // multiples of 64, <= 256
  int n[kChainMaxLayers];                     // multiples of 64, <= 256 (n[l] == k[l + 1])
  int nvalid[kChainMaxLayers];                // real output channels (< n only for the fp32 last layer)
  int act[kChainMaxLayers];
  float slope[kChainMaxLayers];
  uint32_t idesc[kChainMaxLayers];
  const float* bias[kChainMaxLayers];
  int has_out[kChainMaxLayers];               // store the layer's 16-bit output
  uint32_t* bits_out[kChainMaxLayers];        // sign-bit maps (or NULL)
  int64_t ld_bits[kChainMaxLayers];
  unsigned long long* rowmax_key;             // last layer: max over channels instead of an output
  float* out_f32;                             // last layer: fp32 rows [rows, n_f32] (contiguous) instead
  int n_f32;                                  //   of a 16-bit TMA-stored output (n_f32 <= 64 <= n[last])
  int serial;                                 // one tile in flight, both halves split its steps (wide chains)
  int groups;                                 // epilogue groups of four warps = tiles in flight (2, or 3 when every n <= 128)
  int nstages, stage_bytes, h_bytes;
  int bf16;
  // debugging aid (pcadv_debug_chain_trace): clock64 stamps of CTA 0, [warp 0..9][slot 0..kTraceSlots)
  long long* trace;
};
constexpr int kTraceSlots = 256;
#define CHAIN_STAMP(slot_expr)                                                         \
  do {                                                                                 \
    if (tracing && lane == 0) {                                                        \
      const int s_ = (slot_expr);                                                      \
      if (s_ < kTraceSlots) p.trace[warp * kTraceSlots + s_] = clock64();              \
    }                                                                                  \
  } while (0)

struct ChainSmem {
  uint8_t* stages;
  uint8_t* H;                                 // [groups][h_bytes]
  float* bias;                                // [layers][256]
  uint32_t* bits;                             // [4 * groups warps][32 rows][8 words]
  ChainTail* tail;
};

__device__ __forceinline__ ChainSmem carve_chain(uint8_t* raw, const ChainParams& p) {
  ChainSmem L;
  L.stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  L.H = L.stages + p.nstages * p.stage_bytes;
  L.bias = reinterpret_cast<float*>(L.H + p.groups * p.h_bytes);
  L.bits = reinterpret_cast<uint32_t*>(L.bias + kChainMaxLayers * kMaxTileN);
  L.tail = reinterpret_cast<ChainTail*>(L.bits + 4 * p.groups * 32 * kChainBitsWords);
  return L;
}

static size_t chain_smem_bytes(int nstages, int stage_bytes, int h_bytes, int groups = 2) {
  return 1024 + static_cast<size_t>(nstages) * stage_bytes + groups * static_cast<size_t>(h_bytes) +
         kChainMaxLayers * kMaxTileN * 4 + 4 * groups * 32 * kChainBitsWords * 4 + sizeof(ChainTail) + 16;
}

__device__ __forceinline__ float4 c_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void c_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void c_sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 c_lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void c_tma_store(const CUtensorMap* m, uint32_t smem_addr, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <bool kBf16>
__device__ __forceinline__ uint32_t c_gt0_mask(uint32_t packed) {
  uint32_t m;
  if (kBf16) asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(packed), "r"(0u));
  else asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(packed), "r"(0u));
  return m;
}

template <bool kBf16, int kG>
__global__ void __launch_bounds__(chain_threads(kG), 1)
tc_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  constexpr int kChainThreads = chain_threads(kG);
  constexpr uint32_t kGroupCols = kG == 2 ? kMaxTileN : 128;   // TMEM columns of a group's accumulator
  extern __shared__ uint8_t smem_raw[];
  const ChainSmem L = carve_chain(smem_raw, p);
  ChainTail* st = L.tail;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NL = p.num_layers;
  const bool tracing = kG == 2 && p.trace != nullptr && blockIdx.x == 0;   // the trace buffer has ten warps

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x);
    for (int l = 0; l < NL; ++l) {
      tma_prefetch_desc(&maps.w[l]);
      if (p.has_out[l]) tma_prefetch_desc(&maps.out[l]);
    }
    for (int i = 0; i < kChainMaxStages; ++i) {
      mbar_init(&st->full[i], 1);
      mbar_init(&st->empty[i], 1);
    }
    for (int h = 0; h < kG; ++h) {
      mbar_init(&st->acc_full[h], 1);
      mbar_init(&st->in_ready[h], p.serial ? 8 : 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&st->tmem_base, kTmemCols);
  // every layer's bias, once: all tiles cover all columns (n <= 256)
  for (int e = threadIdx.x; e < NL * kMaxTileN; e += kChainThreads) {
    const int l = e / kMaxTileN, c = e - l * kMaxTileN;
    L.bias[e] = (c < p.nvalid[l] && p.bias[l]) ? __ldg(p.bias[l] + c) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = st->tmem_base;
  // tiles of this CTA: i = 0, 1, 2, ... -> global tile blockIdx.x + i * gridDim.x, half i & 1
  const int64_t my_tiles = p.tiles > blockIdx.x ? (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  // serial mode (chains with a 256-wide intermediate: one shared-memory tile only): one tile in
  // flight in accumulator / tile 0, and the two epilogue halves take alternate 64-column steps of it
  const int tiles_in_flight = p.serial ? 1 : kG;

  if (warp == 0) {
    // ================= TMA producer: same (pair, layer, half, chunk) order as the MMA thread =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t i0 = 0; i0 < my_tiles; i0 += tiles_in_flight) {
        for (int l = 0; l < NL; ++l) {
          const int chunks = p.k[l] >> 6;
          const uint32_t tx = static_cast<uint32_t>((l == 0 ? kTileM : 0) + p.n[l]) * kBlockK * 2;
          for (int h = 0; h < tiles_in_flight && i0 + h < my_tiles; ++h) {
            const int32_t m0 = static_cast<int32_t>((blockIdx.x + (i0 + h) * gridDim.x) * kTileM);
            for (int c = 0; c < chunks; ++c) {
              mbar_wait_backoff(&st->empty[stage], phase ^ 1);
              mbar_arrive_expect_tx(&st->full[stage], tx);
              uint8_t* sa = L.stages + stage * p.stage_bytes;
              if (l == 0) tma_load_2d(sa, &maps.x, &st->full[stage], c * kBlockK, m0);
              tma_load_2d(sa + kABytes, &maps.w[l], &st->full[stage], c * kBlockK, 0);
              if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t ready_par[kChainMaxGroups] = {0, 0, 0};   // parity of the next in_ready wait per group
      for (int64_t i0 = 0; i0 < my_tiles; i0 += tiles_in_flight) {
        for (int l = 0; l < NL; ++l) {
          const int chunks = p.k[l] >> 6;
          for (int h = 0; h < tiles_in_flight && i0 + h < my_tiles; ++h) {
            // H[h] holds the previous layer's output (l > 0) and the accumulator has been drained
            const int tslot = static_cast<int>(((i0 / tiles_in_flight) * NL + l) * kG + h) * 3;
            CHAIN_STAMP(tslot);
            uint32_t rp = ready_par[0];
            if (h == 1) rp = ready_par[1];
            else if (h == 2) rp = ready_par[2];
            mbar_wait_backoff(&st->in_ready[h], rp ^ 1);
            CHAIN_STAMP(tslot + 1);
            if (h == 0) ready_par[0] ^= 1;
            else if (h == 1) ready_par[1] ^= 1;
            else ready_par[2] ^= 1;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(h) * kGroupCols;
            const uint32_t h_addr = smem_u32(L.H + h * p.h_bytes);
            for (int c = 0; c < chunks; ++c) {
              mbar_wait_backoff(&st->full[stage], phase);
              tc_fence_after();
              const uint32_t s_addr = smem_u32(L.stages + stage * p.stage_bytes);
              const uint32_t a_addr = l == 0 ? s_addr : h_addr + static_cast<uint32_t>(c) * kABytes;
              mma_chunk_kmajor(d_tmem, a_addr, s_addr + kABytes, p.idesc[l], c == 0);
              umma_commit(&st->empty[stage]);
              if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&st->acc_full[h]);
            CHAIN_STAMP(tslot + 2);
          }
        }
      }
    }
  } else {
    // ================= epilogue: half = tile parity, warp = 32-row TMEM lane quarter =================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int buf = p.serial ? 0 : half;                 // accumulator buffer / H tile this warp works on
    const int lane_row = quarter * 32 + lane;
    const uint32_t h_base = smem_u32(L.H + buf * p.h_bytes);
    // this thread's 128-byte row inside a step's [128 rows x 128 B] tile of H
    const uint32_t row_off = static_cast<uint32_t>(quarter) * 4096u + static_cast<uint32_t>(lane) * 128u;
    const uint32_t bits_s = smem_u32(L.bits + ew * 32 * kChainBitsWords);
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                            static_cast<uint32_t>(buf) * kGroupCols;
    uint32_t full_par = 0;
    for (int64_t i = p.serial ? 0 : half; i < my_tiles; i += tiles_in_flight) {
      const int64_t tm = blockIdx.x + i * gridDim.x;
      const int64_t r = tm * kTileM + lane_row;
      const bool r_ok = r < p.rows;
      const int32_t row_tma = static_cast<int32_t>(tm * kTileM + quarter * 32);
      for (int l = 0; l < NL; ++l) {
        const int steps = p.n[l] >> 6;
        const bool last = l == NL - 1;
        const bool rowmax = last && p.rowmax_key != nullptr;
        const bool f32out = last && p.out_f32 != nullptr;
        const bool want_bits = p.bits_out[l] != nullptr;
        const int act = p.act[l];
        const float slope = p.slope[l];
        const uint32_t bias_a = smem_u32(L.bias + l * kMaxTileN);
        const int eslot = static_cast<int>((i / tiles_in_flight) * NL + l) * 6;
        CHAIN_STAMP(eslot);
        mbar_wait(&st->acc_full[buf], full_par);
        CHAIN_STAMP(eslot + 1);
        full_par ^= 1;
        tc_fence_after();
        // H[half] was the A operand of the MMAs that just completed; the TMA stores issued from it
        // for the previous layer must have finished reading before it is overwritten
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
        // the fp32 row staging of a warp overlaps slabs other warps stored from: everybody's stores
        // must have been read out first
        if (f32out) named_barrier_sync(1 + buf, p.serial ? 256 : 128);
        float best = -INFINITY;
        int best_col = 0;
#pragma unroll 1
        for (int step = 0; step < steps; ++step) {
          if (p.serial && (step & 1) != half) continue;    // the other half's step
          const uint32_t srow = h_base + static_cast<uint32_t>(step) * kABytes + row_off;
          uint32_t raw[2][32];
          tmem_ld32_issue(taddr0 + step * 64, raw[0]);
          tmem_ld32_issue(taddr0 + step * 64 + 32, raw[1]);
          tmem_ld32_wait(raw[0]);
          tmem_ld32_wait(raw[1]);
          if (step == 0) CHAIN_STAMP(eslot + 2);
          uint32_t obits[2] = {0u, 0u};
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[h2][j]);
            const uint32_t ba = bias_a + static_cast<uint32_t>(step * 64 + h2 * 32) * 4u;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b = c_lds128(ba + q * 16);
              v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
            }
            if (rowmax) {
              float m = v[0];
#pragma unroll
              for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
              if (m > best) {
                int jj = 0;
#pragma unroll
                for (int j = 31; j >= 0; --j) jj = (v[j] == m) ? j : jj;
                best = m;
                best_col = step * 64 + h2 * 32 + jj;
              }
              continue;
            }
            if (f32out) {
              // fp32 rows, compact [32 rows][n_f32] staging in this warp's 8 KB of the (now free) tile
              const uint32_t crow = h_base + static_cast<uint32_t>(quarter) * 8192u +
                                    static_cast<uint32_t>(lane * p.n_f32 + step * 64 + h2 * 32) * 4u;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (step * 64 + h2 * 32 + j < p.n_f32)
                  asm volatile("st.shared.f32 [%0], %1;" ::"r"(crow + j * 4), "f"(v[j]) : "memory");
              continue;
            }
            // ReLU: on the packed halves below
            if (act == PCADV_ACT_LEAKY) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pk[j] = kBf16 ? pack_bf16x2(v[2 * j], v[2 * j + 1]) : pack_f16x2_sat(v[2 * j], v[2 * j + 1]);
            if (act == PCADV_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = relu_packed<kBf16>(pk[j]);
            }
            if (want_bits) {
              uint32_t w = 0u;
#pragma unroll
              for (int j = 0; j < 16; ++j) w |= c_gt0_mask<kBf16>(pk[j]) & (0x00010001u << j);
              obits[h2] = w;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              c_sts128(srow + (((h2 * 4 + q) ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
          if (rowmax || f32out) continue;
          if (want_bits) c_sts64(bits_s + (lane * kChainBitsWords + 2 * step) * 4, obits[0], obits[1]);
          if (step == 0) CHAIN_STAMP(eslot + 3);
          fence_proxy_async();                             // slab visible to the TMA store and the next MMA
          __syncwarp();
          if (p.has_out[l] && lane == 0) {
            c_tma_store(&maps.out[l], h_base + static_cast<uint32_t>(step) * kABytes + quarter * 4096u, step * 64,
                        row_tma);
            bulk_commit_group();
          }
        }
        if (rowmax) {
          if (r_ok) {
            const unsigned long long key = pack_key(best, static_cast<uint32_t>(best_col));
            if (p.serial) atomicMax(&p.rowmax_key[r], key);    // the two halves hold alternate steps
            else p.rowmax_key[r] = key;
          }
        } else if (f32out) {
          // n_f32 <= 64: one step, taken by half 0 in serial mode
          if (!p.serial || half == 0) {
            const int64_t wr0 = tm * kTileM + quarter * 32;
            int64_t nrows = p.rows - wr0;
            nrows = nrows > 32 ? 32 : nrows;
            if (nrows > 0) {
              const uint32_t bytes = static_cast<uint32_t>(nrows * p.n_f32 * 4);
              float* gdst = p.out_f32 + wr0 * p.n_f32;
              const uint32_t ssrc = h_base + static_cast<uint32_t>(quarter) * 8192u;
              if ((bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(gdst) & 15) == 0) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(ssrc), "r"(bytes) : "memory");
                  bulk_commit_group();
                }
              } else {
                __syncwarp();
                for (uint32_t e = lane; e < bytes / 4; e += 32) {
                  float val;
                  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val) : "r"(ssrc + e * 4));
                  gdst[e] = val;
                }
                __syncwarp();
              }
            }
          }
        } else if (want_bits) {
          __syncwarp();
          const int64_t wr0 = tm * kTileM + quarter * 32;
          for (int e = lane; e < 32 * steps; e += 32) {
            const int rr = e / steps, uu = e - rr * steps;
            if (p.serial && (uu & 1) != half) continue;    // the other half staged (and flushes) that step
            if (wr0 + rr < p.rows) {
              const uint2 w2 = c_lds64(bits_s + (rr * kChainBitsWords + 2 * uu) * 4);
              *reinterpret_cast<uint2*>(p.bits_out[l] + (wr0 + rr) * p.ld_bits[l] + 2 * uu) = w2;
            }
          }
          __syncwarp();
        }
        // H[half] = this layer's output, accumulator drained: the next MMA of this half may go
        CHAIN_STAMP(eslot + 4);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&st->in_ready[buf]);
        CHAIN_STAMP(eslot + 5);
      }
    }
    if (lane == 0) bulk_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tc
}  // namespace pcadv

using namespace pcadv;

static long long* g_chain_trace = nullptr;
// Debugging aid, not part of the ABI in include/pcadv.h: the next pcadv_chain launches record
// clock64 stamps of CTA 0 into `buf` ([10 warps][256 slots] int64, device memory); NULL switches off.
extern "C" void pcadv_debug_chain_trace(long long* buf) { g_chain_trace = buf; }

extern "C" int pcadv_chain(const pcadv_chain_args* a, void* stream) {
  using namespace tc;
  PCADV_CHECK_ARG(a && a->num_layers >= 2 && a->num_layers <= kChainMaxLayers, "pcadv_chain: 2..4 layers");
  PCADV_CHECK_ARG(a->rows >= 0 && a->x && (a->dtype == PCADV_F16 || a->dtype == PCADV_BF16),
                  "pcadv_chain: 16-bit input required");
  if (a->rows == 0) return 0;
  const int dt = a->dtype;
  ChainMaps maps;
  ChainParams p{};
  p.rows = a->rows; p.num_layers = a->num_layers;
  p.tiles = (a->rows + kTileM - 1) / kTileM;
  p.bf16 = dt == PCADV_BF16 ? 1 : 0;
  p.rowmax_key = a->rowmax_key;
  p.out_f32 = a->out_f32; p.n_f32 = a->n_f32;
  p.trace = g_chain_trace;
  PCADV_CHECK_ARG(!(a->rowmax_key && a->out_f32), "pcadv_chain: rowmax_key and out_f32 are exclusive");
  PCADV_CHECK_ARG(!a->out_f32 || (a->n_f32 >= 1 && a->n_f32 <= 64 && a->layer[a->num_layers - 1].n == 64 &&
                                  (reinterpret_cast<uintptr_t>(a->out_f32) & 3) == 0),
                  "pcadv_chain: out_f32 needs a last layer of (padded) width 64 and n_f32 <= 64");
  int max_n = 0, max_kn = 0;
  for (int l = 0; l < a->num_layers; ++l) {
    const pcadv_chain_layer& Ly = a->layer[l];
    const int k = l == 0 ? a->k0 : a->layer[l - 1].n;
    const bool last = l == a->num_layers - 1;
    PCADV_CHECK_ARG(k % 64 == 0 && k >= 64 && k <= 256 && Ly.n % 64 == 0 && Ly.n >= 64 && Ly.n <= 256,
                    "pcadv_chain: layer %d: k and n must be multiples of 64 in [64, 256] (k=%d n=%d)", l, k, Ly.n);
    PCADV_CHECK_ARG(Ly.w && tma_compatible(Ly.w, dt, Ly.ldw), "pcadv_chain: layer %d weight not TMA-compatible", l);
    PCADV_CHECK_ARG(last || Ly.out, "pcadv_chain: every layer but the last must store its output");
    PCADV_CHECK_ARG(!(last && (a->rowmax_key || a->out_f32)) || (!Ly.out && !Ly.bits_out),
                    "pcadv_chain: a row-max / fp32 last layer has no 16-bit output");
    PCADV_CHECK_ARG(last ? (Ly.out || a->rowmax_key || a->out_f32) : true,
                    "pcadv_chain: last layer needs an output, rowmax_key or out_f32");
    p.k[l] = k; p.n[l] = Ly.n; p.act[l] = Ly.act; p.slope[l] = Ly.slope; p.bias[l] = Ly.bias;
    p.nvalid[l] = (last && a->out_f32) ? a->n_f32 : Ly.n;     // rows the weight / bias really have
    p.idesc[l] = make_idesc(kTileM, Ly.n, dt == PCADV_BF16, false, false);
    if (int rc = encode_tmap_2d(&maps.w[l], Ly.w, dt, p.nvalid[l], k, Ly.ldw, kBlockK, Ly.n)) return rc;
    p.has_out[l] = Ly.out ? 1 : 0;
    if (Ly.out) {
      PCADV_CHECK_ARG(tma_compatible(Ly.out, dt, Ly.ld_out), "pcadv_chain: layer %d output not TMA-storable", l);
      if (int rc = encode_tmap_2d(&maps.out[l], Ly.out, dt, a->rows, Ly.n, Ly.ld_out, 64, 32)) return rc;
    }
    PCADV_CHECK_ARG(!Ly.bits_out || (Ly.ld_bits % 2 == 0 && (reinterpret_cast<uintptr_t>(Ly.bits_out) & 7) == 0),
                    "pcadv_chain: layer %d sign-bit rows must be 8-byte aligned", l);
    p.bits_out[l] = Ly.bits_out; p.ld_bits[l] = Ly.ld_bits;
    if (!(last && (a->rowmax_key || a->out_f32))) max_n = Ly.n > max_n ? Ly.n : max_n;
    max_kn = Ly.n > max_kn ? Ly.n : max_kn;
  }
  PCADV_CHECK_ARG(tma_compatible(a->x, dt, a->ldx), "pcadv_chain: input not TMA-compatible");
  if (int rc = encode_tmap_2d(&maps.x, a->x, dt, a->rows, a->k0, a->ldx, kBlockK, kTileM)) return rc;
  if (a->out_f32 && max_n < 128) max_n = 128;                   // 4 x 8 KB of fp32 row staging
  p.h_bytes = kTileM * (max_n > 64 ? max_n : 64) * 2;           // one [128 x max_n] tile per half
  p.stage_bytes = kABytes + max_kn * kBlockK * 2;
  // two tiles in flight need two tiles of shared memory; chains with a 256-wide intermediate run
  // one tile at a time and let both epilogue halves split its steps
  p.serial = chain_smem_bytes(2, p.stage_bytes, p.h_bytes) > static_cast<size_t>(kChainSmemMax) ? 1 : 0;
  if (p.serial) p.h_bytes /= 2;                                 // carve_chain lays out 2 * h_bytes
  p.groups = 2;
  // A layer of a tile is one dependency chain (MMA -> wake -> epilogue -> wake -> next MMA, ~5 600 cycles,
  // profiles/r01h_chain_trace_cta0.txt) and only whole tiles run side by side: with every accumulator
  // <= 128 columns wide TMEM holds a third tile, and a third group of four epilogue warps takes it.
  static int want_groups = -1;
  if (want_groups < 0) { const char* e = getenv("PCADV_CHAIN_GROUPS"); want_groups = e ? atoi(e) : 3; }   // tuning aid
  if (!p.serial && want_groups >= 3 && max_kn <= 128 &&
      chain_smem_bytes(3, p.stage_bytes, p.h_bytes, 3) <= static_cast<size_t>(kChainSmemMax))
    p.groups = 3;
  // (a fourth group fits for the discriminator's 64-wide chain and measured slower: 0.176 against 0.156 ms)
  int nst = kChainMaxStages;
  while (nst > 2 && chain_smem_bytes(nst, p.stage_bytes, p.h_bytes, p.groups) > static_cast<size_t>(kChainSmemMax)) --nst;
  p.nstages = nst;
  const size_t smem = chain_smem_bytes(nst, p.stage_bytes, p.h_bytes, p.groups);
  PCADV_CHECK_ARG(smem <= static_cast<size_t>(kChainSmemMax), "pcadv_chain: shared memory budget exceeded");
  static bool attr_done = false;
  if (!attr_done) {
    PCADV_CUDA_OK(cudaFuncSetAttribute(tc_chain_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemMax));
    PCADV_CUDA_OK(cudaFuncSetAttribute(tc_chain_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemMax));
    PCADV_CUDA_OK(cudaFuncSetAttribute(tc_chain_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemMax));
    PCADV_CUDA_OK(cudaFuncSetAttribute(tc_chain_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemMax));
    attr_done = true;
  }
  const int grid = static_cast<int>(p.tiles < num_sms() ? p.tiles : num_sms());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.groups == 3) {
    if (dt == PCADV_BF16) tc_chain_kernel<true, 3><<<grid, chain_threads(3), smem, s>>>(maps, p);
    else tc_chain_kernel<false, 3><<<grid, chain_threads(3), smem, s>>>(maps, p);
  } else if (dt == PCADV_BF16) tc_chain_kernel<true, 2><<<grid, chain_threads(2), smem, s>>>(maps, p);
  else tc_chain_kernel<false, 2><<<grid, chain_threads(2), smem, s>>>(maps, p);
  PCADV_LAUNCHED();
  return 0;
}
