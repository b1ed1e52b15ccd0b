// Tensor-core engine, forward / dgrad GEMMs with points on the MMA M axis.
//
//   tc_rows_kernel<act, out>: out[r, c] = epilogue(sum_k x[r, k] w[c, k]) for 128-row tiles.
//
// 320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 =
// epilogue.  The two 256-column TMEM accumulator buffers belong to the two epilogue *halves*
// (warps 2..5 / 6..9): half h drains the CTA's tiles h, h+2, h+4, ... so two tiles are in the
// epilogue at once while the MMA warp fills the next buffer.  Inside a half every warp owns a
// TMEM lane quarter (32 rows) and works alone: it stages its 32 x 128-byte output slab in its
// own shared-memory ring and one lane hands the slab to a TMA store; the activation-derivative
// mask slab of the step arrives by TMA into the SAME slab and is overwritten in place by the
// output.  No block-wide barrier exists in the steady state -- warps only meet when the staged
// bias of their half changes.
#include "tc_pipeline.cuh"

namespace pcadv {
namespace tc {

constexpr int kRowsThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kWarpSlabBytes = 32 * 128;          // one warp's 32 rows x 128 B, 128B-swizzled
constexpr int kMaxSlabs = 3;
constexpr int kRowsMaxStages = 6;
constexpr int kRowsBiasFloats = 2 * 3 * kMaxTileN;   // per half: bias + two per-cloud bias rows
constexpr int kRowsSmemMax = 232448;              // 227 KB opt-in limit of sm_100
constexpr int kBitsWords = kMaxTileN / 32;        // sign-bit words per tile row
constexpr int kBitsBytes = kEpiWarps * 32 * kBitsWords * 4;   // per-warp [32 rows][8 words] staging

struct RowsTail {
  uint64_t full[kRowsMaxStages];
  uint64_t empty[kRowsMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t mask_full[kEpiWarps][kMaxSlabs];
  uint32_t tmem_base;
};

struct RowsParams {
  int64_t rows;
  int n;
  int bn;                 // N-side tile (multiple of 16, <= 256)
  int num_seg;
  int seg_k[PCADV_MAX_SEG];
  int64_t tiles_m, tiles_n;
  uint32_t idesc;
  int nstages, nslabs, stage_bytes;
  float* group_sum;       // lean dgrad variant with column-sum warps: per-cloud column sums of segment 0
  int pair;               // CTA pairs (cluster of 2): each CTA fetches half of a weight tile and multicasts it to both
  int epi_halves;         // lean kernel: 2 = two epilogue halves (one TMEM buffer each), 1 = one half drains both
  // epilogue
  const float* bias;
  const float* group_bias;
  int64_t rows_per_group;
  const float* addend;
  int64_t ld_addend;
  float slope;
  const void* mask;
  int64_t ld_mask;
  int mask_dtype, mask_act;
  float mask_slope;
  const float* out_scale;
  void* out;
  int64_t ld_out;
  unsigned long long* rowmax_key;
  int tma_out;            // output rows are TMA-storable: swizzled slab + tensor-map store
  int tma_mask;           // mask slab fetched by TMA into the output slab (in place)
  int compact_out;        // ld_out == n: a warp's 32 rows are one contiguous span (1-D bulk store)
  uint32_t* bits_out;     // [rows, ld_bits_out] sign bits of the stored output (n % 64 == 0)
  int64_t ld_bits_out;
  const uint32_t* mask_bits;   // 1-bit form of the mask
  int64_t ld_mask_bits;
};

struct RowsSmem {
  uint8_t* stages;
  uint8_t* epi;
  float* bias;
  uint32_t* bits;
  RowsTail* tail;
};

__device__ __forceinline__ RowsSmem carve_rows(uint8_t* raw, const RowsParams& p) {
  RowsSmem L;
  L.stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  L.epi = L.stages + p.nstages * p.stage_bytes;
  L.bias = reinterpret_cast<float*>(L.epi + (p.epi_halves == 1 ? 4 : kEpiWarps) * p.nslabs * kWarpSlabBytes);
  L.bits = reinterpret_cast<uint32_t*>(L.bias + kRowsBiasFloats);
  L.tail = reinterpret_cast<RowsTail*>(reinterpret_cast<uint8_t*>(L.bits) + kBitsBytes);
  return L;
}

static size_t rows_smem_bytes(int nstages, int stage_bytes, int nslabs, int epi_warps = kEpiWarps) {
  return 1024 + static_cast<size_t>(nstages) * stage_bytes + epi_warps * nslabs * kWarpSlabBytes +
         kRowsBiasFloats * 4 + kBitsBytes + sizeof(RowsTail) + 16;
}

// 1-D bulk copy shared -> global (bytes and both addresses multiples of 16)
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void unpack16(const uint4 t4, int dtype, float* m) {
  const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 f;
    if (dtype == PCADV_F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
    else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
    m[2 * e] = f.x; m[2 * e + 1] = f.y;
  }
}

// ---- roles shared by the general and the lean kernel -------------------------------------------
__device__ __forceinline__ uint32_t rows_setup(const TensorMaps& maps, const RowsParams& p, RowsTail* st,
                                               int warp, int lane) {
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_seg; ++s) tma_prefetch_desc(&maps.act[s]);
    tma_prefetch_desc(&maps.w);
    if (p.tma_out) tma_prefetch_desc(&maps.out);
    if (p.tma_mask) tma_prefetch_desc(&maps.mask);
    if (p.pair) tma_prefetch_desc(&maps.w_half);
    for (int i = 0; i < kRowsMaxStages; ++i) {
      mbar_init(&st->full[i], 1);
      // pairs: a stage is refilled (partly by the peer's multicast) once BOTH CTAs have consumed it
      mbar_init(&st->empty[i], (p.pair ? 2 : 1) + (p.group_sum ? 4 : 0));   // + the four column-sum warps
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st->tmem_full[i], 1);
      mbar_init(&st->tmem_empty[i], 4);
    }
    for (int w = 0; w < kEpiWarps; ++w)
      for (int i = 0; i < kMaxSlabs; ++i) mbar_init(&st->mask_full[w][i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&st->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (p.pair) cluster_barrier_sync();                 // the peer's barriers exist before anything is sent to them
  tc_fence_after();
  return st->tmem_base;
}

// Work items of a CTA.  Single CTAs walk the (row tile, column tile) list with stride gridDim.x.  In pair
// mode the two CTAs of a cluster take the row tiles 2 i and 2 i + 1 of the same column tile side by side
// (same weight tile, fetched once per pair), and pairs walk the list of row-tile PAIRS.
struct RowsWork {
  int64_t first, stride, count, tiles_n;
  int pair, rank;
  __device__ __forceinline__ int64_t tm(int64_t t) const { return pair ? 2 * (t / tiles_n) + rank : t / tiles_n; }
  __device__ __forceinline__ int64_t tn(int64_t t) const { return t % tiles_n; }
};
__device__ __forceinline__ RowsWork rows_work(const RowsParams& p) {
  RowsWork w;
  w.pair = p.pair;
  w.tiles_n = p.tiles_n;
  if (p.pair) {
    w.rank = static_cast<int>(cluster_cta_rank());
    w.first = blockIdx.x >> 1;
    w.stride = gridDim.x >> 1;
    w.count = ((p.tiles_m + 1) >> 1) * p.tiles_n;
  } else if (p.group_sum) {
    // column-sum variant: every CTA takes a CONTIGUOUS range of items, so that consecutive row tiles of a CTA lie
    // in the same cloud and the per-cloud sums are reduced and flushed once per cloud, not once per tile
    const int64_t total = p.tiles_m * p.tiles_n;
    const int64_t per = (total + gridDim.x - 1) / gridDim.x;
    w.rank = 0;
    w.first = blockIdx.x * per;
    w.stride = 1;
    w.count = w.first + per < total ? w.first + per : total;
  } else {
    w.rank = 0;
    w.first = blockIdx.x;
    w.stride = gridDim.x;
    w.count = p.tiles_m * p.tiles_n;
  }
  return w;
}

__device__ __forceinline__ void rows_producer(const TensorMaps& maps, const RowsParams& p,
                                              const RowsSmem& L, RowsTail* st, int64_t num_tiles) {
  const RowsWork wk = rows_work(p);
  const uint32_t stage_tx = static_cast<uint32_t>((kTileM + p.bn) * kBlockK * 2);
  const int half_rows = p.bn >> 1;
  int stage = 0;
  uint32_t phase = 0;
  for (int64_t t = wk.first; t < wk.count; t += wk.stride) {
    const int32_t m0 = static_cast<int32_t>(wk.tm(t) * kTileM);   // past the last row in an odd tail: zero fill
    const int32_t n0 = static_cast<int32_t>(wk.tn(t) * p.bn);
    int kg = 0;
    for (int s = 0; s < p.num_seg; ++s) {
      for (int kk = 0; kk < p.seg_k[s]; kk += kBlockK, kg += kBlockK) {
        mbar_wait_backoff(&st->empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&st->full[stage], stage_tx);
        uint8_t* sa = L.stages + stage * p.stage_bytes;
        tma_load_2d(sa, &maps.act[s], &st->full[stage], kk, m0);
        if (p.pair)      // this CTA's half of the weight tile, delivered to both CTAs of the pair
          tma_load_2d_multicast(sa + kABytes + wk.rank * half_rows * 128, &maps.w_half, &st->full[stage], kg,
                                n0 + wk.rank * half_rows, 3);
        else
          tma_load_2d(sa + kABytes, &maps.w, &st->full[stage], kg, n0);
        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
      }
    }
  }
}

__device__ __forceinline__ void rows_mma(const RowsParams& p, const RowsSmem& L, RowsTail* st,
                                         uint32_t tmem_base, int64_t num_tiles) {
  const RowsWork wk = rows_work(p);
  int total_chunks = 0;
  for (int s = 0; s < p.num_seg; ++s) total_chunks += p.seg_k[s] / kBlockK;
  int stage = 0;
  uint32_t phase = 0;
  int buf = 0;
  uint32_t buf_phase = 0;
  for (int64_t t = wk.first; t < wk.count; t += wk.stride) {
    mbar_wait_backoff(&st->tmem_empty[buf], buf_phase ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kMaxTileN);
    for (int c = 0; c < total_chunks; ++c) {
      mbar_wait_backoff(&st->full[stage], phase);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(L.stages + stage * p.stage_bytes);
      mma_chunk_kmajor(d_tmem, a_addr, a_addr + kABytes, p.idesc, c == 0);
      if (p.pair) umma_commit_multicast(&st->empty[stage], 3);
      else umma_commit(&st->empty[stage]);
      if (++stage == p.nstages) { stage = 0; phase ^= 1; }
    }
    umma_commit(&st->tmem_full[buf]);
    if (++buf == 2) { buf = 0; buf_phase ^= 1; }
  }
}

__device__ __forceinline__ void rows_teardown(int warp, uint32_t tmem_base, int pair = 0) {
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_barrier_sync();                   // the peer may still signal this CTA's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// explicit shared-space accesses (the epilogue's pointers are carved from one dynamic buffer and
// would otherwise compile to generic LD / ST)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tma_store_2d_addr(const CUtensorMap* m, uint32_t smem_addr, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_addr), "r"(c0), "r"(c1)
      : "memory");
}

// ---- sign-bit maps -----------------------------------------------------------------------------
// Word w of a row covers columns 32 w .. 32 w + 31; column 2 k sits at bit k, column 2 k + 1 at bit
// 16 + k (k = 0..15): the two halves of packed 16-bit pair k map to the two halves of the word, so
// one HSET2 + one LOP3 per pair builds it and one LOP3 + one FSEL per element applies it.
template <int kOut>
__device__ __forceinline__ uint32_t pair_gt0_mask(uint32_t packed) {
  // 0xffff in each half whose 16-bit value is > 0
  uint32_t m;
  if (kOut == PCADV_F16) asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(packed), "r"(0u));
  else asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(packed), "r"(0u));
  return m;
}

constexpr int kAddNone = 0, kAddBias = 1, kAddBiasGroup = 2;

// =====================================================================================
// Lean variant for the layers that dominate the step: 16-bit TMA-stored output, n % 64 == 0, no
// addend / row-max / output scale.  kAdd: what is added to the accumulator (nothing -- dgrad;
// bias -- forward layers; bias + per-cloud bias -- fc1).  kMaskBits: multiply by act'(.) read from
// the forward layer's sign-bit map (dgrad).  Emits the sign-bit map of its own output on request.
constexpr int kRowsSumThreads = kRowsThreads + 128;   // + warps 10..13: column sums of the A tiles in flight

template <bool kBf16>
__device__ __forceinline__ void rows_add_pair(uint32_t packed, float& a0, float& a1) {
  if (kBf16) {
    asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nadd.rn.f32.bf16 %0, lo, %0;\nadd.rn.f32.bf16 %1, hi, %1;\n}"
        : "+f"(a0), "+f"(a1) : "r"(packed));
  } else {
    asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nadd.rn.f32.f16 %0, lo, %0;\nadd.rn.f32.f16 %1, hi, %1;\n}"
        : "+f"(a0), "+f"(a1) : "r"(packed));
  }
}

// Warps 10..13 of the kGroupSum variant: per-cloud column sums of segment 0 (<= 256 channels), taken from
// its [128 rows][64 channels] tiles while they sit in the ring (first column tile of a row tile only; a row
// tile lies inside one cloud).  Eight consecutive lanes read the eight 16-byte groups of one row (conflict
// free under the 128-byte swizzle), a warp covers 4 rows per trip, the four warps 16; lanes are then folded
// by two shuffles, warps through shared memory (the bias staging area, unused by a dgrad), and one atomic
// per channel and row tile goes to group_sum.
template <int kOut>
__device__ __forceinline__ void rows_group_sums(const RowsParams& p, const RowsSmem& L, RowsTail* st, int warp, int lane) {
  const RowsWork wk = rows_work(p);
  const int cw = warp - 10;
  const uint32_t g8 = static_cast<uint32_t>(lane) & 7u;     // 16-byte group = channels 8 g8 .. 8 g8 + 7
  const int rsub = lane >> 3;                              // row inside the warp's 4-row trip
  const int chunks0 = p.seg_k[0] / kBlockK;                // <= 4
  int total_chunks = 0;
  for (int s = 0; s < p.num_seg; ++s) total_chunks += p.seg_k[s] / kBlockK;
  float* red = L.bias;                                     // [4 warps][256 channels]
  const int64_t rpg = p.rows_per_group;
  int stage = 0;
  uint32_t phase = 0;
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
  int64_t cur_g = -1;
  auto flush = [&]() {                                     // warp-uniform: all four warps walk the same items
    if (cur_g < 0) return;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = acc[c4][e];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[c4][e] = v;
      }
    named_barrier_sync(5, 128);                            // the previous flush's readers are done with red
    if (lane < 8) {
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
        if (c4 < chunks0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) red[cw * 256 + c4 * 64 + lane * 8 + e] = acc[c4][e];
        }
    }
    named_barrier_sync(5, 128);
    for (int ch = cw * 32 + lane; ch < p.seg_k[0]; ch += 128)
      atomicAdd(p.group_sum + cur_g * p.seg_k[0] + ch, red[ch] + red[256 + ch] + red[512 + ch] + red[768 + ch]);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
  };
  for (int64_t t = wk.first; t < wk.count; t += wk.stride) {
    const int64_t tm = wk.tm(t);
    const bool take = wk.tn(t) == 0 && tm * kTileM < p.rows;
    if (take) {
      const int64_t g = (tm * kTileM) / rpg;
      if (g != cur_g) { flush(); cur_g = g; }
    }
    for (int c = 0; c < total_chunks; ++c) {
      mbar_wait(&st->full[stage], phase);
      if (take && c < chunks0) {
        const uint8_t* tile = L.stages + stage * p.stage_bytes;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          if (c4 == c) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t rr = static_cast<uint32_t>(16 * i + 4 * cw + rsub);
              const uint4 q = *reinterpret_cast<const uint4*>(tile + rr * 128 + ((g8 ^ (rr & 7u)) << 4));
              rows_add_pair<kOut == PCADV_BF16>(q.x, acc[c4][0], acc[c4][1]);
              rows_add_pair<kOut == PCADV_BF16>(q.y, acc[c4][2], acc[c4][3]);
              rows_add_pair<kOut == PCADV_BF16>(q.z, acc[c4][4], acc[c4][5]);
              rows_add_pair<kOut == PCADV_BF16>(q.w, acc[c4][6], acc[c4][7]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->empty[stage]);
      if (++stage == p.nstages) { stage = 0; phase ^= 1; }
    }
  }
  flush();
}

template <int kAct, int kOut, int kAdd, bool kMaskBits, bool kRowMax = false, bool kGroupSum = false>
__global__ void __launch_bounds__(kGroupSum ? kRowsSumThreads : kRowsThreads, 1)
tc_rows_lean_kernel(const __grid_constant__ TensorMaps maps, const RowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const RowsSmem L = carve_rows(smem_raw, p);
  RowsTail* st = L.tail;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_tiles = p.tiles_m * p.tiles_n;
  const uint32_t tmem_base = rows_setup(maps, p, st, warp, lane);

  if (warp == 0) {
    if (lane == 0) rows_producer(maps, p, L, st, num_tiles);
  } else if (warp == 1) {
    if (lane == 0) rows_mma(p, L, st, tmem_base, num_tiles);
  } else if (kGroupSum && warp >= 10) {
    rows_group_sums<kOut>(p, L, st, warp, lane);
  } else {
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int lane_row = quarter * 32 + lane;
    const int hid = (ew & 3) * 32 + lane;
    const int steps = p.bn >> 6;                          // bn is a multiple of 64 here
    const int tiles_n = static_cast<int>(p.tiles_n);
    const int n = p.n, bn = p.bn;
    const int halves = p.epi_halves;                      // 1: this half (0) drains both TMEM buffers
    const int nslabs = p.nslabs;
    const uint32_t slab0 = smem_u32(L.epi + ew * nslabs * kWarpSlabBytes) + lane * 128;
    const uint32_t slab_tma0 = smem_u32(L.epi + ew * nslabs * kWarpSlabBytes);
    const uint32_t bits_s = smem_u32(L.bits + ew * 32 * kBitsWords);
    float* bias_s = L.bias + half * 3 * kMaxTileN;
    const uint32_t bias_a = smem_u32(bias_s);
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const float slope = p.slope;
    const float mneg = p.mask_act == PCADV_ACT_LEAKY ? p.mask_slope : 0.f;
    const int64_t rpg = p.rows_per_group > 0 ? p.rows_per_group : p.rows;
    const bool want_bits = p.bits_out != nullptr;
    uint32_t slab = 0;
    int staged_col = -1;
    int64_t staged_g0 = -1, staged_g1 = -1;
    uint32_t use = 0;
    // sign-bit words of this thread's row for a whole tile (<= 4 steps x 2 words), fetched one
    // tile ahead so that their latency hides behind the current tile
    uint2 mbw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) mbw[i] = make_uint2(0u, 0u);
    const RowsWork wk = rows_work(p);
    auto fetch_bits = [&](int64_t tt) {
      if (tt >= wk.count) return;
      const int64_t ttm = wk.tm(tt);
      const int ttn = static_cast<int>(wk.tn(tt));
      const int64_t rr = ttm * kTileM + lane_row;
      if (rr >= p.rows) return;
      const uint2* src = reinterpret_cast<const uint2*>(p.mask_bits + rr * p.ld_mask_bits + ((ttn * bn) >> 5));
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < steps) mbw[i] = __ldg(src + i);
    };
    const int64_t t_begin = half < halves ? wk.first + static_cast<int64_t>(half) * wk.stride : wk.count;
    const int64_t t_stride = static_cast<int64_t>(halves) * wk.stride;
    if (kMaskBits) fetch_bits(t_begin);
    for (int64_t t = t_begin; t < wk.count; t += t_stride, ++use) {
      const int64_t tm = wk.tm(t);
      const int tn = static_cast<int>(wk.tn(t));
      const int64_t r = tm * kTileM + lane_row;
      const bool r_ok = r < p.rows;
      const int col_base = tn * bn;
      const int32_t row_tma = static_cast<int32_t>(tm * kTileM + quarter * 32);
      uint32_t gsel = 0;                                  // byte offset of this row's per-cloud bias
      if (kAdd != kAddNone) {
        // (a pair's odd tail tile lies past the last row: it stages the last cloud's bias and stores nothing)
        const int64_t row_first = tm * kTileM < p.rows ? tm * kTileM : p.rows - 1;
        const int64_t row_last = row_first + kTileM - 1 < p.rows ? row_first + kTileM - 1 : p.rows - 1;
        const int64_t g_first = row_first / rpg, g_last = row_last / rpg;
        const bool restage = col_base != staged_col ||
                             (kAdd == kAddBiasGroup && (g_first != staged_g0 || g_last != staged_g1));
        if (restage) {
          staged_col = col_base; staged_g0 = g_first; staged_g1 = g_last;
          named_barrier_sync(1 + half, 128);
          for (int e = hid; e < kMaxTileN; e += 128) {
            const int c = col_base + e;
            const bool c_ok = e < bn && c < n;
            bias_s[e] = (c_ok && p.bias) ? __ldg(p.bias + c) : 0.f;
            if (kAdd == kAddBiasGroup) {
              bias_s[kMaxTileN + e] = c_ok ? __ldg(p.group_bias + g_first * n + c) : 0.f;
              bias_s[2 * kMaxTileN + e] = (c_ok && g_last != g_first) ? __ldg(p.group_bias + g_last * n + c) : 0.f;
            }
          }
          named_barrier_sync(1 + half, 128);
        }
        if (kAdd == kAddBiasGroup)
          gsel = (r_ok && r / rpg != g_first) ? 2u * kMaxTileN * 4u : 1u * kMaxTileN * 4u;
      }
      uint2 mbc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) mbc[i] = mbw[i];
      if (kMaskBits) fetch_bits(t + t_stride);
      const int buf = halves == 2 ? half : static_cast<int>(use & 1);
      mbar_wait(&st->tmem_full[buf], halves == 2 ? (use & 1) : ((use >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(buf * kMaxTileN);
      float best = -INFINITY;                             // kRowMax: running (max, first column)
      int best_col = 0;
#pragma unroll 1
      for (int step = 0; step < steps; ++step) {
        const uint32_t srow = slab0 + slab * kWarpSlabBytes;
        if (!kRowMax) {
          // the store that last read this slab (two steps ago; the previous one with one slab) is done with it
          if (lane == 0) {
            if (nslabs == 2) bulk_wait_group_read<1>();
            else bulk_wait_group_read<0>();
          }
          __syncwarp();
        }
        uint32_t raw[2][32];
        tmem_ld32_issue(taddr0 + step * 64, raw[0]);
        tmem_ld32_issue(taddr0 + step * 64 + 32, raw[1]);
        uint2 mb = mbc[0];
        if (kMaskBits) {
          if (step == 1) mb = mbc[1];
          else if (step == 2) mb = mbc[2];
          else if (step == 3) mb = mbc[3];
        }
        tmem_ld32_wait(raw[0]);
        tmem_ld32_wait(raw[1]);
        uint32_t obits[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[h][j]);
          if (kAdd != kAddNone) {
            const uint32_t ba = bias_a + static_cast<uint32_t>(step * 64 + h * 32) * 4u;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b = lds128(ba + q * 16);
              v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
            }
            if (kAdd == kAddBiasGroup) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 b = lds128(ba + gsel + q * 16);
                v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
              }
            }
          }
          if (kRowMax) {
            // max over the row's channels of the PRE-activation value, first column on ties
            float m = v[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
            if (m > best) {
              int jj = 0;
#pragma unroll
              for (int j = 31; j >= 0; --j) jj = (v[j] == m) ? j : jj;
              best = m;
              best_col = col_base + step * 64 + h * 32 + jj;
            }
            continue;
          }
          // ReLU is applied to the packed halves below (one HMNMX2 per pair instead of two FMNMX)
          if (kAct == PCADV_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
          }
          if (kMaskBits) {
            const uint32_t word = h == 0 ? mb.x : mb.y;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = (word & (1u << ((j >> 1) + 16 * (j & 1)))) ? v[j] : v[j] * mneg;
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            pk[j] = kOut == PCADV_F16 ? pack_f16x2_sat(v[2 * j], v[2 * j + 1]) : pack_bf16x2(v[2 * j], v[2 * j + 1]);
          if (kAct == PCADV_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = relu_packed<kOut == PCADV_BF16>(pk[j]);
          }
          if (want_bits) {
            uint32_t w = 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) w |= pair_gt0_mask<kOut>(pk[j]) & (0x00010001u << j);
            obits[h] = w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(srow + (((h * 4 + q) ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        if (kRowMax) continue;
        if (want_bits) sts64(bits_s + (lane * kBitsWords + 2 * step) * 4, obits[0], obits[1]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d_addr(&maps.out, slab_tma0 + slab * kWarpSlabBytes, col_base + step * 64, row_tma);
          bulk_commit_group();
        }
        if (nslabs == 2) slab ^= 1;
      }
      if (kRowMax && r_ok) {
        const unsigned long long key = pack_key(best, static_cast<uint32_t>(best_col));
        if (tiles_n == 1) p.rowmax_key[r] = key;          // the tile holds the whole row
        else atomicMax(&p.rowmax_key[r], key);
      }
      if (want_bits) {
        // the warp's [32 rows][bn / 32 words] sign-bit tile, consecutive lanes on consecutive
        // 8-byte pairs of a row
        __syncwarp();
        const int upr = steps;                             // 8-byte pairs per row in this tile
        const int64_t wr0 = tm * kTileM + quarter * 32;
        for (int i = lane; i < 32 * upr; i += 32) {
          const int rr = i / upr, uu = i - rr * upr;
          if (wr0 + rr < p.rows) {
            const uint2 w2 = lds64(bits_s + (rr * kBitsWords + 2 * uu) * 4);
            *reinterpret_cast<uint2*>(p.bits_out + (wr0 + rr) * p.ld_bits_out + (col_base >> 5) + 2 * uu) = w2;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->tmem_empty[buf]);
    }
    if (lane == 0) bulk_wait_group<0>();
  }
  rows_teardown(warp, tmem_base, p.pair);
}

typedef void (*RowsKernel)(const TensorMaps, const RowsParams);

template <int kOut>
static RowsKernel pick_lean(int act, int add, bool maskbits) {
  if (maskbits) return (act == PCADV_ACT_NONE && add == kAddNone) ? tc_rows_lean_kernel<PCADV_ACT_NONE, kOut, kAddNone, true> : nullptr;
  if (act == PCADV_ACT_RELU && add == kAddBias) return tc_rows_lean_kernel<PCADV_ACT_RELU, kOut, kAddBias, false>;
  if (act == PCADV_ACT_RELU && add == kAddBiasGroup) return tc_rows_lean_kernel<PCADV_ACT_RELU, kOut, kAddBiasGroup, false>;
  if (act == PCADV_ACT_LEAKY && add == kAddBias) return tc_rows_lean_kernel<PCADV_ACT_LEAKY, kOut, kAddBias, false>;
  if (act == PCADV_ACT_NONE && add == kAddNone) return tc_rows_lean_kernel<PCADV_ACT_NONE, kOut, kAddNone, false>;
  if (act == PCADV_ACT_NONE && add == kAddBias) return tc_rows_lean_kernel<PCADV_ACT_NONE, kOut, kAddBias, false>;
  return nullptr;
}
template <int kOut>
static RowsKernel lean_group_sum() { return tc_rows_lean_kernel<PCADV_ACT_NONE, kOut, kAddNone, true, false, true>; }
// max over channels only (no stored output): the discriminators' last layer
static RowsKernel lean_rowmax() { return tc_rows_lean_kernel<PCADV_ACT_NONE, PCADV_F16, kAddBias, false, true>; }

// =====================================================================================
// kAct: PCADV_ACT_*;  kOut: PCADV_F32 / PCADV_F16 / PCADV_BF16
template <int kAct, int kOut>
__global__ void __launch_bounds__(kRowsThreads, 1)
tc_rows_kernel(const __grid_constant__ TensorMaps maps, const RowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  const RowsSmem L = carve_rows(smem_raw, p);
  RowsTail* st = L.tail;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_tiles = p.tiles_m * p.tiles_n;

  const uint32_t tmem_base = rows_setup(maps, p, st, warp, lane);

  if (warp == 0) {
    if (lane == 0) rows_producer(maps, p, L, st, num_tiles);
  } else if (warp == 1) {
    if (lane == 0) rows_mma(p, L, st, tmem_base, num_tiles);
  } else {
    // ================= epilogue: two halves x four lane quarters =================
    const int ew = warp - 2;                              // 0..7
    const int quarter = warp & 3;                         // TMEM lane quarter of this warp
    const int half = ew >> 2;                             // which accumulator buffer / tile parity
    const int lane_row = quarter * 32 + lane;             // row of the 128-row tile
    const int hid = (ew & 3) * 32 + lane;                 // 0..127 inside the half
    constexpr int kEsz = kOut == PCADV_F32 ? 4 : 2;
    constexpr int kStepCols = 128 / kEsz;                 // columns per 128-byte slab row
    constexpr int kChunks = kStepCols / 32;               // 32-column TMEM loads per step
    const int steps = (p.bn + kStepCols - 1) / kStepCols;
    const int S = p.nslabs;
    uint8_t* myslabs = L.epi + ew * S * kWarpSlabBytes;
    float* bias_s = L.bias + half * 3 * kMaxTileN;
    float* gb_s = bias_s + kMaxTileN;                     // [2][kMaxTileN]
    const float oscale = p.out_scale ? *p.out_scale : 1.f;
    const float mneg = p.mask_act == PCADV_ACT_LEAKY ? p.mask_slope : 0.f;
    const int64_t rpg = p.rows_per_group > 0 ? p.rows_per_group : p.rows;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);  // swizzle term of this thread's slab row

    int slab = 0;                                         // ring position of the next step
    uint32_t slab_par = 0;                                // parity of the slab's mask barrier
    // mask prefetch iterator (lane 0): S - 1 steps ahead of the consumer
    int64_t pf_t = blockIdx.x + static_cast<int64_t>(half) * gridDim.x;
    int pf_step = 0, pf_slab = 0;
    auto issue_mask = [&]() {
      if (pf_t >= num_tiles) return;
      const int64_t ptm = pf_t / p.tiles_n, ptn = pf_t % p.tiles_n;
      uint64_t* bar = &st->mask_full[ew][pf_slab];
      mbar_arrive_expect_tx(bar, kWarpSlabBytes);
      tma_load_2d(myslabs + pf_slab * kWarpSlabBytes, &maps.mask, bar,
                  static_cast<int32_t>(ptn * p.bn + pf_step * 64),
                  static_cast<int32_t>(ptm * kTileM + quarter * 32));
      if (++pf_slab == S) pf_slab = 0;
      if (++pf_step == steps) { pf_step = 0; pf_t += 2 * static_cast<int64_t>(gridDim.x); }
    };
    if (p.tma_mask && lane == 0)
      for (int i = 0; i < S - 1; ++i) issue_mask();

    int staged_col = -1;
    int64_t staged_g0 = -1, staged_g1 = -1;
    uint32_t use = 0;
    for (int64_t t = blockIdx.x + static_cast<int64_t>(half) * gridDim.x; t < num_tiles;
         t += 2 * static_cast<int64_t>(gridDim.x), ++use) {
      const int64_t tm = t / p.tiles_n, tn = t % p.tiles_n;
      const int64_t r = tm * kTileM + lane_row;
      const bool r_ok = r < p.rows;
      const int col_base = static_cast<int>(tn * p.bn);
      const int64_t row_first = tm * kTileM;
      const int64_t row_last = row_first + kTileM - 1 < p.rows ? row_first + kTileM - 1 : p.rows - 1;
      const int64_t g_first = row_first / rpg, g_last = row_last / rpg;
      const int64_t g = r_ok ? r / rpg : g_first;
      const bool gb_staged = p.group_bias != nullptr && (g_last - g_first) <= 1;

      // ---- stage bias (and the tile's per-cloud bias rows) for this half; skipped while the
      // staged values are still the ones the tile needs
      const bool restage = col_base != staged_col ||
                           (gb_staged && (g_first != staged_g0 || g_last != staged_g1));
      if (restage) {
        staged_col = col_base; staged_g0 = g_first; staged_g1 = g_last;
        named_barrier_sync(1 + half, 128);               // previous tile's readers are done
        for (int e = hid; e < kMaxTileN; e += 128) {
          const int c = col_base + e;
          const bool c_ok = e < p.bn && c < p.n;
          bias_s[e] = c_ok ? (p.bias ? __ldg(p.bias + c) : 0.f) : (p.rowmax_key ? -INFINITY : 0.f);
          if (gb_staged) {
            gb_s[e] = c_ok ? __ldg(p.group_bias + g_first * p.n + c) : 0.f;
            gb_s[kMaxTileN + e] = (c_ok && g_last != g_first) ? __ldg(p.group_bias + g_last * p.n + c) : 0.f;
          }
        }
        named_barrier_sync(1 + half, 128);
      }

      mbar_wait(&st->tmem_full[half], use & 1);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(half * kMaxTileN);
      unsigned long long rkey = 0ull;
      if (p.compact_out) {
        // the warp's previous tile must have left the staging buffer
        if (lane == 0) bulk_wait_group_read<0>();
        __syncwarp();
      }
      for (int step = 0; step < steps; ++step) {
        uint8_t* srow = myslabs + slab * kWarpSlabBytes + lane * 128;
        if (p.tma_out) {
          if (p.tma_mask) {
            mbar_wait(&st->mask_full[ew][slab], slab_par);
          } else {
            // the store that last read this slab (S steps ago) is done with it
            if (lane == 0) {
              if (S == 2) bulk_wait_group_read<1>();
              else bulk_wait_group_read<2>();
            }
            __syncwarp();
          }
        }
        uint32_t raw[kChunks][32];
#pragma unroll
        for (int h = 0; h < kChunks; ++h)
          if (step * kStepCols + h * 32 < p.bn) tmem_ld32_issue(taddr0 + step * kStepCols + h * 32, raw[h]);
#pragma unroll
        for (int h = 0; h < kChunks; ++h) tmem_ld32_wait(raw[h]);

#pragma unroll
        for (int h = 0; h < kChunks; ++h) {
          const int c0 = step * kStepCols + h * 32;       // column inside the tile
          if (c0 >= p.bn) continue;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[h][j]);
          const int cg = col_base + c0;
          const int valid = p.n - cg < 32 ? (p.n - cg > 0 ? p.n - cg : 0) : 32;
          {
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b = b4[q];
              v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
            }
          }
          if (p.group_bias) {
            if (gb_staged) {
              const float4* g4 = reinterpret_cast<const float4*>(gb_s + (g != g_first ? kMaxTileN : 0) + c0);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 b = g4[q];
                v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
              }
            } else if (r_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) v[j] += __ldg(p.group_bias + g * p.n + cg + j);
            }
          }
          if (p.addend && r_ok) {
            const float* ap = p.addend + r * p.ld_addend + cg;
            if (valid == 32 && (reinterpret_cast<uintptr_t>(ap) & 15) == 0) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 b = *reinterpret_cast<const float4*>(ap + 4 * q);
                v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) v[j] += ap[j];
            }
          }
          if (p.rowmax_key) {          // columns >= n carry -inf from bias_s and never win
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const unsigned long long k = pack_key(v[j], static_cast<uint32_t>(cg + j));
              rkey = k > rkey ? k : rkey;
            }
          }
          if (!p.out) continue;
          if (kAct == PCADV_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (kAct == PCADV_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
          }
          if (p.mask && p.mask_act != PCADV_ACT_NONE) {
            if (p.tma_mask) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float m[8];
                unpack16(*reinterpret_cast<const uint4*>(srow + (((h * 4 + q) ^ sw) << 4)), p.mask_dtype, m);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[q * 8 + e] = m[e] > 0.f ? v[q * 8 + e] : v[q * 8 + e] * mneg;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float m = (r_ok && j < valid) ? ld_as_float(p.mask, r * p.ld_mask + cg + j, p.mask_dtype) : 0.f;
                v[j] = m > 0.f ? v[j] : v[j] * mneg;
              }
            }
          }
          if (p.out_scale) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= oscale;
          }
          if (p.tma_out) {
            if (kOut == PCADV_F32) {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(srow + ((static_cast<uint32_t>(q) ^ sw) << 4)) =
                    make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 16; ++j)
                pk[j] = kOut == PCADV_F16 ? pack_f16x2_sat(v[2 * j], v[2 * j + 1])
                                          : pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(srow + (((h * 4 + q) ^ sw) << 4)) =
                    make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
          } else if (p.compact_out) {
            // row-compact staging: [32 rows][n] elements, the warp's rows are contiguous in HBM
            uint8_t* crow = myslabs + static_cast<size_t>(lane) * p.n * kEsz + static_cast<size_t>(c0) * kEsz;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < valid) {
                if (kOut == PCADV_F32) reinterpret_cast<float*>(crow)[j] = v[j];
                else if (kOut == PCADV_F16) reinterpret_cast<__half*>(crow)[j] = __float2half_rn(fminf(fmaxf(v[j], -65504.f), 65504.f));
                else reinterpret_cast<__nv_bfloat16*>(crow)[j] = __float2bfloat16_rn(v[j]);
              }
            }
          } else if (r_ok && valid > 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < valid) st_from_float(p.out, r * p.ld_out + cg + j, kOut, v[j]);
          }
        }
        if (p.tma_out) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&maps.out, myslabs + slab * kWarpSlabBytes, col_base + step * kStepCols,
                         static_cast<int32_t>(tm * kTileM + quarter * 32));
            bulk_commit_group();
            if (p.tma_mask) {
              // the slab filled next-but-(S-2) was last read by the store before this one
              bulk_wait_group_read<1>();
              issue_mask();
            }
          }
          if (++slab == S) { slab = 0; slab_par ^= 1; }
        }
      }
      if (p.compact_out && p.out) {
        const int64_t wr0 = tm * kTileM + quarter * 32;   // first row of this warp
        int64_t nrows = p.rows - wr0;
        nrows = nrows > 32 ? 32 : nrows;
        if (nrows > 0) {
          const uint32_t bytes = static_cast<uint32_t>(nrows * p.n * kEsz);
          uint8_t* gdst = reinterpret_cast<uint8_t*>(p.out) + wr0 * p.n * kEsz;
          if ((bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(gdst) & 15) == 0) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              bulk_store_1d(gdst, myslabs, bytes);
              bulk_commit_group();
            }
          } else {
            __syncwarp();
            const uint32_t words = bytes / 4;             // n * esz * nrows; 16-bit n is even here
            for (uint32_t i = lane; i < words; i += 32)
              reinterpret_cast<uint32_t*>(gdst)[i] = reinterpret_cast<const uint32_t*>(myslabs)[i];
            __syncwarp();
          }
        }
      }
      if (p.rowmax_key && r_ok && rkey) atomicMax(&p.rowmax_key[r], rkey);
      // release the accumulator buffer of this half
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->tmem_empty[half]);
    }
    if (lane == 0) bulk_wait_group<0>();                  // all output slabs have landed
  }
  rows_teardown(warp, tmem_base);
}

template <int kAct>
static RowsKernel pick_out(int out_dtype) {
  switch (out_dtype) {
    case PCADV_F16: return tc_rows_kernel<kAct, PCADV_F16>;
    case PCADV_BF16: return tc_rows_kernel<kAct, PCADV_BF16>;
    default: return tc_rows_kernel<kAct, PCADV_F32>;
  }
}

static RowsKernel pick_rows_kernel(int act, int out_dtype) {
  switch (act) {
    case PCADV_ACT_RELU: return pick_out<PCADV_ACT_RELU>(out_dtype);
    case PCADV_ACT_LEAKY: return pick_out<PCADV_ACT_LEAKY>(out_dtype);
    default: return pick_out<PCADV_ACT_NONE>(out_dtype);
  }
}

static int ensure_rows_smem(const void* kernel) {
  static const void* done[32] = {nullptr};
  for (int i = 0; i < 32; ++i) {
    if (done[i] == kernel) return 0;
    if (done[i] == nullptr) {
      PCADV_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kRowsSmemMax));
      done[i] = kernel;
      return 0;
    }
  }
  return 0;
}

}  // namespace tc

int tc_rows(const pcadv_linear_args& a, cudaStream_t s) {
  using namespace tc;
  const int dt = a.seg[0].dtype;
  TensorMaps maps;
  RowsParams p{};
  p.rows = a.rows; p.n = a.n; p.num_seg = a.num_seg;
  p.bn = (a.n + 15) / 16 * 16;
  if (p.bn > kMaxTileN) p.bn = kMaxTileN;
  int ktot = 0;
  for (int i = 0; i < a.num_seg; ++i) {
    PCADV_CHECK_ARG(a.seg[i].dtype == dt && a.seg[i].k % kBlockK == 0 &&
                        tma_compatible(a.seg[i].ptr, dt, a.seg[i].ld),
                    "tc_linear: segment %d not TMA-compatible (k=%d ld=%lld)", i, a.seg[i].k,
                    (long long)a.seg[i].ld);
    p.seg_k[i] = a.seg[i].k;
    if (int rc = encode_tmap_2d(&maps.act[i], a.seg[i].ptr, dt, a.rows, a.seg[i].k, a.seg[i].ld,
                                kBlockK, kTileM))
      return rc;
    ktot += a.seg[i].k;
  }
  PCADV_CHECK_ARG(tma_compatible(a.w, dt, a.ldw), "tc_linear: weight not TMA-compatible");
  if (int rc = encode_tmap_2d(&maps.w, a.w, dt, a.n, ktot, a.ldw, kBlockK, p.bn)) return rc;
  p.tiles_m = (a.rows + kTileM - 1) / kTileM;
  p.tiles_n = (a.n + p.bn - 1) / p.bn;
  p.idesc = make_idesc(kTileM, p.bn, dt == PCADV_BF16, false, false);
  p.bias = a.bias; p.group_bias = a.group_bias; p.rows_per_group = a.rows_per_group;
  p.addend = a.addend; p.ld_addend = a.ld_addend; p.slope = a.slope;
  p.mask = a.mask; p.ld_mask = a.ld_mask; p.mask_dtype = a.mask_dtype; p.mask_act = a.mask_act;
  p.mask_slope = a.mask_slope; p.out_scale = a.out_scale; p.out = a.out; p.ld_out = a.ld_out;
  p.rowmax_key = a.rowmax_key;
  const int out_dt = a.out ? a.out_dtype : PCADV_F16;
  const int esz = out_dt == PCADV_F32 ? 4 : 2;
  p.tma_out = (a.out && tma_compatible(a.out, out_dt, a.ld_out)) ? 1 : 0;
  if (p.tma_out) {
    if (int rc = encode_tmap_2d(&maps.out, a.out, out_dt, a.rows, a.n, a.ld_out, 128 / esz, 32))
      return rc;
  }
  if (a.bits_out || a.mask_bits) {
    PCADV_CHECK_ARG(a.n % 64 == 0, "tc_linear: bit masks need n %% 64 == 0 (n=%d)", a.n);
    PCADV_CHECK_ARG(!a.bits_out || (a.out && out_dt != PCADV_F32 && a.ld_bits_out % 2 == 0 &&
                                    (reinterpret_cast<uintptr_t>(a.bits_out) & 7) == 0),
                    "tc_linear: bits_out needs a 16-bit output and 8-byte aligned rows");
    PCADV_CHECK_ARG(!a.mask_bits || out_dt != PCADV_F32, "tc_linear: mask_bits needs a 16-bit output");
    PCADV_CHECK_ARG(!a.mask_bits || (a.ld_mask_bits % 2 == 0 && (reinterpret_cast<uintptr_t>(a.mask_bits) & 7) == 0),
                    "tc_linear: mask_bits rows must be 8-byte aligned");
  }
  p.bits_out = a.bits_out; p.ld_bits_out = a.ld_bits_out;
  p.mask_bits = a.mask_act != PCADV_ACT_NONE ? a.mask_bits : nullptr; p.ld_mask_bits = a.ld_mask_bits;
  // ---- the lean kernel takes the common shapes; everything else goes to the general one
  RowsKernel lean = nullptr;
  if (p.tma_out && out_dt != PCADV_F32 && a.n % 64 == 0 && !a.addend && !a.rowmax_key && !a.out_scale &&
      (!a.mask || a.mask_act == PCADV_ACT_NONE || p.mask_bits) &&
      (!a.group_bias || a.rows_per_group >= kTileM)) {
    const int add = a.group_bias ? kAddBiasGroup : (a.bias ? kAddBias : kAddNone);
    lean = out_dt == PCADV_F16 ? pick_lean<PCADV_F16>(a.act, add, p.mask_bits != nullptr)
                               : pick_lean<PCADV_BF16>(a.act, add, p.mask_bits != nullptr);
  }
  if (!a.out && a.rowmax_key && a.n % 64 == 0 && !a.addend && !a.mask && !a.group_bias && !a.out_scale)
    lean = lean_rowmax();
  PCADV_CHECK_ARG(lean || (!a.bits_out && !p.mask_bits),
                  "tc_linear: bit masks are only implemented for the lean layer shapes");
  int threads = kRowsThreads;
  if (a.seg0_group_sum) {
    PCADV_CHECK_ARG(lean && lean != lean_rowmax() && p.mask_bits && a.act == PCADV_ACT_NONE && !a.bias && !a.group_bias &&
                        !a.bits_out && a.rows_per_group > 0 && a.rows_per_group % kTileM == 0 &&
                        a.rows % a.rows_per_group == 0 && a.seg[0].k <= 256,
                    "tc_linear: seg0_group_sum needs the mask_bits dgrad shape, rows_per_group %% 128 == 0 and seg[0].k <= 256");
    lean = out_dt == PCADV_F16 ? lean_group_sum<PCADV_F16>() : lean_group_sum<PCADV_BF16>();
    p.group_sum = a.seg0_group_sum;
    threads = kRowsSumThreads;
  }
  p.tma_mask = (!lean && p.tma_out && a.mask && a.mask_act != PCADV_ACT_NONE && out_dt != PCADV_F32 &&
                a.mask_dtype != PCADV_F32 && tma_compatible(a.mask, a.mask_dtype, a.ld_mask)) ? 1 : 0;
  if (p.tma_mask) {
    if (int rc = encode_tmap_2d(&maps.mask, a.mask, a.mask_dtype, a.rows, a.n, a.ld_mask, 64, 32))
      return rc;
  }
  p.nslabs = p.tma_mask ? 3 : 2;
  if (lean) { p.tma_mask = 0; p.nslabs = 2; }
  // row-compact staging when rows are not TMA-storable but contiguous (fc4: 50 fp32 logits)
  p.compact_out = (a.out && !p.tma_out && a.ld_out == a.n && p.tiles_n == 1 &&
                   (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && (a.n * esz) % 4 == 0 &&
                   32 * a.n * esz <= p.nslabs * kWarpSlabBytes) ? 1 : 0;
  p.stage_bytes = kABytes + p.bn * kBlockK * 2;
  p.epi_halves = 2;
  int nst = kRowsMaxStages;
  while (nst > 2 && rows_smem_bytes(nst, p.stage_bytes, p.nslabs) > static_cast<size_t>(kRowsSmemMax)) --nst;
  int epi_warps = kEpiWarps;
  if (lean && lean != lean_rowmax() && !a.seg0_group_sum && ktot >= 512) {
    // Deep-K layers (fc1: K = 960): a tile's MMAs take several microseconds, so ONE epilogue half with one
    // slab per warp keeps up with both TMEM buffers, and the 48 KB it frees is one more ring stage
    // (measured: fc1 0.564 -> 0.555 ms, the K = 768 dgrad 0.324 -> 0.313 ms per 2^20 points).
    static int deep = -1;
    if (deep < 0) { const char* e = getenv("PCADV_ROWS_DEEP"); deep = (e && atoi(e) == 0) ? 0 : 1; }   // tuning aid
    int nst1 = kRowsMaxStages;
    while (nst1 > 2 && rows_smem_bytes(nst1, p.stage_bytes, 1, 4) > static_cast<size_t>(kRowsSmemMax)) --nst1;
    if (deep && nst1 > nst) { nst = nst1; p.epi_halves = 1; p.nslabs = 1; epi_warps = 4; }
  }
  p.nstages = nst;
  const size_t smem = rows_smem_bytes(p.nstages, p.stage_bytes, p.nslabs, epi_warps);
  PCADV_CHECK_ARG(smem <= static_cast<size_t>(kRowsSmemMax), "tc_linear: shared memory budget exceeded");
  const int64_t tiles = p.tiles_m * p.tiles_n;
  RowsKernel k = lean ? lean : pick_rows_kernel(a.act, out_dt);
  if (int rc = ensure_rows_smem(reinterpret_cast<const void*>(k))) return rc;
  // ---- CTA pairs for wide weight tiles (PCADV_ROWS_PAIR=1; off by default).  Every 128-row tile streams the
  // layer's whole weight tile from L2 (fc1: 480 KB next to 240 KB of activations, ~11 TB/s out of the L2
  // slices at 2^20 points).  Two CTAs of a cluster take two row tiles of the same column tile side by side,
  // each fetches HALF of every weight chunk and TMA multicasts it into both, so the slices serve each weight
  // byte once per 256 rows.  Measured: fc1 alone 0.529 -> 0.509 ms per 2^20 points, the other layers
  // unchanged, the cfg5 step 15.42 -> 15.54 ms (cluster launches next to the discriminator branch of the
  // graph) -- fc1 runs at 1.0 PFLOP/s and 5.0 TB/s at once, i.e. against the power cap, not the L2.
  static int pair_mode = -1;
  if (pair_mode < 0) { const char* e = getenv("PCADV_ROWS_PAIR"); pair_mode = e ? atoi(e) : 0; }   // 0 off, 1 wide weight tiles, 2 all lean launches
  const bool wide_w = static_cast<int64_t>(ktot) * p.bn >= 256 * 256;
  if (lean && lean != lean_rowmax() && !a.seg0_group_sum && pair_mode > 0 && (wide_w || pair_mode == 2) && p.bn % 16 == 0 &&
      p.tiles_m >= 4 && num_sms() >= 2) {
    if (int rc = encode_tmap_2d(&maps.w_half, a.w, dt, a.n, ktot, a.ldw, kBlockK, p.bn / 2)) return rc;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.blockDim = dim3(kRowsThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    // clusters that can be resident at once (a GPC with an odd SM count leaves one SM unpaired)
    static int max_pairs = -1;
    if (max_pairs < 0) {
      cfg.gridDim = dim3(num_sms() / 2 * 2);
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, reinterpret_cast<const void*>(k), &cfg) != cudaSuccess || nc <= 0) {
        cudaGetLastError();
        nc = 0;
      }
      max_pairs = nc < num_sms() / 2 ? nc : num_sms() / 2;
    }
    const int64_t pair_items = ((p.tiles_m + 1) / 2) * p.tiles_n;
    const int pairs = static_cast<int>(pair_items < max_pairs ? pair_items : max_pairs);
    if (pairs >= 1) {
      p.pair = 1;
      cfg.gridDim = dim3(2 * pairs);
      PCADV_CUDA_OK(cudaLaunchKernelEx(&cfg, k, maps, p));
      PCADV_LAUNCHED();
      return 0;
    }
  }
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  k<<<grid, threads, smem, s>>>(maps, p);
  PCADV_LAUNCHED();
  return 0;
}

}  // namespace pcadv
