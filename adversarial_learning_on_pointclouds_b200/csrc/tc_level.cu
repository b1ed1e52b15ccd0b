// Tensor-core engine, one backward LEVEL of a pointwise chain in a single pass over the incoming
// gradient (pcadv_backlevel):
//
//   dz_out[r, c]   = act'(x[r, c]) * sum_i sum_k dz_i[r, k] w[c, koff_i + k]      (dgrad of the level)
//   dw_i[k, c]    += scale * sum_r dz_i[r, k] x[r, c]                             (wgrad of the layers fed by x)
//   dbias_i[k]    += scale * sum_r dz_i[r, k]      dgroup_i[g, k] += sum_{r in cloud g} dz_i[r, k]
//
// The dgrad GEMM (points on M, K-major operands) and the weight-gradient GEMM (channels on M, points
// on K, MN-major operands) consume the SAME shared-memory tile of dz: a [128 points][64 channels]
// TMA box with 128-byte swizzle is the K-major A operand of the first and, paired with its
// neighbour in the ring, the MN-major A operand (M = 128 channels) of the second, so dz is fetched
// from HBM once instead of once per kernel.  The weight-gradient accumulators of all of the level's
// dz channels stay resident in TMEM for the CTA's whole row range (K / 128 tiles of bn columns);
// the dgrad accumulator takes the remaining columns (double-buffered where they fit).  Wide levels
// are split over the output channels: CTA i owns slice i % slices of bn columns and walks the row
// tiles i / slices, i / slices + grid / slices, ... -- the CTAs of one row tile run side by side, but
// nothing keeps them in step and measured the second slice's read of a dz tile misses L2 more often
// than not (K = 768 -> 128: 3.2 GB of DRAM reads for 1.9 GB algorithmic), which is why the models keep
// the wide levels on separate dgrad / wgrad launches.  Narrow levels (K <= 128, what the
// models use) keep up to four row tiles of dz and of x in flight: a tile there is 36-82 KB and its
// round trips, not its bytes, set the pace.
//
// One-hot mode: the single dz segment is the gradient of a max over channels (one non-zero per row)
// and is written into the ring by the four warps that otherwise take the column sums, straight from
// (dy, pooled value, argmax) -- the backward of the discriminators' last layer without gather kernels.
//
// 448 threads: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = dgrad epilogue (two halves,
// as in tc_rows.cu) and, at the end, the drain of the weight-gradient accumulators (vector fp32 RED),
// warps 10..13 = column sums of the dz boxes in flight (bias gradients, per-cloud sums) or, in the
// one-hot mode, the generator of those boxes.
#include <stdlib.h>
#include "tc_pipeline.cuh"

namespace pcadv {
namespace tc {

constexpr int kLvThreads = 448;
constexpr int kLvMaxStages = 8;
constexpr int kLvMaxChunks = 16;                 // K <= 1024 channels of dz per level
constexpr int kLvMaxSum = 8;                     // chunks whose column sums are taken
constexpr int kLvSmemMax = 232448;
constexpr int kLvSlabBytes = 32 * 128;
constexpr int kLvMaxXBuf = 4;                     // x tiles in flight
constexpr int kLvOneHotMax = 256;                // pooled channels of a one-hot segment

struct LvTail {
  uint64_t full[kLvMaxStages];
  uint64_t empty[kLvMaxStages];
  uint64_t x_full[kLvMaxXBuf];
  uint64_t x_empty[kLvMaxXBuf];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t w_full;
  uint64_t wres_full;
  uint32_t tmem_base;
  float hist[kLvOneHotMax];         // one-hot mode: bias gradient per pooled channel
};

struct LvMaps {
  CUtensorMap seg[PCADV_MAX_SEG];   // dz_i [rows, k_i]: box 128 rows x 64 channels
  CUtensorMap w;                    // dgrad weight [n, ktot]: box bn rows x 64
  CUtensorMap x;                    // stored activation [rows, n]: box 128 rows x 64 channels
  CUtensorMap out;                  // dz_out [rows, n]: box 32 rows x 64 channels (TMA store)
};

struct LvParams {
  int64_t rows, tiles_m;
  int n, bn, slices;
  int nchunks;                      // K padded to a multiple of 128, in 64-channel chunks (even)
  int chunk_seg[kLvMaxChunks];      // segment of the chunk, -1 = zero padding
  int chunk_k0[kLvMaxChunks];       // first channel of the chunk inside its segment
  int chunk_kg[kLvMaxChunks];       // column of the chunk inside w
  int chunk_sum[kLvMaxChunks];      // slot of its column sums, -1 = none
  int nsum;
  int pad_seg, pad_k0, pad_kg;      // coordinates outside the tensors: TMA fills zeros
  int nstages, stage_bytes, nbuf, nslabs, epi_warps, nxbuf;
  int wres;                         // the slice's dgrad weight stays resident in shared memory (loaded once)
  uint32_t idesc_d, idesc_w;
  int bf16;
  const uint32_t* mask_bits;
  int64_t ld_mask_bits;
  float mask_neg;                   // act'(.) where the sign bit is clear (0 = ReLU, slope = LeakyReLU)
  float* dw[PCADV_MAX_SEG];
  int64_t ld_dw[PCADV_MAX_SEG];
  float* dbias[PCADV_MAX_SEG];
  float* dgroup[PCADV_MAX_SEG];
  int seg_k[PCADV_MAX_SEG];
  int64_t rows_per_group;
  const float* scale;
  int vec_red;
  // one-hot mode: segment 0 is generated in shared memory (see pcadv_backlevel_args)
  const float* oh_dy;
  const float* oh_val;
  const int32_t* oh_idx;
  int oh_act;
  float oh_slope;
  const float* oh_scale;
  int dbg;                          // tuning aid (PCADV_LEVEL_DBG): 1 = no wgrad MMAs, 2 = no dgrad MMAs, 4 = spinning waits
};

__device__ __forceinline__ void lv_red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void lv_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void lv_tma_store(const CUtensorMap* m, uint32_t smem_addr, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <bool kBf16>
__device__ __forceinline__ void lv_add_pair(uint32_t packed, float2& acc) {
  if (kBf16) {
    asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nadd.rn.f32.bf16 %0, lo, %0;\nadd.rn.f32.bf16 %1, hi, %1;\n}"
        : "+f"(acc.x), "+f"(acc.y) : "r"(packed));
  } else {
    asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nadd.rn.f32.f16 %0, lo, %0;\nadd.rn.f32.f16 %1, hi, %1;\n}"
        : "+f"(acc.x), "+f"(acc.y) : "r"(packed));
  }
}

__device__ __forceinline__ void lv_wait(uint64_t* bar, uint32_t parity, bool spin) {
  if (spin) mbar_wait(bar, parity);
  else mbar_wait_backoff(bar, parity);
}

template <int kOut>
__global__ void __launch_bounds__(kLvThreads, 1)
tc_level_kernel(const __grid_constant__ LvMaps maps, const LvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int x_bytes = p.bn * 256;                                   // [128 points][bn channels] 16-bit
  uint8_t* wres = stages + p.nstages * p.stage_bytes;               // [nchunks][bn rows][64 k] when resident
  const int wchunk_bytes = p.bn * 128;
  uint8_t* xt = wres + (p.wres ? p.nchunks * wchunk_bytes : 0);
  uint8_t* epi = xt + p.nxbuf * x_bytes;
  LvTail* st = reinterpret_cast<LvTail*>(epi + p.epi_warps * p.nslabs * kLvSlabBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.slices;
  const int64_t t_first = blockIdx.x / p.slices;
  const int64_t t_step = gridDim.x / p.slices;
  const int n0 = slice * p.bn;
  const int bn = p.bn;
  const int nchunks = p.nchunks;
  const bool onehot = p.oh_idx != nullptr;
  const bool sums = !onehot && p.nsum > 0 && slice == 0;
  const uint32_t wbase = static_cast<uint32_t>(p.nbuf * bn);       // first weight-gradient column
  const bool spin = (p.dbg & 4) != 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < PCADV_MAX_SEG; ++s)
      if (p.seg_k[s] > 0 && p.oh_idx == nullptr) tma_prefetch_desc(&maps.seg[s]);
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.out);
    for (int i = 0; i < kLvMaxStages; ++i) {
      mbar_init(&st->full[i], onehot ? 4 : 1);     // TMA bytes, or the four generator warps
      mbar_init(&st->empty[i], sums ? 5 : 1);     // MMA commit (+ the four column-sum warps)
    }
    for (int i = 0; i < kLvMaxXBuf; ++i) {
      mbar_init(&st->x_full[i], 1);
      mbar_init(&st->x_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st->tmem_full[i], 1);
      mbar_init(&st->tmem_empty[i], 4);
    }
    mbar_init(&st->w_full, 1);
    mbar_init(&st->wres_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&st->tmem_base, kTmemCols);
  if (onehot)
    for (int i = threadIdx.x; i < kLvOneHotMax; i += kLvThreads) st->hist[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = st->tmem_base;

  if (warp == 0) {
    // ================= TMA producer =================
    // the whole warp: lane 0 arms the barrier, lanes 0 / 1 issue the two boxes of a chunk (a single
    // thread issues one tensor copy every ~140 ns, which would be the pipeline's pace)
    {
      const uint32_t stage_tx = static_cast<uint32_t>((kTileM + (p.wres ? 0 : bn)) * 128);
      int stage = 0, xb = 0;
      uint32_t phase = 0, x_phase = 0;
      if (p.wres && t_first < p.tiles_m) {
        if (lane == 0) mbar_arrive_expect_tx(&st->wres_full, static_cast<uint32_t>(nchunks * wchunk_bytes));
        __syncwarp();
        if (lane < nchunks) {
          const int kg = p.chunk_seg[lane] >= 0 ? p.chunk_kg[lane] : p.pad_kg;
          tma_load_2d(wres + lane * wchunk_bytes, &maps.w, &st->wres_full, kg, n0);
        }
      }
      for (int64_t t = t_first; t < p.tiles_m; t += t_step) {
        const int32_t m0 = static_cast<int32_t>(t * kTileM);
        lv_wait(&st->x_empty[xb], x_phase ^ 1, spin);
        if (lane == 0) mbar_arrive_expect_tx(&st->x_full[xb], static_cast<uint32_t>(x_bytes));
        __syncwarp();
        if (lane < (bn >> 6))
          tma_load_2d(xt + xb * x_bytes + lane * kABytes, &maps.x, &st->x_full[xb], n0 + lane * 64, m0);
        if (++xb == p.nxbuf) { xb = 0; x_phase ^= 1; }
        for (int c = 0; c < nchunks && !onehot; ++c) {
          lv_wait(&st->empty[stage], phase ^ 1, spin);
          if (lane == 0) mbar_arrive_expect_tx(&st->full[stage], stage_tx);
          __syncwarp();
          uint8_t* sa = stages + stage * p.stage_bytes;
          const int sg = p.chunk_seg[c];
          const int seg_i = sg >= 0 ? sg : p.pad_seg;          // padding: coordinates outside the tensors, TMA fills zeros
          const int k0 = sg >= 0 ? p.chunk_k0[c] : p.pad_k0;
          const int kg = sg >= 0 ? p.chunk_kg[c] : p.pad_kg;
          if (lane == 0) tma_load_2d(sa, &maps.seg[seg_i], &st->full[stage], k0, m0);
          else if (lane == 1 && !p.wres) tma_load_2d(sa + kABytes, &maps.w, &st->full[stage], kg, n0);
          if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0, xb = 0, buf = 0;
      uint32_t phase = 0, x_phase = 0, buf_phase = 0;
      bool wfirst = true;
      if (p.wres && t_first < p.tiles_m) lv_wait(&st->wres_full, 0, spin);
      const uint32_t wres_addr = smem_u32(wres);
      for (int64_t t = t_first; t < p.tiles_m; t += t_step) {
        lv_wait(&st->tmem_empty[buf], buf_phase ^ 1, spin);
        lv_wait(&st->x_full[xb], x_phase, spin);
        tc_fence_after();
        const uint32_t d_g = tmem_base + static_cast<uint32_t>(buf * bn);
        const uint32_t x_addr = smem_u32(xt + xb * x_bytes);
        for (int c = 0; c < nchunks; ++c) {
          lv_wait(&st->full[stage], phase, spin);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stages + stage * p.stage_bytes);
          const uint32_t b_addr = p.wres ? wres_addr + static_cast<uint32_t>(c * wchunk_bytes) : a_addr + kABytes;
          if (!(p.dbg & 2) || c == 0) mma_chunk_kmajor(d_g, a_addr, b_addr, p.idesc_d, c == 0);
          if (c & 1) {
            // the chunk pair (stage - 1, stage) = 128 dz channels on M; the 128 points are K
            const uint32_t a0 = a_addr - static_cast<uint32_t>(p.stage_bytes);
            const uint32_t d_w = tmem_base + wbase + static_cast<uint32_t>((c >> 1) * bn);
#pragma unroll
            for (int k = 0; k < kTileM / 16; ++k)
              if (!(p.dbg & 1) || (c == 1 && k == 0)) umma_f16(d_w, make_smem_desc(a0 + k * 2048, static_cast<uint32_t>(p.stage_bytes), 1024),
                       make_smem_desc(x_addr + k * 2048, kABytes, 1024), p.idesc_w,
                       (wfirst && k == 0) ? 0u : 1u);
            umma_commit(&st->empty[stage - 1]);
            umma_commit(&st->empty[stage]);
          }
          if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
        wfirst = false;
        umma_commit(&st->x_empty[xb]);
        umma_commit(&st->tmem_full[buf]);
        if (++xb == p.nxbuf) { xb = 0; x_phase ^= 1; }
        if (++buf == p.nbuf) { buf = 0; buf_phase ^= 1; }
      }
      umma_commit(&st->w_full);
    }
  } else if (warp < 10) {
    // ================= dgrad epilogue (two halves x four lane quarters) =================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int lane_row = quarter * 32 + lane;
    const int steps = bn >> 6;
    const bool has_mask = p.mask_bits != nullptr;
    const float mneg = p.mask_neg;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    if (half < p.nbuf) {
      const int sw_idx = p.nbuf == 2 ? ew : (ew & 3);
      const uint32_t slab_tma0 = smem_u32(epi + sw_idx * p.nslabs * kLvSlabBytes);
      const uint32_t slab0 = slab_tma0 + lane * 128;
      uint32_t slab = 0, use = 0;
      uint2 mbw[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
      auto fetch_bits = [&](int64_t tt) {
        if (tt >= p.tiles_m) return;
        const int64_t rr = tt * kTileM + lane_row;
        if (rr >= p.rows) return;
        const uint2* src = reinterpret_cast<const uint2*>(p.mask_bits + rr * p.ld_mask_bits + (n0 >> 5));
        mbw[0] = __ldg(src);
        if (steps > 1) mbw[1] = __ldg(src + 1);
      };
      const int64_t t0 = t_first + half * t_step;
      const int64_t tstride = p.nbuf * t_step;
      if (has_mask) fetch_bits(t0);
      for (int64_t t = t0; t < p.tiles_m; t += tstride, ++use) {
        const int32_t row_tma = static_cast<int32_t>(t * kTileM + quarter * 32);
        const uint2 mbc0 = mbw[0], mbc1 = mbw[1];
        if (has_mask) fetch_bits(t + tstride);
        mbar_wait(&st->tmem_full[half], use & 1);
        tc_fence_after();
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                static_cast<uint32_t>(half * bn);
#pragma unroll 1
        for (int step = 0; step < steps; ++step) {
          const uint32_t srow = slab0 + slab * kLvSlabBytes;
          if (lane == 0) {
            if (p.nslabs == 2) bulk_wait_group_read<1>();
            else bulk_wait_group_read<0>();
          }
          __syncwarp();
          uint32_t raw[2][32];
          tmem_ld32_issue(taddr0 + step * 64, raw[0]);
          tmem_ld32_issue(taddr0 + step * 64 + 32, raw[1]);
          const uint2 mb = step == 0 ? mbc0 : mbc1;
          tmem_ld32_wait(raw[0]);
          tmem_ld32_wait(raw[1]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[h][j]);
            if (has_mask) {
              const uint32_t word = h == 0 ? mb.x : mb.y;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                v[j] = (word & (1u << ((j >> 1) + 16 * (j & 1)))) ? v[j] : v[j] * mneg;
            }
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pk[j] = kOut == PCADV_F16 ? pack_f16x2_sat(v[2 * j], v[2 * j + 1]) : pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              lv_sts128(srow + (((h * 4 + q) ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            lv_tma_store(&maps.out, slab_tma0 + slab * kLvSlabBytes, n0 + step * 64, row_tma);
            bulk_commit_group();
          }
          if (p.nslabs == 2) slab ^= 1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&st->tmem_empty[half]);
      }
      if (lane == 0) bulk_wait_group<0>();
    }
    // ---- drain of the weight-gradient accumulators: M tile j = dz channels 128 j .. 128 j + 127 on
    // the TMEM lanes, bn columns = this slice's x channels
    if (t_first < p.tiles_m) {
      mbar_wait(&st->w_full, 0);
      tc_fence_after();
      const float sc = p.scale ? *p.scale : 1.f;
      for (int j = half; j < (nchunks >> 1); j += 2) {
        const int c = 2 * j + (quarter >> 1);
        const int sg = p.chunk_seg[c];
        const int ch = p.chunk_k0[c] + (quarter & 1) * 32 + lane;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + wbase +
                               static_cast<uint32_t>(j * bn);
        for (int cc = 0; cc < bn; cc += 32) {
          float v[32];
          tmem_ld32(taddr + cc, v);
          if (sg < 0 || p.dw[sg] == nullptr) continue;
          float* dst = p.dw[sg] + static_cast<int64_t>(ch) * p.ld_dw[sg] + n0 + cc;
          if (p.vec_red) {
#pragma unroll
            for (int q = 0; q < 32; q += 4)
              lv_red_add_v4(dst + q, v[q] * sc, v[q + 1] * sc, v[q + 2] * sc, v[q + 3] * sc);
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) atomicAdd(dst + q, v[q] * sc);
          }
        }
      }
    }
  } else {
    if (onehot) {
      // ================= generator: the one-hot dz boxes, one row per thread =================
      const int gw = warp - 10;
      const uint32_t rr = static_cast<uint32_t>(gw * 32 + lane);     // row of the 128-point box
      const float S = p.oh_scale ? *p.oh_scale : 1.f;
      const float oneg = p.oh_act == PCADV_ACT_LEAKY ? p.oh_slope : 0.f;
      int stage = 0;
      uint32_t phase = 0;
      // (idx, val, dy) of the NEXT tile are requested before this tile's boxes are written
      int n_hot = -1;
      float n_val = 0.f, n_dy = 0.f;
      auto fetch = [&](int64_t tt) {
        const int64_t r = tt * kTileM + rr;
        n_hot = -1; n_val = 0.f; n_dy = 0.f;
        if (tt < p.tiles_m && r < p.rows) {
          n_hot = __ldg(p.oh_idx + r);
          n_dy = __ldg(p.oh_dy + r);
          if (p.oh_act != PCADV_ACT_NONE) n_val = __ldg(p.oh_val + r);
        }
      };
      fetch(t_first);
      for (int64_t t = t_first; t < p.tiles_m; t += t_step) {
        const int hot = n_hot;
        const float d = p.oh_act == PCADV_ACT_NONE ? 1.f : (n_val > 0.f ? 1.f : oneg);
        const float v = hot >= 0 ? n_dy * d * S : 0.f;
        fetch(t + t_step);
        if (slice == 0 && hot >= 0 && hot < nchunks * 64) atomicAdd(&st->hist[hot], v);
        const uint32_t pk = kOut == PCADV_F16 ? pack_f16x2_sat(v, v) : pack_bf16x2(v, v);
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&st->empty[stage], phase ^ 1);
          const uint32_t row_a = smem_u32(stages + stage * p.stage_bytes) + rr * 128;
          const int hc = (hot >= c * 64 && hot < c * 64 + 64) ? hot - c * 64 : -1;   // hot channel inside this chunk
#pragma unroll
          for (uint32_t ph = 0; ph < 8; ++ph) {
            const uint32_t lg = ph ^ (rr & 7);                       // logical 16-byte group at this position
            uint32_t w4[4] = {0u, 0u, 0u, 0u};
            if (hc >= 0 && static_cast<uint32_t>(hc >> 3) == lg) {
              const uint32_t e = static_cast<uint32_t>(hc & 7);
              const uint32_t word = (e & 1) ? (pk & 0xffff0000u) : (pk & 0x0000ffffu);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (static_cast<uint32_t>(q) == (e >> 1)) w4[q] = word;
            }
            lv_sts128(row_a + (ph << 4), w4[0], w4[1], w4[2], w4[3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&st->full[stage]);
          if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
      }
      // bias gradient: the CTA's histogram, once all four generator warps are through
      named_barrier_sync(3, 128);
      if (slice == 0 && p.dbias[0] != nullptr) {
        const float sc = p.scale ? *p.scale : 1.f;
        for (int c = gw * 32 + lane; c < p.seg_k[0]; c += 128) {
          const float h = st->hist[c];
          if (h != 0.f) atomicAdd(p.dbias[0] + c, h * sc);
        }
      }
    }
    // ================= column sums of the dz boxes in flight =================
    if (sums) {
      const int cw = warp - 10;                       // row quarter of the 128-point box
      const uint32_t cc = static_cast<uint32_t>(lane) * 2;   // channel pair inside the 64-channel chunk
      float2 bsum[kLvMaxSum], gsum[kLvMaxSum];
#pragma unroll
      for (int s = 0; s < kLvMaxSum; ++s) bsum[s] = gsum[s] = make_float2(0.f, 0.f);
      int64_t cur_g = -1;
      const int64_t rpg = p.rows_per_group > 0 ? p.rows_per_group : p.rows;
      auto flush_groups = [&]() {
        if (cur_g < 0) return;
        for (int c = 0; c < nchunks; ++c) {
          const int slot = p.chunk_sum[c];
          if (slot < 0) continue;
          const int sg = p.chunk_seg[c];
          if (p.dgroup[sg] == nullptr) continue;
          float2 g = make_float2(0.f, 0.f);
#pragma unroll
          for (int s = 0; s < kLvMaxSum; ++s)
            if (s == slot) { g = gsum[s]; gsum[s] = make_float2(0.f, 0.f); }
          float* dst = p.dgroup[sg] + cur_g * p.seg_k[sg] + p.chunk_k0[c] + cc;
          atomicAdd(dst, g.x);
          atomicAdd(dst + 1, g.y);
        }
      };
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = t_first; t < p.tiles_m; t += t_step) {
        const int64_t g = (t * kTileM) / rpg;
        if (g != cur_g) { flush_groups(); cur_g = g; }
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&st->full[stage], phase);
          const int slot = p.chunk_sum[c];
          if (slot >= 0) {
            const uint8_t* tile = stages + stage * p.stage_bytes;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
              const uint32_t rr = static_cast<uint32_t>(cw * 32 + i);
              const uint32_t v = *reinterpret_cast<const uint32_t*>(
                  tile + rr * 128 + ((((cc >> 3) ^ (rr & 7)) << 4) | ((cc & 7) << 1)));
              if (p.bf16) lv_add_pair<true>(v, acc);
              else lv_add_pair<false>(v, acc);
            }
#pragma unroll
            for (int s = 0; s < kLvMaxSum; ++s)
              if (s == slot) {
                bsum[s].x += acc.x; bsum[s].y += acc.y;
                gsum[s].x += acc.x; gsum[s].y += acc.y;
              }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&st->empty[stage]);
          if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
      }
      flush_groups();
      const float sc = p.scale ? *p.scale : 1.f;
      for (int c = 0; c < nchunks; ++c) {
        const int slot = p.chunk_sum[c];
        if (slot < 0) continue;
        const int sg = p.chunk_seg[c];
        if (p.dbias[sg] == nullptr) continue;
        float2 b = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < kLvMaxSum; ++s)
          if (s == slot) b = bsum[s];
        float* dst = p.dbias[sg] + p.chunk_k0[c] + cc;
        atomicAdd(dst, b.x * sc);
        atomicAdd(dst + 1, b.y * sc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

static size_t lv_smem_bytes(int nstages, int stage_bytes, int bn, int epi_warps, int nslabs, int wres_bytes = 0,
                            int nxbuf = 2) {
  return 1024 + static_cast<size_t>(nstages) * stage_bytes + wres_bytes + nxbuf * static_cast<size_t>(bn) * 256 +
         static_cast<size_t>(epi_warps) * nslabs * kLvSlabBytes + sizeof(LvTail) + 16;
}

}  // namespace tc

// 0 = launched, > 0 = error, -1 = shape not taken (the caller falls back to separate kernels)
int tc_backlevel(const pcadv_backlevel_args& a, cudaStream_t s) {
  using namespace tc;
  const int dt = a.seg[0].dtype;
  if (dt != PCADV_F16 && dt != PCADV_BF16) return -1;
  if (a.n % 64 != 0 || a.rows < kTileM) return -1;
  LvMaps maps;
  LvParams p{};
  p.rows = a.rows; p.n = a.n;
  p.tiles_m = (a.rows + kTileM - 1) / kTileM;
  int nch = 0, ktot = 0, nsum = 0;
  const bool onehot = a.onehot_idx != nullptr;
  if (onehot && (a.num_seg != 1 || !a.onehot_dy || (a.onehot_act != PCADV_ACT_NONE && !a.onehot_val) ||
                 a.seg[0].k > kLvOneHotMax))
    return -1;
  for (int i = 0; i < a.num_seg; ++i) {
    const pcadv_seg& sg = a.seg[i];
    if (sg.dtype != dt || sg.k % 64 != 0 || (!onehot && !tma_compatible(sg.ptr, dt, sg.ld))) return -1;
    if (nch + sg.k / 64 > kLvMaxChunks) return -1;
    const bool want_sum = !onehot && (a.dbias[i] != nullptr || a.dgroup[i] != nullptr);
    if (onehot && a.dgroup[i]) return -1;
    for (int k0 = 0; k0 < sg.k; k0 += 64, ++nch) {
      p.chunk_seg[nch] = i; p.chunk_k0[nch] = k0; p.chunk_kg[nch] = ktot + k0;
      p.chunk_sum[nch] = want_sum ? nsum++ : -1;
    }
    p.seg_k[i] = sg.k;
    p.dw[i] = a.dw[i]; p.ld_dw[i] = a.ld_dw[i]; p.dbias[i] = a.dbias[i]; p.dgroup[i] = a.dgroup[i];
    if (a.dgroup[i] && (a.rows_per_group <= 0 || a.rows_per_group % kTileM != 0)) return -1;
    if (!onehot)
      if (int rc = encode_tmap_2d(&maps.seg[i], sg.ptr, dt, a.rows, sg.k, sg.ld, kBlockK, kTileM)) return rc;
    ktot += sg.k;
  }
  if (nsum > kLvMaxSum) return -1;
  p.nsum = nsum;
  p.pad_seg = a.num_seg - 1; p.pad_k0 = a.seg[a.num_seg - 1].k; p.pad_kg = ktot;
  if (nch & 1) { p.chunk_seg[nch] = -1; p.chunk_k0[nch] = 0; p.chunk_kg[nch] = ktot; p.chunk_sum[nch] = -1; ++nch; }
  p.nchunks = nch;
  const int mt = nch / 2;
  if (!tma_compatible(a.w, dt, a.ldw) || !tma_compatible(a.x, dt, a.ldx) || !tma_compatible(a.dz_out, dt, a.ld_out))
    return -1;
  // slice width and buffering: weight-gradient tiles + dgrad accumulator(s) inside 512 TMEM columns
  int bn = 0, nbuf = 0;
  int force_bn = 0;
  if (const char* e = getenv("PCADV_LEVEL_BN")) force_bn = atoi(e);       // tuning aid
  for (int cand = 128; cand >= 64 && bn == 0; cand >>= 1) {
    if (a.n % cand != 0 || (force_bn && cand != force_bn && a.n % force_bn == 0)) continue;
    if ((2 + mt) * cand <= kTmemCols) { bn = cand; nbuf = 2; }
    else if ((1 + mt) * cand <= kTmemCols) { bn = cand; nbuf = 1; }
  }
  if (bn == 0) return -1;
  if (const char* e = getenv("PCADV_LEVEL_NBUF")) { if (atoi(e) == 1) nbuf = 1; }   // tuning aid
  p.bn = bn; p.nbuf = nbuf; p.slices = a.n / bn;
  if (p.slices > num_sms()) return -1;
  p.epi_warps = nbuf == 2 ? 8 : 4;
  // shared memory: prefer the slice's whole dgrad weight resident (stages carry dz only: half the
  // L2 -> SM traffic and twice the ring depth per byte) when >= 4 stages still fit, else stream it
  int want_wres = 1;
  if (const char* e = getenv("PCADV_LEVEL_WRES")) want_wres = atoi(e);     // tuning aid
  int nst = 0;
  size_t smem = 0;
  // pick the layout with the most dz bytes in flight (the ring's turnaround -- load latency plus two
  // barrier wake-ups, ~2 us -- bounds the pipeline at ring bytes / turnaround); ties go to resident weights
  if (onehot) want_wres = 1;
  // narrow levels (few chunks per row tile) are bound by round trips per tile, so as many whole tiles as
  // possible are kept in flight: score = min(dz tiles in the ring, x tiles buffered), then the ring depth
  int best_score = -1;
  for (int mode = want_wres ? 0 : 1; mode < (onehot ? 1 : 2); ++mode) {
    const int wres_bytes = mode == 0 ? nch * bn * 128 : 0;
    const int stage_bytes = kABytes + (mode == 0 ? 0 : bn * 128);
    for (int nslabs = 2; nslabs >= 1; --nslabs) {
      for (int nx = 2; nx <= kLvMaxXBuf; ++nx) {
        int cand = kLvMaxStages;
        while (cand >= 2 && lv_smem_bytes(cand, stage_bytes, bn, p.epi_warps, nslabs, wres_bytes, nx) > static_cast<size_t>(kLvSmemMax)) cand -= 2;
        if (cand < 2) continue;
        const int tiles16 = cand * 16 / nch;                       // dz tiles in the ring, in 1/16
        const int score = (tiles16 < nx * 16 ? tiles16 : nx * 16) * 64 + cand;
        if (score > best_score) {
          best_score = score;
          nst = cand; p.nslabs = nslabs; p.nxbuf = nx; p.wres = mode == 0 ? 1 : 0; p.stage_bytes = stage_bytes;
          smem = lv_smem_bytes(cand, stage_bytes, bn, p.epi_warps, nslabs, wres_bytes, nx);
        }
      }
    }
  }
  if (nst == 0) return -1;
  if (const char* e = getenv("PCADV_LEVEL_STAGES")) {                     // tuning aid
    const int v = atoi(e);
    if (v >= 2 && v <= nst && v % 2 == 0) nst = v;
  }
  p.nstages = nst;
  if (int rc = encode_tmap_2d(&maps.w, a.w, dt, a.n, ktot, a.ldw, kBlockK, bn)) return rc;
  if (int rc = encode_tmap_2d(&maps.x, a.x, dt, a.rows, a.n, a.ldx, kBlockK, kTileM)) return rc;
  if (int rc = encode_tmap_2d(&maps.out, a.dz_out, dt, a.rows, a.n, a.ld_out, 64, 32)) return rc;
  p.idesc_d = make_idesc(kTileM, bn, dt == PCADV_BF16, false, false);
  p.idesc_w = make_idesc(kTileM, bn, dt == PCADV_BF16, true, true);
  p.bf16 = dt == PCADV_BF16 ? 1 : 0;
  p.mask_bits = a.mask_act != PCADV_ACT_NONE ? a.mask_bits : nullptr;
  p.ld_mask_bits = a.ld_mask_bits;
  p.mask_neg = a.mask_act == PCADV_ACT_LEAKY ? a.mask_slope : 0.f;
  p.rows_per_group = a.rows_per_group;
  p.scale = a.scale;
  if (onehot) {
    if (!p.wres || (nch & 1)) return -1;
    p.oh_dy = a.onehot_dy; p.oh_val = a.onehot_val; p.oh_idx = a.onehot_idx;
    p.oh_act = a.onehot_act; p.oh_slope = a.onehot_slope; p.oh_scale = a.onehot_scale;
  }
  p.vec_red = 1;
  if (const char* e = getenv("PCADV_LEVEL_DBG")) p.dbg = atoi(e);
  for (int i = 0; i < a.num_seg; ++i)
    if (a.dw[i] && (a.ld_dw[i] % 4 != 0 || (reinterpret_cast<uintptr_t>(a.dw[i]) & 15) != 0)) p.vec_red = 0;
  int64_t streams = num_sms() / p.slices;
  if (streams > p.tiles_m) streams = p.tiles_m;
  const int grid = static_cast<int>(streams * p.slices);
  auto kern = dt == PCADV_F16 ? tc_level_kernel<PCADV_F16> : tc_level_kernel<PCADV_BF16>;
  static bool attr_done[2] = {false, false};
  if (!attr_done[p.bf16]) {
    PCADV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kLvSmemMax));
    attr_done[p.bf16] = true;
  }
  kern<<<grid, kLvThreads, smem, s>>>(maps, p);
  PCADV_LAUNCHED();
  return 0;
}

}  // namespace pcadv

using namespace pcadv;

extern "C" int pcadv_backlevel(const pcadv_backlevel_args* a, void* stream) {
  PCADV_CHECK_ARG(a != nullptr, "pcadv_backlevel: null args");
  PCADV_CHECK_ARG(a->rows >= 0 && a->n > 0, "pcadv_backlevel: bad shape rows=%lld n=%d", (long long)a->rows, a->n);
  PCADV_CHECK_ARG(a->num_seg >= 1 && a->num_seg <= PCADV_MAX_SEG, "pcadv_backlevel: num_seg=%d", a->num_seg);
  for (int i = 0; i < a->num_seg; ++i)
    PCADV_CHECK_ARG((a->seg[i].ptr != nullptr || a->onehot_idx != nullptr) && a->seg[i].k > 0,
                    "pcadv_backlevel: bad segment %d", i);
  PCADV_CHECK_ARG(a->w && a->x && a->dz_out, "pcadv_backlevel: w, x and dz_out are required");
  PCADV_CHECK_ARG(a->mask_act == PCADV_ACT_NONE || a->mask_bits != nullptr,
                  "pcadv_backlevel: an activation mask needs mask_bits");
  if (a->rows == 0) return 0;
  const int rc = tc_backlevel(*a, static_cast<cudaStream_t>(stream));
  PCADV_CHECK_ARG(rc >= 0, "pcadv_backlevel: shape not supported (16-bit TMA-compatible operands, every k and n "
                           "a multiple of 64, rows >= 128, sum k <= 1024, rows_per_group %% 128 == 0 with dgroup)");
  return rc;
}
