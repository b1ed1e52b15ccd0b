// Tensor-core engine, weight gradients:  dW[c, koff + k] += scale * sum_r dz[r, c] x[r, k].
//
// The reduction runs over points, so both operands are MN-major in shared memory: the
// M side is dz^T (channels c on M, rows on K), the N side is x^T (k on N, rows on K); both
// are fetched as [64 rows][64 channels] TMA boxes (128 B inner, 128B swizzle) straight from
// the point-major activations -- no transposed copies.  The K-concat of a layer's input
// segments (x1..x5 of fc1) is one list of 64-channel boxes; an N tile is up to eight
// consecutive boxes (512 accumulator columns = all of TMEM, two N <= 256 MMAs per K step), so
// dz is re-read once per N tile and x once per 128-channel M tile.  Work = (tile, row range),
// one wave of CTAs; every CTA accumulates its range in TMEM and adds the tile into dW with
// vector fp32 RED atomics.  A pipeline stage carries 64, 128 or 256 rows (as many as fit in
// ~48 KB) so narrow layers do not pay one barrier round trip per 64 rows.  The bias gradient
// and the per-cloud bias gradient (column sums of dz) are taken from the dz boxes in flight by
// the otherwise idle epilogue warps.
#include <stdlib.h>
#include "tc_pipeline.cuh"

namespace pcadv {
namespace tc {

constexpr int kWgradThreads = 192;
constexpr int kMaxBoxes = 8;                      // N tile <= 512 columns
constexpr int kMaxNTiles = 16;
constexpr int kWgradMaxStages = 6;
constexpr int kWgradSmemMax = 232448;

struct WgradTail {
  uint64_t full[kWgradMaxStages];
  uint64_t empty[kWgradMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

struct WgradParams {
  int64_t rows;
  int n;                          // dz channels
  int num_seg;
  int seg_k[PCADV_MAX_SEG];       // channels of each x segment (multiples of 64)
  int seg_koff[PCADV_MAX_SEG];    // column offset of the segment inside dw
  int seg_box0[PCADV_MAX_SEG + 1];   // first global 64-channel box of each segment
  int tile_box0[kMaxNTiles + 1];  // global box range of each N tile
  int tiles_m, tiles_n;
  int splits;
  int64_t rows_per_split;         // multiple of stage_rows
  int box_rows;                   // rows (K extent) of one TMA box: 32 for wide N tiles, else 64
  int stage_rows;                 // box_rows * sub-chunks per stage: 32 .. 256
  int stage_bytes, nstages, nbuf;
  int bf16;
  float* dw;
  int64_t ld_dw;
  int vec_red;                    // dw rows are 16-byte aligned: red.global.add.v4.f32
  float* dbias;                   // [n] or NULL: summed by one extra N = 16 MMA against a ones column
  float* dgroup_bias;             // [rows / rows_per_group, n] or NULL (NOT scaled)
  int64_t rows_per_group;
  const float* scale;
  // swapped orientation (host): the x segment is the M side and dz the N side, so that every byte of a wide dz is
  // read once; the accumulator then holds dw^T -- `transposed_out` stores it transposed -- and the bias gradient
  // is the column sum of the N-side boxes (`nside_bias`: taken by the epilogue warps from the boxes in flight)
  int transposed_out;
  float* nside_bias;
};

struct WorkItem {
  int tm, tn, box0, nb;
  int64_t r0, r1;
  bool first_n_tile;
};

__device__ __forceinline__ WorkItem decode_work(const WgradParams& p, int64_t w) {
  WorkItem it;
  const int sp = static_cast<int>(w % p.splits);
  const int64_t tile = w / p.splits;
  it.tn = static_cast<int>(tile % p.tiles_n);
  it.tm = static_cast<int>(tile / p.tiles_n);
  it.box0 = p.tile_box0[it.tn];
  it.nb = p.tile_box0[it.tn + 1] - it.box0;
  it.first_n_tile = it.tn == 0;
  it.r0 = sp * p.rows_per_split;
  it.r1 = it.r0 + p.rows_per_split < p.rows ? it.r0 + p.rows_per_split : p.rows;
  return it;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// acc += the two 16-bit values of a packed pair (mixed-precision add: no separate convert)
template <bool kBf16>
__device__ __forceinline__ void add_pair_f32(uint32_t packed, float2& acc) {
  if (kBf16) {
    asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nadd.rn.f32.bf16 %0, lo, %0;\nadd.rn.f32.bf16 %1, hi, %1;\n}"
        : "+f"(acc.x), "+f"(acc.y) : "r"(packed));
  } else {
    asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\nadd.rn.f32.f16 %0, lo, %0;\nadd.rn.f32.f16 %1, hi, %1;\n}"
        : "+f"(acc.x), "+f"(acc.y) : "r"(packed));
  }
}

__global__ void __launch_bounds__(kWgradThreads, 1)
tc_wgrad_kernel(const __grid_constant__ TensorMaps maps, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ones_box = stages + p.nstages * p.stage_bytes;        // [64 rows][64 ch], column 0 = 1
  WgradTail* st = reinterpret_cast<WgradTail*>(ones_box + 8192);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_work = static_cast<int64_t>(p.tiles_m) * p.tiles_n * p.splits;
  const bool sums = p.dgroup_bias != nullptr;          // CUDA-core column sums (per cloud) only
  const bool nsums = p.nside_bias != nullptr;          // column sums of the N-side boxes (swapped orientation)
  const bool bias_mma = p.dbias != nullptr;
  const int box_rows = p.box_rows;
  const int kBoxBytes = box_rows * 128;                // [box_rows][64 ch] 16-bit box
  const int subs = p.stage_rows / box_rows;            // sub-chunks per stage

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_seg; ++s) tma_prefetch_desc(&maps.act[s]);
    tma_prefetch_desc(&maps.w);
    // with column sums the four epilogue warps also read every dz box, so a stage is released
    // by 1 (MMA commit) + 4 (epilogue warps) arrivals
    for (int i = 0; i < kWgradMaxStages; ++i) {
      mbar_init(&st->full[i], 1);
      mbar_init(&st->empty[i], (sums || nsums) ? 5 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st->tmem_full[i], 1);
      mbar_init(&st->tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&st->tmem_base, kTmemCols);
  if (bias_mma) {
    // the constant N-side box of the bias MMA: element (row rr, channel 0) = 1, rest 0 (128B swizzle)
    for (int i = threadIdx.x; i < 8192 / 16; i += kWgradThreads)
      reinterpret_cast<uint4*>(ones_box)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (threadIdx.x < 64) {
      const int rr = threadIdx.x;
      *reinterpret_cast<uint16_t*>(ones_box + rr * 128 + ((rr & 7) << 4)) = p.bf16 ? 0x3F80 : 0x3C00;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = st->tmem_base;

  if (warp == 0) {
    // TMA producer: the whole warp.  A stage is up to ~20 box loads and one thread issues a
    // tensor copy every ~140 ns, which would cap the pipeline below the MMA rate; lane l issues
    // box l of the stage instead.
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      const int per_sub = 2 + it.nb;
      const int sub_bytes = per_sub * kBoxBytes;
      const int nloads = per_sub * subs;
      const uint32_t stage_tx = static_cast<uint32_t>(sub_bytes * subs);
      for (int64_t r = it.r0; r < it.r1; r += p.stage_rows) {
        mbar_wait_backoff(&st->empty[stage], phase ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(&st->full[stage], stage_tx);
        __syncwarp();
        for (int l = lane; l < nloads; l += 32) {
          const int sb = l / per_sub, b = l - sb * per_sub;
          uint8_t* dst = stages + stage * p.stage_bytes + sb * sub_bytes + b * kBoxBytes;
          const int32_t rr = static_cast<int32_t>(r + sb * box_rows);   // rows >= p.rows: zero fill
          if (b < 2) {
            tma_load_2d(dst, &maps.w, &st->full[stage], it.tm * kTileM + b * 64, rr);
          } else {
            const int gb = it.box0 + b - 2;
            int sg = 0;
            while (gb >= p.seg_box0[sg + 1]) ++sg;
            tma_load_2d(dst, &maps.act[sg], &st->full[stage], (gb - p.seg_box0[sg]) * 64, rr);
          }
        }
        __syncwarp();
        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t buf_phase = 0;
      for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
        const WorkItem it = decode_work(p, w);
        const int nb1 = it.nb < 4 ? it.nb : 4, nb2 = it.nb - nb1;
        const uint32_t idesc1 = make_idesc(kTileM, nb1 * 64, p.bf16 != 0, true, true);
        const uint32_t idesc2 = nb2 > 0 ? make_idesc(kTileM, nb2 * 64, p.bf16 != 0, true, true) : 0u;
        const uint32_t idesc_b = make_idesc(kTileM, 16, p.bf16 != 0, true, true);
        const bool do_bias = bias_mma && it.first_n_tile;
        const uint32_t ones_addr = smem_u32(ones_box);
        const int sub_bytes = (2 + it.nb) * kBoxBytes;
        mbar_wait_backoff(&st->tmem_empty[buf], buf_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kMaxTileN);
        bool first = true;
        for (int64_t r = it.r0; r < it.r1; r += p.stage_rows) {
          mbar_wait_backoff(&st->full[stage], phase);
          tc_fence_after();
          for (int sb = 0; sb < subs; ++sb) {
            const uint32_t a_addr = smem_u32(stages + stage * p.stage_bytes + sb * sub_bytes);
            const uint32_t b_addr = a_addr + 2 * kBoxBytes;
            for (int k = 0; k < box_rows / 16; ++k) {
              const uint64_t adesc = make_smem_desc(a_addr + k * 2048, kBoxBytes, 1024);
              const uint32_t accf = first ? 0u : 1u;
              umma_f16(d_tmem, adesc, make_smem_desc(b_addr + k * 2048, kBoxBytes, 1024), idesc1, accf);
              if (nb2 > 0)
                umma_f16(d_tmem + 256, adesc, make_smem_desc(b_addr + 4 * kBoxBytes + k * 2048, kBoxBytes, 1024),
                         idesc2, accf);
              if (do_bias)       // 16 more accumulator columns: column 0 = sum over rows of dz
                umma_f16(d_tmem + it.nb * 64, adesc, make_smem_desc(ones_addr + k * 2048, 8192, 1024),
                         idesc_b, accf);
              first = false;
            }
          }
          umma_commit(&st->empty[stage]);
          if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&st->tmem_full[buf]);
        if (++buf == p.nbuf) { buf = 0; buf_phase ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int lane_row = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                 // 0..127
    int buf = 0;
    uint32_t buf_phase = 0;
    const float sc = p.scale ? *p.scale : 1.f;
    int stage_e = 0;
    uint32_t phase_e = 0;
    // column sums: thread = (channel pair, row half) of the 128-channel dz tile
    const int pi = et & 63;                           // channel pair 0..63
    const int rh = et >> 6;                           // which half of a sub-chunk's rows
    const uint32_t pair_box = static_cast<uint32_t>(pi >> 5) * kBoxBytes;
    const uint32_t cc = static_cast<uint32_t>(pi & 31) * 2;
    for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      const int sub_bytes = (2 + it.nb) * kBoxBytes;
      const int ch = it.tm * kTileM + pi * 2;         // first channel of this thread's pair
      float2 bsum = make_float2(0.f, 0.f);            // dbias partial of the pair
      float2 gsum = make_float2(0.f, 0.f);            // current cloud's partial
      int64_t cur_g = -1;
      auto flush_group = [&]() {
        if (cur_g >= 0 && ch < p.n) {
          atomicAdd(p.dgroup_bias + cur_g * p.n + ch, gsum.x);
          if (ch + 1 < p.n) atomicAdd(p.dgroup_bias + cur_g * p.n + ch + 1, gsum.y);
        }
        gsum = make_float2(0.f, 0.f);
      };
      if (nsums) {
        // swapped orientation: thread = (N box, 16-byte group of 8 channels, half of the sub-chunk's rows);
        // 16-byte shared loads, 8 fp32 partials, added to nside_bias at the end of the work item
        float nacc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) nacc[e] = 0.f;
        const int nbx = et >> 4;                         // box 0..7
        const uint32_t ng = static_cast<uint32_t>(et >> 1) & 7u;   // 16-byte group inside the 128-byte row
        const int rhalf = et & 1;
        const int hrows = box_rows >> 1;
        for (int64_t r = it.r0; r < it.r1; r += p.stage_rows) {
          mbar_wait(&st->full[stage_e], phase_e);
          if (nbx < it.nb) {
            for (int sb = 0; sb < subs; ++sb) {
              if (r + sb * box_rows >= it.r1) break;
              const uint8_t* tile = stages + stage_e * p.stage_bytes + sb * sub_bytes + (2 + nbx) * kBoxBytes +
                                    rhalf * hrows * 128;
#pragma unroll 8
              for (int i = 0; i < hrows; ++i) {          // hrows is a multiple of 8: (row & 7) == (i & 7)
                const uint4 q = *reinterpret_cast<const uint4*>(tile + i * 128 + ((ng ^ (static_cast<uint32_t>(i) & 7u)) << 4));
                const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 acc = make_float2(nacc[2 * e], nacc[2 * e + 1]);
                  if (p.bf16) add_pair_f32<true>(w4[e], acc);
                  else add_pair_f32<false>(w4[e], acc);
                  nacc[2 * e] = acc.x; nacc[2 * e + 1] = acc.y;
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&st->empty[stage_e]);
          if (++stage_e == p.nstages) { stage_e = 0; phase_e ^= 1; }
        }
        if (it.tm == 0 && it.r0 < it.r1 && nbx < it.nb) {   // every M tile streams the same N boxes: count them once
          const int gb = it.box0 + nbx;
          int s2 = 0;
          while (gb >= p.seg_box0[s2 + 1]) ++s2;
          const int kk = p.seg_koff[s2] + (gb - p.seg_box0[s2]) * 64 + static_cast<int>(ng) * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) atomicAdd(p.nside_bias + kk + e, nacc[e] * sc);
        }
      }
      if (sums) {
        for (int64_t r = it.r0; r < it.r1; r += p.stage_rows) {
          mbar_wait(&st->full[stage_e], phase_e);
          if (it.first_n_tile) {
            for (int sb = 0; sb < subs; ++sb) {
              const int64_t rr0 = r + sb * box_rows;
              if (rr0 >= it.r1) break;
              const uint8_t* tile = stages + stage_e * p.stage_bytes + sb * sub_bytes + pair_box;
              float2 acc = make_float2(0.f, 0.f);
              const int half_rows = box_rows >> 1;
#pragma unroll 8
              for (int i = 0; i < half_rows; ++i) {   // exact: 16-bit value + fp32 accumulator
                const int rr = rh * half_rows + i;
                const uint32_t v = *reinterpret_cast<const uint32_t*>(
                    tile + rr * 128 + ((((cc >> 3) ^ (rr & 7)) << 4) | ((cc & 7) << 1)));
                if (p.bf16) add_pair_f32<true>(v, acc);
                else add_pair_f32<false>(v, acc);
              }
              bsum.x += acc.x; bsum.y += acc.y;
              if (p.dgroup_bias) {
                const int64_t g = rr0 / p.rows_per_group;
                if (g != cur_g) { flush_group(); cur_g = g; }
                gsum.x += acc.x; gsum.y += acc.y;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&st->empty[stage_e]);
          if (++stage_e == p.nstages) { stage_e = 0; phase_e ^= 1; }
        }
        if (it.first_n_tile) flush_group();
      }
      mbar_wait(&st->tmem_full[buf], buf_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(buf * kMaxTileN);
      const int c = it.tm * kTileM + lane_row;
      for (int b = 0; b < it.nb; ++b) {
        // box b of the tile: accumulator columns b*64 .. +63; its place in dw
        const int gb = it.box0 + b;
        int s = 0;
        while (gb >= p.seg_box0[s + 1]) ++s;
        const int k0 = (gb - p.seg_box0[s]) * 64;
        const int seg_k = p.seg_k[s];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v[32];
          tmem_ld32(taddr0 + b * 64 + hh * 32, v);
          if (c >= p.n || it.r0 >= it.r1) continue;
          if (p.transposed_out) {
            // dw[k, c]: consecutive lanes are consecutive c, so every RED instruction is one 128-byte line
            float* dt = p.dw + static_cast<int64_t>(p.seg_koff[s] + k0 + hh * 32) * p.ld_dw + c;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k0 + hh * 32 + j < seg_k) atomicAdd(dt + static_cast<int64_t>(j) * p.ld_dw, v[j] * sc);
            continue;
          }
          float* dst = p.dw + static_cast<int64_t>(c) * p.ld_dw + p.seg_koff[s] + k0 + hh * 32;
          if (p.vec_red && k0 + hh * 32 + 32 <= seg_k) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(dst + j, v[j] * sc, v[j + 1] * sc, v[j + 2] * sc, v[j + 3] * sc);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k0 + hh * 32 + j < seg_k) atomicAdd(dst + j, v[j] * sc);
          }
        }
      }
      if (bias_mma && it.first_n_tile) {
        float v[32];
        tmem_ld32(taddr0 + it.nb * 64, v);
        if (c < p.n && it.r0 < it.r1) atomicAdd(p.dbias + c, v[0] * sc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->tmem_empty[buf]);
      if (++buf == p.nbuf) { buf = 0; buf_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tc

int launch_group_colsum(const void* dz, int dz_dtype, int64_t ld, int64_t rows, int n,
                        int64_t rows_per_group, float* out, cudaStream_t s);

int tc_wgrad_pair(const pcadv_wgrad_args& a, cudaStream_t s);   // tc_wgrad2.cu

static int tc_wgrad_impl(const pcadv_wgrad_args& a, cudaStream_t s, float* nside_bias, bool transposed);

int tc_wgrad(const pcadv_wgrad_args& a, cudaStream_t s) {
  using namespace tc;
  const int dt = a.dz_dtype;
  if (dt == PCADV_F16 || dt == PCADV_BF16) {
    const int rc = tc_wgrad_pair(a, s);                  // CTA-pair kernel for 256 x wide-K layers
    if (rc >= 0) return rc;
  }
  // Swapped orientation for a wide dz over a narrow input (conv5: dz 512 x x 128).  With dz on the M side
  // its 128-channel tiles are separate work items that each stream x again (4 x here, and not every re-read
  // hits L2); with x on the M side (one tile) and dz as up to eight N boxes both are read exactly once:
  // 0.538 -> 0.390 ms per 2^21 points (6.9 TB/s).  The accumulator is dw^T (stored transposed) and the bias
  // gradient, now a column sum of the N side, comes from the boxes in flight.
  static int swap_on = -1;
  if (swap_on < 0) { const char* e = getenv("PCADV_WGRAD_SWAP"); swap_on = (e && atoi(e) == 0) ? 0 : 1; }   // tuning aid
  if (swap_on && (dt == PCADV_F16 || dt == PCADV_BF16) && a.num_seg == 1 && a.dw && !a.dgroup_bias &&
      a.n % 64 == 0 && a.n > kTileM && a.n <= 64 * kMaxBoxes * kMaxNTiles && a.seg[0].dtype == dt &&
      a.seg[0].k % 64 == 0 && a.seg[0].k <= kTileM && tma_compatible(a.seg[0].ptr, dt, a.seg[0].ld) &&
      tma_compatible(a.dz, dt, a.ld_dz)) {
    pcadv_wgrad_args b = a;
    b.dz = a.seg[0].ptr; b.ld_dz = a.seg[0].ld; b.n = a.seg[0].k;
    b.seg[0].ptr = a.dz; b.seg[0].ld = a.ld_dz; b.seg[0].k = a.n;
    b.dbias = nullptr;
    return tc_wgrad_impl(b, s, a.dbias, true);
  }
  return tc_wgrad_impl(a, s, nullptr, false);
}

static int tc_wgrad_impl(const pcadv_wgrad_args& a, cudaStream_t s, float* nside_bias, bool transposed) {
  using namespace tc;
  const int dt = a.dz_dtype;
  PCADV_CHECK_ARG(dt == PCADV_F16 || dt == PCADV_BF16, "tc_wgrad: dz must be fp16 / bf16");
  PCADV_CHECK_ARG(a.dw != nullptr && a.num_seg >= 1, "tc_wgrad: dw and segments required");
  PCADV_CHECK_ARG(a.n % 64 == 0 && tma_compatible(a.dz, dt, a.ld_dz),
                  "tc_wgrad: dz not TMA-compatible (n=%d)", a.n);
  TensorMaps maps;
  WgradParams p{};
  p.rows = a.rows; p.n = a.n; p.num_seg = a.num_seg;
  int koff = 0, boxes = 0;
  for (int i = 0; i < a.num_seg; ++i) {
    const pcadv_seg& sg = a.seg[i];
    PCADV_CHECK_ARG(sg.dtype == dt && sg.k % 64 == 0 && tma_compatible(sg.ptr, dt, sg.ld),
                    "tc_wgrad: segment %d not TMA-compatible", i);
    p.seg_k[i] = sg.k;
    p.seg_koff[i] = koff;
    p.seg_box0[i] = boxes;
    boxes += sg.k / 64;
    koff += sg.k;
  }
  p.seg_box0[a.num_seg] = boxes;
  // N tiles: as few as possible (each one re-reads dz), boxes spread evenly over them
  int max_boxes = kMaxBoxes;
  if (const char* e = getenv("PCADV_WGRAD_MAXBOXES")) {     // tuning aid
    const int v = atoi(e);
    if (v >= 1 && v <= kMaxBoxes) max_boxes = v;
  }
  p.tiles_n = (boxes + max_boxes - 1) / max_boxes;
  // the bias MMA needs 16 accumulator columns next to the first tile's: at most 7 boxes there
  if (a.dbias && (boxes + p.tiles_n - 1) / p.tiles_n > kMaxBoxes - 1) ++p.tiles_n;
  PCADV_CHECK_ARG(p.tiles_n <= kMaxNTiles, "tc_wgrad: K too large (%d boxes)", boxes);
  int max_nb = 0;
  for (int t = 0; t <= p.tiles_n; ++t) p.tile_box0[t] = static_cast<int>(static_cast<int64_t>(boxes) * t / p.tiles_n);
  for (int t = 0; t < p.tiles_n; ++t) max_nb = p.tile_box0[t + 1] - p.tile_box0[t] > max_nb ? p.tile_box0[t + 1] - p.tile_box0[t] : max_nb;
  p.tiles_m = (a.n + kTileM - 1) / kTileM;
  p.nbuf = (max_nb * 64 + (a.dbias ? 16 : 0) > kMaxTileN) ? 1 : 2;
  // wide N tiles take 32-row boxes so that the ring still has >= 4 stages to hide the load
  // latency behind the MMAs; narrow ones put several 64-row sub-chunks into one stage
  p.box_rows = 64;
  if (const char* e = getenv("PCADV_WGRAD_BOXROWS")) {      // tuning aid
    if (atoi(e) == 32 && max_nb > 4) p.box_rows = 32;
  }
  if (int rc = encode_tmap_2d(&maps.w, a.dz, dt, a.rows, a.n, a.ld_dz, 64, p.box_rows)) return rc;
  for (int i = 0; i < a.num_seg; ++i)
    if (int rc = encode_tmap_2d(&maps.act[i], a.seg[i].ptr, dt, a.rows, a.seg[i].k, a.seg[i].ld, 64, p.box_rows))
      return rc;
  const int sub_bytes = (2 + max_nb) * p.box_rows * 128;
  const int budget = kWgradSmemMax - 2048 - 8192 - static_cast<int>(sizeof(WgradTail));
  int stage_rows = 256;
  while (stage_rows > p.box_rows && budget / (sub_bytes * (stage_rows / p.box_rows)) < 3) stage_rows >>= 1;
  p.stage_rows = stage_rows;
  p.stage_bytes = sub_bytes * (stage_rows / p.box_rows);
  p.nstages = budget / p.stage_bytes;
  if (p.nstages > kWgradMaxStages) p.nstages = kWgradMaxStages;
  PCADV_CHECK_ARG(p.nstages >= 2, "tc_wgrad: shared memory budget exceeded");
  const int tiles = p.tiles_m * p.tiles_n;
  // one wave: tiles * splits <= #SMs, so no CTA gets a second work item (a 2x tail)
  int64_t splits = num_sms() / tiles;
  const int64_t max_splits = (a.rows + 511) / 512;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rps = (a.rows + splits - 1) / splits;
  rps = (rps + stage_rows - 1) / stage_rows * stage_rows;
  p.splits = static_cast<int>((a.rows + rps - 1) / rps);
  p.rows_per_split = rps;
  p.bf16 = dt == PCADV_BF16 ? 1 : 0;
  p.dw = a.dw; p.ld_dw = a.ld_dw; p.scale = a.scale; p.dbias = a.dbias;
  p.transposed_out = transposed ? 1 : 0;
  p.nside_bias = nside_bias;
  p.vec_red = (a.ld_dw % 4 == 0 && (reinterpret_cast<uintptr_t>(a.dw) & 15) == 0) ? 1 : 0;
  // the per-cloud column sums ride along when no 64-row sub-chunk straddles two clouds
  const bool fuse_group = a.dgroup_bias && a.rows_per_group % p.box_rows == 0;
  p.dgroup_bias = fuse_group ? a.dgroup_bias : nullptr;
  p.rows_per_group = a.rows_per_group;
  const size_t smem = 1024 + static_cast<size_t>(p.nstages) * p.stage_bytes + 8192 + sizeof(WgradTail) + 16;
  static bool attr_done = false;
  if (!attr_done) {
    PCADV_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kWgradSmemMax));
    attr_done = true;
  }
  const int64_t work = static_cast<int64_t>(tiles) * p.splits;
  const int grid = static_cast<int>(work < num_sms() ? work : num_sms());
  tc_wgrad_kernel<<<grid, kWgradThreads, smem, s>>>(maps, p);
  PCADV_LAUNCHED();
  if (a.dgroup_bias && !fuse_group) {
    if (int rc = launch_group_colsum(a.dz, a.dz_dtype, a.ld_dz, a.rows, a.n, a.rows_per_group,
                                     a.dgroup_bias, s))
      return rc;
  }
  return 0;
}

}  // namespace pcadv
