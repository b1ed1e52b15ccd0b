// Tensor-core engine: warp-specialised persistent GEMM kernels on tcgen05.mma with
// TMEM accumulators, fed by TMA through a 4-stage mbarrier pipeline.
//
//   tc_linear_kernel<false>  rows (points) on the MMA M axis, output channels on N:
//                            bias / per-cloud bias / addend / activation / activation-
//                            derivative mask / scale epilogue, optional max over channels.
//   tc_linear_kernel<true>   "swapped": output channels on M, points on N, so the max
//                            over a cloud's points is a per-thread running max over TMEM
//                            columns (conv6 + ReLU + max-pool, never storing the map).
//   tc_wgrad_kernel          dW = dZ^T X with both operands MN-major in shared memory
//                            (the reduction runs over points), split over row ranges,
//                            fp32 RED accumulation into dW.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
// (one elected lane), warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#include "tc_common.cuh"

namespace pcadv {
namespace tc {

constexpr int kStages = 4;
constexpr int kBlockK = 64;                       // elements per K chunk = one 128-byte swizzle row
constexpr int kTileM = 128;
constexpr int kMaxTileN = 256;
constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KB
constexpr int kBBytes = kMaxTileN * kBlockK * 2;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;

struct TensorMaps {
  CUtensorMap act[PCADV_MAX_SEG];   // activation segments [rows, k_i]
  CUtensorMap w;                    // weights [n, ktot]  (wgrad: dz [rows, n])
};

struct LinearParams {
  int64_t rows;
  int n;
  int bn;                 // N-side tile (multiple of 16, <= 256)
  int num_seg;
  int seg_k[PCADV_MAX_SEG];
  int64_t tiles_m, tiles_n;
  uint32_t idesc;
  // epilogue
  const float* bias;
  const float* group_bias;
  int64_t rows_per_group;
  const float* addend;
  int64_t ld_addend;
  int act;
  float slope;
  const void* mask;
  int64_t ld_mask;
  int mask_dtype, mask_act;
  float mask_slope;
  const float* out_scale;
  void* out;
  int64_t ld_out;
  int out_dtype;
  unsigned long long* colmax_key;
  unsigned long long* rowmax_key;
};

struct SharedTail {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// ---- epilogue helpers ------------------------------------------------------------------
__device__ __forceinline__ void store_row_chunk(void* out, int out_dtype, int64_t off, int valid,
                                                const float (&v)[32], bool vec_ok) {
  if (out_dtype == PCADV_F32) {
    float* p = reinterpret_cast<float*>(out) + off;
    if (vec_ok && valid == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) p[j] = v[j];
    }
  } else {
    uint32_t packed[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (out_dtype == PCADV_F16) {
        const float a = fminf(fmaxf(v[2 * j], -65504.f), 65504.f);
        const float b = fminf(fmaxf(v[2 * j + 1], -65504.f), 65504.f);
        __half2 h = __floats2half2_rn(a, b);
        packed[j] = *reinterpret_cast<uint32_t*>(&h);
      } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        packed[j] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    uint16_t* p = reinterpret_cast<uint16_t*>(out) + off;
    if (vec_ok && valid == 32) {
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<uint4*>(p + 2 * j) =
            make_uint4(packed[j], packed[j + 1], packed[j + 2], packed[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) p[j] = static_cast<uint16_t>((packed[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
    }
  }
}

__device__ __forceinline__ void load_row_chunk(const void* src, int dtype, int64_t off, int valid,
                                               float (&v)[32], bool vec_ok) {
  if (dtype == PCADV_F32) {
    const float* p = reinterpret_cast<const float*>(src) + off;
    if (vec_ok && valid == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(p + j);
        v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = j < valid ? p[j] : 0.f;
    }
  } else {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(src) + off;
    if (vec_ok && valid == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const uint4 t = *reinterpret_cast<const uint4*>(p + j);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 f;
          if (dtype == PCADV_F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
          else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[q]));
          v[j + 2 * q] = f.x; v[j + 2 * q + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = j < valid ? ld_as_float(src, off + j, dtype) : 0.f;
    }
  }
}

// =====================================================================================
template <bool kSwapped>
__global__ void __launch_bounds__(kThreads, 1)
tc_linear_kernel(const __grid_constant__ TensorMaps maps, const LinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  SharedTail* st = reinterpret_cast<SharedTail*>(smem + kStages * kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_tiles = p.tiles_m * p.tiles_n;
  int total_chunks = 0;
  for (int s = 0; s < p.num_seg; ++s) total_chunks += p.seg_k[s] / kBlockK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_seg; ++s) tma_prefetch_desc(&maps.act[s]);
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < kStages; ++i) { mbar_init(&st->full[i], 1); mbar_init(&st->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&st->tmem_full[i], 1); mbar_init(&st->tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&st->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = st->tmem_base;

  // M-side tile = 128 rows x 64 K; N-side tile = bn rows x 64 K.  Non-swapped: M = points
  // (activations), N = channels (weights).  Swapped: M = channels, N = points.
  const uint32_t stage_tx = static_cast<uint32_t>((kTileM + p.bn) * kBlockK * 2);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int64_t tm, tn;
        if (kSwapped) { tn = t / p.tiles_m; tm = t % p.tiles_m; }
        else { tm = t / p.tiles_n; tn = t % p.tiles_n; }
        const int32_t m0 = static_cast<int32_t>(tm * kTileM);
        const int32_t n0 = static_cast<int32_t>(tn * p.bn);
        int kg = 0;
        for (int s = 0; s < p.num_seg; ++s) {
          for (int kk = 0; kk < p.seg_k[s]; kk += kBlockK, kg += kBlockK) {
            mbar_wait(&st->empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&st->full[stage], stage_tx);
            uint8_t* sa = smem + stage * kStageBytes;
            uint8_t* sb = sa + kABytes;
            if (kSwapped) {
              tma_load_2d(sa, &maps.w, &st->full[stage], kg, m0);
              tma_load_2d(sb, &maps.act[s], &st->full[stage], kk, n0);
            } else {
              tma_load_2d(sa, &maps.act[s], &st->full[stage], kk, m0);
              tma_load_2d(sb, &maps.w, &st->full[stage], kg, n0);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t buf_phase = 0;
      for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&st->tmem_empty[buf], buf_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kMaxTileN);
        for (int c = 0; c < total_chunks; ++c) {
          mbar_wait(&st->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // K-major, 128B swizzle: 8-row groups 1024 B apart; +32 B per K=16 step
            const uint64_t adesc = make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_f16(d_tmem, adesc, bdesc, p.idesc, (c | k) != 0 ? 1u : 0u);
          }
          umma_commit(&st->empty[stage]);          // frees the smem slot when the MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&st->tmem_full[buf]);          // accumulator complete -> epilogue
        if (++buf == 2) { buf = 0; buf_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (4 warps, 128 TMEM lanes) =================
    const int quarter = warp & 3;
    const int lane_row = quarter * 32 + lane;            // row of the 128-row tile
    int buf = 0;
    uint32_t buf_phase = 0;
    const float oscale = p.out_scale ? *p.out_scale : 1.f;
    const int esz_out = p.out_dtype == PCADV_F32 ? 4 : 2;
    const bool out_vec = p.out && ((p.ld_out * esz_out) % 16 == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const int esz_mask = p.mask_dtype == PCADV_F32 ? 4 : 2;
    const bool mask_vec = p.mask && ((p.ld_mask * esz_mask) % 16 == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0);
    const bool add_vec = p.addend && ((p.ld_addend * 4) % 16 == 0) &&
                         ((reinterpret_cast<uintptr_t>(p.addend) & 15) == 0);
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int64_t tm, tn;
      if (kSwapped) { tn = t / p.tiles_m; tm = t % p.tiles_m; }
      else { tm = t / p.tiles_n; tn = t % p.tiles_n; }
      mbar_wait(&st->tmem_full[buf], buf_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(buf * kMaxTileN);
      if (!kSwapped) {
        const int64_t r = tm * kTileM + lane_row;
        const bool r_ok = r < p.rows;
        const int64_t g = (p.rows_per_group > 0 && r_ok) ? r / p.rows_per_group : 0;
        const int col_base = static_cast<int>(tn * p.bn);
        unsigned long long rkey = 0ull;
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
          float v[32];
          tmem_ld32(taddr0 + c0, v);
          const int cg = col_base + c0;
          const int valid = p.n - cg < 32 ? (p.n - cg > 0 ? p.n - cg : 0) : 32;
          if (!r_ok || valid <= 0) continue;
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < valid) v[j] += __ldg(p.bias + cg + j);
          }
          if (p.group_bias) {
            const float* gb = p.group_bias + g * p.n + cg;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < valid) v[j] += __ldg(gb + j);
          }
          if (p.addend) {
            float a[32];
            load_row_chunk(p.addend, PCADV_F32, r * p.ld_addend + cg, valid, a, add_vec);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += a[j];
          }
          if (p.rowmax_key) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < valid) {
                const unsigned long long k = pack_key(v[j], static_cast<uint32_t>(cg + j));
                rkey = k > rkey ? k : rkey;
              }
            }
          }
          if (p.out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act, p.slope);
            if (p.mask) {
              float m[32];
              load_row_chunk(p.mask, p.mask_dtype, r * p.ld_mask + cg, valid, m, mask_vec);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= act_grad_from_output(m[j], p.mask_act, p.mask_slope);
            }
            if (p.out_scale) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= oscale;
            }
            store_row_chunk(p.out, p.out_dtype, r * p.ld_out + cg, valid, v, out_vec);
          }
        }
        if (p.rowmax_key && r_ok && rkey) atomicMax(&p.rowmax_key[r], rkey);
      } else {
        // thread = output channel; columns = points.  Running (max, first index) per cloud.
        const int ch = static_cast<int>(tm * kTileM) + lane_row;
        const bool ch_ok = ch < p.n;
        const int64_t p0 = tn * p.bn;
        const int64_t rpg = p.rows_per_group;
        int64_t g = p0 / rpg;
        int64_t g_end = (g + 1) * rpg;                  // first row of the next cloud
        float best = -INFINITY;
        int64_t best_row = -1;
        const float bias = (p.bias && ch_ok) ? p.bias[ch] : 0.f;
        for (int c0 = 0; c0 < p.bn; c0 += 32) {
          float v[32];
          tmem_ld32(taddr0 + c0, v);
          const int64_t pr = p0 + c0;
          if (!ch_ok || pr >= p.rows) continue;
          if (pr + 32 <= g_end && pr + 32 <= p.rows) {
            // fast path: the whole chunk lies in one cloud
            float m = v[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
            if (m > best) {
              int jj = 0;
#pragma unroll
              for (int j = 31; j >= 0; --j) jj = (v[j] == m) ? j : jj;
              best = m;
              best_row = pr + jj;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t row = pr + j;
              if (row >= p.rows) break;
              if (row >= g_end) {
                if (best_row >= 0)
                  atomicMax(&p.colmax_key[g * p.n + ch],
                            pack_key(best + bias, static_cast<uint32_t>(best_row - g * rpg)));
                g = row / rpg;
                g_end = (g + 1) * rpg;
                best = -INFINITY;
                best_row = -1;
              }
              if (v[j] > best) { best = v[j]; best_row = row; }
            }
          }
        }
        if (ch_ok && best_row >= 0)
          atomicMax(&p.colmax_key[g * p.n + ch],
                    pack_key(best + bias, static_cast<uint32_t>(best_row - g * rpg)));
      }
      // release the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->tmem_empty[buf]);
      if (++buf == 2) { buf = 0; buf_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// =====================================================================================
// wgrad: dW[c, koff + k] += scale * sum_r dz[r, c] x[r, k]
//   M side = dz^T (channels c on M, rows on K), N side = x^T (k on N, rows on K); both tiles
//   are loaded as [64 rows][64 channels] boxes (128 B inner, 128B swizzle) = MN-major.
// =====================================================================================
struct WgradParams {
  int64_t rows;
  int n;                  // dz channels
  int seg_index;
  int seg_k;              // channels of this x segment (multiple of 64)
  int koff;               // column offset of the segment inside dw
  int bn;                 // N tile (multiple of 64, <= 256)
  int tiles_m, tiles_n;
  int splits;
  int64_t rows_per_split; // multiple of 64
  uint32_t idesc;
  float* dw;
  int64_t ld_dw;
  const float* scale;
};

__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_kernel(const __grid_constant__ TensorMaps maps, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  SharedTail* st = reinterpret_cast<SharedTail*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_work = static_cast<int64_t>(p.tiles_m) * p.tiles_n * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.act[p.seg_index]);
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < kStages; ++i) { mbar_init(&st->full[i], 1); mbar_init(&st->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&st->tmem_full[i], 1); mbar_init(&st->tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&st->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = st->tmem_base;
  const int n_boxes = p.bn / 64;
  const uint32_t stage_tx = static_cast<uint32_t>((kTileM + p.bn) * kBlockK * 2);

  auto decode = [&](int64_t w, int& tm, int& tn, int64_t& r0, int64_t& r1) {
    const int sp = static_cast<int>(w % p.splits);
    const int64_t tile = w / p.splits;
    tn = static_cast<int>(tile % p.tiles_n);
    tm = static_cast<int>(tile / p.tiles_n);
    r0 = sp * p.rows_per_split;
    r1 = r0 + p.rows_per_split < p.rows ? r0 + p.rows_per_split : p.rows;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
        int tm, tn; int64_t r0, r1;
        decode(w, tm, tn, r0, r1);
        for (int64_t r = r0; r < r1; r += kBlockK) {
          mbar_wait(&st->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&st->full[stage], stage_tx);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          // M side: two [64 rows][64 ch] boxes; N side: bn/64 boxes
          tma_load_2d(sa, &maps.w, &st->full[stage], tm * kTileM, static_cast<int32_t>(r));
          tma_load_2d(sa + 8192, &maps.w, &st->full[stage], tm * kTileM + 64, static_cast<int32_t>(r));
          for (int b = 0; b < n_boxes; ++b)
            tma_load_2d(sb + b * 8192, &maps.act[p.seg_index], &st->full[stage],
                        tn * p.bn + b * 64, static_cast<int32_t>(r));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t buf_phase = 0;
      for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
        int tm, tn; int64_t r0, r1;
        decode(w, tm, tn, r0, r1);
        mbar_wait(&st->tmem_empty[buf], buf_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kMaxTileN);
        bool first = true;
        for (int64_t r = r0; r < r1; r += kBlockK) {
          mbar_wait(&st->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // MN-major, 128B swizzle: 64-channel blocks 8192 B apart (LBO), 8-row K groups
            // 1024 B apart (SBO); +16 rows * 128 B per K=16 step
            const uint64_t adesc = make_smem_desc(a_addr + k * 2048, 8192, 1024);
            const uint64_t bdesc = make_smem_desc(b_addr + k * 2048, 8192, 1024);
            umma_f16(d_tmem, adesc, bdesc, p.idesc, (first && k == 0) ? 0u : 1u);
          }
          first = false;
          umma_commit(&st->empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&st->tmem_full[buf]);
        if (++buf == 2) { buf = 0; buf_phase ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int lane_row = quarter * 32 + lane;
    int buf = 0;
    uint32_t buf_phase = 0;
    const float sc = p.scale ? *p.scale : 1.f;
    for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
      int tm, tn; int64_t r0, r1;
      decode(w, tm, tn, r0, r1);
      mbar_wait(&st->tmem_full[buf], buf_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(buf * kMaxTileN);
      const int c = tm * kTileM + lane_row;
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        float v[32];
        tmem_ld32(taddr0 + c0, v);
        const int k0 = tn * p.bn + c0;
        if (c >= p.n || r0 >= r1) continue;
        float* dst = p.dw + static_cast<int64_t>(c) * p.ld_dw + p.koff + k0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (k0 + j < p.seg_k) atomicAdd(dst + j, v[j] * sc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->tmem_empty[buf]);
      if (++buf == 2) { buf = 0; buf_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

int encode_tmap_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int64_t cols,
                   int64_t ld, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  PCADV_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt =
      dtype == PCADV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PCADV_CHECK_ARG(r == CUDA_SUCCESS,
                  "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
                  (long long)rows, (long long)cols, (long long)ld, box_cols, box_rows);
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int ensure_smem(const void* kernel) {
  static const void* done[8] = {nullptr};
  for (int i = 0; i < 8; ++i) {
    if (done[i] == kernel) return 0;
    if (done[i] == nullptr) {
      PCADV_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBytes));
      done[i] = kernel;
      return 0;
    }
  }
  return 0;
}

}  // namespace tc

int tc_linear(const pcadv_linear_args& a, cudaStream_t s) {
  using namespace tc;
  const int dt = a.seg[0].dtype;
  PCADV_CHECK_ARG(dt == PCADV_F16 || dt == PCADV_BF16, "tc_linear: operands must be fp16 / bf16");
  PCADV_CHECK_ARG(a.w_dtype == dt, "tc_linear: weight dtype differs from activations");
  const bool swapped = a.colmax_key != nullptr;
  PCADV_CHECK_ARG(!swapped || (!a.out && !a.rowmax_key && !a.group_bias && !a.addend && !a.mask),
                  "tc_linear: the max-over-points kernel takes bias only");
  TensorMaps maps;
  LinearParams p{};
  p.rows = a.rows; p.n = a.n; p.num_seg = a.num_seg;
  if (swapped) p.bn = kMaxTileN;
  else { p.bn = (a.n + 15) / 16 * 16; if (p.bn > kMaxTileN) p.bn = kMaxTileN; }
  int ktot = 0;
  for (int i = 0; i < a.num_seg; ++i) {
    PCADV_CHECK_ARG(a.seg[i].dtype == dt && a.seg[i].k % kBlockK == 0 && a.seg[i].ld % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(a.seg[i].ptr) & 15) == 0,
                    "tc_linear: segment %d not TMA-compatible (k=%d ld=%lld)", i, a.seg[i].k,
                    (long long)a.seg[i].ld);
    p.seg_k[i] = a.seg[i].k;
    int rc = encode_tmap_2d(&maps.act[i], a.seg[i].ptr, dt, a.rows, a.seg[i].k, a.seg[i].ld, kBlockK,
                            swapped ? p.bn : kTileM);
    if (rc) return rc;
    ktot += a.seg[i].k;
  }
  PCADV_CHECK_ARG(a.ldw % 8 == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0,
                  "tc_linear: weight not TMA-compatible");
  {
    int rc = encode_tmap_2d(&maps.w, a.w, dt, a.n, ktot, a.ldw, kBlockK, swapped ? kTileM : p.bn);
    if (rc) return rc;
  }
  if (swapped) {
    p.tiles_m = (a.n + kTileM - 1) / kTileM;
    p.tiles_n = (a.rows + p.bn - 1) / p.bn;
  } else {
    p.tiles_m = (a.rows + kTileM - 1) / kTileM;
    p.tiles_n = (a.n + p.bn - 1) / p.bn;
  }
  p.idesc = make_idesc(kTileM, p.bn, dt == PCADV_BF16, false, false);
  p.bias = a.bias; p.group_bias = a.group_bias; p.rows_per_group = a.rows_per_group;
  p.addend = a.addend; p.ld_addend = a.ld_addend; p.act = a.act; p.slope = a.slope;
  p.mask = a.mask; p.ld_mask = a.ld_mask; p.mask_dtype = a.mask_dtype; p.mask_act = a.mask_act;
  p.mask_slope = a.mask_slope; p.out_scale = a.out_scale; p.out = a.out; p.ld_out = a.ld_out;
  p.out_dtype = a.out_dtype; p.colmax_key = a.colmax_key; p.rowmax_key = a.rowmax_key;
  const int64_t tiles = p.tiles_m * p.tiles_n;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  if (swapped) {
    if (int rc = ensure_smem(reinterpret_cast<const void*>(&tc_linear_kernel<true>))) return rc;
    tc_linear_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(maps, p);
  } else {
    if (int rc = ensure_smem(reinterpret_cast<const void*>(&tc_linear_kernel<false>))) return rc;
    tc_linear_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(maps, p);
  }
  PCADV_LAUNCHED();
  return 0;
}

int launch_colsum(const void* dz, int dz_dtype, int64_t ld, int64_t rows, int n, const float* scale,
                  float* out, cudaStream_t s);
int launch_group_colsum(const void* dz, int dz_dtype, int64_t ld, int64_t rows, int n,
                        int64_t rows_per_group, float* out, cudaStream_t s);

int tc_wgrad(const pcadv_wgrad_args& a, cudaStream_t s) {
  using namespace tc;
  const int dt = a.dz_dtype;
  PCADV_CHECK_ARG(dt == PCADV_F16 || dt == PCADV_BF16, "tc_wgrad: dz must be fp16 / bf16");
  PCADV_CHECK_ARG(a.dw != nullptr && a.num_seg >= 1, "tc_wgrad: dw and segments required");
  PCADV_CHECK_ARG(a.n % 64 == 0 && a.ld_dz % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(a.dz) & 15) == 0,
                  "tc_wgrad: dz not TMA-compatible (n=%d)", a.n);
  TensorMaps maps;
  if (int rc = encode_tmap_2d(&maps.w, a.dz, dt, a.rows, a.n, a.ld_dz, 64, kBlockK)) return rc;
  if (int rc = ensure_smem(reinterpret_cast<const void*>(&tc_wgrad_kernel))) return rc;
  int koff = 0;
  for (int i = 0; i < a.num_seg; ++i) {
    const pcadv_seg& sg = a.seg[i];
    PCADV_CHECK_ARG(sg.dtype == dt && sg.k % 64 == 0 && sg.ld % 8 == 0 &&
                        (reinterpret_cast<uintptr_t>(sg.ptr) & 15) == 0,
                    "tc_wgrad: segment %d not TMA-compatible", i);
    if (int rc = encode_tmap_2d(&maps.act[i], sg.ptr, dt, a.rows, sg.k, sg.ld, 64, kBlockK)) return rc;
    WgradParams p{};
    p.rows = a.rows; p.n = a.n; p.seg_index = i; p.seg_k = sg.k; p.koff = koff;
    p.bn = sg.k < kMaxTileN ? sg.k : kMaxTileN;
    p.tiles_m = (a.n + kTileM - 1) / kTileM;
    p.tiles_n = (sg.k + p.bn - 1) / p.bn;
    const int tiles = p.tiles_m * p.tiles_n;
    int64_t splits = (num_sms() + tiles - 1) / tiles;
    const int64_t max_splits = (a.rows + 511) / 512;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t rps = (a.rows + splits - 1) / splits;
    rps = (rps + kBlockK - 1) / kBlockK * kBlockK;
    p.splits = static_cast<int>((a.rows + rps - 1) / rps);
    p.rows_per_split = rps;
    p.idesc = make_idesc(kTileM, p.bn, dt == PCADV_BF16, true, true);
    p.dw = a.dw; p.ld_dw = a.ld_dw; p.scale = a.scale;
    const int64_t work = static_cast<int64_t>(tiles) * p.splits;
    const int grid = static_cast<int>(work < num_sms() ? work : num_sms());
    tc_wgrad_kernel<<<grid, kThreads, kSmemBytes, s>>>(maps, p);
    PCADV_LAUNCHED();
    koff += sg.k;
  }
  if (a.dbias) {
    if (int rc = launch_colsum(a.dz, a.dz_dtype, a.ld_dz, a.rows, a.n, a.scale, a.dbias, s)) return rc;
  }
  if (a.dgroup_bias) {
    if (int rc = launch_group_colsum(a.dz, a.dz_dtype, a.ld_dz, a.rows, a.n, a.rows_per_group,
                                     a.dgroup_bias, s))
      return rc;
  }
  return 0;
}

}  // namespace pcadv
