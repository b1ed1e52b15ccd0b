// Tensor-core engine, forward / dgrad GEMMs: warp-specialised persistent kernels on
// tcgen05.mma with TMEM accumulators, fed by TMA through an mbarrier ring.
//
//   tc_rows_kernel<act, out>  rows (points) on the MMA M axis, output channels on N.
//       320 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 =
//       epilogue.  Two epilogue warps share each TMEM lane quarter and split every 64-column
//       step into its two 32-column chunks, so each SM sub-partition has two epilogue warps to
//       interleave.  Bias and the (at most two) per-cloud bias rows of the tile are staged in
//       shared memory; the activation-derivative mask tile arrives by TMA; the output leaves
//       through 128-byte-swizzled shared-memory slabs and TMA stores (full-line writes).
//   tc_colmax_kernel          "swapped": output channels on M, points on N, so the max over a
//       cloud's points is a per-thread running max over TMEM columns (conv6 + ReLU + max-pool
//       without ever storing the B x 2048 x N map).  192 threads.
#include "tc_pipeline.cuh"

namespace pcadv {
namespace tc {

constexpr int kRowsThreads = 320;
constexpr int kColmaxThreads = 192;

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
  float slope;
  const void* mask;
  int64_t ld_mask;
  int mask_dtype, mask_act;
  float mask_slope;
  const float* out_scale;
  void* out;
  int64_t ld_out;
  unsigned long long* colmax_key;
  unsigned long long* rowmax_key;
  int tma_out;            // stage the output through swizzled smem slabs + TMA store
  int tma_mask;           // fetch the mask tile with TMA into smem slabs
};

template <bool kSwapped>
__device__ __forceinline__ void tile_coords(const LinearParams& p, int64_t t, int64_t& tm, int64_t& tn) {
  if (kSwapped) { tn = t / p.tiles_m; tm = t % p.tiles_m; }     // channel tiles fastest
  else { tm = t / p.tiles_n; tn = t % p.tiles_n; }
}

template <bool kSwapped>
__device__ __forceinline__ void producer_loop(const TensorMaps& maps, const LinearParams& p,
                                              const SmemLayout& L, int nstages) {
  SharedTail* st = L.tail;
  const int64_t num_tiles = p.tiles_m * p.tiles_n;
  const uint32_t stage_tx = static_cast<uint32_t>((kTileM + p.bn) * kBlockK * 2);
  int stage = 0;
  uint32_t phase = 0;
  for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    int64_t tm, tn;
    tile_coords<kSwapped>(p, t, tm, tn);
    const int32_t m0 = static_cast<int32_t>(tm * kTileM);
    const int32_t n0 = static_cast<int32_t>(tn * p.bn);
    int kg = 0;
    for (int s = 0; s < p.num_seg; ++s) {
      for (int kk = 0; kk < p.seg_k[s]; kk += kBlockK, kg += kBlockK) {
        mbar_wait_backoff(&st->empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&st->full[stage], stage_tx);
        uint8_t* sa = L.stages + stage * kStageBytes;
        uint8_t* sb = sa + kABytes;
        if (kSwapped) {
          tma_load_2d(sa, &maps.w, &st->full[stage], kg, m0);
          tma_load_2d(sb, &maps.act[s], &st->full[stage], kk, n0);
        } else {
          tma_load_2d(sa, &maps.act[s], &st->full[stage], kk, m0);
          tma_load_2d(sb, &maps.w, &st->full[stage], kg, n0);
        }
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  }
}

__device__ __forceinline__ void mma_loop(const LinearParams& p, const SmemLayout& L,
                                         uint32_t tmem_base, int nstages) {
  SharedTail* st = L.tail;
  const int64_t num_tiles = p.tiles_m * p.tiles_n;
  int total_chunks = 0;
  for (int s = 0; s < p.num_seg; ++s) total_chunks += p.seg_k[s] / kBlockK;
  int stage = 0;
  uint32_t phase = 0;
  int buf = 0;
  uint32_t buf_phase = 0;
  for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    mbar_wait_backoff(&st->tmem_empty[buf], buf_phase ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kMaxTileN);
    for (int c = 0; c < total_chunks; ++c) {
      mbar_wait_backoff(&st->full[stage], phase);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(L.stages + stage * kStageBytes);
      mma_chunk_kmajor(d_tmem, a_addr, a_addr + kABytes, p.idesc, c == 0);
      umma_commit(&st->empty[stage]);          // frees the smem slot when the MMAs retire
      if (++stage == nstages) { stage = 0; phase ^= 1; }
    }
    umma_commit(&st->tmem_full[buf]);          // accumulator complete -> epilogue
    if (++buf == 2) { buf = 0; buf_phase ^= 1; }
  }
}

// ---- generic (unaligned) row access used only by the fallback paths -----------------------
__device__ __forceinline__ void store_row_generic(void* out, int out_dtype, int64_t off, int valid,
                                                  const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < valid) st_from_float(out, off + j, out_dtype, v[j]);
}
__device__ __forceinline__ void load_row_generic(const void* src, int dtype, int64_t off, int valid,
                                                 float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = j < valid ? ld_as_float(src, off + j, dtype) : 0.f;
}

// =====================================================================================
// kAct: PCADV_ACT_*;  kOut: PCADV_F32 / PCADV_F16 / PCADV_BF16
template <int kAct, int kOut>
__global__ void __launch_bounds__(kRowsThreads, 1)
tc_rows_kernel(const __grid_constant__ TensorMaps maps, const LinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemLayout L = carve_smem(smem_raw);
  SharedTail* st = L.tail;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_seg; ++s) tma_prefetch_desc(&maps.act[s]);
    tma_prefetch_desc(&maps.w);
    if (p.tma_out) tma_prefetch_desc(&maps.out);
    if (p.tma_mask) tma_prefetch_desc(&maps.mask);
  }
  const uint32_t tmem_base = pipeline_setup(L, warp, lane, 8);
  const int64_t num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    if (lane == 0) producer_loop<false>(maps, p, L, kStages);
  } else if (warp == 1) {
    if (lane == 0) mma_loop(p, L, tmem_base, kStages);
  } else {
    // ================= epilogue: 8 warps, two per TMEM lane quarter =================
    const int quarter = warp & 3;
    const int hsel = (warp - 2) >> 2;                    // which 32-column chunk of each step
    const int lane_row = quarter * 32 + lane;            // row of the 128-row tile
    const int eid = threadIdx.x - 64;                    // 0..255
    const bool is_issuer = (warp == 2 && lane == 0);
    const float oscale = p.out_scale ? *p.out_scale : 1.f;
    float* bias_s = L.bias;
    float* gb_s = L.bias + kMaxTileN;                    // [2][kMaxTileN]
    const int steps = (p.bn + 63) / 64;
    int buf = 0;
    uint32_t buf_phase = 0;
    uint64_t slab_count = 0;
    int staged_col = -1;                                 // what bias_s / gb_s currently hold
    int64_t staged_g0 = -1, staged_g1 = -1;
    // mask prefetch iterator (issuer only): one 64-column slab ahead of the consumers
    int64_t pf_tile = blockIdx.x;
    int pf_step = 0;
    uint64_t pf_count = 0;
    auto issue_mask = [&]() {
      if (pf_tile >= num_tiles) return;
      const int64_t ptm = pf_tile / p.tiles_n, ptn = pf_tile % p.tiles_n;
      const int b = static_cast<int>(pf_count & 1);
      mbar_arrive_expect_tx(&st->mask_full[b], kSlabBytes);
      tma_load_2d(L.epi + (2 + b) * kSlabBytes, &maps.mask, &st->mask_full[b],
                  static_cast<int32_t>(ptn * p.bn + pf_step * 64), static_cast<int32_t>(ptm * kTileM));
      ++pf_count;
      if (++pf_step == steps) { pf_step = 0; pf_tile += gridDim.x; }
    };
    if (p.tma_mask && is_issuer) issue_mask();

    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int64_t tm = t / p.tiles_n, tn = t % p.tiles_n;
      const int64_t r = tm * kTileM + lane_row;
      const bool r_ok = r < p.rows;
      const int col_base = static_cast<int>(tn * p.bn);
      const int64_t rpg = p.rows_per_group > 0 ? p.rows_per_group : p.rows;
      const int64_t row_first = tm * kTileM;
      const int64_t row_last = row_first + kTileM - 1 < p.rows ? row_first + kTileM - 1 : p.rows - 1;
      const int64_t g_first = row_first / rpg, g_last = row_last / rpg;
      const int64_t g = r_ok ? r / rpg : g_first;
      const bool gb_staged = p.group_bias != nullptr && (g_last - g_first) <= 1;

      // ---- stage bias (and per-cloud bias rows) for this tile's columns; skipped when the
      // staged values are still the ones this tile needs (same columns, same clouds)
      const bool restage = col_base != staged_col || (gb_staged && (g_first != staged_g0 || g_last != staged_g1));
      if (restage) {
        staged_col = col_base; staged_g0 = g_first; staged_g1 = g_last;
        named_barrier_sync(1, 256);                      // previous tile's readers are done
        const int c = col_base + eid;
        const bool c_ok = eid < p.bn && c < p.n;
        bias_s[eid] = c_ok ? (p.bias ? __ldg(p.bias + c) : 0.f)
                           : (p.rowmax_key ? -INFINITY : 0.f);
        if (gb_staged) {
          gb_s[eid] = c_ok ? __ldg(p.group_bias + g_first * p.n + c) : 0.f;
          gb_s[kMaxTileN + eid] = (c_ok && g_last != g_first) ? __ldg(p.group_bias + g_last * p.n + c) : 0.f;
        }
        named_barrier_sync(1, 256);
      }

      mbar_wait(&st->tmem_full[buf], buf_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(buf * kMaxTileN);
      unsigned long long rkey = 0ull;
      for (int step = 0; step < steps; ++step) {
        const int c0 = step * 64 + hsel * 32;
        const bool live = c0 < p.bn;
        const int sb = static_cast<int>(slab_count & 1);
        // 16-bit outputs without a TMA-fetched mask rotate through all four slabs, so three
        // stores can be in flight; otherwise two slabs (one pending store group)
        const bool deep = (kOut != PCADV_F32) && !p.tma_mask;
        const int ob = deep ? static_cast<int>(slab_count & 3) : sb;
        if (p.tma_out) {
          // the slab(s) we are about to fill were handed to TMA stores 2 (4) steps ago
          if (is_issuer) {
            if (deep) bulk_wait_group_read<3>();
            else bulk_wait_group_read<1>();
          }
          named_barrier_sync(1, 256);
          if (p.tma_mask) {
            if (is_issuer) issue_mask();
            mbar_wait(&st->mask_full[sb], static_cast<uint32_t>((slab_count >> 1) & 1));
          }
        }
        if (live) {
          float v[32];
          tmem_ld32(taddr0 + c0, v);
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
              const float4* g4 = reinterpret_cast<const float4*>(
                  gb_s + (g != g_first ? kMaxTileN : 0) + c0);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 b = g4[q];
                v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
              }
            } else if (r_ok) {
              float a[32];
              load_row_generic(p.group_bias, PCADV_F32, g * p.n + cg, valid, a);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += a[j];
            }
          }
          if (p.addend && r_ok) {
            float a[32];
            const float* ap = p.addend + r * p.ld_addend + cg;
            if (valid == 32 && (reinterpret_cast<uintptr_t>(ap) & 15) == 0) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 b = *reinterpret_cast<const float4*>(ap + 4 * q);
                a[4 * q] = b.x; a[4 * q + 1] = b.y; a[4 * q + 2] = b.z; a[4 * q + 3] = b.w;
              }
            } else {
              load_row_generic(p.addend, PCADV_F32, r * p.ld_addend + cg, valid, a);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += a[j];
          }
          if (p.rowmax_key) {          // columns >= n carry -inf from bias_s and never win
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const unsigned long long k = pack_key(v[j], static_cast<uint32_t>(cg + j));
              rkey = k > rkey ? k : rkey;
            }
          }
          if (p.out) {
            if (kAct == PCADV_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            } else if (kAct == PCADV_ACT_LEAKY) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
            }
            if (p.mask) {
              float m[32];
              if (p.tma_mask) {
                const uint8_t* mrow = L.epi + (2 + sb) * kSlabBytes + lane_row * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 t4 = *reinterpret_cast<const uint4*>(
                      mrow + (((hsel * 4 + q) ^ (lane_row & 7)) << 4));
                  const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float2 f;
                    if (p.mask_dtype == PCADV_F16)
                      f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
                    else
                      f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                    m[q * 8 + 2 * e] = f.x; m[q * 8 + 2 * e + 1] = f.y;
                  }
                }
              } else if (r_ok) {
                load_row_generic(p.mask, p.mask_dtype, r * p.ld_mask + cg, valid, m);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) m[j] = 0.f;
              }
              const float neg = p.mask_act == PCADV_ACT_LEAKY ? p.mask_slope : 0.f;
              if (p.mask_act != PCADV_ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = m[j] > 0.f ? v[j] : v[j] * neg;
              }
            }
            if (p.out_scale) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= oscale;
            }
            if (p.tma_out) {
              if (kOut == PCADV_F32) {
                uint8_t* orow = L.epi + (2 * sb + hsel) * kSlabBytes + lane_row * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  *reinterpret_cast<float4*>(orow + ((q ^ (lane_row & 7)) << 4)) =
                      make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              } else {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (kOut == PCADV_F16) {
                    const float a = fminf(fmaxf(v[2 * j], -65504.f), 65504.f);
                    const float b = fminf(fmaxf(v[2 * j + 1], -65504.f), 65504.f);
                    __half2 h = __floats2half2_rn(a, b);
                    pk[j] = *reinterpret_cast<uint32_t*>(&h);
                  } else {
                    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                    pk[j] = *reinterpret_cast<uint32_t*>(&h);
                  }
                }
                uint8_t* orow = L.epi + ob * kSlabBytes + lane_row * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *reinterpret_cast<uint4*>(orow + (((hsel * 4 + q) ^ (lane_row & 7)) << 4)) =
                      make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
              }
            } else if (r_ok && valid > 0) {
              store_row_generic(p.out, kOut, r * p.ld_out + cg, valid, v);
            }
          }
        }
        if (p.tma_out) {
          fence_proxy_async();
          named_barrier_sync(1, 256);
          if (is_issuer) {
            const int32_t row0 = static_cast<int32_t>(tm * kTileM);
            if (kOut == PCADV_F32) {
              tma_store_2d(&maps.out, L.epi + (2 * sb) * kSlabBytes, col_base + step * 64, row0);
              if (step * 64 + 32 < p.bn)
                tma_store_2d(&maps.out, L.epi + (2 * sb + 1) * kSlabBytes, col_base + step * 64 + 32, row0);
            } else {
              tma_store_2d(&maps.out, L.epi + ob * kSlabBytes, col_base + step * 64, row0);
            }
            bulk_commit_group();
          }
          ++slab_count;
        }
      }
      if (p.rowmax_key && r_ok && rkey) atomicMax(&p.rowmax_key[r], rkey);
      // release the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->tmem_empty[buf]);
      if (++buf == 2) { buf = 0; buf_phase ^= 1; }
    }
    if (p.tma_out && is_issuer) bulk_wait_group<0>();   // all output slabs have landed
  }
  pipeline_teardown(warp, tmem_base);
}

// =====================================================================================
__global__ void __launch_bounds__(kColmaxThreads, 1)
tc_colmax_kernel(const __grid_constant__ TensorMaps maps, const LinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemLayout L = carve_smem(smem_raw);
  SharedTail* st = L.tail;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_seg; ++s) tma_prefetch_desc(&maps.act[s]);
    tma_prefetch_desc(&maps.w);
  }
  const uint32_t tmem_base = pipeline_setup(L, warp, lane, 4);
  const int64_t num_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0) {
    if (lane == 0) producer_loop<true>(maps, p, L, kMaxStages);
  } else if (warp == 1) {
    if (lane == 0) mma_loop(p, L, tmem_base, kMaxStages);
  } else {
    // thread = output channel; TMEM columns = points.  Running (max, first index) per cloud.
    const int quarter = warp & 3;
    const int lane_row = quarter * 32 + lane;
    int buf = 0;
    uint32_t buf_phase = 0;
    for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int64_t tm, tn;
      tile_coords<true>(p, t, tm, tn);
      mbar_wait(&st->tmem_full[buf], buf_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(buf * kMaxTileN);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->tmem_empty[buf]);
      if (++buf == 2) { buf = 0; buf_phase ^= 1; }
    }
  }
  pipeline_teardown(warp, tmem_base);
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
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * (dtype == PCADV_F32 ? 4 : 2)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt =
      dtype == PCADV_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                         : (dtype == PCADV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PCADV_CHECK_ARG(r == CUDA_SUCCESS,
                  "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
                  (long long)rows, (long long)cols, (long long)ld, box_cols, box_rows);
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int ensure_smem(const void* kernel) {
  static const void* done[32] = {nullptr};
  for (int i = 0; i < 32; ++i) {
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

typedef void (*RowsKernel)(const TensorMaps, const LinearParams);

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
    PCADV_CHECK_ARG(a.seg[i].dtype == dt && a.seg[i].k % kBlockK == 0 &&
                        tma_compatible(a.seg[i].ptr, dt, a.seg[i].ld),
                    "tc_linear: segment %d not TMA-compatible (k=%d ld=%lld)", i, a.seg[i].k,
                    (long long)a.seg[i].ld);
    p.seg_k[i] = a.seg[i].k;
    if (int rc = encode_tmap_2d(&maps.act[i], a.seg[i].ptr, dt, a.rows, a.seg[i].k, a.seg[i].ld,
                                kBlockK, swapped ? p.bn : kTileM))
      return rc;
    ktot += a.seg[i].k;
  }
  PCADV_CHECK_ARG(tma_compatible(a.w, dt, a.ldw), "tc_linear: weight not TMA-compatible");
  if (int rc = encode_tmap_2d(&maps.w, a.w, dt, a.n, ktot, a.ldw, kBlockK, swapped ? kTileM : p.bn))
    return rc;
  if (swapped) {
    p.tiles_m = (a.n + kTileM - 1) / kTileM;
    p.tiles_n = (a.rows + p.bn - 1) / p.bn;
  } else {
    p.tiles_m = (a.rows + kTileM - 1) / kTileM;
    p.tiles_n = (a.n + p.bn - 1) / p.bn;
  }
  p.idesc = make_idesc(kTileM, p.bn, dt == PCADV_BF16, false, false);
  p.bias = a.bias; p.group_bias = a.group_bias; p.rows_per_group = a.rows_per_group;
  p.addend = a.addend; p.ld_addend = a.ld_addend; p.slope = a.slope;
  p.mask = a.mask; p.ld_mask = a.ld_mask; p.mask_dtype = a.mask_dtype; p.mask_act = a.mask_act;
  p.mask_slope = a.mask_slope; p.out_scale = a.out_scale; p.out = a.out; p.ld_out = a.ld_out;
  p.colmax_key = a.colmax_key; p.rowmax_key = a.rowmax_key;
  p.tma_out = (!swapped && a.out && tma_compatible(a.out, a.out_dtype, a.ld_out)) ? 1 : 0;
  if (p.tma_out) {
    if (int rc = encode_tmap_2d(&maps.out, a.out, a.out_dtype, a.rows, a.n, a.ld_out,
                                a.out_dtype == PCADV_F32 ? 32 : 64, kTileM))
      return rc;
  }
  p.tma_mask = (p.tma_out && a.mask && a.out_dtype != PCADV_F32 && a.mask_dtype != PCADV_F32 &&
                tma_compatible(a.mask, a.mask_dtype, a.ld_mask)) ? 1 : 0;
  if (p.tma_mask) {
    if (int rc = encode_tmap_2d(&maps.mask, a.mask, a.mask_dtype, a.rows, a.n, a.ld_mask, 64, kTileM))
      return rc;
  }
  const int64_t tiles = p.tiles_m * p.tiles_n;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  if (swapped) {
    if (int rc = ensure_smem(reinterpret_cast<const void*>(&tc_colmax_kernel))) return rc;
    tc_colmax_kernel<<<grid, kColmaxThreads, kSmemBytes, s>>>(maps, p);
  } else {
    RowsKernel k = pick_rows_kernel(a.act, a.out ? a.out_dtype : PCADV_F16);
    if (int rc = ensure_smem(reinterpret_cast<const void*>(k))) return rc;
    k<<<grid, kRowsThreads, kSmemBytes, s>>>(maps, p);
  }
  PCADV_LAUNCHED();
  return 0;
}

}  // namespace pcadv
