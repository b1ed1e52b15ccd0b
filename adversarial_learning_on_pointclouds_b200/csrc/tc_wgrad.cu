// Tensor-core engine, weight gradients:  dW[c, koff + k] += scale * sum_r dz[r, c] x[r, k].
//
// The reduction runs over points, so both operands are MN-major in shared memory: the
// M side is dz^T (channels c on M, rows on K), the N side is x^T (k on N, rows on K); both
// are fetched as [64 rows][64 channels] TMA boxes (128 B inner, 128B swizzle) straight from
// the point-major activations -- no transposed copies.  Work = (output tile, row range);
// every K-segment of a concatenated input (x1..x5 of fc1) is a set of N-side tiles of the
// SAME launch, so the CTAs that share a row range run together and dz is read from HBM once.
// Each CTA accumulates its range in TMEM and adds the tile into dW with fp32 RED atomics;
// the bias gradient is summed from the dz tiles in flight by the (otherwise idle) epilogue
// warps.
#include "tc_pipeline.cuh"

namespace pcadv {
namespace tc {

constexpr int kWgradThreads = 192;

struct WgradParams {
  int64_t rows;
  int n;                          // dz channels
  int num_seg;
  int seg_k[PCADV_MAX_SEG];       // channels of each x segment (multiples of 64)
  int seg_koff[PCADV_MAX_SEG];    // column offset of the segment inside dw
  int seg_bn[PCADV_MAX_SEG];      // N tile of the segment (multiple of 64, <= 256)
  int seg_tile0[PCADV_MAX_SEG + 1];  // first N-tile index of each segment (prefix sum)
  int tiles_m, tiles_n;           // tiles_n = seg_tile0[num_seg]
  int splits;
  int64_t rows_per_split;         // multiple of 64
  int bf16;
  float* dw;
  int64_t ld_dw;
  float* dbias;                   // [n] or NULL
  const float* scale;
};

struct WorkItem {
  int tm, seg, tn_local, bn;
  int64_t r0, r1;
  bool first_n_tile;
};

__device__ __forceinline__ WorkItem decode_work(const WgradParams& p, int64_t w) {
  WorkItem it;
  const int sp = static_cast<int>(w % p.splits);
  const int64_t tile = w / p.splits;
  const int tn = static_cast<int>(tile % p.tiles_n);
  it.tm = static_cast<int>(tile / p.tiles_n);
  int s = 0;
  while (s + 1 < p.num_seg && tn >= p.seg_tile0[s + 1]) ++s;
  it.seg = s;
  it.tn_local = tn - p.seg_tile0[s];
  it.bn = p.seg_bn[s];
  it.first_n_tile = tn == 0;
  it.r0 = sp * p.rows_per_split;
  it.r1 = it.r0 + p.rows_per_split < p.rows ? it.r0 + p.rows_per_split : p.rows;
  return it;
}

__global__ void __launch_bounds__(kWgradThreads, 1)
tc_wgrad_kernel(const __grid_constant__ TensorMaps maps, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemLayout L = carve_smem(smem_raw);
  SharedTail* st = L.tail;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_work = static_cast<int64_t>(p.tiles_m) * p.tiles_n * p.splits;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_seg; ++s) tma_prefetch_desc(&maps.act[s]);
    tma_prefetch_desc(&maps.w);
  }
  // with a bias gradient the four epilogue warps also read every dz tile, so a stage is
  // released by 1 (MMA commit) + 4 (epilogue warps) arrivals
  const uint32_t tmem_base = pipeline_setup(L, warp, lane, 4, p.dbias ? 5 : 1);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
        const WorkItem it = decode_work(p, w);
        const int n_boxes = it.bn / 64;
        const uint32_t stage_tx = static_cast<uint32_t>((kTileM + it.bn) * kBlockK * 2);
        for (int64_t r = it.r0; r < it.r1; r += kBlockK) {
          mbar_wait_backoff(&st->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&st->full[stage], stage_tx);
          uint8_t* sa = L.stages + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          // M side: two [64 rows][64 ch] boxes; N side: bn/64 boxes
          tma_load_2d(sa, &maps.w, &st->full[stage], it.tm * kTileM, static_cast<int32_t>(r));
          tma_load_2d(sa + 8192, &maps.w, &st->full[stage], it.tm * kTileM + 64, static_cast<int32_t>(r));
          for (int b = 0; b < n_boxes; ++b)
            tma_load_2d(sb + b * 8192, &maps.act[it.seg], &st->full[stage],
                        it.tn_local * it.bn + b * 64, static_cast<int32_t>(r));
          if (++stage == kMaxStages) { stage = 0; phase ^= 1; }
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
        const WorkItem it = decode_work(p, w);
        const uint32_t idesc = make_idesc(kTileM, it.bn, p.bf16 != 0, true, true);
        mbar_wait_backoff(&st->tmem_empty[buf], buf_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kMaxTileN);
        bool first = true;
        for (int64_t r = it.r0; r < it.r1; r += kBlockK) {
          mbar_wait_backoff(&st->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(L.stages + stage * kStageBytes);
          mma_chunk_mnmajor(d_tmem, a_addr, a_addr + kABytes, idesc, first);
          first = false;
          umma_commit(&st->empty[stage]);
          if (++stage == kMaxStages) { stage = 0; phase ^= 1; }
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
    int stage_e = 0;
    uint32_t phase_e = 0;
    for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
      const WorkItem it = decode_work(p, w);
      float bsum = 0.f;
      if (p.dbias) {
        // bias gradient: thread = dz channel; sum its 64 rows of every stage straight from
        // the swizzled smem tile ([64 rows][64 ch] boxes, 16-byte chunk index ^ (row & 7))
        const uint32_t box_off = (lane_row >> 6) * 8192u;
        const uint32_t cc = lane_row & 63;
        for (int64_t r = it.r0; r < it.r1; r += kBlockK) {
          mbar_wait(&st->full[stage_e], phase_e);
          if (it.first_n_tile) {
            const uint8_t* tile = L.stages + stage_e * kStageBytes + box_off;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < kBlockK; rr += 2) {
              const uint16_t a = *reinterpret_cast<const uint16_t*>(
                  tile + rr * 128 + ((((cc >> 3) ^ (rr & 7)) << 4) | ((cc & 7) << 1)));
              const uint16_t b = *reinterpret_cast<const uint16_t*>(
                  tile + (rr + 1) * 128 + ((((cc >> 3) ^ ((rr + 1) & 7)) << 4) | ((cc & 7) << 1)));
              if (p.bf16) {
                s0 += __bfloat162float(__ushort_as_bfloat16(a));
                s1 += __bfloat162float(__ushort_as_bfloat16(b));
              } else {
                s0 += __half2float(__ushort_as_half(a));
                s1 += __half2float(__ushort_as_half(b));
              }
            }
            bsum += s0 + s1;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&st->empty[stage_e]);
          if (++stage_e == kMaxStages) { stage_e = 0; phase_e ^= 1; }
        }
      }
      mbar_wait(&st->tmem_full[buf], buf_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(buf * kMaxTileN);
      const int c = it.tm * kTileM + lane_row;
      const int seg_k = p.seg_k[it.seg];
      for (int c0 = 0; c0 < it.bn; c0 += 32) {
        float v[32];
        tmem_ld32(taddr0 + c0, v);
        const int k0 = it.tn_local * it.bn + c0;
        if (c >= p.n || it.r0 >= it.r1) continue;
        float* dst = p.dw + static_cast<int64_t>(c) * p.ld_dw + p.seg_koff[it.seg] + k0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (k0 + j < seg_k) atomicAdd(dst + j, v[j] * sc);
      }
      if (p.dbias && it.first_n_tile && c < p.n && it.r0 < it.r1) atomicAdd(p.dbias + c, bsum * sc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->tmem_empty[buf]);
      if (++buf == 2) { buf = 0; buf_phase ^= 1; }
    }
  }
  pipeline_teardown(warp, tmem_base);
}

}  // namespace tc

int launch_group_colsum(const void* dz, int dz_dtype, int64_t ld, int64_t rows, int n,
                        int64_t rows_per_group, float* out, cudaStream_t s);

int tc_wgrad(const pcadv_wgrad_args& a, cudaStream_t s) {
  using namespace tc;
  const int dt = a.dz_dtype;
  PCADV_CHECK_ARG(dt == PCADV_F16 || dt == PCADV_BF16, "tc_wgrad: dz must be fp16 / bf16");
  PCADV_CHECK_ARG(a.dw != nullptr && a.num_seg >= 1, "tc_wgrad: dw and segments required");
  PCADV_CHECK_ARG(a.n % 64 == 0 && tma_compatible(a.dz, dt, a.ld_dz),
                  "tc_wgrad: dz not TMA-compatible (n=%d)", a.n);
  TensorMaps maps;
  if (int rc = encode_tmap_2d(&maps.w, a.dz, dt, a.rows, a.n, a.ld_dz, 64, kBlockK)) return rc;
  if (int rc = ensure_smem(reinterpret_cast<const void*>(&tc_wgrad_kernel))) return rc;
  WgradParams p{};
  p.rows = a.rows; p.n = a.n; p.num_seg = a.num_seg;
  int koff = 0, tiles_n = 0;
  for (int i = 0; i < a.num_seg; ++i) {
    const pcadv_seg& sg = a.seg[i];
    PCADV_CHECK_ARG(sg.dtype == dt && sg.k % 64 == 0 && tma_compatible(sg.ptr, dt, sg.ld),
                    "tc_wgrad: segment %d not TMA-compatible", i);
    if (int rc = encode_tmap_2d(&maps.act[i], sg.ptr, dt, a.rows, sg.k, sg.ld, 64, kBlockK)) return rc;
    p.seg_k[i] = sg.k;
    p.seg_koff[i] = koff;
    p.seg_bn[i] = sg.k < kMaxTileN ? sg.k : kMaxTileN;
    p.seg_tile0[i] = tiles_n;
    tiles_n += (sg.k + p.seg_bn[i] - 1) / p.seg_bn[i];
    koff += sg.k;
  }
  p.seg_tile0[a.num_seg] = tiles_n;
  p.tiles_m = (a.n + kTileM - 1) / kTileM;
  p.tiles_n = tiles_n;
  const int tiles = p.tiles_m * p.tiles_n;
  // one wave: tiles * splits <= #SMs, so no CTA gets a second work item (a 2x tail)
  int64_t splits = num_sms() / tiles;
  const int64_t max_splits = (a.rows + 511) / 512;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rps = (a.rows + splits - 1) / splits;
  rps = (rps + kBlockK - 1) / kBlockK * kBlockK;
  p.splits = static_cast<int>((a.rows + rps - 1) / rps);
  p.rows_per_split = rps;
  p.bf16 = dt == PCADV_BF16 ? 1 : 0;
  p.dw = a.dw; p.ld_dw = a.ld_dw; p.scale = a.scale; p.dbias = a.dbias;
  const int64_t work = static_cast<int64_t>(tiles) * p.splits;
  const int grid = static_cast<int>(work < num_sms() ? work : num_sms());
  tc_wgrad_kernel<<<grid, kWgradThreads, kSmemBytes, s>>>(maps, p);
  PCADV_LAUNCHED();
  if (a.dgroup_bias) {
    if (int rc = launch_group_colsum(a.dz, a.dz_dtype, a.ld_dz, a.rows, a.n, a.rows_per_group,
                                     a.dgroup_bias, s))
      return rc;
  }
  return 0;
}

}  // namespace pcadv
