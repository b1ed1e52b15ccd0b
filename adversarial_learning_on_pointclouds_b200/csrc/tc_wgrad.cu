// Tensor-core engine, weight gradients:  dW[c, koff + k] += scale * sum_r dz[r, c] x[r, k].
//
// The reduction runs over points, so both operands are MN-major in shared memory: the
// M side is dz^T (channels c on M, rows on K), the N side is x^T (k on N, rows on K); both
// are fetched as [64 rows][64 channels] TMA boxes (128 B inner, 128B swizzle) straight from
// the point-major activations -- no transposed copies.  Work = (output tile, row range);
// each CTA accumulates its range in TMEM and adds the tile into dW with fp32 RED atomics.
#include "tc_pipeline.cuh"

namespace pcadv {
namespace tc {

constexpr int kWgradThreads = 192;

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
  int bf16;
  float* dw;
  int64_t ld_dw;
  float* dbias;           // [n] or NULL: column sums of dz, formed from the smem tiles in flight
  const float* scale;
};

__global__ void __launch_bounds__(kWgradThreads, 1)
tc_wgrad_kernel(const __grid_constant__ TensorMaps maps, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemLayout L = carve_smem(smem_raw);
  SharedTail* st = L.tail;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t num_work = static_cast<int64_t>(p.tiles_m) * p.tiles_n * p.splits;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.act[p.seg_index]);
    tma_prefetch_desc(&maps.w);
  }
  // with a bias gradient the four epilogue warps also read every dz tile, so a stage is
  // released by 1 (MMA commit) + 4 (epilogue warps) arrivals
  const uint32_t tmem_base = pipeline_setup(L, warp, lane, 4, p.dbias ? 5 : 1);
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
          mbar_wait_backoff(&st->empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&st->full[stage], stage_tx);
          uint8_t* sa = L.stages + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          // M side: two [64 rows][64 ch] boxes; N side: bn/64 boxes
          tma_load_2d(sa, &maps.w, &st->full[stage], tm * kTileM, static_cast<int32_t>(r));
          tma_load_2d(sa + 8192, &maps.w, &st->full[stage], tm * kTileM + 64, static_cast<int32_t>(r));
          for (int b = 0; b < n_boxes; ++b)
            tma_load_2d(sb + b * 8192, &maps.act[p.seg_index], &st->full[stage],
                        tn * p.bn + b * 64, static_cast<int32_t>(r));
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
        int tm, tn; int64_t r0, r1;
        decode(w, tm, tn, r0, r1);
        mbar_wait_backoff(&st->tmem_empty[buf], buf_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * kMaxTileN);
        bool first = true;
        for (int64_t r = r0; r < r1; r += kBlockK) {
          mbar_wait_backoff(&st->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(L.stages + stage * kStageBytes);
          mma_chunk_mnmajor(d_tmem, a_addr, a_addr + kABytes, p.idesc, first);
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
      int tm, tn; int64_t r0, r1;
      decode(w, tm, tn, r0, r1);
      float bsum = 0.f;
      if (p.dbias) {
        // bias gradient: thread = dz channel; sum its 64 rows of every stage straight from
        // the swizzled smem tile ([64 rows][64 ch] boxes, 16-byte chunk index ^ (row & 7))
        const uint32_t box_off = (lane_row >> 6) * 8192u;
        const uint32_t cc = lane_row & 63;
        for (int64_t r = r0; r < r1; r += kBlockK) {
          mbar_wait(&st->full[stage_e], phase_e);
          if (tn == 0) {
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
      if (p.dbias && tn == 0 && c < p.n && r0 < r1) atomicAdd(p.dbias + c, bsum * sc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&st->tmem_empty[buf]);
      if (++buf == 2) { buf = 0; buf_phase ^= 1; }
    }
  }
  pipeline_teardown(warp, tmem_base);
}

}  // namespace tc

int launch_colsum(const void* dz, int dz_dtype, int64_t ld, int64_t rows, int n, const float* scale,
                  float* out, cudaStream_t s);
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
  int koff = 0;
  for (int i = 0; i < a.num_seg; ++i) {
    const pcadv_seg& sg = a.seg[i];
    PCADV_CHECK_ARG(sg.dtype == dt && sg.k % 64 == 0 && tma_compatible(sg.ptr, dt, sg.ld),
                    "tc_wgrad: segment %d not TMA-compatible", i);
    if (int rc = encode_tmap_2d(&maps.act[i], sg.ptr, dt, a.rows, sg.k, sg.ld, 64, kBlockK)) return rc;
    WgradParams p{};
    p.rows = a.rows; p.n = a.n; p.seg_index = i; p.seg_k = sg.k; p.koff = koff;
    p.bn = sg.k < kMaxTileN ? sg.k : kMaxTileN;
    p.tiles_m = (a.n + kTileM - 1) / kTileM;
    p.tiles_n = (sg.k + p.bn - 1) / p.bn;
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
    p.idesc = make_idesc(kTileM, p.bn, dt == PCADV_BF16, true, true);
    p.dw = a.dw; p.ld_dw = a.ld_dw; p.scale = a.scale;
    p.bf16 = dt == PCADV_BF16 ? 1 : 0;
    p.dbias = i == 0 ? a.dbias : nullptr;
    const int64_t work = static_cast<int64_t>(tiles) * p.splits;
    const int grid = static_cast<int>(work < num_sms() ? work : num_sms());
    tc_wgrad_kernel<<<grid, kWgradThreads, kSmemBytes, s>>>(maps, p);
    PCADV_LAUNCHED();
    koff += sg.k;
  }
  if (a.dgroup_bias) {
    if (int rc = launch_group_colsum(a.dz, a.dz_dtype, a.ld_dz, a.rows, a.n, a.rows_per_group,
                                     a.dgroup_bias, s))
      return rc;
  }
  return 0;
}

}  // namespace pcadv
