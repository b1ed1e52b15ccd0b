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

}  // namespace tc

int tc_rows(const pcadv_linear_args& a, cudaStream_t s);   // tc_rows.cu

int tc_linear(const pcadv_linear_args& a, cudaStream_t s) {
  using namespace tc;
  const int dt = a.seg[0].dtype;
  PCADV_CHECK_ARG(dt == PCADV_F16 || dt == PCADV_BF16, "tc_linear: operands must be fp16 / bf16");
  PCADV_CHECK_ARG(a.w_dtype == dt, "tc_linear: weight dtype differs from activations");
  if (a.colmax_key == nullptr) return tc_rows(a, s);
  PCADV_CHECK_ARG(!a.out && !a.rowmax_key && !a.group_bias && !a.addend && !a.mask,
                  "tc_linear: the max-over-points kernel takes bias only");
  TensorMaps maps;
  LinearParams p{};
  p.rows = a.rows; p.n = a.n; p.num_seg = a.num_seg;
  p.bn = kMaxTileN;
  int ktot = 0;
  for (int i = 0; i < a.num_seg; ++i) {
    PCADV_CHECK_ARG(a.seg[i].dtype == dt && a.seg[i].k % kBlockK == 0 &&
                        tma_compatible(a.seg[i].ptr, dt, a.seg[i].ld),
                    "tc_linear: segment %d not TMA-compatible (k=%d ld=%lld)", i, a.seg[i].k,
                    (long long)a.seg[i].ld);
    p.seg_k[i] = a.seg[i].k;
    if (int rc = encode_tmap_2d(&maps.act[i], a.seg[i].ptr, dt, a.rows, a.seg[i].k, a.seg[i].ld,
                                kBlockK, p.bn))
      return rc;
    ktot += a.seg[i].k;
  }
  PCADV_CHECK_ARG(tma_compatible(a.w, dt, a.ldw), "tc_linear: weight not TMA-compatible");
  if (int rc = encode_tmap_2d(&maps.w, a.w, dt, a.n, ktot, a.ldw, kBlockK, kTileM)) return rc;
  p.tiles_m = (a.n + kTileM - 1) / kTileM;
  p.tiles_n = (a.rows + p.bn - 1) / p.bn;
  p.idesc = make_idesc(kTileM, p.bn, dt == PCADV_BF16, false, false);
  p.bias = a.bias; p.rows_per_group = a.rows_per_group;
  p.colmax_key = a.colmax_key;
  const int64_t tiles = p.tiles_m * p.tiles_n;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  if (int rc = ensure_smem(reinterpret_cast<const void*>(&tc_colmax_kernel))) return rc;
  tc_colmax_kernel<<<grid, kColmaxThreads, kSmemBytes, s>>>(maps, p);
  PCADV_LAUNCHED();
  return 0;
}

}  // namespace pcadv
