// Row-wise softmax heads over the segmentation logits (HBM-bound, one pass each):
//
//   pcadv_softmax_head     logits -> softmax / log_softmax (16-bit, padded, point-major: exactly
//                          what the discriminator's first layer multiplies), and for the labelled
//                          batch the cross-entropy sum and its gradient (softmax - onehot).
//   pcadv_logsoftmax_bwd   dz = dy - exp(lp) * sum_c dy        (backward of log_softmax)
//
// They stand in for F.softmax / F.log_softmax / CrossEntropyLoss and their autograd backward on
// the B x 50 x N logits (utils/trainer.py:899-901, :914, :927-929), which the reference runs as
// separate ATen passes over a strided view.
#include "common.cuh"

namespace pcadv {
namespace {

constexpr int kHeadMaxN = 128;     // classes per row the kernels accept

// A warp owns 32 consecutive rows: the [32, n] fp32 logits tile is fetched coalesced into
// shared memory (odd row stride: conflict-free row-per-thread reads); then thread = row.
__global__ void __launch_bounds__(256) softmax_head_kernel(const pcadv_head_args a) {
  extern __shared__ float sm[];
  __shared__ float red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n;
  const int ldt = n | 1;                                   // odd stride
  float* tile = sm + static_cast<size_t>(warp) * 32 * ldt;
  const int64_t nblk = (a.rows + 31) / 32;
  float loss = 0.f;
  for (int64_t blk = static_cast<int64_t>(blockIdx.x) * 8 + warp; blk < nblk;
       blk += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t row0 = blk * 32;
    const int nrows = a.rows - row0 < 32 ? static_cast<int>(a.rows - row0) : 32;
    __syncwarp();
    if (a.ld == n) {
      const float* src = a.logits + row0 * n;
      const int total = nrows * n;
      for (int i = lane; i < total; i += 32) {
        const int r = i / n, c = i - r * n;
        tile[r * ldt + c] = __ldg(src + i);
      }
    } else {
      for (int r = 0; r < nrows; ++r)
        for (int c = lane; c < n; c += 32) tile[r * ldt + c] = __ldg(a.logits + (row0 + r) * a.ld + c);
    }
    __syncwarp();
    if (lane < nrows) {
      float* t = tile + lane * ldt;
      const int64_t row = row0 + lane;
      float m = t[0];
      for (int c = 1; c < n; ++c) m = fmaxf(m, t[c]);
      float s = 0.f;
      for (int c = 0; c < n; ++c) s += expf(t[c] - m);
      const float lse = m + logf(s);
      int label = -1;
      if (a.labels) {
        label = static_cast<int>(a.labels[row]);
        if (label >= 0 && label < n) loss += lse - t[label];
      }
      // outputs, 8 columns at a time (vector stores when the row address allows)
      const int cols_p = a.probs ? a.probs_cols : 0;
      const int cols_d = a.dz ? a.dz_cols : 0;
      const int cmax = cols_p > cols_d ? cols_p : cols_d;
      for (int c0 = 0; c0 < cmax; c0 += 8) {
        float p[8], d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + j;
          const float lp = c < n ? t[c] - lse : 0.f;
          const float pr = c < n ? expf(lp) : 0.f;
          p[j] = a.mode == PCADV_HEAD_LSM ? lp : pr;
          d[j] = c < n ? a.dz_gain * (pr - (c == label ? 1.f : 0.f)) : 0.f;
        }
        if (c0 < cols_p) {
          if (c0 + 8 <= cols_p) store8(a.probs, row * a.ld_probs + c0, a.probs_dtype, p);
          else for (int j = 0; c0 + j < cols_p; ++j) st_from_float(a.probs, row * a.ld_probs + c0 + j, a.probs_dtype, p[j]);
        }
        if (c0 < cols_d) {
          if (c0 + 8 <= cols_d) store8(a.dz, row * a.ld_dz + c0, a.dz_dtype, d);
          else for (int j = 0; c0 + j < cols_d; ++j) st_from_float(a.dz, row * a.ld_dz + c0 + j, a.dz_dtype, d[j]);
        }
      }
    }
  }
  if (a.loss_sum) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) red[warp] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w];
      atomicAdd(a.loss_sum, s);
    }
  }
}

// thread = row; lp and dy rows are read 8 columns at a time
__global__ void __launch_bounds__(256) logsoftmax_bwd_kernel(
    const void* __restrict__ lp, int lp_dtype, int64_t ld_lp, const void* __restrict__ dy, int dy_dtype,
    int64_t ld_dy, int64_t rows, int n, const float* scale, void* dz, int dz_dtype, int64_t ld_dz,
    int dz_cols) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float sc = scale ? *scale : 1.f;
  float sum = 0.f;
  for (int c0 = 0; c0 < n; c0 += 8) {
    float g[8];
    if (c0 + 8 <= n) load8(dy, r * ld_dy + c0, dy_dtype, g);
    else for (int j = 0; j < 8; ++j) g[j] = c0 + j < n ? ld_as_float(dy, r * ld_dy + c0 + j, dy_dtype) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += g[j];
  }
  for (int c0 = 0; c0 < dz_cols; c0 += 8) {
    float g[8], l[8], o[8];
    if (c0 + 8 <= n) {
      load8(dy, r * ld_dy + c0, dy_dtype, g);
      load8(lp, r * ld_lp + c0, lp_dtype, l);
    } else {
      for (int j = 0; j < 8; ++j) {
        const bool ok = c0 + j < n;
        g[j] = ok ? ld_as_float(dy, r * ld_dy + c0 + j, dy_dtype) : 0.f;
        l[j] = ok ? ld_as_float(lp, r * ld_lp + c0 + j, lp_dtype) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = c0 + j < n ? (g[j] - expf(l[j]) * sum) * sc : 0.f;
    if (c0 + 8 <= dz_cols) store8(dz, r * ld_dz + c0, dz_dtype, o);
    else for (int j = 0; c0 + j < dz_cols; ++j) st_from_float(dz, r * ld_dz + c0 + j, dz_dtype, o[j]);
  }
}

}  // namespace
}  // namespace pcadv

using namespace pcadv;

extern "C" int pcadv_softmax_head(const pcadv_head_args* a, void* stream) {
  PCADV_CHECK_ARG(a && a->logits && a->rows >= 0 && a->n > 0 && a->n <= kHeadMaxN,
                  "pcadv_softmax_head: bad args (1 <= n <= %d)", kHeadMaxN);
  PCADV_CHECK_ARG(a->mode == PCADV_HEAD_CE || a->mode == PCADV_HEAD_LSM, "pcadv_softmax_head: bad mode");
  PCADV_CHECK_ARG(!a->probs || a->probs_cols >= a->n, "pcadv_softmax_head: probs_cols < n");
  PCADV_CHECK_ARG(!a->dz || (a->dz_cols >= a->n && a->labels), "pcadv_softmax_head: dz needs labels, dz_cols >= n");
  if (a->rows == 0) return 0;
  const size_t smem = static_cast<size_t>(8) * 32 * (a->n | 1) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    PCADV_CUDA_OK(cudaFuncSetAttribute(softmax_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       8 * 32 * (kHeadMaxN | 1) * 4));
    attr_done = true;
  }
  const int64_t nblk = (a->rows + 31) / 32;
  int64_t grid = (nblk + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  softmax_head_kernel<<<static_cast<unsigned>(grid), 256, smem, static_cast<cudaStream_t>(stream)>>>(*a);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_logsoftmax_bwd(const void* lp, int32_t lp_dtype, int64_t ld_lp, const void* dy,
                                    int32_t dy_dtype, int64_t ld_dy, int64_t rows, int32_t n,
                                    const float* scale, void* dz, int32_t dz_dtype, int64_t ld_dz,
                                    int32_t dz_cols, void* stream) {
  PCADV_CHECK_ARG(lp && dy && dz && rows >= 0 && n > 0 && dz_cols >= n, "pcadv_logsoftmax_bwd: bad args");
  if (rows == 0) return 0;
  logsoftmax_bwd_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(lp, lp_dtype, ld_lp, dy, dy_dtype, ld_dy,
                                                               rows, n, scale, dz, dz_dtype, ld_dz, dz_cols);
  PCADV_LAUNCHED();
  return 0;
}
