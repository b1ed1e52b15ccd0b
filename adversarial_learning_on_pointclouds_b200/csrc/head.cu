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

// 16-bit point-major [rows, 64] outputs leave through a per-warp [32 rows][128 B] staging tile and
// fully coalesced 16-byte stores (a warp's 32 rows are 4 KB contiguous in HBM).
constexpr int kPitchW = 33;        // staging row pitch in 32-bit words: conflict-free rows and columns
__device__ __forceinline__ void flush_tile16(const uint32_t* stage, void* dst, int64_t row0, int nrows, int lane) {
  uint32_t* g = reinterpret_cast<uint32_t*>(dst) + row0 * 32;
  for (int r = 0; r < nrows; ++r) g[r * 32 + lane] = stage[r * kPitchW + lane];
}

// A warp owns 32 consecutive rows: the [32, n] fp32 logits tile is fetched coalesced into
// shared memory (odd row stride: conflict-free row-per-thread reads); then thread = row.
__global__ void __launch_bounds__(256) softmax_head_kernel(const pcadv_head_args a) {
  extern __shared__ float sm[];
  __shared__ float red[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n;
  const int ldt = n | 1;                                   // odd stride
  float* tile = sm + static_cast<size_t>(warp) * (32 * ldt + 32 * kPitchW);
  uint32_t* stage = reinterpret_cast<uint32_t*>(tile + 32 * ldt);     // [32 rows][33 words]
  // packed fast path: 16-bit outputs of exactly 64 columns with contiguous rows
  const bool packed_p = a.probs && a.probs_dtype != PCADV_F32 && a.probs_cols == 64 && a.ld_probs == 64 &&
                        (reinterpret_cast<uintptr_t>(a.probs) & 15) == 0 && n <= 64;
  const bool packed_d = a.dz && a.dz_dtype != PCADV_F32 && a.dz_cols == 64 && a.ld_dz == 64 &&
                        (reinterpret_cast<uintptr_t>(a.dz) & 15) == 0 && n <= 64;
  const int64_t nblk = (a.rows + 31) / 32;
  float loss = 0.f, valid = 0.f;
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
    float lse = 0.f;
    int label = -1;
    float* t = tile + lane * ldt;
    const int64_t row = row0 + lane;
    if (lane < nrows) {
      float m = t[0];
      for (int c = 1; c < n; ++c) m = fmaxf(m, t[c]);
      float s = 0.f;
      for (int c = 0; c < n; ++c) s += expf(t[c] - m);
      lse = m + logf(s);
      if (a.labels) {
        const int64_t l64 = a.labels[row];
        label = (l64 >= 0 && l64 < n) ? static_cast<int>(l64) : -1;     // outside [0, n): ignored row
        if (label >= 0) { loss += lse - t[label]; valid += 1.f; }
      }
      // in place: t[c] <- log_softmax
      for (int c = 0; c < n; ++c) t[c] -= lse;
    }
    // ---- probs (softmax, or log_softmax in LSM mode)
    if (a.probs) {
      if (packed_p) {
        if (lane < nrows) {
          uint32_t* srow = stage + lane * kPitchW;
#pragma unroll 4
          for (int c = 0; c < 64; c += 2) {
            float v0 = c < n ? t[c] : 0.f, v1 = c + 1 < n ? t[c + 1] : 0.f;
            if (a.mode == PCADV_HEAD_CE) { v0 = c < n ? expf(v0) : 0.f; v1 = c + 1 < n ? expf(v1) : 0.f; }
            srow[c >> 1] = a.probs_dtype == PCADV_F16 ? pack_f16x2_sat(v0, v1) : pack_bf16x2(v0, v1);
          }
        }
        __syncwarp();
        flush_tile16(stage, a.probs, row0, nrows, lane);
        __syncwarp();
      } else if (lane < nrows) {
        for (int c = 0; c < a.probs_cols; ++c) {
          float v = c < n ? t[c] : 0.f;
          if (a.mode == PCADV_HEAD_CE && c < n) v = expf(v);
          st_from_float(a.probs, row * a.ld_probs + c, a.probs_dtype, v);
        }
      }
    }
    // ---- dz = gain * (softmax - onehot)
    if (a.dz) {
      if (packed_d) {
        if (lane < nrows) {
          uint32_t* srow = stage + lane * kPitchW;
#pragma unroll 4
          for (int c = 0; c < 64; c += 2) {
            const float v0 = (c < n && label >= 0) ? a.dz_gain * (expf(t[c]) - (c == label ? 1.f : 0.f)) : 0.f;
            const float v1 = (c + 1 < n && label >= 0) ? a.dz_gain * (expf(t[c + 1]) - (c + 1 == label ? 1.f : 0.f)) : 0.f;
            srow[c >> 1] = a.dz_dtype == PCADV_F16 ? pack_f16x2_sat(v0, v1) : pack_bf16x2(v0, v1);
          }
        }
        __syncwarp();
        flush_tile16(stage, a.dz, row0, nrows, lane);
        __syncwarp();
      } else if (lane < nrows) {
        for (int c = 0; c < a.dz_cols; ++c) {
          const float v = (c < n && label >= 0) ? a.dz_gain * (expf(t[c]) - (c == label ? 1.f : 0.f)) : 0.f;
          st_from_float(a.dz, row * a.ld_dz + c, a.dz_dtype, v);
        }
      }
    }
  }
  if (a.loss_sum || a.valid_count) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      loss += __shfl_xor_sync(0xffffffffu, loss, o);
      valid += __shfl_xor_sync(0xffffffffu, valid, o);
    }
    if (lane == 0) { red[warp] = loss; red[8 + warp] = valid; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f, v = 0.f;
      for (int w = 0; w < 8; ++w) { s += red[w]; v += red[8 + w]; }
      if (a.loss_sum) atomicAdd(a.loss_sum, s);
      if (a.valid_count) atomicAdd(a.valid_count, v);
    }
  }
}

// Packed path (n <= 64 classes, contiguous fp32 logits rows with an even n, 16-bit [rows, 64]
// outputs): EIGHT lanes per row, four rows per warp and trip, two trips in flight.  Lane `sub` of a
// row owns the column pairs sub + 8 j (j < 4): its 8-byte loads and the 4-byte words it stores are
// 32-byte (full sector) segments per row, and the row's max / sum need three shuffle steps for four
// rows at once instead of five per row (the 32-lanes-per-row version of round 1 was bound by its
// instruction count: 74 % issue-slot utilisation at 42 % of the HBM rate).
template <bool kBf16>
__global__ void __launch_bounds__(256) softmax_head_rows_kernel(const pcadv_head_args a) {
  __shared__ float red[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane & 7, rsel = lane >> 3;
  const int n = a.n, pairs = n >> 1;
  const int64_t gwarp = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * 8;
  uint32_t* probs = reinterpret_cast<uint32_t*>(a.probs);
  uint32_t* dz = reinterpret_cast<uint32_t*>(a.dz);
  const bool lsm = a.mode == PCADV_HEAD_LSM;
  constexpr float kLog2e = 1.4426950408889634f;
  float loss = 0.f, valid = 0.f;
  for (int64_t r0 = gwarp * 8; r0 < a.rows; r0 += nwarps * 8) {
    float2 t[2][4];
    int label[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + 4 * u + rsel;
      label[u] = -1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = sub + 8 * j;
        t[u][j] = make_float2(-INFINITY, -INFINITY);
        if (r < a.rows && p < pairs) t[u][j] = __ldg(reinterpret_cast<const float2*>(a.logits + r * n) + p);
      }
      if (r < a.rows && a.labels) {
        const int64_t l64 = __ldg(a.labels + r);
        label[u] = (l64 >= 0 && l64 < n) ? static_cast<int>(l64) : -1;   // outside [0, n): ignored row
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + 4 * u + rsel;
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) m = fmaxf(m, fmaxf(t[u][j].x, t[u][j].y));
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      // e = exp(t - m) (0 for the padding pairs: exp2(-inf)); the row's sum over the eight lanes
      float2 e[4];
      float s = 0.f;
      const float mb = m * kLog2e;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        e[j].x = exp2f(fmaf(t[u][j].x, kLog2e, -mb));
        e[j].y = exp2f(fmaf(t[u][j].y, kLog2e, -mb));
        s += e[j].x + e[j].y;
      }
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (r >= a.rows) continue;                          // (after the shuffles: warp-uniform control above)
      const float lse = m + __logf(s), inv = 1.f / s;
      const bool live = label[u] >= 0;
      if (sub == 0 && live) valid += 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = sub + 8 * j;
        const bool ok = p < pairs;
        const float l0 = t[u][j].x - lse, l1 = t[u][j].y - lse;
        const float p0 = e[j].x * inv, p1 = e[j].y * inv;
        if (label[u] == 2 * p) loss -= l0;
        else if (label[u] == 2 * p + 1) loss -= l1;
        if (probs) {
          const float o0 = ok ? (lsm ? l0 : p0) : 0.f, o1 = ok ? (lsm ? l1 : p1) : 0.f;
          probs[r * 32 + p] = kBf16 ? pack_bf16x2(o0, o1) : pack_f16x2_sat(o0, o1);
        }
        if (dz) {
          const float d0 = (ok && live) ? a.dz_gain * (p0 - (label[u] == 2 * p ? 1.f : 0.f)) : 0.f;
          const float d1 = (ok && live) ? a.dz_gain * (p1 - (label[u] == 2 * p + 1 ? 1.f : 0.f)) : 0.f;
          dz[r * 32 + p] = kBf16 ? pack_bf16x2(d0, d1) : pack_f16x2_sat(d0, d1);
        }
      }
    }
  }
  if (a.loss_sum || a.valid_count) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      loss += __shfl_xor_sync(0xffffffffu, loss, o);
      valid += __shfl_xor_sync(0xffffffffu, valid, o);
    }
    if (lane == 0) { red[warp] = loss; red[8 + warp] = valid; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float sum = 0.f, v = 0.f;
      for (int w = 0; w < 8; ++w) { sum += red[w]; v += red[8 + w]; }
      if (a.loss_sum) atomicAdd(a.loss_sum, sum);
      if (a.valid_count) atomicAdd(a.valid_count, v);
    }
  }
}

// Packed path ([rows, 64] 16-bit lp / dy / dz with contiguous rows): a warp stages its 32 rows of
// lp and dy with coalesced 16-byte loads, thread = row computes in place, coalesced store.
__global__ void __launch_bounds__(128) logsoftmax_bwd16_kernel(const uint32_t* __restrict__ lp,
                                                              const uint32_t* __restrict__ dy, int64_t rows,
                                                              int n, const float* scale, uint32_t* dz, int bf16) {
  __shared__ uint32_t sm_lp[4][32 * kPitchW];
  __shared__ uint32_t sm_dy[4][32 * kPitchW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float sc = scale ? *scale : 1.f;
  const int64_t nblk = (rows + 31) / 32;
  for (int64_t blk = static_cast<int64_t>(blockIdx.x) * 4 + warp; blk < nblk;
       blk += static_cast<int64_t>(gridDim.x) * 4) {
    const int64_t row0 = blk * 32;
    const int nrows = rows - row0 < 32 ? static_cast<int>(rows - row0) : 32;
    __syncwarp();
#pragma unroll 8
    for (int r = 0; r < nrows; ++r) {
      sm_lp[warp][r * kPitchW + lane] = __ldg(lp + (row0 + r) * 32 + lane);
      sm_dy[warp][r * kPitchW + lane] = __ldg(dy + (row0 + r) * 32 + lane);
    }
    __syncwarp();
    if (lane < nrows) {
      uint32_t* l32 = &sm_lp[warp][lane * kPitchW];
      uint32_t* d32 = &sm_dy[warp][lane * kPitchW];
      float sum = 0.f;
      for (int c = 0; c < 32; ++c) {
        float2 g;
        if (bf16) g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&d32[c]));
        else g = __half22float2(*reinterpret_cast<const __half2*>(&d32[c]));
        sum += (2 * c < n ? g.x : 0.f) + (2 * c + 1 < n ? g.y : 0.f);
      }
      for (int c = 0; c < 32; ++c) {
        float2 g, l;
        if (bf16) {
          g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&d32[c]));
          l = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l32[c]));
        } else {
          g = __half22float2(*reinterpret_cast<const __half2*>(&d32[c]));
          l = __half22float2(*reinterpret_cast<const __half2*>(&l32[c]));
        }
        const float o0 = 2 * c < n ? (g.x - __expf(l.x) * sum) * sc : 0.f;
        const float o1 = 2 * c + 1 < n ? (g.y - __expf(l.y) * sum) * sc : 0.f;
        d32[c] = bf16 ? pack_bf16x2(o0, o1) : pack_f16x2_sat(o0, o1);
      }
    }
    __syncwarp();
#pragma unroll 8
    for (int r = 0; r < nrows; ++r) dz[(row0 + r) * 32 + lane] = sm_dy[warp][r * kPitchW + lane];
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


// Column sums of what a 16-bit conversion drops: out[c] += sum_r (v - round16(v)), v = src[r, c] * scale.
// A bias gradient is a plain sum over rows of dz; when dz is stored in 16 bits and its entries barely
// vary from row to row (softmax - onehot of a freshly initialised head: every entry ~1/k), the
// roundings are coherent and their sum grows like the row count instead of its square root.  The
// caller adds (*inv_scale) * out to the bias gradient formed from the 16-bit dz, which makes it the
// exact fp32 column sum the reference's autograd computes.
__global__ void __launch_bounds__(256) round_residual_kernel(const float* __restrict__ src, int64_t ld,
                                                             int64_t rows, int cols, const float* scale,
                                                             int dtype, float* out) {
  __shared__ float red[4][64];
  const int c = threadIdx.x & 63, sub = threadIdx.x >> 6;          // 64 columns x 4 row lanes
  const float sc = scale ? *scale : 1.f;
  float acc = 0.f;
  if (c < cols) {
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * 4 + sub; r < rows; r += static_cast<int64_t>(gridDim.x) * 4) {
      const float v = __ldg(src + r * ld + c) * sc;
      float q;
      if (dtype == PCADV_F16) q = __half2float(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
      else q = __bfloat162float(__float2bfloat16_rn(v));
      acc += v - q;
    }
  }
  red[sub][c] = acc;
  __syncthreads();
  if (sub == 0 && c < cols) atomicAdd(out + c, red[0][c] + red[1][c] + red[2][c] + red[3][c]);
}


// StackDiscNet's custom activation (models/discriminator.py:153-159): per row of the shape logits
// [rows, S], z = logsumexp_c x[r, c], y[r] = z / (z + 1); backward dx[r, c] = dy[r] * softmax(x[r, :])[c]
// / (z + 1)^2.  One thread per row (S <= 64): rows are short and contiguous.
__global__ void __launch_bounds__(256) lse_ratio_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int S,
                                                        const float* __restrict__ dy, float* __restrict__ y,
                                                        float* __restrict__ dx, int64_t ld_dx) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* xr = x + r * ld;
  float m = xr[0];
  for (int c = 1; c < S; ++c) m = fmaxf(m, xr[c]);
  float s = 0.f;
  for (int c = 0; c < S; ++c) s += expf(xr[c] - m);
  const float z = m + logf(s);
  if (y) y[r] = z / (z + 1.f);
  if (dx) {
    const float g = dy[r] / ((z + 1.f) * (z + 1.f) * s);
    for (int c = 0; c < S; ++c) dx[r * ld_dx + c] = g * expf(xr[c] - m);
  }
}

}  // namespace
}  // namespace pcadv

using namespace pcadv;

extern "C" int pcadv_lse_ratio(const float* x, int64_t ld, int64_t rows, int32_t S, const float* dy, float* y,
                               float* dx, int64_t ld_dx, void* stream) {
  PCADV_CHECK_ARG(x && rows >= 0 && S > 0 && S <= 64 && (y || (dx && dy)), "pcadv_lse_ratio: bad args (S <= 64)");
  if (rows == 0) return 0;
  lse_ratio_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ld, rows, S, dy, y, dx, ld_dx);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_round_residual(const float* src, int64_t ld, int64_t rows, int32_t cols, const float* scale,
                                    int32_t dtype, float* out, void* stream) {
  PCADV_CHECK_ARG(src && out && rows >= 0 && cols > 0 && cols <= 64 && (dtype == PCADV_F16 || dtype == PCADV_BF16),
                  "pcadv_round_residual: bad args (1 <= cols <= 64, 16-bit dtype)");
  if (rows == 0) return 0;
  int64_t grid = (rows + 3) / 4;
  if (grid > 148 * 8) grid = 148 * 8;
  round_residual_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, ld, rows, cols, scale, dtype, out);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_softmax_head(const pcadv_head_args* a, void* stream) {
  PCADV_CHECK_ARG(a && a->logits && a->rows >= 0 && a->n > 0 && a->n <= kHeadMaxN,
                  "pcadv_softmax_head: bad args (1 <= n <= %d)", kHeadMaxN);
  PCADV_CHECK_ARG(a->mode == PCADV_HEAD_CE || a->mode == PCADV_HEAD_LSM, "pcadv_softmax_head: bad mode");
  PCADV_CHECK_ARG(!a->probs || a->probs_cols >= a->n, "pcadv_softmax_head: probs_cols < n");
  PCADV_CHECK_ARG(!a->dz || (a->dz_cols >= a->n && a->labels), "pcadv_softmax_head: dz needs labels, dz_cols >= n");
  if (a->rows == 0) return 0;
  {
    auto packed16 = [](const void* p, int64_t ld, int dtype, int cols) {
      return p == nullptr || (dtype != PCADV_F32 && cols == 64 && ld == 64 && (reinterpret_cast<uintptr_t>(p) & 3) == 0);
    };
    const bool same_dt = !a->probs || !a->dz || a->probs_dtype == a->dz_dtype;
    if (a->n <= 64 && a->n % 2 == 0 && a->ld == a->n && (reinterpret_cast<uintptr_t>(a->logits) & 7) == 0 &&
        (a->probs || a->dz) && packed16(a->probs, a->ld_probs, a->probs_dtype, a->probs_cols) &&
        packed16(a->dz, a->ld_dz, a->dz_dtype, a->dz_cols) && same_dt) {
      const int dt = a->probs ? a->probs_dtype : a->dz_dtype;
      int64_t grid = (a->rows + 63) / 64;               // 8 warps x 8 rows per trip
      if (grid > 148 * 8) grid = 148 * 8;
      if (dt == PCADV_BF16)
        softmax_head_rows_kernel<true><<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
      else
        softmax_head_rows_kernel<false><<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
      PCADV_LAUNCHED();
      return 0;
    }
  }
  const size_t smem = static_cast<size_t>(8) * (32 * (a->n | 1) + 32 * kPitchW) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    PCADV_CUDA_OK(cudaFuncSetAttribute(softmax_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       8 * (32 * (kHeadMaxN | 1) + 32 * kPitchW) * 4));
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
  const bool packed = lp_dtype != PCADV_F32 && dy_dtype == lp_dtype && dz_dtype == lp_dtype && n <= 64 &&
                      ld_lp == 64 && ld_dy == 64 && ld_dz == 64 && dz_cols == 64 &&
                      ((reinterpret_cast<uintptr_t>(lp) | reinterpret_cast<uintptr_t>(dy) |
                        reinterpret_cast<uintptr_t>(dz)) & 15) == 0;
  if (packed) {
    int64_t grid = ((rows + 31) / 32 + 3) / 4;
    if (grid > 148 * 12) grid = 148 * 12;
    logsoftmax_bwd16_kernel<<<static_cast<unsigned>(grid), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint32_t*>(lp), reinterpret_cast<const uint32_t*>(dy), rows, n, scale,
        reinterpret_cast<uint32_t*>(dz), lp_dtype == PCADV_BF16 ? 1 : 0);
    PCADV_LAUNCHED();
    return 0;
  }
  logsoftmax_bwd_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(lp, lp_dtype, ld_lp, dy, dy_dtype, ld_dy,
                                                               rows, n, scale, dz, dz_dtype, ld_dz, dz_cols);
  PCADV_LAUNCHED();
  return 0;
}
