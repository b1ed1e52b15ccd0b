// HBM-bound helper kernels: max-key unpack, sparse max-pool backward, max-over-
// channels backward, gradient amax / power-of-two scale, dtype convert + pad,
// weight transpose.
#include "common.cuh"

namespace pcadv {
namespace {

__global__ void max_finalize_kernel(const unsigned long long* __restrict__ key, int64_t count,
                                    int act, float slope, float* __restrict__ val,
                                    int32_t* __restrict__ idx) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const unsigned long long k = key[i];
  float v = key_value(k);
  uint32_t id = key_index(k);
  if (k == 0ull) { v = 0.f; id = 0; }      // empty group (never written)
  if (act == PCADV_ACT_RELU && !(v > 0.f)) { v = 0.f; id = 0; }
  else v = apply_act(v, act, slope);
  val[i] = v;
  if (idx) idx[i] = static_cast<int32_t>(id);
}

// ---- sparse max-pool backward ---------------------------------------------------
// dW: one CTA per channel c (no atomics): dw[c, :] = scale * sum_g dzc[g] * x[row(g, c), :]
__global__ void __launch_bounds__(128) maxbwd_dw_kernel(const pcadv_maxbwd_args a) {
  const int c = blockIdx.x;
  const float sc = a.scale ? *a.scale : 1.f;
  float bsum = 0.f;
  for (int k0 = threadIdx.x; k0 < a.k; k0 += blockDim.x * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int g = 0; g < a.groups; ++g) {
      const int64_t gc = static_cast<int64_t>(g) * a.n + c;
      const float dz = a.dg[gc] * act_grad_from_output(a.gval[gc], a.act, a.slope);
      if (dz == 0.f) continue;
      const int64_t r = static_cast<int64_t>(g) * a.rows_per_group + a.idx[gc];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + q * blockDim.x;
        if (k < a.k) acc[q] = fmaf(dz, ld_as_float(a.x, r * a.ldx + k, a.x_dtype), acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + q * blockDim.x;
      if (k < a.k && a.dw) a.dw[static_cast<int64_t>(c) * a.ld_dw + k] += acc[q] * sc;
    }
  }
  if (threadIdx.x == 0 && a.dbias) {
    for (int g = 0; g < a.groups; ++g) {
      const int64_t gc = static_cast<int64_t>(g) * a.n + c;
      bsum += a.dg[gc] * act_grad_from_output(a.gval[gc], a.act, a.slope);
    }
    a.dbias[c] += bsum * sc;
  }
}

// dX: one warp per (g, c): dx_acc[row, :] += dzc * w[c, :]   (fp32 RED atomics)
__global__ void __launch_bounds__(256) maxbwd_dx_kernel(const pcadv_maxbwd_args a) {
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total = static_cast<int64_t>(a.groups) * a.n;
  if (warp >= total) return;
  const int g = static_cast<int>(warp / a.n), c = static_cast<int>(warp % a.n);
  const float dz = a.dg[warp] * act_grad_from_output(a.gval[warp], a.act, a.slope);
  if (dz == 0.f) return;
  const int64_t r = static_cast<int64_t>(g) * a.rows_per_group + a.idx[warp];
  float* dst = a.dx_acc + r * a.ld_dx;
  for (int k = lane; k < a.k; k += 32)
    atomicAdd(dst + k, dz * ld_as_float(a.w, static_cast<int64_t>(c) * a.ldw + k, a.w_dtype));
}

__global__ void rowmax_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ val,
                                  const int32_t* __restrict__ idx, int64_t rows, int n, int act,
                                  float slope, const float* scale, void* dz, int64_t ld_dz,
                                  int dz_dtype) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * n) return;
  const int64_t r = i / n;
  const int c = static_cast<int>(i - r * n);
  float v = 0.f;
  if (c == idx[r]) v = dy[r] * (scale ? *scale : 1.f) * act_grad_from_output(val[r], act, slope);
  st_from_float(dz, r * ld_dz + c, dz_dtype, v);
}

__global__ void amax_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ld,
                            unsigned int* ws) {
  float m = 0.f;
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    const float v = fabsf(x[r * ld + (i - r * cols)]);
    if (v < INFINITY) m = fmaxf(m, v);     // ignores NaN / inf
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(ws, __float_as_uint(m));
}

__global__ void scale_from_amax_kernel(const unsigned int* ws, float target, float* scale2) {
  const float amax = __uint_as_float(*ws);
  float s = 1.f;
  if (amax > 0.f) {
    int e = ilogbf(target / amax);
    e = e > 60 ? 60 : (e < -60 ? -60 : e);
    s = ldexpf(1.f, e);
  }
  scale2[0] = s;
  scale2[1] = 1.f / s;
}

__global__ void convert_kernel(const void* src, int src_dtype, int64_t ld_src, int64_t rows, int cols,
                               void* dst, int dst_dtype, int64_t ld_dst, int cols_pad,
                               const float* scale, const void* mask, int64_t ld_mask, int mask_dtype,
                               int mask_act, float mask_slope) {
  const float sc = scale ? *scale : 1.f;
  const int64_t total = rows * cols_pad;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols_pad;
    const int c = static_cast<int>(i - r * cols_pad);
    float v = c < cols ? ld_as_float(src, r * ld_src + c, src_dtype) * sc : 0.f;
    if (mask && c < cols)
      v *= act_grad_from_output(ld_as_float(mask, r * ld_mask + c, mask_dtype), mask_act, mask_slope);
    st_from_float(dst, r * ld_dst + c, dst_dtype, v);
  }
}

__global__ void transpose_kernel(const void* src, int src_dtype, int64_t ld_src, int rows, int cols,
                                 void* dst, int dst_dtype, int64_t ld_dst) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols)
                               ? ld_as_float(src, static_cast<int64_t>(r) * ld_src + c, src_dtype)
                               : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;   // dst[c, r]
    if (r < rows && c < cols)
      st_from_float(dst, static_cast<int64_t>(c) * ld_dst + r, dst_dtype, tile[threadIdx.x][i]);
  }
}

inline unsigned grid_for(int64_t total, int block, int64_t cap = 148 * 16) {
  int64_t g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace
}  // namespace pcadv

using namespace pcadv;

extern "C" int pcadv_max_finalize(const unsigned long long* key, int64_t count, int32_t act,
                                  float slope, float* val, int32_t* idx, void* stream) {
  PCADV_CHECK_ARG(key && val && count >= 0, "pcadv_max_finalize: bad args");
  if (count == 0) return 0;
  max_finalize_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(key, count, act, slope, val, idx);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_maxpool_bwd(const pcadv_maxbwd_args* a, void* stream) {
  PCADV_CHECK_ARG(a && a->dg && a->gval && a->idx, "pcadv_maxpool_bwd: null input");
  PCADV_CHECK_ARG(a->groups > 0 && a->n > 0 && a->k > 0 && a->rows_per_group > 0,
                  "pcadv_maxpool_bwd: bad shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->dw || a->dbias) {
    PCADV_CHECK_ARG(!a->dw || a->x, "pcadv_maxpool_bwd: dw needs x");
    maxbwd_dw_kernel<<<a->n, 128, 0, s>>>(*a);
    PCADV_LAUNCHED();
  }
  if (a->dx_acc) {
    PCADV_CHECK_ARG(a->w, "pcadv_maxpool_bwd: dx needs w");
    const int64_t warps = static_cast<int64_t>(a->groups) * a->n;
    maxbwd_dx_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, s>>>(*a);
    PCADV_LAUNCHED();
  }
  return 0;
}

extern "C" int pcadv_rowmax_bwd(const float* dy, const float* val, const int32_t* idx, int64_t rows,
                                int32_t n, int32_t act, float slope, const float* scale, void* dz,
                                int64_t ld_dz, int32_t dz_dtype, void* stream) {
  PCADV_CHECK_ARG(dy && val && idx && dz && rows >= 0 && n > 0, "pcadv_rowmax_bwd: bad args");
  if (rows == 0) return 0;
  const int64_t total = rows * n;
  rowmax_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                      static_cast<cudaStream_t>(stream)>>>(dy, val, idx, rows, n, act, slope, scale,
                                                           dz, ld_dz, dz_dtype);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_amax_scale(const float* x, int64_t rows, int32_t cols, int64_t ld, float target,
                                unsigned int* workspace, float* scale2, void* stream) {
  PCADV_CHECK_ARG(x && workspace && scale2 && rows >= 0 && cols > 0 && target > 0.f,
                  "pcadv_amax_scale: bad args");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows > 0) {
    amax_kernel<<<grid_for(rows * cols, 256), 256, 0, s>>>(x, rows, cols, ld, workspace);
    PCADV_LAUNCHED();
  }
  scale_from_amax_kernel<<<1, 1, 0, s>>>(workspace, target, scale2);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_convert(const void* src, int32_t src_dtype, int64_t ld_src, int64_t rows,
                             int32_t cols, void* dst, int32_t dst_dtype, int64_t ld_dst,
                             int32_t cols_pad, const float* scale, const void* mask,
                             int64_t ld_mask, int32_t mask_dtype, int32_t mask_act, float mask_slope,
                             void* stream) {
  PCADV_CHECK_ARG(src && dst && rows >= 0 && cols > 0 && cols_pad >= cols, "pcadv_convert: bad args");
  if (rows == 0) return 0;
  convert_kernel<<<grid_for(rows * cols_pad, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, src_dtype, ld_src, rows, cols, dst, dst_dtype, ld_dst, cols_pad, scale, mask, ld_mask,
      mask_dtype, mask_act, mask_slope);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_transpose(const void* src, int32_t src_dtype, int64_t ld_src, int32_t rows,
                               int32_t cols, void* dst, int32_t dst_dtype, int64_t ld_dst,
                               void* stream) {
  PCADV_CHECK_ARG(src && dst && rows > 0 && cols > 0, "pcadv_transpose: bad args");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(src, src_dtype, ld_src, rows,
                                                                         cols, dst, dst_dtype, ld_dst);
  PCADV_LAUNCHED();
  return 0;
}
