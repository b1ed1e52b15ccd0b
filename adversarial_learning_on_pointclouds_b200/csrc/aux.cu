// Auxiliary HBM-bound kernels around the hot path:
//
//   pcadv_query_workspace   sizes of the caller-owned scratch buffers (lower face)
//   pcadv_jitter            on-device point jitter (dataset/modelNetData.py:80-91), counter-based RNG
//   pcadv_bn_*              OPTIONAL BatchNorm over point-major rows: per-channel statistics as
//                           warp-shuffle + shared-memory reductions, normalise (+ReLU), backward.
//                           The reference has no BatchNorm (the only mentions are
//                           commented out, models/pointnet.py:100-103); this layer is default-off and is
//                           checked against torch.nn.BatchNorm1d, never enabled in a parity run.
#include "common.cuh"

namespace pcadv {
namespace {

// ---------------------------------------------------------------------------------- Philox-4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__device__ __forceinline__ void philox4x32_10(uint64_t counter, uint64_t seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {static_cast<uint32_t>(counter), static_cast<uint32_t>(counter >> 32), 0u, 0u};
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

// 24-bit uniform in (0, 1): exact in fp32, never 0 (the logarithm below stays finite)
__device__ __forceinline__ float u01(uint32_t x) { return (static_cast<float>(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// dst[i] = src[i] + clamp(sigma * z_i, -clip, clip), z_i ~ N(0, 1): element i takes normal (i & 3) of
// Philox counter (offset + (i >> 2)) -- Box-Muller on the pairs (u0, u1) and (u2, u3).
__global__ void __launch_bounds__(256) jitter_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                     int64_t count, float sigma, float clip, uint64_t seed,
                                                     uint64_t offset) {
  const int64_t quads = (count + 3) >> 2;
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < quads;
       q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    uint32_t r[4];
    philox4x32_10(offset + static_cast<uint64_t>(q), seed, r);
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float rad = sqrtf(-2.0f * logf(u01(r[2 * h])));
      float s, c;
      sincosf(6.283185307179586f * u01(r[2 * h + 1]), &s, &c);
      z[2 * h] = rad * c;
      z[2 * h + 1] = rad * s;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = q * 4 + j;
      if (i < count) dst[i] = src[i] + fminf(fmaxf(sigma * z[j], -clip), clip);
    }
  }
}

// ---------------------------------------------------------------------------------- BatchNorm rows
// Layout of every bn kernel: a thread owns 8 consecutive channels (one 16-byte load of 16-bit data,
// two of fp32); G = C / 8 channel groups; a block of 256 threads covers 256 / G rows per trip
// (G <= 256, 256 % G == 0), lanes of a warp that own the same channel group are G apart.
constexpr int kBnThreads = 256;

// sums of (x - pivot) and (x - pivot)^2 per channel, pivot = x[0, c]: conditions the variance
// against a large mean (E[x^2] - E[x]^2 cancels otherwise).  Optional second operand: with `dy`
// given the kernel instead accumulates sum(dy') and sum(dy' * xhat), dy' = dy * [y > 0] when a ReLU
// follows the normalisation -- the two reductions of the backward.
__global__ void __launch_bounds__(kBnThreads) bn_reduce_kernel(
    const void* __restrict__ x, int x_dtype, int64_t ld_x, int64_t rows, int C, const void* __restrict__ dy,
    int dy_dtype, int64_t ld_dy, const void* __restrict__ y, int y_dtype, int64_t ld_y,
    const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ out_a,
    float* __restrict__ out_b) {
  __shared__ float red[kBnThreads / 32][2][8 * 32];     // [warp][quantity][lane-local 8 channels x up to 32 groups]
  const int G = C >> 3;
  const int grp = threadIdx.x % G, rsub = threadIdx.x / G, rows_per_trip = kBnThreads / G;
  const int c0 = grp * 8;
  float a[8], b[8], p0[8], p1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  if (dy == nullptr) {
    load8(x, c0, x_dtype, p0);                            // pivot: row 0
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { p0[j] = mean[c0 + j]; p1[j] = rstd[c0 + j]; }
  }
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * rows_per_trip + rsub; r < rows;
       r += static_cast<int64_t>(gridDim.x) * rows_per_trip) {
    float xv[8];
    load8(x, r * ld_x + c0, x_dtype, xv);
    if (dy == nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = xv[j] - p0[j]; a[j] += d; b[j] = fmaf(d, d, b[j]); }
    } else {
      float g[8];
      load8(dy, r * ld_dy + c0, dy_dtype, g);
      if (y != nullptr) {
        float yv[8];
        load8(y, r * ld_y + c0, y_dtype, yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] += g[j]; b[j] = fmaf(g[j], (xv[j] - p0[j]) * p1[j], b[j]); }
    }
  }
  // lanes that own the same channel group inside a warp are G apart (G < 32): shuffle-reduce them
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (G < 32) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      for (int o = 16; o >= G; o >>= 1) {
        a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
        b[j] += __shfl_xor_sync(0xffffffffu, b[j], o);
      }
  }
  const int owners = G < 32 ? G : 32;                     // lanes of a warp holding distinct groups
  if (lane < owners) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[warp][0][lane * 8 + j] = a[j]; red[warp][1][lane * 8 + j] = b[j]; }
  }
  __syncthreads();
  // warps that hold the same channel groups: with G >= 32 warp w holds groups 32 (w % (G / 32)) ..,
  // with G < 32 every warp holds all groups
  const int span = G < 32 ? 1 : G / 32;                   // distinct group blocks among the warps
  for (int i = threadIdx.x; i < C; i += kBnThreads) {
    const int gblock = (i >> 3) / owners, within = (i >> 3) % owners;
    float sa = 0.f, sb = 0.f;
    for (int w = gblock % span; w < kBnThreads / 32; w += span) {
      sa += red[w][0][within * 8 + (i & 7)];
      sb += red[w][1][within * 8 + (i & 7)];
    }
    atomicAdd(out_a + i, sa);
    atomicAdd(out_b + i, sb);
  }
}

// mean / rstd from the pivoted sums; running statistics as nn.BatchNorm1d updates them (momentum,
// unbiased variance).  One thread per channel.
__global__ void bn_finalize_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ s,
                                   const float* __restrict__ ss, int64_t rows, int C, float eps, float momentum,
                                   float* mean, float* rstd, float* running_mean, float* running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float pivot = ld_as_float(x, c, x_dtype);
  const float inv = 1.0f / static_cast<float>(rows);
  const float m = s[c] * inv;
  float var = ss[c] * inv - m * m;
  var = var > 0.f ? var : 0.f;
  mean[c] = pivot + m;
  rstd[c] = rsqrtf(var + eps);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (pivot + m);
  if (running_var) {
    const float unbiased = rows > 1 ? var * static_cast<float>(rows) / static_cast<float>(rows - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// y = act((x - mean) * rstd * gamma + beta);  with dy: dx = gamma * rstd * (dy' - dbeta / R - xhat * dgamma / R)
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(
    const void* __restrict__ x, int x_dtype, int64_t ld_x, int64_t rows, int C, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta, int act,
    const void* __restrict__ dy, int dy_dtype, int64_t ld_dy, const void* __restrict__ y_in, int y_dtype,
    int64_t ld_yin, const float* __restrict__ dgamma, const float* __restrict__ dbeta, void* __restrict__ out,
    int out_dtype, int64_t ld_out) {
  const int G = C >> 3;
  const int grp = threadIdx.x % G, rsub = threadIdx.x / G, rows_per_trip = kBnThreads / G;
  const int c0 = grp * 8;
  float m[8], rs[8], ga[8], be[8], dg[8], db[8];
  const float invR = 1.0f / static_cast<float>(rows);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m[j] = mean[c0 + j]; rs[j] = rstd[c0 + j];
    ga[j] = gamma ? gamma[c0 + j] : 1.f;
    be[j] = beta ? beta[c0 + j] : 0.f;
    dg[j] = dgamma ? dgamma[c0 + j] * invR : 0.f;
    db[j] = dbeta ? dbeta[c0 + j] * invR : 0.f;
  }
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * rows_per_trip + rsub; r < rows;
       r += static_cast<int64_t>(gridDim.x) * rows_per_trip) {
    float xv[8], o[8];
    load8(x, r * ld_x + c0, x_dtype, xv);
    if (dy == nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = fmaf((xv[j] - m[j]) * rs[j], ga[j], be[j]);
        o[j] = act == PCADV_ACT_RELU ? fmaxf(v, 0.f) : v;
      }
    } else {
      float g[8];
      load8(dy, r * ld_dy + c0, dy_dtype, g);
      if (y_in != nullptr) {
        float yv[8];
        load8(y_in, r * ld_yin + c0, y_dtype, yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = ga[j] * rs[j] * (g[j] - db[j] - (xv[j] - m[j]) * rs[j] * dg[j]);
    }
    store8(out, r * ld_out + c0, out_dtype, o);
  }
}

bool bn_shape_ok(int C) {
  const int G = C >> 3;
  return C > 0 && C % 8 == 0 && G <= kBnThreads && kBnThreads % G == 0;
}

unsigned bn_grid(int64_t rows, int C) {
  const int rows_per_trip = kBnThreads / (C >> 3);
  int64_t g = (rows + rows_per_trip - 1) / rows_per_trip;
  if (g > 148 * 8) g = 148 * 8;
  return static_cast<unsigned>(g < 1 ? 1 : g);
}

}  // namespace
}  // namespace pcadv

using namespace pcadv;

extern "C" long long pcadv_query_workspace(int32_t op, int64_t groups, int64_t rows_per_group, int32_t n) {
  switch (op) {
    case PCADV_WS_MAXPOOL_BWD_INPLACE:                    // pcadv_maxbwd_args.workspace with dz_inout:
      // per cloud [ends N | list n | dz n | row n | pad to 16 B | n packed 16-byte entries] (misc.cu)
      return groups * ((rows_per_group + 3ll * n + 3) / 4 * 4 + 4ll * n) * 4ll;
    case PCADV_WS_AMAX_SCALE:                             // pcadv_amax_scale workspace (one uint32)
      return 4;
    default:
      set_error("pcadv_query_workspace: unknown op %d", op);
      return -1;
  }
}

extern "C" int pcadv_jitter(const float* src, float* dst, int64_t count, float sigma, float clip,
                            unsigned long long seed, unsigned long long offset, void* stream) {
  PCADV_CHECK_ARG(src && dst && count >= 0 && clip > 0.f, "pcadv_jitter: bad args (clip > 0)");
  if (count == 0) return 0;
  int64_t grid = ((count + 3) / 4 + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  jitter_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, count, sigma,
                                                                                           clip, seed, offset);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_bn_stats(const void* x, int32_t x_dtype, int64_t ld_x, int64_t rows, int32_t C, float eps,
                              float momentum, float* workspace, float* mean, float* rstd, float* running_mean,
                              float* running_var, void* stream) {
  PCADV_CHECK_ARG(x && workspace && mean && rstd && rows > 0 && bn_shape_ok(C),
                  "pcadv_bn_stats: bad args (C %% 8 == 0, C / 8 a divisor of 256)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bn_reduce_kernel<<<bn_grid(rows, C), kBnThreads, 0, s>>>(x, x_dtype, ld_x, rows, C, nullptr, 0, 0, nullptr, 0, 0,
                                                           nullptr, nullptr, workspace, workspace + C);
  PCADV_LAUNCHED();
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(x, x_dtype, workspace, workspace + C, rows, C, eps, momentum,
                                                     mean, rstd, running_mean, running_var);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_bn_apply(const void* x, int32_t x_dtype, int64_t ld_x, int64_t rows, int32_t C,
                              const float* mean, const float* rstd, const float* gamma, const float* beta,
                              int32_t act, void* y, int32_t y_dtype, int64_t ld_y, void* stream) {
  PCADV_CHECK_ARG(x && y && mean && rstd && rows >= 0 && bn_shape_ok(C) && (act == PCADV_ACT_NONE || act == PCADV_ACT_RELU),
                  "pcadv_bn_apply: bad args");
  if (rows == 0) return 0;
  bn_apply_kernel<<<bn_grid(rows, C), kBnThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_dtype, ld_x, rows, C, mean, rstd, gamma, beta, act, nullptr, 0, 0, nullptr, 0, 0, nullptr, nullptr, y,
      y_dtype, ld_y);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_bn_bwd(const void* x, int32_t x_dtype, int64_t ld_x, const void* dy, int32_t dy_dtype,
                            int64_t ld_dy, const void* y, int32_t y_dtype, int64_t ld_y, int64_t rows, int32_t C,
                            const float* mean, const float* rstd, const float* gamma, float* dgamma, float* dbeta,
                            void* dx, int32_t dx_dtype, int64_t ld_dx, void* stream) {
  PCADV_CHECK_ARG(x && dy && mean && rstd && dgamma && dbeta && rows > 0 && bn_shape_ok(C), "pcadv_bn_bwd: bad args");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bn_reduce_kernel<<<bn_grid(rows, C), kBnThreads, 0, s>>>(x, x_dtype, ld_x, rows, C, dy, dy_dtype, ld_dy, y, y_dtype,
                                                           ld_y, mean, rstd, dbeta, dgamma);
  PCADV_LAUNCHED();
  if (dx) {
    bn_apply_kernel<<<bn_grid(rows, C), kBnThreads, 0, s>>>(x, x_dtype, ld_x, rows, C, mean, rstd, gamma, nullptr,
                                                            PCADV_ACT_NONE, dy, dy_dtype, ld_dy, y, y_dtype, ld_y,
                                                            dgamma, dbeta, dx, dx_dtype, ld_dx);
    PCADV_LAUNCHED();
  }
  return 0;
}
