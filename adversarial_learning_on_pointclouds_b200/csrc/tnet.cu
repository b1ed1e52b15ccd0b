// T-Net transform kernels (fp32, CUDA cores; k <= 128): the per-cloud k x k matrices are far too
// small for the tensor-core tiles and the work is a few GFLOP per batch, so these are
// shared-memory tiled FFMA kernels with one launch per op instead of one torch.bmm per cloud.
//
//   pcadv_bmm          y[b, n, :] = x[b, n, :] @ T[b]   (or @ T[b]^T: the backward dx)
//                      torch.bmm(x^T, trans) at models/pointnet.py:120-122, :231, :238
//   pcadv_bmm_tgrad    dT[b] += x[b]^T @ dy[b]           (backward with respect to the transform)
//   pcadv_ortho_reg    diff[b] = T[b] T[b]^T - I,  norms[b] = ||diff[b]||_F
//                      feature_transform_regularizer, models/pointnet.py:345-353
//   pcadv_ortho_reg_bwd  dT[b] = (2 * dloss / (B * norms[b])) * diff[b] @ T[b]
#include "common.cuh"

namespace pcadv {
namespace {

constexpr int kTnMaxK = 128;
constexpr int kBmmRows = 64;                 // rows of x per tile
constexpr int kBmmTiles = 8;                 // tiles per CTA (T[b] is staged once for all of them)

// grid (ceil(N / 64), B); 256 threads = 16 row groups (4 rows) x 16 column groups (<= 8 columns)
__global__ void __launch_bounds__(256) bmm_kernel(const float* __restrict__ x, const float* __restrict__ T,
                                                  float* __restrict__ y, int N, int k, int transpose_t) {
  extern __shared__ float sm[];
  float* Ts = sm;                                        // [k][k + 1]   Ts[j][c] = T_eff[j, c]
  float* xs = sm + k * (k + 1);                          // [64][k + 1]
  const int b = blockIdx.y, t = threadIdx.x;
  const float* Tb = T + static_cast<int64_t>(b) * k * k;
  for (int e = t; e < k * k; e += 256) {
    const int j = e / k, c = e - j * k;
    Ts[j * (k + 1) + c] = transpose_t ? Tb[c * k + j] : Tb[e];
  }
  const int ty = t >> 4, tx = t & 15;
  const int cpt = (k + 15) / 16;                         // columns per thread (<= 8)
  for (int tile = 0; tile < kBmmTiles; ++tile) {
  const int n0 = (blockIdx.x * kBmmTiles + tile) * kBmmRows;
  if (n0 >= N) break;
  const float* xb = x + (static_cast<int64_t>(b) * N + n0) * k;
  const int rows = N - n0 < kBmmRows ? N - n0 : kBmmRows;
  __syncthreads();                                       // previous tile's readers are done
  for (int e = t; e < rows * k; e += 256) {
    const int r = e / k, j = e - r * k;
    xs[r * (k + 1) + j] = xb[e];
  }
  __syncthreads();
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
  for (int j = 0; j < k; ++j) {
    float xv[4], tv[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) xv[i] = xs[(ty * 4 + i) * (k + 1) + j];
#pragma unroll
    for (int c = 0; c < 8; ++c) tv[c] = (c < cpt && tx + 16 * c < k) ? Ts[j * (k + 1) + tx + 16 * c] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(xv[i], tv[c], acc[i][c]);
  }
  float* yb = y + (static_cast<int64_t>(b) * N + n0) * k;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int col = tx + 16 * c;                       // consecutive lanes, consecutive columns
      if (c < cpt && col < k) yb[r * k + col] = acc[i][c];
    }
  }
  }
}

constexpr int kTgChunk = 64;                 // rows staged per trip

// grid (splits, B); each CTA reduces a row range of cloud b into a k x k tile (8 x 8 per thread
// for k = 128) and adds it into dT[b] with fp32 atomics
__global__ void __launch_bounds__(256) bmm_tgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        float* __restrict__ dT, int N, int k, int rows_per_split) {
  extern __shared__ float sm[];
  float* xs = sm;                                        // [kTgChunk][k]
  float* ds = sm + kTgChunk * k;                         // [kTgChunk][k]
  const int b = blockIdx.y, t = threadIdx.x;
  const int r0 = blockIdx.x * rows_per_split;
  const int r1 = r0 + rows_per_split < N ? r0 + rows_per_split : N;
  const int ty = t >> 4, tx = t & 15;                    // rows j = ty + 16 i, cols c = tx + 16 q
  const int per = (k + 15) / 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] = 0.f;
  const float* xb = x + static_cast<int64_t>(b) * N * k;
  const float* db = dy + static_cast<int64_t>(b) * N * k;
  for (int r = r0; r < r1; r += kTgChunk) {
    const int rows = r1 - r < kTgChunk ? r1 - r : kTgChunk;
    __syncthreads();
    for (int e = t; e < kTgChunk * k; e += 256) {
      const bool ok = e < rows * k;
      xs[e] = ok ? xb[static_cast<int64_t>(r) * k + e] : 0.f;
      ds[e] = ok ? db[static_cast<int64_t>(r) * k + e] : 0.f;
    }
    __syncthreads();
    for (int rr = 0; rr < kTgChunk; ++rr) {
      float xv[8], dv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = (i < per && ty + 16 * i < k) ? xs[rr * k + ty + 16 * i] : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) dv[q] = (q < per && tx + 16 * q < k) ? ds[rr * k + tx + 16 * q] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[i][q] = fmaf(xv[i], dv[q], acc[i][q]);
    }
  }
  float* out = dT + static_cast<int64_t>(b) * k * k;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = ty + 16 * i, c = tx + 16 * q;
      if (i < per && q < per && j < k && c < k) atomicAdd(out + j * k + c, acc[i][q]);
    }
}

// one CTA per cloud: out[i, c] = alpha_b * sum_j A[i, j] B[j, c]  (+ optional "- I" and Frobenius norm)
//   mode 0: A = T, B = T^T  -> diff = T T^T - I, norms[b] = ||diff||_F
//   mode 1: A = diff, B = T -> dT = (2 * dloss / (nb * norms[b])) * diff @ T
__global__ void __launch_bounds__(256) ortho_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                    float* __restrict__ out, float* norms,
                                                    const float* dloss, int nb, int d, int mode) {
  extern __shared__ float sm[];
  float* As = sm;                                        // [d][d + 1]
  float* Bs = sm + d * (d + 1);                          // [d][d + 1]   Bs[j][c]
  __shared__ float red[8];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* Ab = A + static_cast<int64_t>(b) * d * d;
  const float* Bb = Bm + static_cast<int64_t>(b) * d * d;
  for (int e = t; e < d * d; e += 256) {
    const int i = e / d, j = e - i * d;
    As[i * (d + 1) + j] = Ab[e];
    if (mode == 0) Bs[j * (d + 1) + i] = Bb[e];          // B = T^T: Bs[j][c] = T[c][j]
    else Bs[i * (d + 1) + j] = Bb[e];
  }
  __syncthreads();
  const int ty = t >> 4, tx = t & 15;
  const int per = (d + 15) / 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] = 0.f;
  for (int j = 0; j < d; ++j) {
    float av[8], bv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = (i < per && ty + 16 * i < d) ? As[(ty + 16 * i) * (d + 1) + j] : 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) bv[q] = (q < per && tx + 16 * q < d) ? Bs[j * (d + 1) + tx + 16 * q] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[i][q] = fmaf(av[i], bv[q], acc[i][q]);
  }
  float alpha = 1.f;
  if (mode == 1) {
    const float nrm = norms[b];
    alpha = nrm > 1e-30f ? 2.f * (*dloss) / (static_cast<float>(nb) * nrm) : 0.f;
  }
  float ss = 0.f;
  float* ob = out + static_cast<int64_t>(b) * d * d;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = ty + 16 * i, c = tx + 16 * q;
      if (i < per && q < per && r < d && c < d) {
        float v = acc[i][q];
        if (mode == 0) { v -= (r == c) ? 1.f : 0.f; ss += v * v; }
        else v *= alpha;
        ob[r * d + c] = v;
      }
    }
  if (mode == 0) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((t & 31) == 0) red[t >> 5] = ss;
    __syncthreads();
    if (t == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w];
      norms[b] = sqrtf(s);
    }
  }
}

int set_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    PCADV_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  return 0;
}

}  // namespace
}  // namespace pcadv

using namespace pcadv;

extern "C" int pcadv_bmm(const float* x, const float* T, float* y, int32_t groups, int64_t rows_per_group,
                         int32_t k, int32_t transpose_t, void* stream) {
  PCADV_CHECK_ARG(x && T && y && groups >= 0 && rows_per_group >= 0 && k >= 1 && k <= kTnMaxK,
                  "pcadv_bmm: bad args (1 <= k <= %d)", kTnMaxK);
  if (groups == 0 || rows_per_group == 0) return 0;
  const size_t smem = (static_cast<size_t>(k) * (k + 1) + kBmmRows * (k + 1)) * sizeof(float);
  if (int rc = set_smem(reinterpret_cast<const void*>(bmm_kernel), smem)) return rc;
  dim3 grid(static_cast<unsigned>((rows_per_group + kBmmRows * kBmmTiles - 1) / (kBmmRows * kBmmTiles)), groups);
  bmm_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, T, y, static_cast<int>(rows_per_group),
                                                                     k, transpose_t);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_bmm_tgrad(const float* x, const float* dy, float* dT, int32_t groups,
                               int64_t rows_per_group, int32_t k, void* stream) {
  PCADV_CHECK_ARG(x && dy && dT && groups >= 0 && rows_per_group >= 0 && k >= 1 && k <= kTnMaxK,
                  "pcadv_bmm_tgrad: bad args (1 <= k <= %d)", kTnMaxK);
  if (groups == 0 || rows_per_group == 0) return 0;
  int splits = static_cast<int>((148 * 4 + groups - 1) / groups);
  const int max_splits = static_cast<int>((rows_per_group + 127) / 128);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rps = static_cast<int>((rows_per_group + splits - 1) / splits);
  rps = (rps + kTgChunk - 1) / kTgChunk * kTgChunk;
  splits = static_cast<int>((rows_per_group + rps - 1) / rps);
  const size_t smem = static_cast<size_t>(2 * kTgChunk) * k * sizeof(float);
  if (int rc = set_smem(reinterpret_cast<const void*>(bmm_tgrad_kernel), smem)) return rc;
  dim3 grid(splits, groups);
  bmm_tgrad_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, dy, dT, static_cast<int>(rows_per_group),
                                                                           k, rps);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_ortho_reg(const float* T, int32_t groups, int32_t d, float* diff, float* norms, void* stream) {
  PCADV_CHECK_ARG(T && diff && norms && groups >= 0 && d >= 1 && d <= kTnMaxK, "pcadv_ortho_reg: bad args");
  if (groups == 0) return 0;
  const size_t smem = static_cast<size_t>(2) * d * (d + 1) * sizeof(float);
  if (int rc = set_smem(reinterpret_cast<const void*>(ortho_kernel), smem)) return rc;
  ortho_kernel<<<groups, 256, smem, static_cast<cudaStream_t>(stream)>>>(T, T, diff, norms, nullptr, groups, d, 0);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_ortho_reg_bwd(const float* diff, const float* T, const float* norms, const float* dloss,
                                   int32_t groups, int32_t d, float* dT, void* stream) {
  PCADV_CHECK_ARG(diff && T && norms && dloss && dT && groups >= 0 && d >= 1 && d <= kTnMaxK,
                  "pcadv_ortho_reg_bwd: bad args");
  if (groups == 0) return 0;
  const size_t smem = static_cast<size_t>(2) * d * (d + 1) * sizeof(float);
  if (int rc = set_smem(reinterpret_cast<const void*>(ortho_kernel), smem)) return rc;
  ortho_kernel<<<groups, 256, smem, static_cast<cudaStream_t>(stream)>>>(diff, T, dT, const_cast<float*>(norms),
                                                                         dloss, groups, d, 1);
  PCADV_LAUNCHED();
  return 0;
}
