// fp32-FFMA (CUDA-core) GEMM kernels: the fp32-accumulate verification mode of
// every layer, and the production path of the layers that are too small or too
// unaligned for tensor cores (K = 3 first layer, per-cloud heads with M = B
// rows, K = 50 discriminator inputs).
#include "common.cuh"

namespace pcadv {

namespace {

constexpr int BM = 128;   // rows (points) per CTA
constexpr int BN = 64;    // output channels per CTA
constexpr int BK = 16;    // K chunk
constexpr int AS_LD = BM + 4;
constexpr int BS_LD = BN + 4;

// ---- tile loaders: 8 (A) / 4 (W) consecutive k elements per thread ---------------
template <int NV>
__device__ __forceinline__ void load_k_run(const void* base, int dtype, int64_t row_off, int k0,
                                           int kmax, bool row_ok, float (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = 0.f;
  if (!row_ok) return;
  if (dtype == PCADV_F32) {
    const float* p = reinterpret_cast<const float*>(base) + row_off;
    if (k0 + NV <= kmax && ((reinterpret_cast<uintptr_t>(p + k0) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(p + k0 + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (k0 + i < kmax) v[i] = __ldg(p + k0 + i);
    }
  } else if (dtype == PCADV_F16) {
    const __half* p = reinterpret_cast<const __half*>(base) + row_off;
    if (k0 + NV <= kmax && ((reinterpret_cast<uintptr_t>(p + k0) & (NV * 2 - 1)) == 0)) {
      if constexpr (NV == 8) {
        uint4 t = *reinterpret_cast<const uint4*>(p + k0);
        const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
      } else {
        uint2 t = *reinterpret_cast<const uint2*>(p + k0);
        const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
        for (int i = 0; i < 2; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (k0 + i < kmax) v[i] = __half2float(p[k0 + i]);
    }
  } else {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + row_off;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (k0 + i < kmax) v[i] = __bfloat162float(p[k0 + i]);
  }
}

// =====================================================================================
// out[r, c] = post(sum_k A[r, k] W[c, k] + ...), see pcadv_linear in pcadv.h
// =====================================================================================
__global__ void __launch_bounds__(256) simt_linear_kernel(const pcadv_linear_args a) {
  __shared__ __align__(16) float As[BK][AS_LD];
  __shared__ __align__(16) float Bs[BK][BS_LD];
  __shared__ unsigned long long red_keys[16][BN];

  const int t = threadIdx.x;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * BM;
  const int col0 = blockIdx.y * BN;

  const int arow = t >> 1, ak = (t & 1) * 8;
  const int wn = t >> 2, wk = (t & 3) * 4;
  const int ty = t >> 4, tx = t & 15;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t grow = row0 + arow;
  const bool arow_ok = grow < a.rows;
  const int gcol = col0 + wn;
  const bool wcol_ok = gcol < a.n;

  int koff = 0;
  for (int s = 0; s < a.num_seg; ++s) {
    const pcadv_seg sg = a.seg[s];
    for (int k0 = 0; k0 < sg.k; k0 += BK) {
      float av[8], wv[4];
      load_k_run<8>(sg.ptr, sg.dtype, grow * sg.ld, k0 + ak, sg.k, arow_ok, av);
      load_k_run<4>(a.w, a.w_dtype, static_cast<int64_t>(gcol) * a.ldw + koff, k0 + wk, sg.k,
                    wcol_ok, wv);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) As[ak + i][arow] = av[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[wk + i][wn] = wv[i];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float br[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
    }
    koff += sg.k;
  }

  // ---- epilogue -----------------------------------------------------------------
  const float oscale = a.out_scale ? *a.out_scale : 1.f;
  const int64_t rpg = a.rows_per_group > 0 ? a.rows_per_group : a.rows;
  const int64_t last_row = (row0 + BM - 1 < a.rows ? row0 + BM - 1 : a.rows - 1);
  const bool one_group = (row0 / rpg) == (last_row / rpg);

  unsigned long long ckey[4] = {0ull, 0ull, 0ull, 0ull};
  int64_t cgroup = -1;

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = row0 + ty * 8 + i;
    const bool r_ok = r < a.rows;
    const int64_t g = r_ok ? r / rpg : 0;
    if (a.colmax_key && r_ok && !one_group && g != cgroup) {
      if (cgroup >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = col0 + tx * 4 + j;
          if (c < a.n && ckey[j]) atomicMax(&a.colmax_key[cgroup * a.n + c], ckey[j]);
          ckey[j] = 0ull;
        }
      }
      cgroup = g;
    }
    unsigned long long rkey = 0ull;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (!r_ok || c >= a.n) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[c];
      if (a.group_bias) v += a.group_bias[g * a.n + c];
      if (a.addend) v += a.addend[r * a.ld_addend + c];
      const float pre = v;
      v = apply_act(v, a.act, a.slope);
      if (a.mask)
        v *= act_grad_from_output(ld_as_float(a.mask, r * a.ld_mask + c, a.mask_dtype), a.mask_act,
                                  a.mask_slope);
      v *= oscale;
      if (a.out) st_from_float(a.out, r * a.ld_out + c, a.out_dtype, v);
      if (a.colmax_key) {
        const unsigned long long k = pack_key(pre, static_cast<uint32_t>(r - g * rpg));
        ckey[j] = k > ckey[j] ? k : ckey[j];
      }
      if (a.rowmax_key) {
        const unsigned long long k = pack_key(pre, static_cast<uint32_t>(c));
        rkey = k > rkey ? k : rkey;
      }
    }
    if (a.rowmax_key) {
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, rkey, o);
        rkey = other > rkey ? other : rkey;
      }
      if (tx == 0 && r_ok && rkey) atomicMax(&a.rowmax_key[r], rkey);
    }
  }

  if (a.colmax_key) {
    if (one_group) {
      const int64_t g = row0 / rpg;
#pragma unroll
      for (int j = 0; j < 4; ++j) red_keys[ty][tx * 4 + j] = ckey[j];
      __syncthreads();
      if (t < BN) {
        unsigned long long k = 0ull;
#pragma unroll
        for (int q = 0; q < 16; ++q) k = red_keys[q][t] > k ? red_keys[q][t] : k;
        const int c = col0 + t;
        if (c < a.n && k) atomicMax(&a.colmax_key[g * a.n + c], k);
      }
    } else if (cgroup >= 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = col0 + tx * 4 + j;
        if (c < a.n && ckey[j]) atomicMax(&a.colmax_key[cgroup * a.n + c], ckey[j]);
      }
    }
  }
}

// =====================================================================================
// dw[c, koff + k] += scale * sum_r dz[r, c] * x[r, k]   (split over rows, fp32 atomics)
// =====================================================================================
constexpr int WT = 64;   // output tile: 64 (c) x 64 (k)
constexpr int WR = 16;   // rows per smem stage

__global__ void __launch_bounds__(256) simt_wgrad_kernel(const pcadv_wgrad_args a, int seg_index,
                                                        int koff, int64_t rows_per_split) {
  __shared__ __align__(16) float Zs[WR][WT + 4];
  __shared__ __align__(16) float Xs[WR][WT + 4];
  const pcadv_seg sg = a.seg[seg_index];
  const int t = threadIdx.x;
  const int c0 = blockIdx.x * WT;
  const int k0 = blockIdx.y * WT;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.z) * rows_per_split;
  const int64_t r_end = r_begin + rows_per_split < a.rows ? r_begin + rows_per_split : a.rows;

  const int lr = t >> 4, l4 = (t & 15) * 4;   // loader: row lr, 4 consecutive columns
  const int ty = t >> 4, tx = t & 15;         // compute: c = ty*4+i, k = tx*4+j
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = a.dbias != nullptr && blockIdx.y == 0 && seg_index == 0 && tx == 0;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += WR) {
    float zv[4], xv[4];
    const int64_t r = r0 + lr;
    const bool ok = r < r_end;
    load_k_run<4>(a.dz, a.dz_dtype, r * a.ld_dz, c0 + l4, a.n, ok, zv);
    load_k_run<4>(sg.ptr, sg.dtype, r * sg.ld, k0 + l4, sg.k, ok, xv);
    __syncthreads();
    *reinterpret_cast<float4*>(&Zs[lr][l4]) = make_float4(zv[0], zv[1], zv[2], zv[3]);
    *reinterpret_cast<float4*>(&Xs[lr][l4]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < WR; ++q) {
      const float4 z = *reinterpret_cast<const float4*>(&Zs[q][ty * 4]);
      const float4 x = *reinterpret_cast<const float4*>(&Xs[q][tx * 4]);
      const float zr[4] = {z.x, z.y, z.z, z.w};
      const float xr[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(zr[i], xr[j], acc[i][j]);
        bsum[i] += zr[i];
      }
    }
  }
  const float sc = a.scale ? *a.scale : 1.f;
  if (a.dw) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty * 4 + i;
      if (c >= a.n) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + tx * 4 + j;
        if (k < sg.k) atomicAdd(&a.dw[static_cast<int64_t>(c) * a.ld_dw + koff + k], acc[i][j] * sc);
      }
    }
  }
  if (do_bias) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty * 4 + i;
      if (c < a.n) atomicAdd(&a.dbias[c], bsum[i] * sc);
    }
  }
}

// dbias only (no segments): column sums of dz, split over rows.
__global__ void __launch_bounds__(256) colsum_kernel(const void* dz, int dz_dtype, int64_t ld,
                                                    int64_t rows, int n, int64_t rows_per_split,
                                                    const float* scale, float* out) {
  __shared__ float red[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int sub = threadIdx.x >> 6;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.y) * rows_per_split;
  const int64_t r_end = r_begin + rows_per_split < rows ? r_begin + rows_per_split : rows;
  float s = 0.f;
  if (c < n)
    for (int64_t r = r_begin + sub; r < r_end; r += 4) s += ld_as_float(dz, r * ld + c, dz_dtype);
  red[sub][threadIdx.x & 63] = s;
  __syncthreads();
  if (sub == 0 && c < n) {
    s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    atomicAdd(&out[c], s * (scale ? *scale : 1.f));
  }
}

// dgroup_bias[g, c] += sum_{r in cloud g} dz[r, c]
__global__ void __launch_bounds__(256) group_colsum_kernel(const void* dz, int dz_dtype, int64_t ld,
                                                          int n, int64_t rows_per_group,
                                                          float* out) {
  __shared__ float red[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int sub = threadIdx.x >> 6;
  const int64_t g = blockIdx.y;
  const int64_t r_begin = g * rows_per_group, r_end = r_begin + rows_per_group;
  float s = 0.f;
  if (c < n)
    for (int64_t r = r_begin + sub; r < r_end; r += 4) s += ld_as_float(dz, r * ld + c, dz_dtype);
  red[sub][threadIdx.x & 63] = s;
  __syncthreads();
  if (sub == 0 && c < n) {
    s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    out[g * n + c] += s;
  }
}

// 16-bit dz with 16-byte aligned rows: thread = (8-column group, row lane), 16-byte loads, four
// rows in flight per thread
__global__ void __launch_bounds__(256) group_colsum16_kernel(const uint16_t* __restrict__ dz, int bf16,
                                                            int64_t ld, int n, int64_t rows_per_group,
                                                            float* out) {
  __shared__ float red[32][65];
  const int cgi = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cgi * 8;
  const int64_t g = blockIdx.y;
  const uint16_t* base = dz + g * rows_per_group * ld + c0;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c0 < n) {
    for (int64_t r = rl; r < rows_per_group; r += 128) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = r + 32 * u < rows_per_group ? __ldg(reinterpret_cast<const uint4*>(base + (r + 32 * u) * ld))
                                           : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f;
          if (bf16) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
          else f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
          acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][cgi * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float sum = 0.f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) sum += red[i][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < n) out[g * n + c] += sum;
  }
}

// =====================================================================================
// First layer (K <= 4, e.g. Conv1d(3, 64, 1) at models/pointnet.py:115, :291): HBM-bound.
// One thread = one point x 8 output channels: the point's xyz is a broadcast load, the
// 8 x K weights and 8 biases sit in registers, and the 8 threads of a point write one
// contiguous, fully coalesced output row.
// =====================================================================================
template <int K>
__global__ void __launch_bounds__(256, 4) first_layer_kernel(const pcadv_linear_args a) {
  const int groups = a.n >> 3;                       // threads per point
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int cg = static_cast<int>(tid % groups) * 8;
  float w[8][K], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    b[i] = a.bias ? a.bias[cg + i] : 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k)
      w[i][k] = ld_as_float(a.w, static_cast<int64_t>(cg + i) * a.ldw + k, a.w_dtype);
  }
  const float* x = reinterpret_cast<const float*>(a.seg[0].ptr);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x / groups;
  const int gidx = static_cast<int>(tid % groups);
  const bool relu16 = a.act == PCADV_ACT_RELU && a.out_dtype != PCADV_F32;
  // two points per trip: both xyz loads are in flight before either is used.  The trip count is
  // warp-uniform (bounded by the warp's first point) because the sign-bit words are combined
  // with full-warp shuffles.
  const int64_t warp_first = (tid - (threadIdx.x & 31)) / groups;
  for (int64_t off = 0; warp_first + off < a.rows; off += 2 * stride) {
    const int64_t r0 = tid / groups + off;
    float xv[2][K];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + u * stride;
#pragma unroll
      for (int k = 0; k < K; ++k) xv[u][k] = r < a.rows ? __ldg(x + r * a.seg[0].ld + k) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + u * stride;
      const bool r_ok = r < a.rows;                  // uniform over the threads of a point
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float acc = b[i];
#pragma unroll
        for (int k = 0; k < K; ++k) acc = fmaf(xv[u][k], w[i][k], acc);
        v[i] = relu16 ? acc : apply_act(acc, a.act, a.slope);   // 16-bit ReLU: applied on the packed halves below
      }
      if (a.out_dtype == PCADV_F32) {
        if (r_ok) {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + r * a.ld_out + cg);
          o[0] = make_float4(v[0], v[1], v[2], v[3]);
          o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      } else {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          pk[i] = a.out_dtype == PCADV_F16 ? pack_f16x2_sat(v[2 * i], v[2 * i + 1]) : pack_bf16x2(v[2 * i], v[2 * i + 1]);
        if (relu16) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            pk[i] = a.out_dtype == PCADV_F16 ? relu_packed<false>(pk[i]) : relu_packed<true>(pk[i]);
        }
        if (r_ok)
          *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.out) + r * a.ld_out + cg) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (a.bits_out) {
          // sign-bit map (layout in pcadv.h): the four threads that share a 32-column word OR
          // their 4 + 4 bits together; packed pair i of this thread is pair (gidx % 4) * 4 + i
          uint32_t wbits = 0u;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t m;
            if (a.out_dtype == PCADV_F16) asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(pk[i]), "r"(0u));
            else asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(pk[i]), "r"(0u));
            wbits |= m & (0x00010001u << ((gidx & 3) * 4 + i));
          }
          wbits |= __shfl_xor_sync(0xffffffffu, wbits, 1);
          wbits |= __shfl_xor_sync(0xffffffffu, wbits, 2);
          if (r_ok && (gidx & 3) == 0) a.bits_out[r * a.ld_bits_out + (gidx >> 2)] = wbits;
        }
      }
    }
  }
}

// Conv1d(3, 64) + ReLU with 16-bit output and sign-bit map, the shape every network here starts
// with (models/pointnet.py:17, :86, :268): the instruction-lean form of first_layer_kernel.  A warp
// takes 32 consecutive points per super-trip: their 96 coordinates arrive with three coalesced loads
// and are re-read from shared memory (immediate offsets), the eight lanes of a point own eight
// channels each and write one contiguous 128-byte row; every address of the unrolled 8 x 4-point
// loop is a constant offset from one per-thread base (the general kernel spent 118 instructions per
// point and thread, 57 % issue-slot utilisation at 2.7 TB/s; this one about 50).
template <bool kBf16, bool kRelu>
__global__ void __launch_bounds__(256, 4) first_layer64_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, int64_t rows,
                                                               uint16_t* __restrict__ out, uint32_t* __restrict__ bits) {
  __shared__ float sx[8][96];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gidx = lane & 7, psel = lane >> 3, cg = gidx * 8;
  float wr[8][3], br[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    br[i] = bias ? bias[cg + i] : 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) wr[i][k] = w[(cg + i) * 3 + k];
  }
  const int64_t tiles = (rows + 31) >> 5, total = rows * 3;
  const float* sp = &sx[warp][psel * 3];
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * 8 + warp; tile < tiles; tile += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t p0 = tile << 5, f0 = p0 * 3;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 3; ++j) sx[warp][lane + 32 * j] = f0 + lane + 32 * j < total ? __ldg(x + f0 + lane + 32 * j) : 0.f;
    __syncwarp();
    uint16_t* op = out + (p0 + psel) * 64 + cg;
    uint32_t* bp = bits ? bits + (p0 + psel) * 2 + (gidx >> 2) : nullptr;
    const int64_t left = rows - p0 - psel;              // this thread's point t * 4 exists while t * 4 < left
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float x0 = sp[t * 12], x1 = sp[t * 12 + 1], x2 = sp[t * 12 + 2];
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a0 = fmaf(x2, wr[2 * i][2], fmaf(x1, wr[2 * i][1], fmaf(x0, wr[2 * i][0], br[2 * i])));
        const float a1 = fmaf(x2, wr[2 * i + 1][2], fmaf(x1, wr[2 * i + 1][1], fmaf(x0, wr[2 * i + 1][0], br[2 * i + 1])));
        pk[i] = kBf16 ? pack_bf16x2(a0, a1) : pack_f16x2_sat(a0, a1);
        if (kRelu) pk[i] = relu_packed<kBf16>(pk[i]);
      }
      const bool ok = t * 4 < left;
      if (ok) *reinterpret_cast<uint4*>(op + t * 256) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if (bits) {
        uint32_t wbits = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t m;
          if (kBf16) asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(pk[i]), "r"(0u));
          else asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(pk[i]), "r"(0u));
          wbits |= m & (0x00010001u << ((gidx & 3) * 4 + i));
        }
        wbits |= __shfl_xor_sync(0xffffffffu, wbits, 1);
        wbits |= __shfl_xor_sync(0xffffffffu, wbits, 2);
        if (ok && (gidx & 3) == 0) bp[t * 8] = wbits;
      }
    }
  }
}

// First-layer weight gradient (K <= 4): dw[c, k] = scale * sum_r dz[r, c] * x[r, k], HBM-bound.
// Same mapping as first_layer_kernel: 8 threads per point, each owning 8 channels; per-thread
// register accumulators over a grid-stride range of points, then shuffle + shared-memory
// reduction and one atomicAdd per (CTA, element).
template <int K>
__global__ void __launch_bounds__(256) first_layer_wgrad_kernel(const pcadv_wgrad_args a) {
  __shared__ float red[8][32][8 * (K + 1)];          // [warp][channel group][8 x (K weights + bias)]
  const int groups = a.n >> 3;                       // threads per point (divides 32)
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int grp = static_cast<int>(tid % groups);
  const int cg = grp * 8;
  const float* x = reinterpret_cast<const float*>(a.seg[0].ptr);
  float acc[8][K + 1];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k <= K; ++k) acc[i][k] = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x / groups;
  // kU rows per trip with every load issued before the first FMA: the kernel is latency-bound (one
  // 16-byte load per row and thread), so the loads in flight per thread are what sets its rate
  constexpr int kU = 4;
  auto load_dz = [&](int64_t r, float (&dz)[8]) {
    if (a.dz_dtype == PCADV_F32) {
      const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.dz) + r * a.ld_dz + cg);
      const float4 u = p[0], v = p[1];
      dz[0] = u.x; dz[1] = u.y; dz[2] = u.z; dz[3] = u.w; dz[4] = v.x; dz[5] = v.y; dz[6] = v.z; dz[7] = v.w;
    } else {
      const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.dz) + r * a.ld_dz + cg));
      const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f;
        if (a.dz_dtype == PCADV_F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
        else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
        dz[2 * e] = f.x; dz[2 * e + 1] = f.y;
      }
    }
  };
  int64_t r = tid / groups;
  for (; r + (kU - 1) * stride < a.rows; r += kU * stride) {
    float xv[kU][K];
    uint4 raw[kU];
    if (a.dz_dtype != PCADV_F32) {
#pragma unroll
      for (int u = 0; u < kU; ++u)
        raw[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.dz) + (r + u * stride) * a.ld_dz + cg));
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
#pragma unroll
      for (int k = 0; k < K; ++k) xv[u][k] = __ldg(x + (r + u * stride) * a.seg[0].ld + k);
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float dz[8];
      if (a.dz_dtype == PCADV_F32) {
        load_dz(r + u * stride, dz);
      } else {
        const uint32_t w4[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f;
          if (a.dz_dtype == PCADV_F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
          else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
          dz[2 * e] = f.x; dz[2 * e + 1] = f.y;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[i][k] = fmaf(dz[i], xv[u][k], acc[i][k]);
        acc[i][K] += dz[i];
      }
    }
  }
  for (; r < a.rows; r += stride) {
    float xv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) xv[k] = __ldg(x + r * a.seg[0].ld + k);
    float dz[8];
    load_dz(r, dz);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int k = 0; k < K; ++k) acc[i][k] = fmaf(dz[i], xv[k], acc[i][k]);
      acc[i][K] += dz[i];
    }
  }
  // lanes that share a channel group are `groups` apart
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k <= K; ++k) {
      float v = acc[i][k];
      for (int o = 16; o >= groups; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[i][k] = v;
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < groups) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k <= K; ++k) red[warp][lane][i * (K + 1) + k] = acc[i][k];
  }
  __syncthreads();
  const float sc = a.scale ? *a.scale : 1.f;
  for (int e = threadIdx.x; e < groups * 8 * (K + 1); e += 256) {
    const int gq = e / (8 * (K + 1)), rem = e % (8 * (K + 1));
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][gq][rem];
    const int c = gq * 8 + rem / (K + 1), k = rem % (K + 1);
    if (k < K) { if (a.dw) atomicAdd(&a.dw[static_cast<int64_t>(c) * a.ld_dw + k], v * sc); }
    else if (a.dbias) atomicAdd(&a.dbias[c], v * sc);
  }
}

// =====================================================================================
// Thin GEMM for the per-cloud layers (rows = clouds, a few hundred at most; K up to a few
// thousand): the fold of the tiled global feature into a per-cloud bias (models/pointnet.py
// :304-309), the T-Net / classifier heads.  One CTA = an 8 x 8 output tile with the whole K
// range split over its 256 threads (consecutive lanes on consecutive k: coalesced operand
// reads), then a shuffle + shared-memory reduction -- enough CTAs to fill the GPU where the
// 128 x 64 tiling of simt_linear_kernel would launch a handful.
// =====================================================================================
constexpr int TR = 8, TC = 8;

__global__ void __launch_bounds__(256) thin_linear_kernel(const pcadv_linear_args a) {
  __shared__ float red[8][TR * TC];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * TR;
  const int col0 = blockIdx.y * TC;
  float acc[TR][TC];
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;
  int koff = 0;
  for (int s = 0; s < a.num_seg; ++s) {
    const pcadv_seg sg = a.seg[s];
    for (int k = t; k < sg.k; k += 256) {
      float av[TR], wv[TC];
#pragma unroll
      for (int i = 0; i < TR; ++i)
        av[i] = row0 + i < a.rows ? ld_as_float(sg.ptr, (row0 + i) * sg.ld + k, sg.dtype) : 0.f;
#pragma unroll
      for (int j = 0; j < TC; ++j)
        wv[j] = col0 + j < a.n ? ld_as_float(a.w, static_cast<int64_t>(col0 + j) * a.ldw + koff + k, a.w_dtype) : 0.f;
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    koff += sg.k;
  }
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      float v = acc[i][j];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][i * TC + j] = v;
    }
  __syncthreads();
  if (t < TR * TC) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][t];
    const int64_t r = row0 + t / TC;
    const int c = col0 + t % TC;
    if (r < a.rows && c < a.n) {
      if (a.bias) v += a.bias[c];
      v = apply_act(v, a.act, a.slope);
      if (a.out_scale) v *= *a.out_scale;
      st_from_float(a.out, r * a.ld_out + c, a.out_dtype, v);
    }
  }
}

}  // namespace

static bool first_layer_wgrad_eligible(const pcadv_wgrad_args& a) {
  const int esz = a.dz_dtype == PCADV_F32 ? 4 : 2;
  return a.num_seg == 1 && a.seg[0].k >= 1 && a.seg[0].k <= 4 && a.seg[0].dtype == PCADV_F32 &&
         a.n % 8 == 0 && a.n <= 256 && 32 % (a.n / 8) == 0 && !a.dgroup_bias &&
         (a.ld_dz * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(a.dz) & 15) == 0 && a.rows >= 1024;
}

static bool first_layer_eligible(const pcadv_linear_args& a) {
  const int esz = a.out_dtype == PCADV_F32 ? 4 : 2;
  return a.num_seg == 1 && a.seg[0].k >= 1 && a.seg[0].k <= 4 && a.seg[0].dtype == PCADV_F32 &&
         a.n % 8 == 0 && a.n <= 256 && 256 % (a.n / 8) == 0 && (!a.bits_out || (a.n % 32 == 0 && a.out_dtype != PCADV_F32)) &&
         !a.mask_bits && a.out && !a.group_bias && !a.addend && !a.mask &&
         !a.out_scale && !a.colmax_key && !a.rowmax_key && (a.ld_out * esz) % 16 == 0 &&
         (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && a.rows >= 1024;
}

int simt_linear(const pcadv_linear_args& a, cudaStream_t s) {
  PCADV_CHECK_ARG((!a.bits_out && !a.mask_bits) || first_layer_eligible(a),
                  "pcadv_linear: on the CUDA-core engine only the first-layer kernel writes sign bits");
  if (first_layer_eligible(a) && a.seg[0].k == 3 && a.seg[0].ld == 3 && a.n == 64 && a.ld_out == 64 &&
      a.out_dtype != PCADV_F32 && a.w_dtype == PCADV_F32 && a.ldw == 3 && (a.act == PCADV_ACT_RELU || a.act == PCADV_ACT_NONE) &&
      (!a.bits_out || a.ld_bits_out == 2)) {
    const int64_t tiles = (a.rows + 31) / 32;
    int64_t blocks = (tiles + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    const float* x = reinterpret_cast<const float*>(a.seg[0].ptr);
    const float* w = reinterpret_cast<const float*>(a.w);
    uint16_t* o = reinterpret_cast<uint16_t*>(a.out);
    const unsigned g = static_cast<unsigned>(blocks);
    const bool bf = a.out_dtype == PCADV_BF16, relu = a.act == PCADV_ACT_RELU;
    if (bf && relu) first_layer64_kernel<true, true><<<g, 256, 0, s>>>(x, w, a.bias, a.rows, o, a.bits_out);
    else if (bf) first_layer64_kernel<true, false><<<g, 256, 0, s>>>(x, w, a.bias, a.rows, o, a.bits_out);
    else if (relu) first_layer64_kernel<false, true><<<g, 256, 0, s>>>(x, w, a.bias, a.rows, o, a.bits_out);
    else first_layer64_kernel<false, false><<<g, 256, 0, s>>>(x, w, a.bias, a.rows, o, a.bits_out);
    PCADV_LAUNCHED();
    return 0;
  }
  if (first_layer_eligible(a)) {
    const int64_t threads = a.rows * (a.n / 8);
    int64_t blocks = (threads + 255) / 256;
    const int64_t cap = 148 * 16;
    if (blocks > cap) blocks = cap;
    // n/8 divides the block size, so every thread keeps its channel group across the stride loop
    switch (a.seg[0].k) {
      case 1: first_layer_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, s>>>(a); break;
      case 2: first_layer_kernel<2><<<static_cast<unsigned>(blocks), 256, 0, s>>>(a); break;
      case 3: first_layer_kernel<3><<<static_cast<unsigned>(blocks), 256, 0, s>>>(a); break;
      default: first_layer_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, s>>>(a); break;
    }
    PCADV_LAUNCHED();
    return 0;
  }
  int ktot = 0;
  for (int i = 0; i < a.num_seg; ++i) ktot += a.seg[i].k;
  if (a.rows <= 1024 && ktot >= 1024 && a.out && !a.group_bias && !a.addend && !a.mask && !a.colmax_key &&
      !a.rowmax_key) {
    dim3 grid(static_cast<unsigned>((a.rows + TR - 1) / TR), (a.n + TC - 1) / TC);
    thin_linear_kernel<<<grid, 256, 0, s>>>(a);
    PCADV_LAUNCHED();
    return 0;
  }
  const int64_t tiles_m = (a.rows + BM - 1) / BM;
  PCADV_CHECK_ARG(tiles_m <= 0x7fffffffLL, "pcadv_linear: too many rows");
  dim3 grid(static_cast<unsigned>(tiles_m), (a.n + BN - 1) / BN);
  simt_linear_kernel<<<grid, 256, 0, s>>>(a);
  PCADV_LAUNCHED();
  return 0;
}

int launch_group_colsum(const void* dz, int dz_dtype, int64_t ld, int64_t rows, int n,
                        int64_t rows_per_group, float* out, cudaStream_t s) {
  PCADV_CHECK_ARG(rows % rows_per_group == 0, "rows (%lld) not a multiple of rows_per_group (%lld)",
                  (long long)rows, (long long)rows_per_group);
  dim3 grid((n + 63) / 64, static_cast<unsigned>(rows / rows_per_group));
  if (dz_dtype != PCADV_F32 && n % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0)
    group_colsum16_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint16_t*>(dz), dz_dtype == PCADV_BF16 ? 1 : 0,
                                               ld, n, rows_per_group, out);
  else
    group_colsum_kernel<<<grid, 256, 0, s>>>(dz, dz_dtype, ld, n, rows_per_group, out);
  PCADV_LAUNCHED();
  return 0;
}

int launch_colsum(const void* dz, int dz_dtype, int64_t ld, int64_t rows, int n, const float* scale,
                  float* out, cudaStream_t s) {
  int64_t splits = (rows + 4095) / 4096;
  if (splits > 1024) splits = 1024;
  const int64_t rps = (rows + splits - 1) / splits;
  dim3 grid((n + 63) / 64, static_cast<unsigned>((rows + rps - 1) / rps));
  colsum_kernel<<<grid, 256, 0, s>>>(dz, dz_dtype, ld, rows, n, rps, scale, out);
  PCADV_LAUNCHED();
  return 0;
}

int simt_wgrad(const pcadv_wgrad_args& a, cudaStream_t s) {
  if ((a.dw || a.dbias) && first_layer_wgrad_eligible(a)) {
    const unsigned blocks = 148 * 3;      // 80 registers x 256 threads: three CTAs per SM, one wave
    switch (a.seg[0].k) {
      case 1: first_layer_wgrad_kernel<1><<<blocks, 256, 0, s>>>(a); break;
      case 2: first_layer_wgrad_kernel<2><<<blocks, 256, 0, s>>>(a); break;
      case 3: first_layer_wgrad_kernel<3><<<blocks, 256, 0, s>>>(a); break;
      default: first_layer_wgrad_kernel<4><<<blocks, 256, 0, s>>>(a); break;
    }
    PCADV_LAUNCHED();
    return 0;
  }
  if (a.dw) {
    int koff = 0;
    for (int i = 0; i < a.num_seg; ++i) {
      const int tiles = ((a.n + WT - 1) / WT) * ((a.seg[i].k + WT - 1) / WT);
      int64_t splits = (148 * 8 + tiles - 1) / tiles;
      const int64_t max_splits = (a.rows + 255) / 256;
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
      int64_t rps = (a.rows + splits - 1) / splits;
      rps = (rps + WR - 1) / WR * WR;
      dim3 grid((a.n + WT - 1) / WT, (a.seg[i].k + WT - 1) / WT,
                static_cast<unsigned>((a.rows + rps - 1) / rps));
      simt_wgrad_kernel<<<grid, 256, 0, s>>>(a, i, koff, rps);
      PCADV_LAUNCHED();
      koff += a.seg[i].k;
    }
  } else if (a.dbias) {
    int rc = launch_colsum(a.dz, a.dz_dtype, a.ld_dz, a.rows, a.n, a.scale, a.dbias, s);
    if (rc) return rc;
  }
  if (a.dgroup_bias) {
    int rc = launch_group_colsum(a.dz, a.dz_dtype, a.ld_dz, a.rows, a.n, a.rows_per_group,
                                 a.dgroup_bias, s);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace pcadv
