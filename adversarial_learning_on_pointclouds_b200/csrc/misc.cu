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

// ---- fast sparse max-pool backward (k % 64 == 0, k <= 1024, rows_per_group <= 8192) -------
// pair loads: lane owns elements 2*lane + 64*j
__device__ __forceinline__ float2 ld_pair(const void* p, int64_t i, int dtype) {
  if (dtype == PCADV_F32) return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(p) + i);
  if (dtype == PCADV_F16) return __half22float2(*reinterpret_cast<const __half2*>(reinterpret_cast<const __half*>(p) + i));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(p) + i));
}
__device__ __forceinline__ void st_pair(void* p, int64_t i, int dtype, float2 v) {
  if (dtype == PCADV_F32) {
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(p) + i) = v;
  } else if (dtype == PCADV_F16) {
    *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(p) + i) =
        __floats2half2_rn(fminf(fmaxf(v.x, -65504.f), 65504.f), fminf(fmaxf(v.y, -65504.f), 65504.f));
  } else {
    *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = __floats2bfloat162_rn(v.x, v.y);
  }
}

constexpr int kMaxPairs = 16;     // k <= 1024

// dW / dbias: one CTA per channel c, 8 warps stride over the clouds; no atomics.
__global__ void __launch_bounds__(256) maxbwd_dw_fast_kernel(const pcadv_maxbwd_args a) {
  __shared__ float red[8][1024];
  __shared__ float bred[8];
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pairs = a.k >> 6;
  float2 acc[kMaxPairs];
#pragma unroll
  for (int j = 0; j < kMaxPairs; ++j) acc[j] = make_float2(0.f, 0.f);
  float bsum = 0.f;
  // four clouds per trip: the (dz, argmax row) lookups and then 4 x pairs row gathers are
  // independent, so the dependent-load latency is paid once per four clouds
  for (int g0 = warp; g0 < a.groups; g0 += 32) {
    float dz[4];
    int64_t r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = g0 + 8 * u;
      dz[u] = 0.f;
      r[u] = 0;
      if (g < a.groups) {
        const int64_t gc = static_cast<int64_t>(g) * a.n + c;
        dz[u] = a.dg[gc] * act_grad_from_output(a.gval[gc], a.act, a.slope);
        r[u] = static_cast<int64_t>(g) * a.rows_per_group + a.idx[gc];
      }
      bsum += dz[u];
    }
    if (a.dw) {
#pragma unroll
      for (int j = 0; j < kMaxPairs; ++j) {
        if (j < pairs) {
          float2 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            x[u] = dz[u] != 0.f ? ld_pair(a.x, r[u] * a.ldx + 2 * lane + 64 * j, a.x_dtype)
                                : make_float2(0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[j].x = fmaf(dz[u], x[u].x, acc[j].x);
            acc[j].y = fmaf(dz[u], x[u].y, acc[j].y);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kMaxPairs; ++j) {
    if (j < pairs) {
      red[warp][2 * lane + 64 * j] = acc[j].x;
      red[warp][2 * lane + 64 * j + 1] = acc[j].y;
    }
  }
  if (lane == 0) bred[warp] = bsum;
  __syncthreads();
  const float sc = a.scale ? *a.scale : 1.f;
  if (a.dw) {
    for (int k = threadIdx.x; k < a.k; k += 256) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][k];
      a.dw[static_cast<int64_t>(c) * a.ld_dw + k] += s * sc;
    }
  }
  if (threadIdx.x == 0 && a.dbias) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += bred[w];
    a.dbias[c] += s * sc;
  }
}

// per-cloud workspace of the in-place path, in 32-bit words (see maxbwd_rows_kernel)
__host__ __device__ __forceinline__ int64_t maxbwd_ws_entry_offset(int64_t N, int64_t n) {
  return (N + 3 * n + 3) / 4 * 4;
}
__host__ __device__ __forceinline__ int64_t maxbwd_ws_ints(int64_t N, int64_t n) {
  return maxbwd_ws_entry_offset(N, n) + 4 * n;
}

// dz of the previous layer: one CTA per cloud.  Channels are bucketed by their argmax row
// in shared memory (count, scan, fill); each warp then sums one touched row in fp32 and
// adds it once into dz_inout through the previous layer's activation mask.
__global__ void __launch_bounds__(256) maxbwd_rows_kernel(const pcadv_maxbwd_args a) {
  extern __shared__ int sm_i[];
  int* ends = sm_i;                                        // [rows_per_group]
  int* list = sm_i + a.rows_per_group;                     // [n] channels ordered by row
  float* dzv = reinterpret_cast<float*>(list + a.n);       // [n]
  int* rowl = list + 2 * a.n;                              // [n] argmax row of every list entry
  __shared__ int part[256];
  const int g = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int N = static_cast<int>(a.rows_per_group);
  for (int r = t; r < N; r += 256) ends[r] = 0;
  __syncthreads();
  for (int c = t; c < a.n; c += 256) {
    const int64_t gc = static_cast<int64_t>(g) * a.n + c;
    const float dz = a.dg[gc] * act_grad_from_output(a.gval[gc], a.act, a.slope);
    dzv[c] = dz;
    if (dz != 0.f) atomicAdd(&ends[a.idx[gc]], 1);
  }
  __syncthreads();
  // exclusive scan of the counts: thread t owns a contiguous chunk of rows
  const int chunk = (N + 255) / 256;
  const int lo = t * chunk, hi = lo + chunk < N ? lo + chunk : N;
  int s = 0;
  for (int r = lo; r < hi; ++r) s += ends[r];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int i = 0; i < 256; ++i) { const int v = part[i]; part[i] = run; run += v; }
  }
  __syncthreads();
  int run = part[t];
  for (int r = lo; r < hi; ++r) { const int v = ends[r]; ends[r] = run; run += v; }
  __syncthreads();
  // fill: ends[r] walks from the row's start to its end
  for (int c = t; c < a.n; c += 256) {
    if (dzv[c] != 0.f) {
      const int row = a.idx[static_cast<int64_t>(g) * a.n + c];
      const int pos = atomicAdd(&ends[row], 1);
      list[pos] = c;
      rowl[pos] = row;
    }
  }
  __syncthreads();
  // publish the bucket table per cloud: [ends (N) | list (n) | dz (n) | row of list entry (n) | pad |
  // packed entries (n x int4 {row | -1, channel, dz bits, 0}, 16-byte aligned)]: the 16-bit apply
  // kernel reads one packed entry per list position instead of three dependent words
  int* ws = reinterpret_cast<int*>(a.workspace) + static_cast<int64_t>(g) * maxbwd_ws_ints(N, a.n);
  const int cnt = ends[N - 1];
  for (int r = t; r < N; r += 256) ws[r] = ends[r];
  int4* ent = reinterpret_cast<int4*>(ws + maxbwd_ws_entry_offset(N, a.n));
  for (int c = t; c < a.n; c += 256) {
    ws[N + c] = list[c];
    reinterpret_cast<float*>(ws + N + a.n)[c] = dzv[c];
    ws[N + 2 * a.n + c] = c < cnt ? rowl[c] : -1;
    const int ch = c < cnt ? list[c] : 0;
    ent[c] = make_int4(c < cnt ? rowl[c] : -1, ch, c < cnt ? __float_as_int(dzv[ch]) : 0, 0);
  }
}

// one warp per row of the whole batch; rows no channel points at leave after two loads
__global__ void __launch_bounds__(256) maxbwd_rows_apply_kernel(const pcadv_maxbwd_args a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = static_cast<int>(a.rows_per_group);
  const int64_t grow = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (grow >= static_cast<int64_t>(a.groups) * N) return;
  const int g = static_cast<int>(grow / N), r = static_cast<int>(grow - static_cast<int64_t>(g) * N);
  const int* ws = reinterpret_cast<const int*>(a.workspace) + static_cast<int64_t>(g) * maxbwd_ws_ints(N, a.n);
  const int* ends = ws;
  const int* list = ws + N;
  const float* dzv = reinterpret_cast<const float*>(ws + N + a.n);
  const int pairs = a.k >> 6;
  {
    const int beg = r == 0 ? 0 : ends[r - 1], end = ends[r];
    if (beg == end) return;
    float2 acc[kMaxPairs];
#pragma unroll
    for (int j = 0; j < kMaxPairs; ++j) acc[j] = make_float2(0.f, 0.f);
    // a point that is the argmax of hundreds of channels makes this loop long: four
    // channels per trip keep 4 x pairs independent gathers in flight
    int q = beg;
    for (; q + 4 <= end; q += 4) {
      int c[4];
      float dz[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { c[u] = list[q + u]; dz[u] = dzv[c[u]]; }
#pragma unroll
      for (int j = 0; j < kMaxPairs; ++j) {
        if (j < pairs) {
          float2 w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            w[u] = ld_pair(a.w, static_cast<int64_t>(c[u]) * a.ldw + 2 * lane + 64 * j, a.w_dtype);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[j].x = fmaf(dz[u], w[u].x, acc[j].x);
            acc[j].y = fmaf(dz[u], w[u].y, acc[j].y);
          }
        }
      }
    }
    for (; q < end; ++q) {
      const int c = list[q];
      const float dz = dzv[c];
#pragma unroll
      for (int j = 0; j < kMaxPairs; ++j) {
        if (j < pairs) {
          const float2 w = ld_pair(a.w, static_cast<int64_t>(c) * a.ldw + 2 * lane + 64 * j, a.w_dtype);
          acc[j].x = fmaf(dz, w.x, acc[j].x);
          acc[j].y = fmaf(dz, w.y, acc[j].y);
        }
      }
    }
    const int64_t row = static_cast<int64_t>(g) * a.rows_per_group + r;
#pragma unroll
    for (int j = 0; j < kMaxPairs; ++j) {
      if (j < pairs) {
        const int64_t kk = 2 * lane + 64 * j;
        const float2 x = ld_pair(a.x, row * a.ldx + kk, a.x_dtype);
        float2 d = ld_pair(a.dz_inout, row * a.ld_dz + kk, a.dz_dtype);
        d.x += acc[j].x * act_grad_from_output(x.x, a.prev_act, a.prev_slope);
        d.y += acc[j].y * act_grad_from_output(x.y, a.prev_act, a.prev_slope);
        st_pair(a.dz_inout, row * a.ld_dz + kk, a.dz_dtype, d);
      }
    }
  }
}

// ---- 16-bit fast path of the apply step ----------------------------------------------------
// Work item = kSegEntries consecutive entries of a cloud's row-sorted channel list; a warp owns
// every run (equal argmax row) that STARTS inside its segment and follows it to its end, so no
// row is touched by two warps and no atomics are needed, while a point that is the argmax of
// hundreds of channels costs one warp a long run instead of serialising a whole cloud.  A lane
// owns 8-element vectors lane + 32 j of the k-wide rows (16-byte gathers of w6 rows from L2).
constexpr int kMaxVec = 4;        // k <= 1024
constexpr int kSegEntries = 8;

template <bool kBf16>
__device__ __forceinline__ void fma8(const uint4 raw, float dz, float (&acc)[8]) {
  const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float2 f;
    if (kBf16) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
    else f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
    acc[2 * e] = fmaf(dz, f.x, acc[2 * e]);
    acc[2 * e + 1] = fmaf(dz, f.y, acc[2 * e + 1]);
  }
}

template <bool kBf16, int KV>
__global__ void __launch_bounds__(256) maxbwd_rows_apply16_kernel(const pcadv_maxbwd_args a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = static_cast<int>(a.rows_per_group);
  const int segs = (a.n + kSegEntries - 1) / kSegEntries;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (wid >= static_cast<int64_t>(a.groups) * segs) return;
  const int g = static_cast<int>(wid / segs);
  const int seg = static_cast<int>(wid - static_cast<int64_t>(g) * segs);
  const int* ws = reinterpret_cast<const int*>(a.workspace) + static_cast<int64_t>(g) * maxbwd_ws_ints(N, a.n);
  const int4* ent = reinterpret_cast<const int4*>(ws + maxbwd_ws_entry_offset(N, a.n));   // row -1 beyond the last entry
  // one round trip for the segment's rows and its left neighbour: lanes 0..7 the entries, lane 8
  // the entry before the segment; a run starts where the row differs from the entry before it
  const int q0 = seg * kSegEntries;
  int myrow = -1;
  if (lane < kSegEntries) myrow = q0 + lane < a.n ? __ldg(&ent[q0 + lane].x) : -1;
  else if (lane == kSegEntries) myrow = q0 > 0 ? __ldg(&ent[q0 - 1].x) : -2;
  const int left = __shfl_sync(0xffffffffu, myrow, lane == 0 ? kSegEntries : (lane - 1) & 31);
  const unsigned valid = __ballot_sync(0xffffffffu, lane < kSegEntries && myrow >= 0);
  const unsigned starts = __ballot_sync(0xffffffffu, lane < kSegEntries && myrow >= 0 && myrow != left);
  if (starts == 0u) return;                      // nothing starts here (or the segment is empty)
  int q = q0 + __ffs(starts) - 1;
  const int seg_end = q0 + __popc(valid);        // valid entries are a prefix of the segment
  const int cnt = a.n;                           // windows below stop at row -1
  const int nvec = a.k >> 3;
  const uint4* wbase = reinterpret_cast<const uint4*>(a.w);
  const int64_t ldw4 = a.ldw >> 3;
  while (q < seg_end) {
    const int row = __shfl_sync(0xffffffffu, myrow, q - q0);      // the run starts inside the segment
    float acc[KV][8];
#pragma unroll
    for (int j = 0; j < KV; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
    // the touched row of x (mask) and dz: fetched now, used after the run's gathers
    const int64_t grow = static_cast<int64_t>(g) * a.rows_per_group + row;
    const uint4* xrow = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(a.x) + grow * a.ldx);
    uint4* drow = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(a.dz_inout) + grow * a.ld_dz);
    uint4 xr[KV], dr[KV];
#pragma unroll
    for (int j = 0; j < KV; ++j) {
      const int v = lane + 32 * j;
      if (v < nvec) { xr[j] = xrow[v]; dr[j] = drow[v]; }
    }
    bool more = true;
    while (more) {
      // the next 32 list entries, one packed entry per lane; the run is a prefix of the window
      const int mq = q + lane;
      int4 m = make_int4(-1, 0, 0, 0);
      if (mq < cnt) m = __ldg(ent + mq);
      const int mc = m.y;
      const float mdz = __int_as_float(m.z);
      const unsigned same = __ballot_sync(0xffffffffu, m.x == row);
      const int len = same == 0xffffffffu ? 32 : __ffs(~same) - 1;
      more = len == 32;
      for (int i = 0; i < len; i += 4) {               // four gathers in flight
        int cs[4];
        float ds[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int iu = i + u < len ? i + u : i;
          cs[u] = __shfl_sync(0xffffffffu, mc, iu);
          const float d = __shfl_sync(0xffffffffu, mdz, iu);
          ds[u] = i + u < len ? d : 0.f;
        }
#pragma unroll
        for (int j = 0; j < KV; ++j) {
          const int v = lane + 32 * j;
          if (v < nvec) {
            uint4 rw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) rw[u] = __ldg(wbase + static_cast<int64_t>(cs[u]) * ldw4 + v);
#pragma unroll
            for (int u = 0; u < 4; ++u) fma8<kBf16>(rw[u], ds[u], acc[j]);
          }
        }
      }
      q += len;
    }
    // one read-modify-write of the touched row, through the previous layer's activation mask
#pragma unroll
    for (int j = 0; j < KV; ++j) {
      const int v = lane + 32 * j;
      if (v < nvec) {
        const uint32_t x4[4] = {xr[j].x, xr[j].y, xr[j].z, xr[j].w};
        const uint32_t d4[4] = {dr[j].x, dr[j].y, dr[j].z, dr[j].w};
        uint32_t o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 xf, df;
          if (kBf16) {
            xf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&x4[e]));
            df = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&d4[e]));
          } else {
            xf = __half22float2(*reinterpret_cast<const __half2*>(&x4[e]));
            df = __half22float2(*reinterpret_cast<const __half2*>(&d4[e]));
          }
          df.x += acc[j][2 * e] * act_grad_from_output(xf.x, a.prev_act, a.prev_slope);
          df.y += acc[j][2 * e + 1] * act_grad_from_output(xf.y, a.prev_act, a.prev_slope);
          o4[e] = kBf16 ? pack_bf16x2(df.x, df.y) : pack_f16x2_sat(df.x, df.y);
        }
        drow[v] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
}

template <bool kBf16>
void launch_apply16(const pcadv_maxbwd_args& a, unsigned blocks, cudaStream_t s) {
  const int kv = (a.k / 8 + 31) / 32;             // 16-byte vectors per lane
  if (kv <= 1) maxbwd_rows_apply16_kernel<kBf16, 1><<<blocks, 256, 0, s>>>(a);
  else if (kv == 2) maxbwd_rows_apply16_kernel<kBf16, 2><<<blocks, 256, 0, s>>>(a);
  else if (kv == 3) maxbwd_rows_apply16_kernel<kBf16, 3><<<blocks, 256, 0, s>>>(a);
  else maxbwd_rows_apply16_kernel<kBf16, 4><<<blocks, 256, 0, s>>>(a);
}

// dW / dbias, 16-bit x rows gathered with 16-byte loads: one CTA per channel c, 8 warps stride
// over the clouds (four clouds per trip), no atomics.
// kV = 16-byte vector columns per lane (k <= 256 kV: 1 / 2 / 4), so that k = 512 keeps 16 accumulators,
// not 32, and three CTAs stay resident with eight gathers in flight per lane
template <bool kBf16, int kV>
__global__ void __launch_bounds__(256) maxbwd_dw16_kernel(const pcadv_maxbwd_args a) {
  __shared__ float red[8][1024];
  __shared__ float bred[8];
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = a.k >> 3;
  float acc[kV][8];
#pragma unroll
  for (int j = 0; j < kV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  float bsum = 0.f;
  const uint4* xbase = reinterpret_cast<const uint4*>(a.x);
  const int64_t ldx4 = a.ldx >> 3;
  // the (dz, argmax row) of a trip's four clouds are fetched one trip ahead, so that the row gathers
  // of this trip and the metadata of the next are in flight together
  float ndz[4];
  int64_t nr[4];
  auto fetch = [&](int g0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = g0 + 8 * u;
      ndz[u] = 0.f;
      nr[u] = 0;
      if (g < a.groups) {
        const int64_t gc = static_cast<int64_t>(g) * a.n + c;
        ndz[u] = __ldg(a.dg + gc) * act_grad_from_output(__ldg(a.gval + gc), a.act, a.slope);
        nr[u] = static_cast<int64_t>(g) * a.rows_per_group + __ldg(a.idx + gc);
      }
    }
  };
  fetch(warp);
  for (int g0 = warp; g0 < a.groups; g0 += 32) {
    float dz[4];
    int64_t r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { dz[u] = ndz[u]; r[u] = nr[u]; bsum += dz[u]; }
    if (g0 + 32 < a.groups) fetch(g0 + 32);
    if (a.dw) {
      // the row gathers of two vector columns (eight 16-byte loads per lane) are issued before the first
      // FMA: the kernel is latency-bound (22 cycles per issued instruction at 31 % issue utilisation)
#pragma unroll
      for (int j0 = 0; j0 < kV; j0 += 2) {
        uint4 x[2][4];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int v = lane + 32 * (j0 + jj);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            x[jj][u] = (j0 + jj < kV && v < nvec && dz[u] != 0.f) ? __ldg(xbase + r[u] * ldx4 + v)
                                                                      : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          if (j0 + jj < kV && lane + 32 * (j0 + jj) < nvec) {
#pragma unroll
            for (int u = 0; u < 4; ++u) fma8<kBf16>(x[jj][u], dz[u], acc[j0 + jj]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kV; ++j) {
    const int v = lane + 32 * j;
    if (v < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp][8 * v + e] = acc[j][e];
    }
  }
  if (lane == 0) bred[warp] = bsum;
  __syncthreads();
  const float sc = a.scale ? *a.scale : 1.f;
  if (a.dw) {
    for (int k = threadIdx.x; k < a.k; k += 256) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][k];
      a.dw[static_cast<int64_t>(c) * a.ld_dw + k] += s * sc;
    }
  }
  if (threadIdx.x == 0 && a.dbias) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += bred[w];
    a.dbias[c] += s * sc;
  }
}

__global__ void rowmax_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ val,
                                  const int32_t* __restrict__ idx, int64_t rows, int n, int act,
                                  float slope, const float* scale, void* dz, int64_t ld_dz,
                                  int dz_dtype) {
  // one thread = 8 consecutive columns of one row (n % 8 == 0 is checked by the caller)
  const int per_row = n >> 3;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * per_row) return;
  const int64_t r = i / per_row;
  const int c0 = static_cast<int>(i - r * per_row) << 3;
  const int hit = idx[r] - c0;                     // 0..7 when the row's argmax is in this group
  float v = 0.f;
  if (hit >= 0 && hit < 8) v = dy[r] * (scale ? *scale : 1.f) * act_grad_from_output(val[r], act, slope);
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = (j == hit) ? v : 0.f;
  store8(dz, r * ld_dz + c0, dz_dtype, o);
}

// Backward of "layer + activation + max over channels" (models/discriminator.py:67-72) without
// the dense one-hot dz: per row only channel idx[r] carries gradient s[r] = dy[r] * act'(val[r]).
//   dgrad:  dz_prev[r, :] = prev_act'(y_prev[r, :]) * (scale * s[r]) * w[idx[r], :]
__global__ void rowmax_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ val,
                                    const int32_t* __restrict__ idx, int64_t rows, int k, int act,
                                    float slope, const float* scale, const void* w, int64_t ldw,
                                    int w_dtype, const void* yprev, int64_t ld_y, int y_dtype,
                                    int prev_act, float prev_slope, void* dz, int64_t ld_dz,
                                    int dz_dtype) {
  const int per_row = k >> 3;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * per_row) return;
  const int64_t r = i / per_row;
  const int c0 = static_cast<int>(i - r * per_row) << 3;
  const float s = dy[r] * (scale ? *scale : 1.f) * act_grad_from_output(val[r], act, slope);
  float wv[8], yv[8], o[8];
  load8(w, static_cast<int64_t>(idx[r]) * ldw + c0, w_dtype, wv);
  load8(yprev, r * ld_y + c0, y_dtype, yv);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = s * wv[j] * act_grad_from_output(yv[j], prev_act, prev_slope);
  store8(dz, r * ld_dz + c0, dz_dtype, o);
}

//   wgrad:  dw[idx[r], :] += s[r] * y_prev[r, :],  dbias[idx[r]] += s[r]   (fp32, no scaling
//   needed).  Per-CTA shared-memory accumulators [n x k] take the scatter; one global atomicAdd
//   per (CTA, element) at the end.
__global__ void __launch_bounds__(256) rowmax_wgrad_kernel(const float* __restrict__ dy,
                                                          const float* __restrict__ val,
                                                          const int32_t* __restrict__ idx,
                                                          int64_t rows, int n, int k, int act,
                                                          float slope, const void* yprev, int64_t ld_y,
                                                          int y_dtype, float* dw, int64_t ld_dw,
                                                          float* dbias) {
  extern __shared__ float acc_s[];                       // [n * k] + [n]
  float* bias_s = acc_s + static_cast<size_t>(n) * k;
  for (int e = threadIdx.x; e < n * k + n; e += blockDim.x) acc_s[e] = 0.f;
  __syncthreads();
  const int per_row = k >> 3;
  const int64_t total = rows * per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / per_row;
    const int c0 = static_cast<int>(i - r * per_row) << 3;
    const float s = dy[r] * act_grad_from_output(val[r], act, slope);
    if (s == 0.f) continue;
    const int ch = idx[r];
    float yv[8];
    load8(yprev, r * ld_y + c0, y_dtype, yv);
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&acc_s[ch * k + c0 + j], s * yv[j]);
    if (c0 == 0) atomicAdd(&bias_s[ch], s);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * k; e += blockDim.x) {
    const float v = acc_s[e];
    if (v != 0.f && dw) atomicAdd(&dw[static_cast<int64_t>(e / k) * ld_dw + (e % k)], v);
  }
  if (dbias)
    for (int e = threadIdx.x; e < n; e += blockDim.x)
      if (bias_s[e] != 0.f) atomicAdd(&dbias[e], bias_s[e]);
}

// Sorted variant for 16-bit y rows (k in {64, 128, 256}): every CTA takes 2048-row chunks, buckets
// the rows of a chunk by their argmax channel in shared memory (count, scan, fill -- integer
// atomics only), then the eight warps walk equal slices of the channel-sorted row list with the
// running channel's sums in registers; a shared-memory float add happens only when the channel
// changes (<= n + 8 times per chunk instead of once per row and column).
constexpr int kRwChunk = 2048;

template <bool kBf16>
__global__ void __launch_bounds__(256) rowmax_wgrad_sorted_kernel(
    const float* __restrict__ dy, const float* __restrict__ val, const int32_t* __restrict__ idx,
    int64_t rows, int n, int k, int act, float slope, const uint16_t* __restrict__ yprev, int64_t ld_y,
    float* dw, int64_t ld_dw, float* dbias) {
  extern __shared__ float acc_s[];                       // [n * k] + [n] + chunk tables
  float* bias_s = acc_s + static_cast<size_t>(n) * k;
  float* s_s = bias_s + n;                               // [kRwChunk]
  int* cnt = reinterpret_cast<int*>(s_s + kRwChunk);     // [n + 1] starts after the scan
  int* fillp = cnt + n + 1;                              // [n]
  uint16_t* ch_s = reinterpret_cast<uint16_t*>(fillp + n);   // [kRwChunk]
  uint16_t* list_s = ch_s + kRwChunk;                    // [kRwChunk] rows ordered by channel
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int epl = k >> 5;                                // y elements per lane: 2, 4 or 8
  for (int e = t; e < n * k + n; e += 256) acc_s[e] = 0.f;
  const int64_t nchunks = (rows + kRwChunk - 1) / kRwChunk;
  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int64_t base = chunk * kRwChunk;
    const int m = rows - base < kRwChunk ? static_cast<int>(rows - base) : kRwChunk;
    __syncthreads();                                     // previous chunk's tables are free
    for (int c = t; c <= n; c += 256) cnt[c] = 0;
    __syncthreads();
    for (int i = t; i < m; i += 256) {
      const float sv = dy[base + i] * act_grad_from_output(val[base + i], act, slope);
      const int ch = idx[base + i];
      s_s[i] = sv;
      ch_s[i] = static_cast<uint16_t>(ch);
      if (sv != 0.f) atomicAdd(&cnt[ch], 1);
    }
    __syncthreads();
    if (warp == 0) {                                     // exclusive scan of the n counts
      const int per = (n + 31) / 32;
      int local = 0;
      for (int c = lane * per; c < (lane + 1) * per && c < n; ++c) local += cnt[c];
      int incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int run = incl - local;
      for (int c = lane * per; c < (lane + 1) * per && c < n; ++c) {
        const int v = cnt[c];
        cnt[c] = run;
        fillp[c] = run;
        run += v;
      }
      if (lane == 31) cnt[n] = incl;                     // total
    }
    __syncthreads();
    for (int i = t; i < m; i += 256)
      if (s_s[i] != 0.f) list_s[atomicAdd(&fillp[ch_s[i]], 1)] = static_cast<uint16_t>(i);
    __syncthreads();
    const int total = cnt[n];
    const int q0 = static_cast<int>(static_cast<int64_t>(total) * warp / 8);
    const int q1 = static_cast<int>(static_cast<int64_t>(total) * (warp + 1) / 8);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    float bsum = 0.f;
    int cur = -1;
    auto flush = [&]() {
      if (cur >= 0) {
        float* dst = acc_s + static_cast<size_t>(cur) * k + lane * epl;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (e < epl) atomicAdd(dst + e, acc[e]);
        if (lane == 0) atomicAdd(&bias_s[cur], bsum);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      bsum = 0.f;
    };
    for (int q = q0; q < q1; q += 4) {
      int ri[4], chs[4];
      float sv[4];
      uint4 yv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = q + u < q1;
        ri[u] = ok ? list_s[q + u] : 0;
        chs[u] = ok ? ch_s[ri[u]] : -1;
        sv[u] = ok ? s_s[ri[u]] : 0.f;
        yv[u] = make_uint4(0u, 0u, 0u, 0u);
        if (ok) {
          const uint16_t* yp = yprev + (base + ri[u]) * ld_y + lane * epl;
          if (epl == 2) yv[u].x = *reinterpret_cast<const uint32_t*>(yp);
          else if (epl == 4) { const uint2 v2 = *reinterpret_cast<const uint2*>(yp); yv[u].x = v2.x; yv[u].y = v2.y; }
          else yv[u] = *reinterpret_cast<const uint4*>(yp);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (chs[u] < 0) continue;
        if (chs[u] != cur) { flush(); cur = chs[u]; }
        const uint32_t w4[4] = {yv[u].x, yv[u].y, yv[u].z, yv[u].w};
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          if (2 * e2 < epl) {
            float2 f;
            if (kBf16) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e2]));
            else f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e2]));
            acc[2 * e2] = fmaf(sv[u], f.x, acc[2 * e2]);
            acc[2 * e2 + 1] = fmaf(sv[u], f.y, acc[2 * e2 + 1]);
          }
        }
        bsum += sv[u];
      }
    }
    flush();
  }
  __syncthreads();
  for (int e = t; e < n * k; e += 256) {
    const float v = acc_s[e];
    if (v != 0.f && dw) atomicAdd(&dw[static_cast<int64_t>(e / k) * ld_dw + (e % k)], v);
  }
  if (dbias)
    for (int e = t; e < n; e += 256)
      if (bias_s[e] != 0.f) atomicAdd(&dbias[e], bias_s[e]);
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
  // one thread = 8 consecutive output columns of one row (cols_pad % 8 == 0: 16-byte stores);
  // the generic tail handles any other width
  const float sc = scale ? *scale : 1.f;
  if ((cols_pad & 7) == 0) {
    const int per_row = cols_pad >> 3;
    const int64_t total = rows * per_row;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t r = i / per_row;
      const int c0 = static_cast<int>(i - r * per_row) << 3;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        float v = c < cols ? ld_as_float(src, r * ld_src + c, src_dtype) * sc : 0.f;
        if (mask && c < cols)
          v *= act_grad_from_output(ld_as_float(mask, r * ld_mask + c, mask_dtype), mask_act, mask_slope);
        o[j] = v;
      }
      store8(dst, r * ld_dst + c0, dst_dtype, o);
    }
    return;
  }
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

// channel-major B x C x N fp32 (what torch's softmax / log_softmax / their backward produce)
// -> point-major [B*N, cols_pad] with scale, conversion and zero padding; a 32 x 32 tile
// transpose through shared memory keeps both sides coalesced.
__global__ void __launch_bounds__(256) convert_cm_kernel(const float* __restrict__ src,
                                                        int64_t batch_stride, int64_t chan_stride,
                                                        int64_t rows_per_group, int cols, void* dst,
                                                        int dst_dtype, int64_t ld_dst, int cols_pad,
                                                        const float* scale) {
  __shared__ float tile[32][33];
  const float sc = scale ? *scale : 1.f;
  const int64_t b = blockIdx.z;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * 32;     // point within the cloud
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc;
    const int64_t i = i0 + tx;
    tile[cc][tx] = (c < cols && i < rows_per_group)
                       ? src[b * batch_stride + c * chan_stride + i] * sc : 0.f;
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int64_t i = i0 + rr;
    const int c = c0 + tx;
    if (i < rows_per_group && c < cols_pad)
      st_from_float(dst, (b * rows_per_group + i) * ld_dst + c, dst_dtype, tile[tx][rr]);
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
  auto pair_ok = [](const void* p, int64_t ld, int dtype) {
    return p == nullptr || (ld % 2 == 0 && (reinterpret_cast<uintptr_t>(p) & (dtype == PCADV_F32 ? 7 : 3)) == 0);
  };
  const bool fast = a->k % 64 == 0 && a->k <= 64 * kMaxPairs && pair_ok(a->x, a->ldx, a->x_dtype) &&
                    pair_ok(a->w, a->ldw, a->w_dtype) && pair_ok(a->dz_inout, a->ld_dz, a->dz_dtype);
  if (a->dw || a->dbias) {
    PCADV_CHECK_ARG(!a->dw || a->x, "pcadv_maxpool_bwd: dw needs x");
    const bool x16 = a->x && a->x_dtype != PCADV_F32 && a->k % 8 == 0 && a->k <= 8 * 32 * kMaxVec &&
                     a->ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(a->x) & 15) == 0;
    const int kv16 = (a->k / 8 + 31) / 32;               // 16-byte vector columns per lane
    if (x16 && a->x_dtype == PCADV_BF16) {
      if (kv16 <= 1) maxbwd_dw16_kernel<true, 1><<<a->n, 256, 0, s>>>(*a);
      else if (kv16 == 2) maxbwd_dw16_kernel<true, 2><<<a->n, 256, 0, s>>>(*a);
      else maxbwd_dw16_kernel<true, 4><<<a->n, 256, 0, s>>>(*a);
    } else if (x16) {
      if (kv16 <= 1) maxbwd_dw16_kernel<false, 1><<<a->n, 256, 0, s>>>(*a);
      else if (kv16 == 2) maxbwd_dw16_kernel<false, 2><<<a->n, 256, 0, s>>>(*a);
      else maxbwd_dw16_kernel<false, 4><<<a->n, 256, 0, s>>>(*a);
    }
    else if (fast) maxbwd_dw_fast_kernel<<<a->n, 256, 0, s>>>(*a);
    else maxbwd_dw_kernel<<<a->n, 128, 0, s>>>(*a);
    PCADV_LAUNCHED();
  }
  if (a->dz_inout) {
    PCADV_CHECK_ARG(a->w && a->x, "pcadv_maxpool_bwd: dz_inout needs w and x");
    PCADV_CHECK_ARG(fast && a->rows_per_group <= 8192 && a->n <= 4096,
                    "pcadv_maxpool_bwd: dz_inout needs k %% 64 == 0, k <= 1024, rows_per_group <= 8192, "
                    "n <= 4096 (got k=%d rpg=%lld n=%d)", a->k, (long long)a->rows_per_group, a->n);
    const size_t smem = (static_cast<size_t>(a->rows_per_group) + 3 * static_cast<size_t>(a->n)) * 4;
    static bool attr_done = false;
    if (!attr_done) {
      PCADV_CUDA_OK(cudaFuncSetAttribute(maxbwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (8192 + 3 * 4096) * 4));
      attr_done = true;
    }
    PCADV_CHECK_ARG(a->workspace != nullptr, "pcadv_maxpool_bwd: dz_inout needs a workspace");
    maxbwd_rows_kernel<<<a->groups, 256, smem, s>>>(*a);
    PCADV_LAUNCHED();
    auto vec_ok = [](const void* p, int64_t ld) {
      return ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    };
    const bool all16 = a->x_dtype != PCADV_F32 && a->w_dtype == a->x_dtype && a->dz_dtype == a->x_dtype &&
                       a->k % 8 == 0 && vec_ok(a->x, a->ldx) && vec_ok(a->w, a->ldw) &&
                       vec_ok(a->dz_inout, a->ld_dz);
    if (all16) {
      const int64_t items = static_cast<int64_t>(a->groups) * ((a->n + kSegEntries - 1) / kSegEntries);
      const unsigned blocks = static_cast<unsigned>((items + 7) / 8);
      if (a->x_dtype == PCADV_BF16) launch_apply16<true>(*a, blocks, s);
      else launch_apply16<false>(*a, blocks, s);
    } else {
      const int64_t total_rows = static_cast<int64_t>(a->groups) * a->rows_per_group;
      maxbwd_rows_apply_kernel<<<static_cast<unsigned>((total_rows + 7) / 8), 256, 0, s>>>(*a);
    }
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
  PCADV_CHECK_ARG(dy && val && idx && dz && rows >= 0 && n > 0 && n % 8 == 0,
                  "pcadv_rowmax_bwd: bad args (n must be a multiple of 8)");
  if (rows == 0) return 0;
  const int64_t total = rows * (n / 8);
  rowmax_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                      static_cast<cudaStream_t>(stream)>>>(dy, val, idx, rows, n, act, slope, scale,
                                                           dz, ld_dz, dz_dtype);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_rowmax_dgrad(const float* dy, const float* val, const int32_t* idx, int64_t rows,
                                  int32_t k, int32_t act, float slope, const float* scale,
                                  const void* w, int64_t ldw, int32_t w_dtype, const void* yprev,
                                  int64_t ld_y, int32_t y_dtype, int32_t prev_act, float prev_slope,
                                  void* dz, int64_t ld_dz, int32_t dz_dtype, void* stream) {
  PCADV_CHECK_ARG(dy && val && idx && w && yprev && dz && rows >= 0 && k > 0 && k % 8 == 0,
                  "pcadv_rowmax_dgrad: bad args (k must be a multiple of 8)");
  if (rows == 0) return 0;
  const int64_t total = rows * (k / 8);
  rowmax_dgrad_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(dy, val, idx, rows, k, act, slope, scale, w,
                                                             ldw, w_dtype, yprev, ld_y, y_dtype,
                                                             prev_act, prev_slope, dz, ld_dz, dz_dtype);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_rowmax_wgrad(const float* dy, const float* val, const int32_t* idx, int64_t rows,
                                  int32_t n, int32_t k, int32_t act, float slope, const void* yprev,
                                  int64_t ld_y, int32_t y_dtype, float* dw, int64_t ld_dw,
                                  float* dbias, void* stream) {
  PCADV_CHECK_ARG(dy && val && idx && yprev && rows >= 0 && n > 0 && k > 0 && k % 8 == 0,
                  "pcadv_rowmax_wgrad: bad args (k must be a multiple of 8)");
  const size_t smem = (static_cast<size_t>(n) * k + n) * sizeof(float);
  PCADV_CHECK_ARG(smem <= 200 * 1024, "pcadv_rowmax_wgrad: n * k too large (%d x %d)", n, k);
  if (rows == 0) return 0;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    PCADV_CUDA_OK(cudaFuncSetAttribute(rowmax_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
    attr = 200 * 1024;
  }
  if (y_dtype != PCADV_F32 && (k == 64 || k == 128 || k == 256) && n <= 1024 && ld_y % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(yprev) & 15) == 0) {
    const size_t smem2 = smem + kRwChunk * sizeof(float) + (2 * static_cast<size_t>(n) + 1) * sizeof(int) +
                         2 * kRwChunk * sizeof(uint16_t) + 16;
    if (smem2 <= 200 * 1024) {
      static bool attr2 = false;
      if (!attr2) {
        PCADV_CUDA_OK(cudaFuncSetAttribute(rowmax_wgrad_sorted_kernel<false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        PCADV_CUDA_OK(cudaFuncSetAttribute(rowmax_wgrad_sorted_kernel<true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr2 = true;
      }
      int64_t blocks2 = (rows + kRwChunk - 1) / kRwChunk;
      const int64_t per_sm = (200 * 1024) / static_cast<int64_t>(smem2 + 1024);
      const int64_t cap2 = 148 * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
      if (blocks2 > cap2) blocks2 = cap2;
      const uint16_t* y16 = reinterpret_cast<const uint16_t*>(yprev);
      if (y_dtype == PCADV_BF16)
        rowmax_wgrad_sorted_kernel<true><<<static_cast<unsigned>(blocks2), 256, smem2, static_cast<cudaStream_t>(stream)>>>(
            dy, val, idx, rows, n, k, act, slope, y16, ld_y, dw, ld_dw, dbias);
      else
        rowmax_wgrad_sorted_kernel<false><<<static_cast<unsigned>(blocks2), 256, smem2, static_cast<cudaStream_t>(stream)>>>(
            dy, val, idx, rows, n, k, act, slope, y16, ld_y, dw, ld_dw, dbias);
      PCADV_LAUNCHED();
      return 0;
    }
  }
  const int64_t total = rows * (k / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  rowmax_wgrad_kernel<<<static_cast<unsigned>(blocks), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      dy, val, idx, rows, n, k, act, slope, yprev, ld_y, y_dtype, dw, ld_dw, dbias);
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
  const int64_t work = (cols_pad & 7) == 0 ? rows * (cols_pad / 8) : rows * cols_pad;
  convert_kernel<<<grid_for(work, 256, 148 * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, src_dtype, ld_src, rows, cols, dst, dst_dtype, ld_dst, cols_pad, scale, mask, ld_mask,
      mask_dtype, mask_act, mask_slope);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_convert_cm(const float* src, int64_t batch_stride, int64_t chan_stride,
                                int64_t groups, int64_t rows_per_group, int32_t cols, void* dst,
                                int32_t dst_dtype, int64_t ld_dst, int32_t cols_pad,
                                const float* scale, void* stream) {
  PCADV_CHECK_ARG(src && dst && groups >= 0 && rows_per_group > 0 && cols > 0 && cols_pad >= cols,
                  "pcadv_convert_cm: bad args");
  PCADV_CHECK_ARG(groups <= 65535, "pcadv_convert_cm: too many clouds (%lld)", (long long)groups);
  if (groups == 0) return 0;
  dim3 grid(static_cast<unsigned>((rows_per_group + 31) / 32), (cols_pad + 31) / 32,
            static_cast<unsigned>(groups));
  convert_cm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, batch_stride, chan_stride, rows_per_group, cols, dst, dst_dtype, ld_dst, cols_pad, scale);
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
