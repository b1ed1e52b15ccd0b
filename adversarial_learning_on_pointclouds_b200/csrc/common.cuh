// Shared helpers for libpcadv.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/pcadv.h"

namespace pcadv {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define PCADV_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      pcadv::set_error(__VA_ARGS__);          \
      return 1;                               \
    }                                         \
  } while (0)

#define PCADV_CUDA_OK(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      pcadv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                       __FILE__, __LINE__);                                          \
      return 2;                                                                      \
    }                                                                                \
  } while (0)

#define PCADV_LAUNCHED()                                                             \
  do {                                                                               \
    pcadv::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      pcadv::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                       __FILE__, __LINE__);                                          \
      return 3;                                                                      \
    }                                                                                \
  } while (0)

// ---- dtype-generic scalar access -------------------------------------------------
__device__ __forceinline__ float ld_as_float(const void* p, int64_t i, int dtype) {
  if (dtype == PCADV_F32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == PCADV_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

__device__ __forceinline__ void st_from_float(void* p, int64_t i, int dtype, float v) {
  if (dtype == PCADV_F32) {
    reinterpret_cast<float*>(p)[i] = v;
  } else if (dtype == PCADV_F16) {
    // saturate instead of producing inf: fp16 overflows at 65504
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  } else {
    reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  }
}

// two floats -> one packed 16-bit pair (lo in the low half); fp16 saturates to +-65504 instead
// of producing inf (one F2FP.SATFINITE instruction)
// ReLU on a packed pair: max(x, +0) per half.  Same bits as packing fmaxf(x, 0.f) for every finite
// or infinite x: rounding and saturation are monotone and keep the sign, and -0 becomes +0 either way.
template <bool kBf16>
__device__ __forceinline__ uint32_t relu_packed(uint32_t pk) {
  uint32_t r;
  if (kBf16) asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(pk), "r"(0u));
  else asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(pk), "r"(0u));
  return r;
}

__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// 8 consecutive elements starting at element offset `off` as floats (vector loads when aligned)
__device__ __forceinline__ void load8(const void* p, int64_t off, int dtype, float (&o)[8]) {
  if (dtype == PCADV_F32) {
    const float* q = reinterpret_cast<const float*>(p) + off;
    if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
      const float4 a = reinterpret_cast<const float4*>(q)[0], b = reinterpret_cast<const float4*>(q)[1];
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = q[j];
    }
    return;
  }
  const uint16_t* q = reinterpret_cast<const uint16_t*>(p) + off;
  if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
    const uint4 t4 = *reinterpret_cast<const uint4*>(q);
    const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f;
      if (dtype == PCADV_F16) f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
      else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
      o[2 * e] = f.x; o[2 * e + 1] = f.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = ld_as_float(p, off + j, dtype);
  }
}

// 8 consecutive elements starting at element offset `off`: one / two 16-byte stores when the
// address allows it, scalar stores otherwise
__device__ __forceinline__ void store8(void* p, int64_t off, int dtype, const float (&o)[8]) {
  if (dtype == PCADV_F32) {
    float* q = reinterpret_cast<float*>(p) + off;
    if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
      reinterpret_cast<float4*>(q)[0] = make_float4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<float4*>(q)[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) q[j] = o[j];
    }
    return;
  }
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (dtype == PCADV_F16) {
      __half2 h = __floats2half2_rn(fminf(fmaxf(o[2 * j], -65504.f), 65504.f),
                                    fminf(fmaxf(o[2 * j + 1], -65504.f), 65504.f));
      pk[j] = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  uint16_t* q = reinterpret_cast<uint16_t*>(p) + off;
  if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
    *reinterpret_cast<uint4*>(q) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = static_cast<uint16_t>((pk[j >> 1] >> ((j & 1) * 16)) & 0xffffu);
  }
}

__host__ __device__ __forceinline__ int dtype_size(int dtype) {
  return dtype == PCADV_F32 ? 4 : 2;
}

// ---- activation helpers ----------------------------------------------------------
__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == PCADV_ACT_RELU) return fmaxf(v, 0.f);
  if (act == PCADV_ACT_LEAKY) return v > 0.f ? v : v * slope;
  return v;
}
// derivative expressed through the saved OUTPUT y = act(z): ReLU' = [y > 0],
// LeakyReLU' = [y > 0] ? 1 : slope (y > 0 <=> z > 0 for slope > 0).
__device__ __forceinline__ float act_grad_from_output(float y, int act, float slope) {
  if (act == PCADV_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == PCADV_ACT_LEAKY) return y > 0.f ? 1.f : slope;
  return 1.f;
}

// ---- packed (value, first index) max keys ----------------------------------------
// Order-preserving map float -> uint32, then (bits << 32) | ~index so that a
// 64-bit unsigned max picks the largest value and, among equals, the smallest
// index (torch.max's first-occurrence rule).  Key 0 is below every finite value.
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned long long pack_key(float v, uint32_t idx) {
  return (static_cast<unsigned long long>(float_to_ordered(v)) << 32) |
         static_cast<unsigned long long>(0xffffffffu - idx);
}
__device__ __forceinline__ float key_value(unsigned long long k) {
  return ordered_to_float(static_cast<uint32_t>(k >> 32));
}
__device__ __forceinline__ uint32_t key_index(unsigned long long k) {
  return 0xffffffffu - static_cast<uint32_t>(k & 0xffffffffull);
}

}  // namespace pcadv
