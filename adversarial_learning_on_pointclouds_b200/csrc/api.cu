// C-ABI entry points: argument validation and engine dispatch.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace pcadv {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int simt_linear(const pcadv_linear_args& a, cudaStream_t s);
int simt_wgrad(const pcadv_wgrad_args& a, cudaStream_t s);
int tc_linear(const pcadv_linear_args& a, cudaStream_t s);
int tc_wgrad(const pcadv_wgrad_args& a, cudaStream_t s);

static bool valid_fdtype(int d) { return d == PCADV_F32 || d == PCADV_F16 || d == PCADV_BF16; }

}  // namespace pcadv

using namespace pcadv;

extern "C" int pcadv_version(void) { return PCADV_VERSION; }

extern "C" const char* pcadv_last_error(void) { return g_err; }

extern "C" long long pcadv_launch_count(void) { return g_launches.load(); }

extern "C" int pcadv_device_check(void) {
  int dev = 0;
  PCADV_CUDA_OK(cudaGetDevice(&dev));
  int major = 0;
  PCADV_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  PCADV_CHECK_ARG(major == 10, "libpcadv is built for sm_100a only; device %d is sm_%d*", dev, major);
  return 0;
}

extern "C" int pcadv_linear(const pcadv_linear_args* a, void* stream) {
  PCADV_CHECK_ARG(a != nullptr, "pcadv_linear: null args");
  PCADV_CHECK_ARG(a->rows >= 0 && a->n > 0, "pcadv_linear: bad shape rows=%lld n=%d",
                  (long long)a->rows, a->n);
  PCADV_CHECK_ARG(a->num_seg >= 1 && a->num_seg <= PCADV_MAX_SEG, "pcadv_linear: num_seg=%d",
                  a->num_seg);
  for (int i = 0; i < a->num_seg; ++i) {
    PCADV_CHECK_ARG(a->seg[i].ptr != nullptr && a->seg[i].k > 0 && valid_fdtype(a->seg[i].dtype),
                    "pcadv_linear: bad segment %d", i);
  }
  PCADV_CHECK_ARG(a->w != nullptr && valid_fdtype(a->w_dtype), "pcadv_linear: bad weight");
  PCADV_CHECK_ARG(!(a->group_bias || a->colmax_key) || a->rows_per_group > 0,
                  "pcadv_linear: rows_per_group required with group_bias / colmax_key");
  PCADV_CHECK_ARG(a->out || a->colmax_key || a->rowmax_key, "pcadv_linear: no output requested");
  PCADV_CHECK_ARG(!a->seg0_group_sum || (a->engine == PCADV_ENGINE_TC && a->rows_per_group > 0),
                  "pcadv_linear: seg0_group_sum needs the tensor-core engine and rows_per_group");
  PCADV_CHECK_ARG(!a->out || valid_fdtype(a->out_dtype), "pcadv_linear: bad out dtype");
  if (a->rows == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->engine == PCADV_ENGINE_TC) return tc_linear(*a, s);
  PCADV_CHECK_ARG(a->engine == PCADV_ENGINE_SIMT, "pcadv_linear: unknown engine %d", a->engine);
  return simt_linear(*a, s);
}

extern "C" int pcadv_wgrad(const pcadv_wgrad_args* a, void* stream) {
  PCADV_CHECK_ARG(a != nullptr, "pcadv_wgrad: null args");
  PCADV_CHECK_ARG(a->rows >= 0 && a->n > 0, "pcadv_wgrad: bad shape");
  PCADV_CHECK_ARG(a->dz != nullptr && valid_fdtype(a->dz_dtype), "pcadv_wgrad: bad dz");
  PCADV_CHECK_ARG(a->num_seg >= 0 && a->num_seg <= PCADV_MAX_SEG, "pcadv_wgrad: num_seg=%d",
                  a->num_seg);
  PCADV_CHECK_ARG(!a->dw || a->num_seg >= 1, "pcadv_wgrad: dw needs at least one segment");
  for (int i = 0; i < a->num_seg; ++i) {
    PCADV_CHECK_ARG(a->seg[i].ptr != nullptr && a->seg[i].k > 0 && valid_fdtype(a->seg[i].dtype),
                    "pcadv_wgrad: bad segment %d", i);
  }
  PCADV_CHECK_ARG(!a->dgroup_bias || a->rows_per_group > 0,
                  "pcadv_wgrad: rows_per_group required with dgroup_bias");
  if (a->rows == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->engine == PCADV_ENGINE_TC) return tc_wgrad(*a, s);
  PCADV_CHECK_ARG(a->engine == PCADV_ENGINE_SIMT, "pcadv_wgrad: unknown engine %d", a->engine);
  return simt_wgrad(*a, s);
}
