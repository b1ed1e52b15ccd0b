// Part-IoU evaluation on the device: the per-cloud intersection / union counts
// of utils/metric.py:20-39 and the argmax + accuracy count of utils/trainer.py:100-110, which the
// reference computes with numpy on the host after copying every prediction back.
//
//   pcadv_part_counts   per cloud: pred = argmax_c logits[p, c] (first maximum), then
//                       counts[g][0][l] = #(pred == l & gt == l), [1][l] = #(pred == l),
//                       [2][l] = #(gt == l), correct[g] = #(pred == gt).  Integer work, HBM-bound:
//                       4*C + 8 bytes per point read, nothing written but the counters.
//   pcadv_part_iou      per cloud: category = argmax of the one-hot row, parts [begin[cat], begin[cat+1]),
//                       iou = mean_l (both empty ? 1 : inter / union) in float64, in the order numpy adds.
#include "common.cuh"

namespace pcadv {
namespace {

constexpr int kPcRows = 128;                 // points per tile (one per thread)
constexpr int kPcMaxC = 64;

// grid (chunks, groups), 128 threads.  Each CTA walks tiles chunk, chunk + gridDim.x, ... of its cloud.
__global__ void __launch_bounds__(kPcRows) part_counts_kernel(
    const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ pred_in,
    const int64_t* __restrict__ labels, int64_t N, int C, int* __restrict__ counts,
    int* __restrict__ correct, int64_t* __restrict__ pred_out) {
  extern __shared__ float tile[];                        // [128][pitch], pitch odd -> conflict-free row scans
  __shared__ int hist[3 * kPcMaxC + 1];
  const int t = threadIdx.x, g = blockIdx.y;
  const bool flat8 = logits != nullptr && ld == C && (C & 1) == 0 &&
                     (reinterpret_cast<uintptr_t>(logits) & 7) == 0;
  const int pitch = flat8 ? C : (C | 1);
  for (int e = t; e < 3 * kPcMaxC + 1; e += kPcRows) hist[e] = 0;
  const int64_t tiles = (N + kPcRows - 1) / kPcRows;
  for (int64_t tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
    const int64_t n0 = tl * kPcRows;
    const int rows = N - n0 < kPcRows ? static_cast<int>(N - n0) : kPcRows;
    const int64_t p0 = static_cast<int64_t>(g) * N + n0;
    int pred = 0;
    __syncthreads();                                     // hist zeroed / previous tile scanned
    if (logits != nullptr) {
      const float* src = logits + p0 * ld;
      if (flat8) {
        // contiguous rows, 8-byte aligned: the tile is one flat copy, every 8-byte piece in flight
        // at once (cp.async), unpadded rows (pitch C: a 2-way bank conflict in the scan below)
        const int pieces = rows * C / 2;
        const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
        for (int e = t; e < pieces; e += kPcRows)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + e * 8), "l"(src + 2 * e) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else if (ld == C) {                              // contiguous rows: flat coalesced copy
        const int total = rows * C;
        int r = t / C, c = t - r * C;                    // walk (r, c) without a division per element
        const int dr = kPcRows / C, dc = kPcRows - dr * C;
        for (int e = t; e < total; e += kPcRows) {
          tile[r * pitch + c] = src[e];
          r += dr; c += dc;
          if (c >= C) { c -= C; ++r; }
        }
      } else {
        for (int r = 0; r < rows; ++r)
          for (int c = t; c < C; c += kPcRows) tile[r * pitch + c] = src[r * ld + c];
      }
      __syncthreads();
      if (t < rows) {
        const float* row = tile + t * pitch;
        float best = row[0];
        for (int c = 1; c < C; ++c) {
          const float v = row[c];
          // torch.max: the first maximum; a NaN wins over any number
          if (v > best || (v != v && best == best)) { best = v; pred = c; }
        }
      }
    } else if (t < rows) {
      pred = static_cast<int>(pred_in[p0 + t]);
    }
    const bool live = t < rows;
    int lab = -1;
    if (live) {
      lab = static_cast<int>(labels[p0 + t]);
      if (pred_out != nullptr) pred_out[p0 + t] = pred;
    }
    // warp-aggregated histogram updates: real clouds have 2-6 parts, so most lanes collide
    const unsigned act = __ballot_sync(0xffffffffu, live);
    if (live) {
      const int lane = t & 31;
      const bool pv = pred >= 0 && pred < C, lv = lab >= 0 && lab < C;
      const unsigned mp = __match_any_sync(act, pv ? pred : -1);
      if (pv && (__ffs(mp) - 1) == lane) atomicAdd(&hist[kPcMaxC + pred], __popc(mp));
      const unsigned ml = __match_any_sync(act, lv ? lab : -1);
      if (lv && (__ffs(ml) - 1) == lane) atomicAdd(&hist[2 * kPcMaxC + lab], __popc(ml));
      const bool hit = pv && pred == lab;
      const unsigned mi = __match_any_sync(act, hit ? pred : -1);
      if (hit && (__ffs(mi) - 1) == lane) {
        atomicAdd(&hist[pred], __popc(mi));
      }
      const unsigned mh = __ballot_sync(act, hit);
      if ((__ffs(act) - 1) == lane && mh) atomicAdd(&hist[3 * kPcMaxC], __popc(mh));
    }
  }
  __syncthreads();
  int* out = counts + static_cast<int64_t>(g) * 3 * C;
  for (int e = t; e < 3 * C; e += kPcRows) {
    const int k = e / C, c = e - k * C;
    const int v = hist[k * kPcMaxC + c];
    if (v) atomicAdd(&out[e], v);
  }
  if (t == 0 && hist[3 * kPcMaxC]) atomicAdd(&correct[g], hist[3 * kPcMaxC]);
}

// one thread per cloud
__global__ void part_iou_kernel(const int* __restrict__ counts, const float* __restrict__ onehot,
                                int64_t ld_onehot, int ncat, const int* __restrict__ part_begin,
                                int groups, int C, double* __restrict__ iou, int* __restrict__ cat_out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  const float* oh = onehot + g * ld_onehot;
  int cat = 0;
  float best = oh[0];
  for (int c = 1; c < ncat; ++c)
    if (oh[c] > best) { best = oh[c]; cat = c; }         // np.argmax: the first maximum
  const int* cnt = counts + static_cast<int64_t>(g) * 3 * C;
  const int lo = part_begin[cat], hi = part_begin[cat + 1];
  double sum = 0.0;
  for (int l = lo; l < hi; ++l) {
    const int inter = cnt[l], np_ = cnt[C + l], ng = cnt[2 * C + l];
    const double v = (np_ == 0 && ng == 0) ? 1.0
                                           : static_cast<double>(inter) / static_cast<double>(np_ + ng - inter);
    sum += v;
  }
  iou[g] = sum / static_cast<double>(hi - lo);
  cat_out[g] = cat;
}

}  // namespace
}  // namespace pcadv

extern "C" int pcadv_part_counts(const float* logits, int64_t ld, const int64_t* pred_in,
                                 const int64_t* labels, int32_t groups, int64_t N, int32_t C,
                                 int32_t* counts, int32_t* correct, int64_t* pred_out, void* stream) {
  using namespace pcadv;
  PCADV_CHECK_ARG((logits != nullptr) != (pred_in != nullptr),
                  "part_counts: give either logits or predictions");
  PCADV_CHECK_ARG(labels && counts && correct, "part_counts: null pointer");
  PCADV_CHECK_ARG(C >= 1 && C <= kPcMaxC, "part_counts: C=%d outside [1, %d]", C, kPcMaxC);
  PCADV_CHECK_ARG(groups >= 0 && N >= 0 && groups <= 65535, "part_counts: groups=%d, N=%lld", groups,
                  static_cast<long long>(N));
  PCADV_CHECK_ARG(logits == nullptr || ld >= C, "part_counts: ld=%lld < C=%d", static_cast<long long>(ld), C);
  PCADV_CHECK_ARG(N < (1ll << 31), "part_counts: N=%lld does not fit the int32 counters",
                  static_cast<long long>(N));
  if (groups == 0 || N == 0) return 0;
  const int64_t tiles = (N + kPcRows - 1) / kPcRows;
  // enough CTAs to fill the machine a few times over, never more than the tiles of a cloud
  int64_t chunks = (148 * 8 + groups - 1) / groups;
  if (chunks > tiles) chunks = tiles;
  if (chunks < 1) chunks = 1;
  const size_t smem = logits ? static_cast<size_t>(kPcRows) * (C | 1) * sizeof(float) : 0;
  part_counts_kernel<<<dim3(static_cast<unsigned>(chunks), groups), kPcRows, smem,
                       static_cast<cudaStream_t>(stream)>>>(logits, ld, pred_in, labels, N, C, counts,
                                                            correct, pred_out);
  PCADV_LAUNCHED();
  return 0;
}

extern "C" int pcadv_part_iou(const int32_t* counts, const float* onehot, int64_t ld_onehot, int32_t ncat,
                              const int32_t* part_begin, int32_t groups, int32_t C, double* iou,
                              int32_t* cat, void* stream) {
  using namespace pcadv;
  PCADV_CHECK_ARG(counts && onehot && part_begin && iou && cat, "part_iou: null pointer");
  PCADV_CHECK_ARG(ncat >= 1 && ld_onehot >= ncat && C >= 1 && groups >= 0, "part_iou: bad sizes");
  if (groups == 0) return 0;
  part_iou_kernel<<<(groups + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      counts, onehot, ld_onehot, ncat, part_begin, groups, C, iou, cat);
  PCADV_LAUNCHED();
  return 0;
}
