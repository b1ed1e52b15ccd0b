// Tensor-core weight gradient on CTA PAIRS (tcgen05.mma.cta_group::2) for layers with 256 dz
// channels and a wide K-concat (fc1: dz 256 x x 960).
//
// One cluster of two CTAs owns a dW tile [256 dz channels x 512 x channels] over a row range: CTA r
// holds dz channels 128 r .. 128 r + 127 (the M half, A operand, MN-major) and HALF of the tile's
// x boxes (the N half, B operand); the leader CTA issues M = 256 MMAs that read both CTAs' shared
// memory and write both CTAs' tensor memory.  Compared with the single-CTA kernel every x box is
// fetched by one SM of the pair instead of two (L2 -> SM traffic 4.9 -> 2.9 KB / point for fc1) and
// a stage is 48 KB instead of 80 KB, so the ring has four stages instead of two.
//
// Protocol (per stage): both producers wait for their own `empty` barrier, the leader's producer
// posts arrive.expect_tx for BOTH CTAs' bytes on its `full` barrier, both issue TMA loads that
// complete_tx on the leader's `full` (cta_group::2 form); the leader's MMA thread waits `full`,
// issues the MMAs and commits with a multicast arrive to both CTAs' `empty`.
//
// The per-cloud column sums of dz (the gradient of fc1's per-cloud bias) are NOT taken from the
// boxes in flight as in the single-CTA kernel: the peer CTA would have to be told that its boxes
// have landed (a cross-CTA arrive per stage) before its sum warps may run and release the stage;
// measured, that hand-off costs 0.28 ms on fc1 against 0.1 ms for a separate pass over dz
// (group_colsum16_kernel), so the host launches that pass instead.
#include <stdlib.h>
#include "tc_pipeline.cuh"

namespace pcadv {
namespace tc {

constexpr int kW2Threads = 192;
constexpr int kW2Stages = 4;
constexpr int kW2BoxBytes = 8192;                  // [64 rows][64 ch]
constexpr int kW2BoxesB = 4;                       // x boxes per CTA per stage (8 per pair = 512 columns)
constexpr int kW2StageBytes = (2 + kW2BoxesB) * kW2BoxBytes;   // 48 KB
constexpr int kW2SmemMax = 232448;

struct W2Tail {
  uint64_t full[kW2Stages];       // leader only: 1 arrival + both CTAs' bytes
  uint64_t empty[kW2Stages];      // per CTA: the leader's multicast MMA commit
  uint64_t tmem_full;
  uint32_t tmem_base;
};

struct W2Params {
  int64_t rows;
  int num_seg;
  int seg_k[PCADV_MAX_SEG];
  int seg_koff[PCADV_MAX_SEG];
  int seg_box0[PCADV_MAX_SEG + 1];
  int boxes;                      // real 64-channel boxes of the K-concat
  int tiles_n;                    // tiles of 8 boxes
  int splits;
  int64_t rows_per_split;         // multiple of 64
  int bf16;
  float* dw;
  int64_t ld_dw;
  int vec_red;
  const float* scale;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// TMA load of the 2-CTA form: lands in the executing CTA's shared memory, signals `mbar_cluster`
// (a shared::cluster address, here the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive (once) on the barrier at the same offset in both CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void red_add_v4_w2(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
// tile box j (0..7) of the pair's N tile -> which CTA holds it and where: the two N = 256 MMAs read
// local boxes {0, 1} and {2, 3} of both CTAs, CTA 0's columns first
__device__ __forceinline__ int tile_box_of(int rank, int local) { return (local >> 1) * 4 + rank * 2 + (local & 1); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kW2Threads, 1)
tc_wgrad_pair_kernel(const __grid_constant__ TensorMaps maps, const W2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  W2Tail* st = reinterpret_cast<W2Tail*>(stages + kW2Stages * kW2StageBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  // one work item per pair (host guarantees tiles_n * splits <= pairs)
  const int num_work = p.tiles_n * p.splits;
  const bool has_work = pair < num_work;
  const int tn = pair / p.splits;
  const int sp = pair - tn * p.splits;
  const int64_t r0 = sp * p.rows_per_split;
  const int64_t r1 = r0 + p.rows_per_split < p.rows ? r0 + p.rows_per_split : p.rows;
  const int nstage_iters = has_work && r1 > r0 ? static_cast<int>((r1 - r0 + 63) / 64) : 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_seg; ++s) tma_prefetch_desc(&maps.act[s]);
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < kW2Stages; ++i) {
      mbar_init(&st->full[i], 1);
      mbar_init(&st->empty[i], 1);
    }
    mbar_init(&st->tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(&st->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // both CTAs' barriers exist before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = st->tmem_base;

  if (warp == 0) {
    // ---------------- TMA producer: the whole warp, lane l issues box l of the stage ----------------
    uint32_t phase = 0;
    int stage = 0;
    for (int it = 0; it < nstage_iters; ++it) {
      const int32_t rr = static_cast<int32_t>(r0 + static_cast<int64_t>(it) * 64);
      mbar_wait_backoff(&st->empty[stage], phase ^ 1);
      if (leader && lane == 0) mbar_arrive_expect_tx(&st->full[stage], 2u * kW2StageBytes);
      __syncwarp();
      const uint32_t full_leader = mapa_rank(smem_u32(&st->full[stage]), 0);
      if (lane < 2 + kW2BoxesB) {
        uint8_t* dst = stages + stage * kW2StageBytes + lane * kW2BoxBytes;
        if (lane < 2) {
          tma_load_2d_pair(dst, &maps.w, full_leader, static_cast<int32_t>(rank) * 128 + lane * 64, rr);
        } else {
          const int gb = tn * 8 + tile_box_of(static_cast<int>(rank), lane - 2);
          if (gb < p.boxes) {
            int sg = 0;
            while (gb >= p.seg_box0[sg + 1]) ++sg;
            tma_load_2d_pair(dst, &maps.act[sg], full_leader, (gb - p.seg_box0[sg]) * 64, rr);
          } else {
            // padding box of the last tile: columns past the tensor -> zero fill (still counts bytes)
            tma_load_2d_pair(dst, &maps.act[0], full_leader, 1 << 30, rr);
          }
        }
      }
      __syncwarp();
      if (++stage == kW2Stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: one thread of the leader CTA ----------------
    if (leader && lane == 0 && nstage_iters > 0) {
      const uint32_t idesc = make_idesc(256, 256, p.bf16 != 0, true, true);
      uint32_t phase = 0;
      int stage = 0;
      bool first = true;
      for (int it = 0; it < nstage_iters; ++it) {
        mbar_wait_backoff(&st->full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(stages + stage * kW2StageBytes);
        const uint32_t b_addr = a_addr + 2 * kW2BoxBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adesc = make_smem_desc(a_addr + k * 2048, kW2BoxBytes, 1024);
          const uint32_t accf = first ? 0u : 1u;
          umma_f16_pair(tmem_base, adesc, make_smem_desc(b_addr + k * 2048, kW2BoxBytes, 1024), idesc, accf);
          umma_f16_pair(tmem_base + 256, adesc,
                        make_smem_desc(b_addr + 2 * kW2BoxBytes + k * 2048, kW2BoxBytes, 1024), idesc, accf);
          first = false;
        }
        umma_commit_pair(&st->empty[stage]);
        if (++stage == kW2Stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_pair(&st->tmem_full);
    }
  } else {
    // ---------------- epilogue warps (both CTAs): this CTA's 128 x 512 half of the tile ----------------
    const int quarter = warp & 3;
    const int lane_row = quarter * 32 + lane;
    const float sc = p.scale ? *p.scale : 1.f;
    if (nstage_iters > 0) {
      mbar_wait(&st->tmem_full, 0);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const int c = static_cast<int>(rank) * 128 + lane_row;      // dz channel of this thread
      for (int b = 0; b < 8; ++b) {
        const int gb = tn * 8 + b;
        if (gb >= p.boxes) break;                     // warp-uniform
        int sg = 0;
        while (gb >= p.seg_box0[sg + 1]) ++sg;
        const int k0 = (gb - p.seg_box0[sg]) * 64;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v[32];
          tmem_ld32(taddr0 + b * 64 + hh * 32, v);
          float* dst = p.dw + static_cast<int64_t>(c) * p.ld_dw + p.seg_koff[sg] + k0 + hh * 32;
          if (p.vec_red) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4_w2(dst + j, v[j] * sc, v[j + 1] * sc, v[j + 2] * sc, v[j + 3] * sc);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j] * sc);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // nobody leaves while the peer may still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace tc

int launch_group_colsum(const void* dz, int dz_dtype, int64_t ld, int64_t rows, int n,
                        int64_t rows_per_group, float* out, cudaStream_t s);

// Returns 0 when the pair kernel took the call, -1 when the shape is not its business (the caller
// then runs the single-CTA kernel), > 0 on error.
int tc_wgrad_pair(const pcadv_wgrad_args& a, cudaStream_t s) {
  using namespace tc;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("PCADV_WGRAD_PAIR");
    enabled = (e && atoi(e) == 0) ? 0 : 1;      // PCADV_WGRAD_PAIR=0 falls back to the single-CTA kernel
  }
  if (!enabled) return -1;
  const int dt = a.dz_dtype;
  if (a.n != 256 || a.dbias != nullptr || a.dw == nullptr || a.num_seg < 1) return -1;
  int boxes = 0;
  for (int i = 0; i < a.num_seg; ++i) {
    if (a.seg[i].dtype != dt || a.seg[i].k % 64 || !tma_compatible(a.seg[i].ptr, dt, a.seg[i].ld)) return -1;
    boxes += a.seg[i].k / 64;
  }
  if (boxes < 8 || !tma_compatible(a.dz, dt, a.ld_dz)) return -1;     // narrow layers: single-CTA kernel
  if (a.dgroup_bias && (a.rows_per_group <= 0 || a.rows % a.rows_per_group != 0)) return -1;
  TensorMaps maps;
  W2Params p{};
  p.rows = a.rows; p.num_seg = a.num_seg;
  if (int rc = encode_tmap_2d(&maps.w, a.dz, dt, a.rows, a.n, a.ld_dz, 64, 64)) return rc;
  int koff = 0, b0 = 0;
  for (int i = 0; i < a.num_seg; ++i) {
    if (int rc = encode_tmap_2d(&maps.act[i], a.seg[i].ptr, dt, a.rows, a.seg[i].k, a.seg[i].ld, 64, 64)) return rc;
    p.seg_k[i] = a.seg[i].k; p.seg_koff[i] = koff; p.seg_box0[i] = b0;
    koff += a.seg[i].k; b0 += a.seg[i].k / 64;
  }
  p.seg_box0[a.num_seg] = b0;
  p.boxes = boxes;
  p.tiles_n = (boxes + 7) / 8;
  const int pairs = num_sms() / 2;
  if (p.tiles_n > pairs) return -1;
  int64_t splits = pairs / p.tiles_n;
  const int64_t max_splits = (a.rows + 511) / 512;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rps = (a.rows + splits - 1) / splits;
  rps = (rps + 63) / 64 * 64;
  p.splits = static_cast<int>((a.rows + rps - 1) / rps);
  p.rows_per_split = rps;
  p.bf16 = dt == PCADV_BF16 ? 1 : 0;
  p.dw = a.dw; p.ld_dw = a.ld_dw; p.scale = a.scale;
  p.vec_red = (a.ld_dw % 4 == 0 && (reinterpret_cast<uintptr_t>(a.dw) & 15) == 0) ? 1 : 0;

  const size_t smem = 1024 + static_cast<size_t>(kW2Stages) * kW2StageBytes + sizeof(W2Tail) + 16;
  static bool attr_done = false;
  if (!attr_done) {
    PCADV_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kW2SmemMax));
    attr_done = true;
  }
  const int grid = 2 * p.tiles_n * p.splits;            // one work item per CTA pair, one wave
  tc_wgrad_pair_kernel<<<grid, kW2Threads, smem, s>>>(maps, p);
  PCADV_LAUNCHED();
  if (a.dgroup_bias) {                                   // per-cloud column sums: a separate pass (see top)
    if (int rc = launch_group_colsum(a.dz, a.dz_dtype, a.ld_dz, a.rows, a.n, a.rows_per_group, a.dgroup_bias, s))
      return rc;
  }
  return 0;
}

}  // namespace pcadv
