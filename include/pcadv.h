/*
 * pcadv.h -- C ABI of libpcadv.so, the sm_100a kernel library behind the
 * PointNet + discriminator adversarial train step.
 *
 * The reference (YiruS/Adversarial_Learning_on_PointClouds) has no FFI layer:
 * its hot path is the Python nn.Module API of models/pointnet.py and
 * models/discriminator.py, whose arithmetic is ATen library calls
 * (SURVEY.md 8b).  Each entry point below replaces one family of those calls;
 * the reference line it stands in for is cited beside it.  The only callers are
 * the torch.autograd.Function classes in
 * adversarial_learning_on_pointclouds_b200/models/ (via ctypes, see
 * INTEGRATION.md).
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every pointer is a DEVICE pointer owned by
 *     the caller (torch.empty); the library allocates nothing persistent.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*).  No
 *     entry point synchronises the device.
 *   - Return 0 on success, non-zero on failure; pcadv_last_error() returns a
 *     thread-local message.  Nothing throws across the ABI, nothing calls exit.
 *   - "rows" are points (B*N of them, point-major storage [rows, channels]).
 *     A "group" is one cloud: rows_per_group consecutive rows.
 *   - dtype codes: PCADV_F32 / PCADV_F16 / PCADV_BF16.  engine codes:
 *     PCADV_ENGINE_SIMT = fp32 FFMA CUDA-core kernels (the fp32-accumulate
 *     verification mode and the small / unaligned layers), PCADV_ENGINE_TC =
 *     tcgen05 + TMEM + TMA kernels (16-bit operands, fp32 accumulate).
 */
#ifndef PCADV_H_
#define PCADV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCADV_VERSION 100

enum { PCADV_F32 = 0, PCADV_F16 = 1, PCADV_BF16 = 2, PCADV_I32 = 3, PCADV_I64 = 4 };
enum { PCADV_ACT_NONE = 0, PCADV_ACT_RELU = 1, PCADV_ACT_LEAKY = 2 };
enum { PCADV_ENGINE_SIMT = 0, PCADV_ENGINE_TC = 1 };
#define PCADV_MAX_SEG 6

/* One K-segment of the left operand: a [rows, k] row-major matrix with leading
 * dimension `ld` (elements).  Several segments stand for the channel concat of
 * models/pointnet.py:306 (torch.cat of x1..x5) without materialising it. */
typedef struct pcadv_seg {
  const void* ptr;
  int64_t ld;
  int32_t k;
  int32_t dtype;
} pcadv_seg;

/*
 * pcadv_linear: out[r, c] = post( sum_seg sum_k seg[r, k] * w[c, koff_seg + k]
 *                                  + bias[c] + group_bias[r / rows_per_group, c]
 *                                  + addend[r, c] )
 *   post(v) = act(v) * act'(mask[r, c]) * (*out_scale)      (mask, or its 1-bit form mask_bits)
 * Replaces: Conv1d(k=1) + F.relu        models/pointnet.py:115-128, :291-301
 *           nn.Linear on B x N x C      models/pointnet.py:309-314
 *           the discriminator convs     models/discriminator.py:22-24, :44-48, :64-67
 *           their dgrad in autograd (mask = the forward layer's saved output).
 *   group_bias carries the folded global-feature / class-one-hot part of the
 *   3024-wide concat (models/pointnet.py:304-309), one row per cloud.
 * Optional fused reductions (out may be NULL when only these are wanted):
 *   colmax_key[g, c] : packed (value, first index) max over the rows of cloud g
 *                      of the PRE-activation value -- torch.max(x, 2) at
 *                      models/pointnet.py:31, :64, :129, :303; unpack with
 *                      pcadv_max_finalize.  Must be zero-filled by the caller.
 *   rowmax_key[r]    : same over the columns of row r -- torch.max over channels
 *                      at models/discriminator.py:71, :135, :169.
 */
typedef struct pcadv_linear_args {
  int64_t rows;
  int32_t n;
  int32_t num_seg;
  pcadv_seg seg[PCADV_MAX_SEG];
  const void* w;              /* [n, ktot] row-major, ktot = sum seg[i].k */
  int64_t ldw;
  int32_t w_dtype;
  int32_t engine;
  const float* bias;          /* [n] or NULL */
  const float* group_bias;    /* [rows / rows_per_group, n] or NULL */
  int64_t rows_per_group;     /* required with group_bias / colmax_key */
  const float* addend;        /* fp32 [rows, n] or NULL */
  int64_t ld_addend;
  int32_t act;
  float slope;
  const void* mask;           /* saved forward activation [rows, n] or NULL */
  int64_t ld_mask;
  int32_t mask_dtype;
  int32_t mask_act;
  float mask_slope;
  int32_t out_dtype;
  const float* out_scale;     /* device scalar or NULL */
  void* out;                  /* [rows, n] or NULL */
  int64_t ld_out;
  unsigned long long* colmax_key;
  unsigned long long* rowmax_key;
  /* Sign bits of the activation, so that the backward reads 1 bit instead of 16 per element
   * (tensor-core engine, 16-bit TMA-storable out, n a multiple of 64, no addend / rowmax / out_scale):
   *   bits_out[r, c / 32], bit ((c % 32) / 2 + 16 * (c % 2)) = [stored out[r, c] > 0]   (written with out)
   *   mask_bits: the same map of the forward layer, used INSTEAD of mask: act'(.) = bit ? 1 : mask_slope
   * Row-major uint32 words with leading dimensions ld_bits_out / ld_mask_bits (words, even). */
  uint32_t* bits_out;
  int64_t ld_bits_out;
  const uint32_t* mask_bits;
  int64_t ld_mask_bits;
  /* Optional by-product (tensor-core engine, the mask_bits dgrad shape, rows_per_group % 128 == 0, seg[0].k <= 256):
   *   seg0_group_sum[g, k] += sum_{r in cloud g} seg[0][r, k]        (fp32 [rows / rows_per_group, seg[0].k], NOT scaled)
   * taken from the tiles of segment 0 while they pass through shared memory -- for PointNetSeg the gradient of
   * fc1's per-cloud bias (models/pointnet.py:304-309 under autograd), without a separate pass over dz_fc1. */
  float* seg0_group_sum;
} pcadv_linear_args;

int pcadv_linear(const pcadv_linear_args* a, void* stream);

/*
 * pcadv_chain: up to four consecutive pointwise layers in one kernel,
 *   y_0 = act_0(x W_0^T + b_0),  y_l = act_l(y_{l-1} W_l^T + b_l)
 * -- the F.relu(self.convK(x)) chains of models/pointnet.py:292-300 and
 * models/discriminator.py:64-67.  Every width (k0 and each n) is a multiple of 64 in [64, 256],
 * x / weights / outputs are 16-bit (`dtype`), accumulation is fp32.  Each layer's output is written
 * once (`out`, with its sign-bit map `bits_out`, see pcadv_linear) but never read back: the next
 * layer multiplies the tile straight out of shared memory.  With `rowmax_key` the LAST layer stores
 * no output; its pre-activation maximum over the channels goes to rowmax_key[r] (packed key, see
 * pcadv_linear / pcadv_max_finalize) -- torch.max over channels at models/discriminator.py:71.
 * With `out_f32` the last layer's rows are stored as fp32 instead (fc4's logits, models/pointnet.py:314).
 * Chains whose widest stored output is 256 run one tile at a time (one shared-memory tile).
 */
#define PCADV_CHAIN_MAX_LAYERS 4
typedef struct pcadv_chain_layer {
  const void* w;              /* [n, k] row-major, k = previous layer's n (k0 for layer 0) */
  int64_t ldw;
  int32_t n;
  int32_t act;
  float slope;
  const float* bias;          /* [n] fp32 or NULL */
  void* out;                  /* [rows, n] 16-bit; may be NULL only for the last layer */
  int64_t ld_out;
  uint32_t* bits_out;         /* [rows, n / 32] or NULL */
  int64_t ld_bits;
} pcadv_chain_layer;

typedef struct pcadv_chain_args {
  int64_t rows;
  const void* x;              /* [rows, k0] */
  int64_t ldx;
  int32_t k0;
  int32_t dtype;              /* PCADV_F16 / PCADV_BF16 */
  int32_t num_layers;         /* 2 .. PCADV_CHAIN_MAX_LAYERS */
  pcadv_chain_layer layer[PCADV_CHAIN_MAX_LAYERS];
  unsigned long long* rowmax_key;   /* [rows] or NULL */
  float* out_f32;             /* or: the last layer's rows as fp32 [rows, n_f32] (contiguous, n_f32 <= 64; the */
  int32_t n_f32;              /*   layer's n is 64 and its weight has n_f32 rows) -- the segmentation logits   */
} pcadv_chain_args;

int pcadv_chain(const pcadv_chain_args* a, void* stream);

/*
 * pcadv_wgrad: dw[c, koff_seg + k] += scale * sum_r dz[r, c] * seg[r, k]
 *              dbias[c]            += scale * sum_r dz[r, c]
 *              dgroup_bias[g, c]   += sum_{r in cloud g} dz[r, c]   (NOT scaled)
 * Replaces the weight / bias gradients autograd derives for the layers above.
 * Outputs are fp32 and are accumulated into (caller zero-fills).
 */
typedef struct pcadv_wgrad_args {
  int64_t rows;
  int32_t n;
  int32_t num_seg;
  const void* dz;             /* [rows, n] */
  int64_t ld_dz;
  int32_t dz_dtype;
  int32_t engine;
  pcadv_seg seg[PCADV_MAX_SEG];
  float* dw;                  /* [n, ktot] or NULL */
  int64_t ld_dw;
  float* dbias;               /* [n] or NULL */
  float* dgroup_bias;         /* [rows / rows_per_group, n] or NULL */
  int64_t rows_per_group;
  const float* scale;         /* device scalar or NULL */
} pcadv_wgrad_args;

int pcadv_wgrad(const pcadv_wgrad_args* a, void* stream);

/*
 * pcadv_backlevel: one backward LEVEL of a pointwise chain in a single pass over the incoming
 * gradient.  x is a stored activation; seg[i] are the gradients dz_i (w.r.t. the pre-activation
 * outputs) of the layers that consumed x -- for PointNetSeg's trunk level k: dz_{k+1} and dz_fc1,
 * because x_k feeds conv_{k+1} and, through the concat of models/pointnet.py:304-309, fc1:
 *   dz_out[r, c]    = act'(x[r, c]) * sum_i sum_k seg_i[r, k] * w[c, koff_i + k]
 *   dw[i][k, c]    += scale * sum_r seg_i[r, k] * x[r, c]        (= d(weight of consumer i)[k, c])
 *   dbias[i][k]    += scale * sum_r seg_i[r, k]
 *   dgroup[i][g, k] += sum_{r in cloud g} seg_i[r, k]            (NOT scaled; rows_per_group % 128 == 0)
 * i.e. what autograd derives with one dgrad GEMM per level and one wgrad GEMM per consumer, each of
 * which re-reads the dz matrices; here every dz tile is fetched once and feeds both tensor-core GEMMs.
 * Tensor-core engine only: seg / w / x / dz_out share one 16-bit dtype, every k_i and n is a multiple
 * of 64, sum k_i <= 1024, rows >= 128.  w = [n, sum k_i] (the K-concat of the transposed forward
 * weights), mask_bits = the sign-bit map of x (see pcadv_linear), fp32 outputs are accumulated into.
 */
typedef struct pcadv_backlevel_args {
  int64_t rows;
  int32_t n;
  int32_t num_seg;
  pcadv_seg seg[PCADV_MAX_SEG];
  const void* w;
  int64_t ldw;
  const void* x;              /* [rows, n] */
  int64_t ldx;
  const uint32_t* mask_bits;  /* [rows, ld_mask_bits] words or NULL with mask_act == PCADV_ACT_NONE */
  int64_t ld_mask_bits;
  int32_t mask_act;
  float mask_slope;
  void* dz_out;               /* [rows, n] */
  int64_t ld_out;
  float* dw[PCADV_MAX_SEG];   /* [k_i, n] or NULL */
  int64_t ld_dw[PCADV_MAX_SEG];
  float* dbias[PCADV_MAX_SEG];    /* [k_i] or NULL */
  float* dgroup[PCADV_MAX_SEG];   /* [rows / rows_per_group, k_i] or NULL */
  int64_t rows_per_group;
  const float* scale;         /* device scalar or NULL */
  /* Optional (onehot_idx != NULL; num_seg == 1, seg[0].ptr == NULL, seg[0].k = pooled channels, a multiple
   * of 64): segment 0 is not read from memory -- it is the gradient of a max over channels
   * (models/discriminator.py:71), built in shared memory one element per row:
   *   dz_0[r, c] = (c == onehot_idx[r]) ? onehot_dy[r] * act'(onehot_val[r]) * (*onehot_scale) : 0
   * (rounded to the 16-bit dtype; the bias gradient sums the unrounded values).  This is pcadv_rowmax_wgrad +
   * pcadv_rowmax_dgrad in one pass over the pooled layer's input x. */
  const float* onehot_dy;     /* [rows] */
  const float* onehot_val;    /* [rows] pooled post-activation value */
  const int32_t* onehot_idx;  /* [rows] argmax channel */
  int32_t onehot_act;
  float onehot_slope;
  const float* onehot_scale;  /* device scalar or NULL */
} pcadv_backlevel_args;

int pcadv_backlevel(const pcadv_backlevel_args* a, void* stream);

/* Unpack `count` packed max keys: val = act(value), idx = first index (0 when a
 * ReLU layer's maximum is <= 0, as every post-ReLU value then ties at 0 and
 * torch.max returns the first).  idx may be NULL. */
int pcadv_max_finalize(const unsigned long long* key, int64_t count, int32_t act, float slope,
                       float* val, int32_t* idx, void* stream);

/*
 * pcadv_maxpool_bwd: backward of "layer + activation + max over the cloud's
 * points" (models/pointnet.py:301-303) through the saved argmax only.
 *   dzc = dg[g, c] * act'(gval[g, c]);  r = g * rows_per_group + idx[g, c]
 *   dw[c, :]    += scale * dzc * x[r, :]
 *   dbias[c]    += scale * dzc
 *   dx_acc[r,:] += dzc * w[c, :]                (fp32 scatter-add, NOT scaled)
 * or, instead of dx_acc, folded straight into the dz of the previous layer:
 *   dz_inout[r,:] += prev_act'(x[r, :]) * sum_{c: row(g,c) = r} dzc * w[c, :]
 *   (each touched row is summed in fp32 and added once; only argmax rows are touched)
 * Replaces the dense B x C x N scatter + dgrad + wgrad autograd runs.
 */
typedef struct pcadv_maxbwd_args {
  int32_t groups;
  int32_t n;                  /* channels of the pooled layer */
  int32_t k;                  /* its input channels */
  int32_t act;
  float slope;
  int32_t x_dtype;
  int32_t w_dtype;
  int32_t prev_act;           /* activation that produced x (for dz_inout) */
  float prev_slope;
  int32_t dz_dtype;
  void* dz_inout;             /* [rows, k] or NULL */
  int64_t ld_dz;
  void* workspace;            /* with dz_inout: pcadv_query_workspace(PCADV_WS_MAXPOOL_BWD_INPLACE, ...) bytes, 16-byte aligned */
  int64_t rows_per_group;
  const float* dg;            /* [groups, n] */
  const float* gval;          /* [groups, n] pooled post-activation value */
  const int32_t* idx;         /* [groups, n] */
  const void* x;              /* [groups * rows_per_group, k] */
  int64_t ldx;
  const void* w;              /* [n, k] */
  int64_t ldw;
  float* dw;                  /* [n, k] or NULL */
  int64_t ld_dw;
  float* dbias;               /* [n] or NULL */
  float* dx_acc;              /* [rows, k] fp32 or NULL */
  int64_t ld_dx;
  const float* scale;
} pcadv_maxbwd_args;

int pcadv_maxpool_bwd(const pcadv_maxbwd_args* a, void* stream);

/* Backward of the max over channels: dz[r, c] = (c == idx[r]) ? dy[r] * (*scale) *
 * act'(val[r]) : 0, written densely as `dz_dtype`. */
int pcadv_rowmax_bwd(const float* dy, const float* val, const int32_t* idx, int64_t rows, int32_t n,
                     int32_t act, float slope, const float* scale, void* dz, int64_t ld_dz,
                     int32_t dz_dtype, void* stream);

/* The same backward without materialising the one-hot dz (models/discriminator.py:67-72),
 * with s[r] = dy[r] * act'(val[r]):
 *   pcadv_rowmax_dgrad: dz_prev[r, :] = prev_act'(yprev[r, :]) * (*scale) * s[r] * w[idx[r], :]
 *   pcadv_rowmax_wgrad: dw[idx[r], :] += s[r] * yprev[r, :];  dbias[idx[r]] += s[r]   (fp32)
 * w is the pooled layer's [n, k] weight, yprev its [rows, k] input (k a multiple of 8). */
int pcadv_rowmax_dgrad(const float* dy, const float* val, const int32_t* idx, int64_t rows, int32_t k,
                       int32_t act, float slope, const float* scale, const void* w, int64_t ldw,
                       int32_t w_dtype, const void* yprev, int64_t ld_y, int32_t y_dtype,
                       int32_t prev_act, float prev_slope, void* dz, int64_t ld_dz, int32_t dz_dtype,
                       void* stream);
int pcadv_rowmax_wgrad(const float* dy, const float* val, const int32_t* idx, int64_t rows, int32_t n,
                       int32_t k, int32_t act, float slope, const void* yprev, int64_t ld_y,
                       int32_t y_dtype, float* dw, int64_t ld_dw, float* dbias, void* stream);

/* scale2[0] = S = 2^floor(log2(target / max|x|)) (1 when x is all zero),
 * scale2[1] = 1 / S.  x is a [rows, cols] fp32 matrix with leading dimension ld.
 * `workspace` is one zero-filled uint32. */
int pcadv_amax_scale(const float* x, int64_t rows, int32_t cols, int64_t ld, float target,
                     unsigned int* workspace, float* scale2, void* stream);

/* dst[r, c] = convert(src[r, c] * (*scale) * act'(mask[r, c])) for c < cols, 0 for
 * cols <= c < cols_pad.  scale and mask are optional (NULL). */
int pcadv_convert(const void* src, int32_t src_dtype, int64_t ld_src, int64_t rows, int32_t cols,
                  void* dst, int32_t dst_dtype, int64_t ld_dst, int32_t cols_pad, const float* scale,
                  const void* mask, int64_t ld_mask, int32_t mask_dtype, int32_t mask_act,
                  float mask_slope, void* stream);

/* Same as pcadv_convert for a CHANNEL-MAJOR fp32 source: element (cloud b, channel c, point i)
 * at src[b * batch_stride + c * chan_stride + i] -- the B x C x N layout torch's softmax /
 * log_softmax (utils/trainer.py:901, :914) and their backward hand over -- written point-major
 * to dst[(b * rows_per_group + i), c]. */
int pcadv_convert_cm(const float* src, int64_t batch_stride, int64_t chan_stride, int64_t groups,
                     int64_t rows_per_group, int32_t cols, void* dst, int32_t dst_dtype,
                     int64_t ld_dst, int32_t cols_pad, const float* scale, void* stream);

/*
 * pcadv_softmax_head: one pass over the fp32 segmentation logits [rows, n] (n <= 128):
 *   mode PCADV_HEAD_CE  (labelled batch, utils/trainer.py:899-901)
 *     probs[r, c]  = softmax(logits[r, :])[c]                       F.softmax(pred, dim=1)
 *     dz[r, c]     = dz_gain * (probs[r, c] - [c == labels[r]])     d CrossEntropyLoss / d logits,
 *                                                                   up to the 1 / rows mean factor
 *     *loss_sum   += sum_r (logsumexp(logits[r, :]) - logits[r, labels[r]])   over rows with a label in [0, n)
 *     *valid_count += the number of such rows (the mean's denominator)
 *   mode PCADV_HEAD_LSM (unlabelled batch, utils/trainer.py:914)
 *     probs[r, c]  = log_softmax(logits[r, :])[c]
 * probs / dz are point-major [rows, *_cols] matrices of *_dtype, zero-filled in the columns
 * n <= c < *_cols (the K padding the discriminator's / fc4-dgrad's tensor-core GEMM wants).
 * Any of probs, dz, loss_sum may be NULL.
 */
enum { PCADV_HEAD_CE = 0, PCADV_HEAD_LSM = 1 };
typedef struct pcadv_head_args {
  int64_t rows;
  int32_t n;
  int32_t mode;
  const float* logits;
  int64_t ld;
  const int64_t* labels;      /* [rows] or NULL */
  void* probs;
  int64_t ld_probs;
  int32_t probs_dtype;
  int32_t probs_cols;
  void* dz;
  int64_t ld_dz;
  int32_t dz_dtype;
  int32_t dz_cols;
  float dz_gain;
  float* loss_sum;            /* device scalar, accumulated into, or NULL */
  float* valid_count;         /* device scalar += #rows with labels[r] in [0, n), or NULL.  Rows whose label
                                 is outside [0, n) are IGNORED rows (nn.CrossEntropyLoss's ignore_index,
                                 -100 by default): no loss term, dz row = 0. */
} pcadv_head_args;

int pcadv_softmax_head(const pcadv_head_args* a, void* stream);

/* Backward of log_softmax over the columns (utils/trainer.py:914 under :927-929):
 *   dz[r, c] = (*scale) * (dy[r, c] - exp(lp[r, c]) * sum_j dy[r, j]),  c < n;  0 for n <= c < dz_cols.
 * lp is the saved log_softmax output, dy its incoming gradient; any float dtype each. */
int pcadv_logsoftmax_bwd(const void* lp, int32_t lp_dtype, int64_t ld_lp, const void* dy,
                         int32_t dy_dtype, int64_t ld_dy, int64_t rows, int32_t n, const float* scale,
                         void* dz, int32_t dz_dtype, int64_t ld_dz, int32_t dz_cols, void* stream);

/* StackDiscNet.custom_activation (models/discriminator.py:153-159) over point-major shape logits
 * x [rows, S] (fp32, S <= 64): z = logsumexp_c x[r, c];  y[r] = z / (z + 1)  (y nullable);
 * with dy and dx given: dx[r, c] = dy[r] * softmax(x[r, :])[c] / (z + 1)^2. */
int pcadv_lse_ratio(const float* x, int64_t ld, int64_t rows, int32_t S, const float* dy, float* y, float* dx,
                    int64_t ld_dx, void* stream);

/* out[c] += sum_r (v - round_dtype(v)), v = src[r, c] * (*scale): the column sums of what the 16-bit
 * conversion of a gradient matrix drops (dtype = PCADV_F16 | PCADV_BF16, cols <= 64; the caller
 * zero-fills out).  Added to a bias gradient formed from the 16-bit dz it restores the exact fp32
 * column sum of the reference's autograd (the `.sum(0)` of a Linear / Conv1d bias gradient). */
int pcadv_round_residual(const float* src, int64_t ld, int64_t rows, int32_t cols, const float* scale,
                         int32_t dtype, float* out, void* stream);

/*
 * T-Net transforms (fp32, k <= 128), one launch per op instead of one torch.bmm per cloud.
 *   pcadv_bmm:       y[g, r, :] = x[g, r, :] @ T[g]  (transpose_t = 0)  or  @ T[g]^T  (= 1, the backward dx)
 *                    -- torch.bmm(x^T, trans) at models/pointnet.py:120-122, :231, :238
 *   pcadv_bmm_tgrad: dT[g] += x[g]^T @ dy[g]          (caller zero-fills dT)
 * x, y, dy: [groups, rows_per_group, k] contiguous; T, dT: [groups, k, k] contiguous.
 */
int pcadv_bmm(const float* x, const float* T, float* y, int32_t groups, int64_t rows_per_group, int32_t k,
              int32_t transpose_t, void* stream);
int pcadv_bmm_tgrad(const float* x, const float* dy, float* dT, int32_t groups, int64_t rows_per_group,
                    int32_t k, void* stream);

/*
 * feature_transform_regularizer (models/pointnet.py:345-353), one CTA per cloud:
 *   pcadv_ortho_reg:     diff[g] = T[g] T[g]^T - I,  norms[g] = ||diff[g]||_F     (loss = mean_g norms[g])
 *   pcadv_ortho_reg_bwd: dT[g] = (2 * (*dloss) / (groups * norms[g])) * diff[g] @ T[g]
 * T, diff, dT: [groups, d, d] fp32 contiguous, d <= 128; dloss: device scalar.
 */
int pcadv_ortho_reg(const float* T, int32_t groups, int32_t d, float* diff, float* norms, void* stream);
int pcadv_ortho_reg_bwd(const float* diff, const float* T, const float* norms, const float* dloss,
                        int32_t groups, int32_t d, float* dT, void* stream);

/*
 * Part-IoU evaluation on the device: utils/metric.py:20-39 (get_iou /
 * batch_get_iou) and the argmax + accuracy count of run_testing_seg, utils/trainer.py:100-110,
 * which the reference computes with numpy on the host from a copy of every prediction.
 *
 * pcadv_part_counts: for each cloud g of N points, pred = argmax over the C columns of
 *   logits[p, :] (fp32, row stride ld; the first maximum, as pred.max(1)[1]) -- or pred_in[p] when
 *   logits is NULL (exactly one of the two is given) -- then, with gt = labels[p] (int64),
 *     counts[g, 0, l] += #(pred == l && gt == l)   counts[g, 1, l] += #(pred == l)
 *     counts[g, 2, l] += #(gt == l)                correct[g]      += #(pred == gt)
 *   counts: int32 [groups, 3, C], correct: int32 [groups]; the caller zero-fills both.
 *   pred_out (nullable): int64 [groups * N] receives the predictions.  C <= 64, groups <= 65535.
 * pcadv_part_iou: cat[g] = argmax of onehot[g, 0..ncat) (first maximum, np.argmax), parts
 *   l in [part_begin[cat], part_begin[cat + 1]) (device int32 [ncat + 1]; seg_classes of
 *   utils/metric.py:5), iou[g] = mean_l (pred and gt both empty ? 1 : inter / union) in float64,
 *   summed in index order as np.mean does for fewer than 8 terms.
 */
int pcadv_part_counts(const float* logits, int64_t ld, const int64_t* pred_in, const int64_t* labels,
                      int32_t groups, int64_t N, int32_t C, int32_t* counts, int32_t* correct,
                      int64_t* pred_out, void* stream);
int pcadv_part_iou(const int32_t* counts, const float* onehot, int64_t ld_onehot, int32_t ncat,
                   const int32_t* part_begin, int32_t groups, int32_t C, double* iou, int32_t* cat,
                   void* stream);

/*
 * Scratch sizes (lower face): bytes of the caller-owned workspace an entry point wants
 * for the given shape, or -1 (and pcadv_last_error) for an unknown op.
 *   PCADV_WS_MAXPOOL_BWD_INPLACE  pcadv_maxbwd_args.workspace when dz_inout is set
 *   PCADV_WS_AMAX_SCALE           pcadv_amax_scale's workspace (zero-filled by the caller)
 */
enum { PCADV_WS_MAXPOOL_BWD_INPLACE = 0, PCADV_WS_AMAX_SCALE = 1 };
long long pcadv_query_workspace(int32_t op, int64_t groups, int64_t rows_per_group, int32_t n);

/*
 * On-device point jitter -- jitter_point_cloud of dataset/modelNetData.py:80-91:
 *   dst[i] = src[i] + clamp(sigma * z_i, -clip, clip),  z_i ~ N(0, 1)
 * over `count` floats (N x 3 coordinates).  z comes from Philox-4x32-10 keyed by `seed`: element i
 * uses normal (i & 3) of counter (offset + (i >> 2)), Box-Muller on the 24-bit uniform pairs, so a
 * (seed, offset) pair reproduces the same jitter on any launch geometry.  numpy's generator (the
 * reference draws np.random.randn on the host) cannot be reproduced; the distribution is the contract.
 * src == dst is allowed.
 */
int pcadv_jitter(const float* src, float* dst, int64_t count, float sigma, float clip,
                 unsigned long long seed, unsigned long long offset, void* stream);

/*
 * OPTIONAL BatchNorm over point-major rows [rows, C] (C in {8, 16, ..., 2048}, a power of two).  The
 * reference has NO BatchNorm (commented out at models/pointnet.py:100-103) -- this is
 * the default-off layer BASELINE.json's north_star words, checked against torch.nn.BatchNorm1d only.
 *   pcadv_bn_stats: per-channel batch mean and 1 / sqrt(var + eps) (biased variance), optionally the
 *                   running statistics update of nn.BatchNorm1d (momentum, unbiased variance).
 *                   workspace: 2 * C floats, zero-filled by the caller.
 *   pcadv_bn_apply: y = act((x - mean) * rstd * gamma + beta), act in {NONE, RELU}; gamma / beta nullable.
 *   pcadv_bn_bwd:   dbeta[c] += sum_r dy', dgamma[c] += sum_r dy' * xhat (caller zero-fills both),
 *                   dx = gamma * rstd * (dy' - dbeta / rows - xhat * dgamma / rows)  (dx nullable);
 *                   dy' = dy * [y > 0] when y (the ReLU'd output) is given, else dy.
 */
int pcadv_bn_stats(const void* x, int32_t x_dtype, int64_t ld_x, int64_t rows, int32_t C, float eps,
                   float momentum, float* workspace, float* mean, float* rstd, float* running_mean,
                   float* running_var, void* stream);
int pcadv_bn_apply(const void* x, int32_t x_dtype, int64_t ld_x, int64_t rows, int32_t C, const float* mean,
                   const float* rstd, const float* gamma, const float* beta, int32_t act, void* y,
                   int32_t y_dtype, int64_t ld_y, void* stream);
int pcadv_bn_bwd(const void* x, int32_t x_dtype, int64_t ld_x, const void* dy, int32_t dy_dtype, int64_t ld_dy,
                 const void* y, int32_t y_dtype, int64_t ld_y, int64_t rows, int32_t C, const float* mean,
                 const float* rstd, const float* gamma, float* dgamma, float* dbeta, void* dx, int32_t dx_dtype,
                 int64_t ld_dx, void* stream);

/* dst[c, r] = src[r, c] with conversion: builds the [k, n] copy of a weight
 * matrix that dgrad consumes. */
int pcadv_transpose(const void* src, int32_t src_dtype, int64_t ld_src, int32_t rows, int32_t cols,
                    void* dst, int32_t dst_dtype, int64_t ld_dst, void* stream);

int pcadv_version(void);
/* 0 when the current device is compute capability 10.x, else non-zero. */
int pcadv_device_check(void);
const char* pcadv_last_error(void);
/* Number of kernel launches issued by this library since process start. */
long long pcadv_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif  /* PCADV_H_ */
