"""Forward / backward of a chain of pointwise layers on libpcadv ops.

A *chain* is what the reference writes as ``F.relu(self.convK(x))`` /
``F.relu(self.fcK(x))`` sequences (models/pointnet.py:291-301, :308-314;
models/discriminator.py:22-24, :44-48, :64-67).  Activations are point-major
``[rows, channels]`` matrices; ``dz`` always means the gradient with respect
to a layer's PRE-activation output, which is what dgrad and wgrad consume.
"""
import os

import torch

from .. import ops
from ..ops import ACT_NONE, ENGINE_TC


class ZeroPool:
    """One zero-filled fp32 buffer per backward pass, carved into the dW / dbias accumulators the
    wgrad kernels add into: one fill kernel instead of one per tensor.  Slices start on 16-byte
    boundaries (the vector RED path of tc_wgrad wants that)."""

    def __init__(self, numel, device):
        self.buf = torch.zeros(int(numel), dtype=torch.float32, device=device)
        self.off = 0

    def take(self, *shape):
        n = 1
        for d in shape:
            n *= int(d)
        if self.off + n > self.buf.numel():
            return torch.zeros(shape, dtype=torch.float32, device=self.buf.device)
        t = self.buf[self.off:self.off + n].view(shape)
        self.off += (n + 3) // 4 * 4
        return t

    @staticmethod
    def size_for(layers_nk):
        """Elements needed for [(n_pad, k_total)] layers (dW + dbias each), with alignment slack."""
        return sum(n * k + n + 8 for n, k in layers_nk)


def _zeros(pool, shape, device):
    return pool.take(*shape) if pool is not None else torch.zeros(shape, dtype=torch.float32, device=device)


class Layer:
    """One pointwise layer: y = act(x W^T + b)."""

    __slots__ = ("w", "b", "act", "slope")

    def __init__(self, w, b, act=ACT_NONE, slope=0.0):
        self.w = w.reshape(w.shape[0], -1)      # Conv1d [Cout, Cin, 1] -> [Cout, Cin]
        self.b = b
        self.act = act
        self.slope = slope


# ---- per-step cache of the engine copies of the weights ------------------------------------------
# Inside ``with weight_cache():`` (the trainer's step functions) a parameter is converted /
# transposed / K-concatenated once and reused by every pass of the step that needs the same copy
# (G runs two forward and two backward passes per step, D three and two and a half).  Keys carry
# the tensor's storage, view geometry and autograd version counter, so an in-place optimizer update
# inside the scope invalidates the entry; the cache dies with the scope, so nothing captured into a
# CUDA graph is ever looked up from eager code later.
_WCACHE = None


class weight_cache:
    def __enter__(self):
        global _WCACHE
        self._prev, _WCACHE = _WCACHE, ({} if _WCACHE is None else _WCACHE)
        return self

    def __exit__(self, *exc):
        global _WCACHE
        _WCACHE = self._prev


def _wkey(kind, prec, tensors, extra):
    # The stream is part of the key: an engine copy built on one stream must not be picked up by
    # work issued on another one without an event in between (the step functions run the
    # discriminator phase on a second stream, trainer.py); each stream builds its own copy.
    stream = torch.cuda.current_stream().cuda_stream if tensors and tensors[0].is_cuda else 0
    return (kind, prec.name, stream,
            tuple((t.data_ptr(), tuple(t.shape), t.stride(), t._version) for t in tensors), tuple(extra))


def compute_weight(prec, w32, k_list, n):
    """The copy of a weight matrix the engine multiplies with: a 16-bit copy when
    the tensor-core engine can take the shape, else the fp32 master."""
    if _WCACHE is not None:
        key = _wkey("cw", prec, [w32], list(k_list) + [n])
        hit = _WCACHE.get(key)
        if hit is None:
            hit = _WCACHE[key] = _compute_weight(prec, w32, k_list, n)
        return hit
    return _compute_weight(prec, w32, k_list, n)


def _compute_weight(prec, w32, k_list, n):
    ktot = sum(k_list)
    if ktot != w32.shape[1]:                     # input carries zero-padded columns (K = 50 -> 64)
        w32 = pad_cols(w32, ktot)
    if prec.engine == ENGINE_TC and all(k % 64 == 0 for k in k_list):
        return w32.to(prec.act_dtype)
    return w32


_CHAIN_ON = os.environ.get("PCADV_CHAIN", "1") != "0"      # tuning aid: per-layer launches only
# widest dz for which a backward level goes through pcadv_backlevel (one weight-gradient tile in TMEM;
# wider levels measured slower than separate dgrad / wgrad launches, tools/level_ab.py)
_LEVEL_MAX_K = int(os.environ.get("PCADV_LEVEL_MAX_K", "128"))


def _chain_run(prec, x, run, rowmax=False, want_bits=True, last_f32=False):
    """``run`` consecutive layers on x through one pcadv_chain launch."""
    ws, k = [], x.shape[1]
    for L in run:
        ws.append((compute_weight(prec, L.w, [k], L.w.shape[0]), L.b, L.act, L.slope))
        k = (L.w.shape[0] + 63) // 64 * 64
    return ops.chain(x, ws, rowmax=rowmax, want_bits=want_bits, last_f32=last_f32)


def chain_run_length(prec, x, layers, start, final_fp32=False):
    """(how many layers from ``start`` on can go into one chained launch (0 = none), whether the
    run ends with the fp32 last layer): tensor-core engine, 16-bit single-segment input, every
    width a multiple of 64 within the kernel's budget; an fp32 last layer may be up to 64 wide."""
    if prec.engine != ENGINE_TC or prec.act_dtype == torch.float32 or not _CHAIN_ON:
        return 0, False
    widths = [x.shape[1]]
    best, best_f32 = 0, False
    for j in range(start, min(len(layers), start + 4)):
        n = layers[j].w.shape[0]
        f32 = final_fp32 and j == len(layers) - 1
        if f32:
            if n > 64 or layers[j].act != ACT_NONE:
                break
            n = 64
        widths.append(n)
        if len(widths) >= 3 and ops.chain_eligible(x, widths, last_f32=f32):
            best, best_f32 = j - start + 1, f32
        if n % 64:
            break
    return best, best_f32


def chain_forward(prec, x_segs, layers, final_fp32=False, rows_per_group=0, group_bias=None,
                  bits=None):
    """Run ``layers`` on the K-concat of ``x_segs``.  ``group_bias`` (per-cloud
    bias) applies to the first layer.  Returns the list of layer outputs.  ``bits``: a list that
    receives, per layer, the 1-bit map [y > 0] of an activated output (or None where the layer
    cannot emit it) -- what the backward reads instead of the 16-bit activation.  Consecutive
    narrow layers (widths <= 256, multiples of 64) go through ``pcadv_chain`` in one launch."""
    ys = []
    segs = list(x_segs)
    i = 0
    while i < len(layers):
        L = layers[i]
        last = i == len(layers) - 1
        if len(segs) == 1 and not (i == 0 and group_bias is not None):
            run, run_f32 = chain_run_length(prec, segs[0], layers, i, final_fp32=final_fp32)
            if run >= 2:
                outs, bts, _ = _chain_run(prec, segs[0], layers[i:i + run], want_bits=bits is not None,
                                          last_f32=run_f32)
                ys.extend(outs)
                if bits is not None:
                    bits.extend(bts)
                segs = [outs[-1]]
                i += run
                continue
        n = L.w.shape[0]
        w = compute_weight(prec, L.w, [s.shape[1] for s in segs], n)
        out_dtype = torch.float32 if (last and final_fp32) else prec.act_dtype
        b = None
        if bits is not None and L.act != ACT_NONE and out_dtype != torch.float32 and \
                ops.bits_eligible(prec, segs, w, n):
            b = ops.new_bits(segs[0].shape[0], n, segs[0].device)
        y, _, _ = ops.linear(segs, w, bias=L.b, act=L.act, slope=L.slope, out_dtype=out_dtype,
                             engine=prec.engine, rows_per_group=rows_per_group if i == 0 else 0,
                             group_bias=group_bias if i == 0 else None, bits_out=b)
        ys.append(y)
        if bits is not None:
            bits.append(b)
        segs = [y]
        i += 1
    return ys


def pad_cols(t, width):
    """Zero-pad the columns of a small weight matrix to ``width``."""
    n = t.shape[1]
    if n == width:
        return t
    out = t.new_zeros((t.shape[0], width))
    out[:, :n] = t
    return out


def prepare_dz(prec, dy, scale2, mask=None, mask_act=ACT_NONE, mask_slope=0.0):
    """Turn an incoming fp32 gradient [rows, n] (with respect to a layer's OUTPUT)
    into the dz the backward chain consumes: multiplied by act'(mask) when the
    layer has an activation; in 16-bit mode also scaled by S, converted to the
    activation dtype and zero-padded to a multiple of 64 columns."""
    if not prec.scaled:
        if mask is None or mask_act == ACT_NONE:
            return dy
        return ops.convert(dy, torch.float32, mask=mask, mask_act=mask_act, mask_slope=mask_slope)
    n = dy.shape[1]
    if mask_act == ACT_NONE:
        mask = None
    return ops.convert(dy, prec.act_dtype, cols_pad=(n + 63) // 64 * 64, scale=scale2[0:1],
                       mask=mask, mask_act=mask_act, mask_slope=mask_slope)


def dgrad_weight(prec, blocks, n_out, widths):
    """Weight of a dgrad GEMM: the K-concat of transposed forward weights.
    ``blocks`` = list of [Cout_i, n_out] forward-weight slices (fp32), ``widths`` the
    column counts of the dz matrices they multiply (>= Cout_i when dz carries
    zero padding); the result is [n_out, sum(widths)] in the compute dtype."""
    if _WCACHE is not None:
        key = _wkey("dg", prec, blocks, [n_out] + list(widths))
        hit = _WCACHE.get(key)
        if hit is None:
            hit = _WCACHE[key] = _dgrad_weight(prec, blocks, n_out, widths)
        return hit
    return _dgrad_weight(prec, blocks, n_out, widths)


def _dgrad_weight(prec, blocks, n_out, widths):
    parts = []
    for wb, width in zip(blocks, widths):
        parts.append(pad_cols(wb.t(), width))
    wt = parts[0] if len(parts) == 1 else torch.cat(parts, 1)
    wt = wt.contiguous()
    if prec.engine == ENGINE_TC and wt.shape[1] % 64 == 0:
        return wt.to(prec.act_dtype)
    return wt


def layer_wgrad(prec, dz, x_segs, w_shape, need_w, need_b, scale2, extra_cols=0, pool=None):
    """fp32 (dw, db) of one layer from its dz and input segments.  ``dz`` may carry
    zero-padded columns beyond w_shape[0]."""
    if not (need_w or need_b):
        return None, None
    n_pad = dz.shape[1]
    dev = dz.device
    ktot = sum(s.shape[1] for s in x_segs)
    dw = _zeros(pool, (n_pad, ktot + extra_cols), dev) if need_w else None
    db = _zeros(pool, (n_pad,), dev) if need_b else None
    inv = scale2[1:2] if scale2 is not None else None
    ops.wgrad(dz, x_segs if need_w else [], dw=dw[:, :ktot] if need_w else None, dbias=db, scale=inv,
              engine=prec.engine)
    n, k = w_shape[0], w_shape[1]
    return (dw[:n, :k] if need_w else None), (db[:n] if need_b else None)


def chain_backward(prec, dz_last, x_segs, ys, layers, need_w, need_b, need_x, scale2, addends=None,
                   dx_packed=False, bits=None, pool=None):
    """Backward through a chain.  ``dz_last``: dz of the last layer ([rows, pad(n)]).
    ``need_w[i]`` / ``need_b[i]``: which parameter gradients to form (frozen
    discriminators skip wgrad, utils/trainer.py:885-886).  Returns
    (list of (dw, db), dx, dz0) where dx is the fp32, UNSCALED gradient of the chain
    input (only when ``need_x``; single input segment) and dz0 the (scaled) dz of
    the first layer.  With ``dx_packed`` dx keeps the scale and has the dtype and (padded)
    width of the chain input instead."""
    addends = addends or {}
    grads = [None] * len(layers)
    dz = dz_last
    inv = scale2[1:2] if scale2 is not None else None
    for i in range(len(layers) - 1, -1, -1):
        L = layers[i]
        xin = x_segs if i == 0 else [ys[i - 1]]
        if i > 0 and need_w[i] and need_b[i] and bits is not None and addends.get(i - 1) is None and \
                dz.shape[1] <= _LEVEL_MAX_K and ops.backlevel_eligible(prec, [dz.shape[1]], ys[i - 1], bits[i - 1]):
            # one pass over dz: this layer's weight / bias gradient and the dz of the layer below
            P = layers[i - 1]
            wt = dgrad_weight(prec, [L.w], L.w.shape[1], [dz.shape[1]])
            dw = _zeros(pool, (dz.shape[1], ys[i - 1].shape[1]), dz.device)
            db = _zeros(pool, (dz.shape[1],), dz.device)
            dz = ops.backlevel([dz], wt, ys[i - 1], mask_bits=bits[i - 1], mask_act=P.act, mask_slope=P.slope,
                               dws=[dw], dbiases=[db], scale=inv)
            grads[i] = (dw[:L.w.shape[0], :L.w.shape[1]], db[:L.w.shape[0]])
            continue
        grads[i] = layer_wgrad(prec, dz, xin, L.w.shape, need_w[i], need_b[i], scale2, pool=pool)
        if i == 0:
            break
        P = layers[i - 1]
        wt = dgrad_weight(prec, [L.w], L.w.shape[1], [dz.shape[1]])
        dz, _, _ = ops.linear([dz], wt, mask=ys[i - 1], mask_act=P.act, mask_slope=P.slope,
                              out_dtype=prec.act_dtype, engine=prec.engine,
                              addend=addends.get(i - 1),
                              mask_bits=bits[i - 1] if bits is not None else None)
    dx = None
    if need_x:
        L = layers[0]
        k_in = L.w.shape[1]
        wt = dgrad_weight(prec, [L.w], k_in, [dz.shape[1]])
        if dx_packed:
            k_pad = x_segs[0].shape[1]
            wt_p = wt.new_zeros((k_pad, wt.shape[1]))
            wt_p[:k_in] = wt
            dx, _, _ = ops.linear([dz], wt_p, out_dtype=prec.act_dtype, engine=prec.engine)
            return grads, dx, dz
        if prec.scaled and k_in % 4:
            # pad the output width so its rows are 16-byte aligned (TMA store); slice afterwards
            k_pad = (k_in + 63) // 64 * 64
            wt_p = wt.new_zeros((k_pad, wt.shape[1]))
            wt_p[:k_in] = wt
            wt = wt_p
        inv = scale2[1:2] if scale2 is not None else None
        dx, _, _ = ops.linear([dz], wt, out_dtype=torch.float32, out_scale=inv, engine=prec.engine)
        dx = dx[:, :k_in]
    return grads, dx, dz
