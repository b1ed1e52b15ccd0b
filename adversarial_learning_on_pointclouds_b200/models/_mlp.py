"""Generic autograd Functions on libpcadv ops: a chain of pointwise layers with
an optional fused reduction (max over the cloud's points, or over channels), the
per-cloud T-Net transform, and the orthogonality regulariser.

Tensors that cross a Function boundary are fp32; inside, activations live in the
precision mode's storage dtype and gradients carry a dynamic power-of-two scale.
"""
import os

import torch

from .. import ops
from . import _chain
from ..ops import ACT_NONE, ENGINE_SIMT
from ._chain import (Layer, ZeroPool, chain_forward, chain_backward, compute_weight, prepare_dz,
                     layer_wgrad, dgrad_weight)


# Backward of "layer + activation + max over channels" through pcadv_backlevel's one-hot mode instead of the
# two gather kernels: 0.30 -> 0.22 ms per 2^20 rows back to back (tools/level_ab.py), but the cfg5 step gets
# SLOWER with it (15.58 -> 15.76 ms same box): the discriminator phase shares the GPU with the generator's
# backward, where the gather kernels' small CTAs fill SMs as they free up and a 148-CTA persistent kernel
# with a static row partition waits for whole SMs.  Off by default; PCADV_ONEHOT_LEVEL=1 enables it.
_ONEHOT_LEVEL = os.environ.get("PCADV_ONEHOT_LEVEL", "0") == "1"

# Tests set this to a list: every PointMLPFunction.forward appends the activations it saved (the
# branch decisions of the pass) -- tests/parity.py builds the oracle's ``branch`` dict from it.
DEBUG_TAPE = None


def _tape(spec, x_in, ys, red_val, red_idx, out, hit):
    if DEBUG_TAPE is not None:
        DEBUG_TAPE.append(dict(acts=list(spec.acts), reduce=spec.reduce, group=spec.group, x_in=x_in,
                               ys=list(ys), red_val=red_val, red_idx=red_idx, out=out, cache_hit=hit))


class MLPSpec:
    """Static description of a PointMLPFunction call.

    acts    : [(act, slope)] per layer
    reduce  : None | "points" (max over each cloud's ``group`` rows of the last
              layer, models/pointnet.py:31/:129/:303) | "channels" (max over the
              last layer's channels per row, models/discriminator.py:71)
    group   : rows per cloud (needed for reduce="points")
    tap     : index of a layer whose output is also returned (fp32) -- the
              ``pointfeat`` of models/pointnet.py:124 -- or None
    A per-cloud bias ``gb`` [rows / group, n0] may be added to the first layer (the
    folded tiled-global-feature columns of models/pointnet.py:135-136).
    """

    def __init__(self, acts, reduce=None, group=0, tap=None, box=None, extra_segs=0):
        self.acts, self.reduce, self.group, self.tap = list(acts), reduce, int(group), tap
        # the first layer reads the K-concat [x | extra segments] (models/pointnet.py:246-251 without
        # building the concatenated map): the extra [rows, k_i] fp32 tensors follow the parameters
        self.extra_segs = int(extra_segs)
        # GradBox of a packed 16-bit input (models/_seg.py): the input gradient is then returned
        # in the input's dtype, still carrying this backward's power-of-two scale, which is left
        # in the box for the producer of the input
        self.box = box


class PointMLPFunction(torch.autograd.Function):
    """(x [rows, k] fp32, gb | None, *[w_i, b_i]) -> out fp32 (+ tap fp32).

    out is [rows, n] (reduce=None), [rows / group, n] (reduce="points") or [rows]
    (reduce="channels")."""

    @staticmethod
    def forward(ctx, prec, spec, x, gb, *params):
        if not x.is_cuda:
            raise RuntimeError("libpcadv layers need CUDA tensors; there is no CPU path")
        nl = len(spec.acts)
        extra = list(params[2 * nl:])
        params = params[:2 * nl]
        if len(extra) != spec.extra_segs:
            raise ValueError("PointMLPFunction: spec.extra_segs does not match the arguments")
        layers = [Layer(params[2 * i], params[2 * i + 1], *spec.acts[i]) for i in range(nl)]
        if extra:
            return PointMLPFunction._forward_segments(ctx, prec, spec, [x] + extra, gb, layers, params)
        # Inside a step scope (``weight_cache``) a forward over the very same input and weights is
        # not repeated: the trainer evaluates D(log_softmax(pred_nogt)) in the G phase and again on
        # the detached tensor in the D phase (utils/trainer.py:916, :951-953).
        cache = _chain._WCACHE if gb is None and spec.tap is None else None
        fkey = None
        x_key = x
        if cache is not None:
            ident = lambda t: None if t is None else (t.data_ptr(), tuple(t.shape), t.stride(), t._version, t.dtype)
            fkey = ("fwd", prec.name, tuple(spec.acts), spec.reduce, spec.group, ident(x),
                    tuple(ident(p_) for p_ in params))
            hit = cache.get(fkey)
            if hit is not None:
                x_in, ys, ybits, red_val, red_idx, ctx.bcn, ctx.packed_in, out = hit[:8]
                ctx.prec, ctx.spec = prec, spec
                ctx.n_ys = len(ys)
                ctx.has_red = red_val is not None
                ctx.has_gb = False
                ctx.bit_slots = [i for i, t in enumerate(ybits) if t is not None]
                ctx.save_for_backward(*([x_in] + ys + ([red_val, red_idx] if ctx.has_red else []) +
                                        list(params) + [ybits[i] for i in ctx.bit_slots]))
                _tape(spec, x_in, ys, red_val, red_idx, out, True)
                return out.detach()
        ctx.bcn = None
        ctx.packed_in = False
        x_in = None
        if x.dim() == 2 and prec.scaled and x.dtype == prec.act_dtype and x.shape[1] % 64 == 0 \
                and x.stride(1) == 1 and x.stride(0) % 8 == 0:
            # packed point-major 16-bit rows (the fused softmax heads' output): consumed as is
            x_in = x
            ctx.packed_in = True
        elif x.dim() == 3:
            # B x C x N map (what the trainers hand the discriminators).  Channel-major memory
            # (torch softmax / log_softmax output) is transposed, converted and padded by one
            # kernel; a transposed view of point-major storage (the generator's logits) is free.
            B_, C_, N_ = x.shape
            ctx.bcn = (B_, C_, N_)
            to16 = prec.scaled and C_ >= 16 and B_ * N_ >= 128
            if ops.is_channel_major(x):
                x_in = ops.convert_cm(x, prec.act_dtype if to16 else torch.float32,
                                      cols_pad=(C_ + 63) // 64 * 64 if to16 else C_)
            else:
                x = x.transpose(1, 2).reshape(B_ * N_, C_)
        elif x.dim() != 2:
            raise ValueError("PointMLPFunction expects [rows, k] or B x C x N")
        if x_in is None:
            if x.dtype != torch.float32 or x.stride(1) != 1:
                x = x.contiguous().float()
            x_in = x
            if prec.scaled and x.shape[1] >= 16 and x.shape[0] >= 128:
                # 16-bit copy, K zero-padded to a multiple of 64, feeds the tensor cores
                x_in = ops.convert(x, prec.act_dtype, cols_pad=(x.shape[1] + 63) // 64 * 64)
        body = layers if spec.reduce is None else layers[:-1]
        if gb is not None:
            gb = gb.contiguous().float()
        ybits = []
        red_val = red_idx = None
        fused_red = False
        if spec.reduce == "channels" and gb is None and body and \
                _chain.chain_run_length(prec, x_in, layers, 0)[0] == len(layers):
            # the whole discriminator trunk + max over channels in one chained launch
            outs, bts, rkey = _chain._chain_run(prec, x_in, layers, rowmax=True)
            ys, ybits = outs[:-1], bts[:-1]
            red_val, red_idx = ops.max_finalize(rkey, layers[-1].act, layers[-1].slope)
            out = red_val
            fused_red = True
        else:
            ys = chain_forward(prec, [x_in], body, final_fp32=(spec.reduce is None),
                               rows_per_group=spec.group if gb is not None else 0,
                               group_bias=gb, bits=ybits) if body else []
        if fused_red:
            pass
        elif spec.reduce is not None:
            L = layers[-1]
            src = ys[-1] if ys else x_in
            w = compute_weight(prec, L.w, [src.shape[1]], L.w.shape[0])
            _, ckey, rkey = ops.linear([src], w, bias=L.b, want_out=False,
                                       colmax=spec.reduce == "points",
                                       rowmax=spec.reduce == "channels",
                                       rows_per_group=spec.group, engine=prec.engine)
            red_val, red_idx = ops.max_finalize(ckey if ckey is not None else rkey, L.act, L.slope)
            out = red_val
        else:
            out = ys[-1]
        tap = ys[spec.tap].float() if spec.tap is not None else None
        ctx.prec, ctx.spec = prec, spec
        ctx.n_ys = len(ys)
        ctx.has_red = red_val is not None
        ctx.has_gb = gb is not None
        ctx.bit_slots = [i for i, t in enumerate(ybits) if t is not None]
        saved = [x_in] + ys + ([red_val, red_idx] if ctx.has_red else []) + list(params) + \
            [ybits[i] for i in ctx.bit_slots]
        ctx.save_for_backward(*saved)
        _tape(spec, x_in, ys, red_val, red_idx, out, False)
        if fkey is not None:
            # the entry holds ``x`` itself (not only the converted copy): while it lives, the address
            # the key was built from cannot be handed to another same-shaped temporary
            cache[fkey] = (x_in, ys, ybits, red_val, red_idx, ctx.bcn, ctx.packed_in, out, x_key)
            return out.detach()
        if tap is not None:
            return out, tap
        return out

    @staticmethod
    def _forward_segments(ctx, prec, spec, xs, gb, layers, params):
        """Forward with a multi-segment first layer: every segment is a [rows, k_i] fp32 matrix
        (k_i a multiple of 64 in the 16-bit modes); no reduction, no tap."""
        if spec.reduce is not None or spec.tap is not None:
            raise ValueError("a multi-segment input supports plain chains only")
        segs = []
        for x in xs:
            if x.dim() != 2:
                raise ValueError("input segments must be [rows, k] matrices")
            if x.dtype != torch.float32 or x.stride(1) != 1:
                x = x.contiguous().float()
            if prec.scaled and x.shape[0] >= 128:
                x = ops.convert(x, prec.act_dtype, cols_pad=(x.shape[1] + 63) // 64 * 64)
            segs.append(x)
        if gb is not None:
            gb = gb.contiguous().float()
        ybits = []
        ys = chain_forward(prec, segs, layers, final_fp32=True, rows_per_group=spec.group if gb is not None else 0,
                           group_bias=gb, bits=ybits)
        ctx.prec, ctx.spec = prec, spec
        ctx.bcn, ctx.packed_in = None, False
        ctx.n_x, ctx.n_ys = len(segs), len(ys)
        ctx.seg_widths = [x.shape[1] for x in xs]
        ctx.has_red, ctx.has_gb = False, gb is not None
        ctx.bit_slots = [i for i, t in enumerate(ybits) if t is not None]
        ctx.save_for_backward(*(segs + ys + list(params) + [ybits[i] for i in ctx.bit_slots]))
        _tape(spec, segs[0], ys, None, None, ys[-1], False)
        return ys[-1]

    @staticmethod
    def backward(ctx, d_out, d_tap=None):
        prec, spec = ctx.prec, ctx.spec
        if spec.extra_segs:
            return PointMLPFunction._backward_segments(ctx, d_out)
        sv = list(ctx.saved_tensors)
        nb = len(ctx.bit_slots)
        ybits = [None] * ctx.n_ys
        for slot, t in zip(ctx.bit_slots, sv[len(sv) - nb:]):
            ybits[slot] = t
        sv = sv[:len(sv) - nb]
        x_in, ys = sv[0], sv[1:1 + ctx.n_ys]
        pos = 1 + ctx.n_ys
        red_val = red_idx = None
        if ctx.has_red:
            red_val, red_idx = sv[pos], sv[pos + 1]
            pos += 2
        params = sv[pos:]
        nl = len(spec.acts)
        layers = [Layer(params[2 * i], params[2 * i + 1], *spec.acts[i]) for i in range(nl)]
        need = ctx.needs_input_grad[4:]
        need_w = [need[2 * i] for i in range(nl)]
        need_b = [need[2 * i + 1] for i in range(nl)]
        need_x = ctx.needs_input_grad[2]
        if ctx.packed_in and spec.box is None:
            # a packed 16-bit map that reaches the discriminator as a fresh leaf (the history pool's
            # clone, utils/image_pool.py:53-55): nobody consumes its gradient, so none is formed
            need_x = False
        dev = x_in.device
        grads = [(None, None)] * nl

        srcs = []
        if d_out is not None:
            d_out = d_out.contiguous().float()
            srcs.append(d_out.reshape(d_out.shape[0], d_out.numel() // max(d_out.shape[0], 1)))
        if d_tap is not None:
            d_tap = d_tap.contiguous().float()
            srcs.append(d_tap)
        if not srcs:
            return (None,) * (4 + 2 * nl)
        scale2 = ops.amax_scale(srcs) if prec.scaled else None
        S = scale2[0:1] if prec.scaled else None
        inv = scale2[1:2] if prec.scaled else None

        pool = None
        if any(need_w) or any(need_b):
            pad = lambda n: (n + 63) // 64 * 64
            pool = ZeroPool(ZeroPool.size_for([(pad(L.w.shape[0]), pad(L.w.shape[1])) for L in layers]), dev)
        addends = {}
        if d_tap is not None:
            addends[spec.tap] = d_tap * S if prec.scaled else d_tap

        body = layers if spec.reduce is None else layers[:-1]
        dz_last = None
        dx = dz0 = None
        db_residual = None
        if spec.reduce is None:
            if d_out is not None:
                L = layers[-1]
                dz_last = prepare_dz(prec, d_out, scale2, mask=ys[-1], mask_act=L.act,
                                     mask_slope=L.slope)
                if prec.scaled and L.act == ACT_NONE and need_b[-1] and d_out.dim() == 2 and d_out.shape[1] <= 64:
                    # the last bias gradient is a plain column sum of d_out: give back what the
                    # 16-bit dz dropped, so that it is the exact fp32 sum (coherent roundings of
                    # nearly constant columns would otherwise add up over the rows)
                    db_residual = ops.round_residual(d_out, S, prec.act_dtype)
        elif d_out is not None:
            L = layers[-1]
            src = ys[-1] if ys else x_in
            if spec.reduce == "channels" and body and src.shape[1] % 8 == 0 and \
                    L.w.shape[0] * src.shape[1] <= 48000 and (len(body) - 1) not in addends:
                # only channel idx[r] of a row carries gradient: gather / scatter kernels instead
                # of a dense one-hot dz and two GEMMs over it
                n, k_true = L.w.shape
                dy = d_out.reshape(-1)
                lb = len(body) - 1
                if need_w[-1] and need_b[-1] and prec.scaled and n % 64 == 0 and n <= 256 and \
                        _ONEHOT_LEVEL and ops.backlevel_eligible(prec, [n], src, ybits[lb]):
                    # one pass over the pooled layer's input: its weight / bias gradient and the dz of the
                    # layer below, the one-hot dz built inside the kernel (pcadv_backlevel, one-hot mode)
                    P_ = body[-1]
                    dw = pool.take(n, src.shape[1])
                    db = pool.take(n)
                    wt = dgrad_weight(prec, [L.w], src.shape[1], [n])
                    dz_last = ops.backlevel(None, wt, src, mask_bits=ybits[lb], mask_act=P_.act,
                                            mask_slope=P_.slope, dws=[dw], dbiases=[db], scale=inv,
                                            onehot=(dy.contiguous().float(), red_val.reshape(-1), red_idx.reshape(-1),
                                                    n, L.act, L.slope, S))
                    grads[-1] = (dw[:, :k_true], db)
                elif need_w[-1] or need_b[-1]:
                    dw = pool.take(n, src.shape[1]) if need_w[-1] else None
                    db = pool.take(n) if need_b[-1] else None
                    ops.rowmax_wgrad(dy, red_val, red_idx, src, n, act=L.act, slope=L.slope, dw=dw, dbias=db)
                    grads[-1] = (dw[:, :k_true] if dw is not None else None, db)
                if dz_last is None:
                    P_ = body[-1]
                    w_pad = compute_weight(prec, L.w, [src.shape[1]], n)
                    dz_last = ops.rowmax_dgrad(dy, red_val, red_idx, w_pad, src, act=L.act, slope=L.slope,
                                           scale=S, prev_act=P_.act, prev_slope=P_.slope,
                                           out_dtype=prec.act_dtype)
            elif spec.reduce == "channels":
                n = L.w.shape[0]
                dz_red = ops.rowmax_bwd(d_out.reshape(-1), red_val, red_idx, n, act=L.act,
                                        slope=L.slope, scale=S, out_dtype=prec.act_dtype)
                grads[-1] = layer_wgrad(prec, dz_red, [src], L.w.shape, need_w[-1], need_b[-1], scale2)
                if body or need_x:
                    wt = dgrad_weight(prec, [L.w], L.w.shape[1], [dz_red.shape[1]])
                    if body:
                        P_ = body[-1]
                        dz_last, _, _ = ops.linear([dz_red], wt, mask=src, mask_act=P_.act,
                                                   mask_slope=P_.slope, out_dtype=prec.act_dtype,
                                                   engine=prec.engine,
                                                   addend=addends.pop(len(body) - 1, None))
                    else:
                        dx, _, _ = ops.linear([dz_red], wt, out_scale=inv, engine=prec.engine)
            else:   # max over the cloud's points: sparse backward through the saved argmax
                n, k = L.w.shape
                dg = d_out * S if prec.scaled else d_out
                dw = torch.zeros((n, k), dtype=torch.float32, device=dev) if need_w[-1] else None
                db = torch.zeros((n,), dtype=torch.float32, device=dev) if need_b[-1] else None
                dx_acc = None
                if body or need_x:
                    dx_acc = torch.zeros((src.shape[0], k), dtype=torch.float32, device=dev)
                ops.maxpool_bwd(dg.contiguous(), red_val, red_idx, src, L.w, spec.group, act=L.act,
                                slope=L.slope, dw=dw, dbias=db, dx_acc=dx_acc, scale=inv)
                grads[-1] = (dw, db)
                if body:
                    P_ = body[-1]
                    ad = addends.pop(len(body) - 1, None)
                    if ad is not None:
                        dx_acc = dx_acc + ad
                    dz_last = ops.convert(dx_acc, prec.act_dtype, mask=src, mask_act=P_.act,
                                          mask_slope=P_.slope)
                elif need_x:
                    dx = dx_acc * inv if prec.scaled else dx_acc
        if body and dz_last is None and addends:
            # only the tap carries gradient: start the chain at the tapped layer
            t = spec.tap
            P_ = body[t]
            dz_t = ops.convert(addends.pop(t), prec.act_dtype, mask=ys[t], mask_act=P_.act,
                               mask_slope=P_.slope)
            g2, dx, dz0 = chain_backward(prec, dz_t, [x_in], ys[:t + 1], body[:t + 1],
                                         need_w[:t + 1], need_b[:t + 1], need_x, scale2,
                                         dx_packed=ctx.packed_in, bits=ybits[:t + 1], pool=pool)
            for i, gr in enumerate(g2):
                grads[i] = gr
        elif body and dz_last is not None:
            g2, dx, dz0 = chain_backward(prec, dz_last, [x_in], ys[:len(body)], body,
                                         need_w[:len(body)], need_b[:len(body)], need_x, scale2,
                                         addends=addends, dx_packed=ctx.packed_in, bits=ybits[:len(body)],
                                         pool=pool)
            for i, gr in enumerate(g2):
                grads[i] = gr
        if db_residual is not None and grads[-1] is not None and grads[-1][1] is not None:
            grads[-1][1].add_(db_residual[:grads[-1][1].shape[0]] * inv)
        dgb = None
        if ctx.has_gb and ctx.needs_input_grad[3] and dz0 is not None:
            rows, n0 = dz0.shape[0], layers[0].w.shape[0]
            dgb = torch.zeros((rows // spec.group, dz0.shape[1]), dtype=torch.float32, device=dev)
            ops.wgrad(dz0, [], dgroup_bias=dgb, rows_per_group=spec.group)
            dgb = dgb[:, :n0] * inv if prec.scaled else dgb[:, :n0]
        flat = []
        for i in range(nl):
            dw, db = grads[i] if grads[i] is not None else (None, None)
            flat.append(dw.reshape(params[2 * i].shape) if dw is not None else None)
            flat.append(db)
        if ctx.packed_in and need_x:
            if dx is None or dx.dtype != x_in.dtype:
                raise RuntimeError("a packed 16-bit input that requires grad needs a layer chain in "
                                   "front of the reduction")
            spec.box.scale2 = scale2
        if dx is not None and ctx.bcn is not None:
            B_, C_, N_ = ctx.bcn                                # back to B x C x N (a view)
            dx = dx.reshape(B_, N_, dx.shape[1])[:, :, :C_].transpose(1, 2)
        return (None, None, dx, dgb, *flat)


    @staticmethod
    def _backward_segments(ctx, d_out):
        prec, spec = ctx.prec, ctx.spec
        sv = list(ctx.saved_tensors)
        nb = len(ctx.bit_slots)
        ybits = [None] * ctx.n_ys
        for slot, t in zip(ctx.bit_slots, sv[len(sv) - nb:]):
            ybits[slot] = t
        sv = sv[:len(sv) - nb]
        segs, ys = sv[:ctx.n_x], sv[ctx.n_x:ctx.n_x + ctx.n_ys]
        params = sv[ctx.n_x + ctx.n_ys:]
        nl = len(spec.acts)
        layers = [Layer(params[2 * i], params[2 * i + 1], *spec.acts[i]) for i in range(nl)]
        need = ctx.needs_input_grad[4:4 + 2 * nl]
        need_w = [need[2 * i] for i in range(nl)]
        need_b = [need[2 * i + 1] for i in range(nl)]
        need_x = ctx.needs_input_grad[2] or any(ctx.needs_input_grad[4 + 2 * nl:])
        n_out = 4 + 2 * nl + spec.extra_segs
        if d_out is None:
            return (None,) * n_out
        d_out = d_out.contiguous().float()
        scale2 = ops.amax_scale(d_out) if prec.scaled else None
        S = scale2[0:1] if prec.scaled else None
        inv = scale2[1:2] if prec.scaled else None
        pool = None
        if any(need_w) or any(need_b):
            pad = lambda n: (n + 63) // 64 * 64
            pool = ZeroPool(ZeroPool.size_for([(pad(L.w.shape[0]), pad(L.w.shape[1])) for L in layers]), d_out.device)
        L = layers[-1]
        dz_last = prepare_dz(prec, d_out, scale2, mask=ys[-1], mask_act=L.act, mask_slope=L.slope)
        db_residual = None
        if prec.scaled and L.act == ACT_NONE and need_b[-1] and d_out.shape[1] <= 64:
            db_residual = ops.round_residual(d_out, S, prec.act_dtype)
        grads, dx, dz0 = chain_backward(prec, dz_last, list(segs), ys, layers, need_w, need_b, need_x, scale2,
                                        bits=ybits, pool=pool)
        if db_residual is not None and grads[-1] is not None and grads[-1][1] is not None:
            grads[-1][1].add_(db_residual[:grads[-1][1].shape[0]] * inv)
        dgb = None
        if ctx.has_gb and ctx.needs_input_grad[3]:
            rows, n0 = dz0.shape[0], layers[0].w.shape[0]
            dgb = torch.zeros((rows // spec.group, dz0.shape[1]), dtype=torch.float32, device=d_out.device)
            ops.wgrad(dz0, [], dgroup_bias=dgb, rows_per_group=spec.group)
            dgb = dgb[:, :n0] * inv if prec.scaled else dgb[:, :n0]
        flat = []
        for i in range(nl):
            dw, db = grads[i] if grads[i] is not None else (None, None)
            flat.append(dw.reshape(params[2 * i].shape) if dw is not None else None)
            flat.append(db)
        dxs, off = [], 0
        for w_ in ctx.seg_widths:                       # the K-concat's gradient, segment by segment (views)
            dxs.append(dx[:, off:off + w_] if dx is not None else None)
            off += w_
        return (None, None, dxs[0], dgb, *flat, *dxs[1:])


def point_mlp(prec, x, layers, acts, reduce=None, group=0, tap=None, group_bias=None, box=None):
    """Convenience wrapper: ``layers`` are nn.Conv1d / nn.Linear modules or
    (weight, bias | None) pairs.  ``x`` may be a list of [rows, k_i] matrices: the first layer then
    reads their K-concat without it being built."""
    params = []
    for m in layers:
        params += [m.weight, m.bias] if hasattr(m, "weight") else [m[0], m[1]]
    if isinstance(x, (list, tuple)):
        if len(x) > 1:
            return PointMLPFunction.apply(prec, MLPSpec(acts, reduce, group, tap, box, extra_segs=len(x) - 1), x[0],
                                          group_bias, *params, *x[1:])
        x = x[0]
    return PointMLPFunction.apply(prec, MLPSpec(acts, reduce, group, tap, box), x, group_bias, *params)


_TNET_MAX_K = 128


class BmmFunction(torch.autograd.Function):
    """y[b] = x[b] @ T[b]: torch.bmm(x^T, trans) of models/pointnet.py:120-122 /
    :231 / :238 on point-major x [B, N, k] (fp32) and T [B, k, k] -- one batched kernel per
    direction (``pcadv_bmm`` / ``pcadv_bmm_tgrad``)."""

    @staticmethod
    def forward(ctx, x, trans):
        B, N, k = x.shape
        if k > _TNET_MAX_K:
            raise ValueError("T-Net transforms are implemented for k <= %d" % _TNET_MAX_K)
        x = x.contiguous().float()
        trans = trans.contiguous().float()
        ctx.save_for_backward(x, trans)
        return ops.bmm(x, trans)

    @staticmethod
    def backward(ctx, dy):
        x, trans = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = ops.bmm(dy, trans, transpose_t=True) if ctx.needs_input_grad[0] else None   # dy @ T^T
        dt = ops.bmm_tgrad(x, dy) if ctx.needs_input_grad[1] else None                   # x^T @ dy
        return dx, dt


class RegularizerFunction(torch.autograd.Function):
    """mean_b || T T^T - I ||_F (models/pointnet.py:345-353): one CTA per cloud forward
    (``pcadv_ortho_reg``) and backward (``pcadv_ortho_reg_bwd``)."""

    @staticmethod
    def forward(ctx, trans):
        t = trans.contiguous().float()
        diff, norms = ops.ortho_reg(t)
        ctx.save_for_backward(t, diff, norms)
        return norms.mean()

    @staticmethod
    def backward(ctx, dloss):
        t, diff, norms = ctx.saved_tensors
        # d||M||_F / dM = M / ||M||;  M = T T^T - I symmetric  =>  dT = 2 (dloss / (B ||M||)) M T
        return ops.ortho_reg_bwd(diff, t, norms, dloss.contiguous().float())


class LogSoftmaxRowsFunction(torch.autograd.Function):
    """log_softmax over the columns of point-major logits [rows, n] (fp32), forward and backward on
    the loss-head kernels (``pcadv_softmax_head`` / ``pcadv_logsoftmax_bwd``): the
    ``F.log_softmax(x.view(-1, k), dim=-1)`` of PointNetDenseCls (models/pointnet.py:341)."""

    @staticmethod
    def forward(ctx, x):
        if x.stride(1) != 1 or x.dtype != torch.float32:
            x = x.contiguous().float()
        lp, _ = ops.softmax_head(x, ops.HEAD_LSM, out_dtype=torch.float32)
        ctx.save_for_backward(lp)
        return lp

    @staticmethod
    def backward(ctx, dy):
        (lp,) = ctx.saved_tensors
        if dy.stride(1) != 1 or dy.dtype != torch.float32:
            dy = dy.contiguous().float()
        return ops.logsoftmax_bwd(lp, dy, lp.shape[1])


class LseRatioFunction(torch.autograd.Function):
    """StackDiscNet.custom_activation on point-major shape logits [rows, S]: z / (z + 1) with
    z = logsumexp over S (models/discriminator.py:153-159), one kernel each way (``pcadv_lse_ratio``)."""

    @staticmethod
    def forward(ctx, x):
        if x.stride(1) != 1 or x.dtype != torch.float32:
            x = x.contiguous().float()
        ctx.save_for_backward(x)
        return ops.lse_ratio(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.lse_ratio(x, dy.contiguous().float())


__all__ = ["MLPSpec", "PointMLPFunction", "point_mlp", "BmmFunction", "RegularizerFunction",
           "LogSoftmaxRowsFunction", "LseRatioFunction", "ACT_NONE"]
