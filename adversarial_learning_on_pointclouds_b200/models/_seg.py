"""PointNetSeg forward / backward as one autograd Function on libpcadv ops.

Mirrors models/pointnet.py:282-317 of the reference with three structural
changes that leave the mathematics unchanged:

* the 2048-wide conv6 output is never written: the layer, its ReLU and the max
  over the cloud's points run as one kernel that emits (max, first argmax);
* the 3024-channel concat is never built: fc1 reads x1..x5 as five K-segments
  and the tiled global-feature / class-one-hot columns become a per-cloud bias
  ``fc1.W[:, 960:] @ [g; cls] + fc1.b``;
* the backward through the max uses the saved argmax only (sparse scatter), and
  each trunk layer's dgrad is one GEMM over the K-concat [dz_next | dz_fc1].
"""
import os

import torch

from .. import ops
from ..ops import ACT_NONE, ACT_RELU, ENGINE_SIMT, HEAD_CE, HEAD_LSM
from ._chain import (Layer, ZeroPool, chain_forward, compute_weight, dgrad_weight, layer_wgrad, prepare_dz)

_TRUNK = ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6")
_HEAD = ("fc1", "fc2", "fc3", "fc4")
PARAM_NAMES = tuple(n + s for n in _TRUNK + _HEAD for s in (".weight", ".bias"))
_SLICES = ((0, 64), (64, 192), (192, 320), (320, 448), (448, 960))   # x1..x5 inside the concat
_G0, _G1, _C1 = 960, 3008, 3024                                       # global / class columns


# Backward levels of PointNetSeg that run as one pcadv_backlevel launch (dgrad + the weight gradients of
# the layers the level's activation feeds); the others keep separate dgrad / wgrad launches.  Measured
# back to back at 2^21 points (tools/level_ab.py): the narrow head levels win (fc4 0.34 -> 0.27 ms, fc3
# 0.55 -> 0.50 ms); fc2 (K = n = 256) and the trunk levels with the fc1 K-segment (K = 384 / 768) do not --
# their weight-gradient tiles fill TMEM, so the level runs sliced or with one dgrad accumulator and a
# two-slot ring, and it loses to two launches that each stream at the HBM rate.  Tuning aid: PCADV_LEVELS
# (e.g. "fc4,fc3,fc2,l4,l3,l2,l1").
_LEVELS = frozenset(x for x in os.environ.get("PCADV_LEVELS", "fc4,fc3").split(",") if x)

# d(cbias) as a by-product of the dz5 dgrad launch instead of a separate column-sum pass (tuning aid: PCADV_DCB_FROM_DGRAD=0)
_DCB_FROM_DGRAD = os.environ.get("PCADV_DCB_FROM_DGRAD", "1") != "0"

HEAD_GAIN = 256.0        # the fused CE head stores 256 * (softmax - onehot) as the 16-bit dz


class GradBox:
    """Side channel between the log-softmax head of the generator and the discriminator that
    consumes its packed 16-bit output: the discriminator's backward leaves the power-of-two
    scale of the gradient it returns here (``scale2`` = device tensor [S, 1/S])."""

    __slots__ = ("scale2",)

    def __init__(self):
        self.scale2 = None


def pad64(n):
    return (n + 63) // 64 * 64


class SegFunction(torch.autograd.Function):
    """(pts [B,N,3], cls [B,1,16], *params) -> (head output, global [B,2048] fp32).

    head = "logits": logits [B,N,k] fp32 (models/pointnet.py:314-317).
    head = "ce"    : (CrossEntropyLoss(pred, labels) as a 0-d tensor, softmax(pred) as a
                     non-differentiable discriminator input) -- utils/trainer.py:899-901 in one
                     pass over the logits; the loss gradient (softmax - onehot) is kept instead
                     of the logits.
    head = "lsm"   : log_softmax(pred) as a discriminator input -- utils/trainer.py:914.
    Discriminator inputs are packed point-major [B, N, pad64(k)] 16-bit tensors in the 16-bit
    precision modes and B x k x N fp32 views in the fp32 mode."""

    @staticmethod
    def forward(ctx, prec, debug, head, labels, pts, cls, *params):
        if not pts.is_cuda:
            raise RuntimeError("PointNetSeg (libpcadv) needs CUDA tensors; there is no CPU path")
        p = dict(zip(PARAM_NAMES, params))
        B, N, _ = pts.shape
        P = B * N
        pts2 = pts.reshape(P, 3).contiguous().float()
        cls2 = cls.reshape(B, cls.shape[-1] * (cls.shape[1] if cls.dim() == 3 else 1)).contiguous().float()
        W = {n: p[n + ".weight"].reshape(p[n + ".weight"].shape[0], -1) for n in _TRUNK + _HEAD}
        b = {n: p[n + ".bias"] for n in _TRUNK + _HEAD}

        # trunk: conv1 (K=3, CUDA-core, HBM-bound) then conv2..conv5
        xbits, hbits = [], []
        xs = chain_forward(prec, [pts2], [Layer(W[n], b[n], ACT_RELU) for n in _TRUNK[:5]], bits=xbits)
        x5 = xs[4]
        # conv6 + ReLU + max over the cloud, fused; the B x 2048 x N map is never stored
        w6 = compute_weight(prec, W["conv6"], [512], 2048)
        _, key, _ = ops.linear([x5], w6, bias=b["conv6"], want_out=False, colmax=True,
                               rows_per_group=N, engine=prec.engine)
        g, idx = ops.max_finalize(key, ACT_RELU)
        # fold the tiled global feature and class one-hot into a per-cloud bias
        cbias, _, _ = ops.linear([g, cls2], W["fc1"][:, _G0:_C1], bias=b["fc1"], engine=ENGINE_SIMT)
        hs = chain_forward(prec, xs, [Layer(W["fc1"][:, :_G0], None, ACT_RELU),
                                      Layer(W["fc2"], b["fc2"], ACT_RELU),
                                      Layer(W["fc3"], b["fc3"], ACT_RELU),
                                      Layer(W["fc4"], b["fc4"], ACT_NONE)],
                           final_fp32=True, rows_per_group=N, group_bias=cbias, bits=hbits)
        logits = hs[3].view(B, N, hs[3].shape[1])
        k_out = logits.shape[2]
        cols = pad64(k_out) if prec.scaled else k_out
        ctx.prec, ctx.shape, ctx.head, ctx.box = prec, (B, N), head, None
        if head == "logits":
            out, extra, keep = logits, None, None
        elif head == "ce":
            # [CE sum, number of rows with a label in [0, k)]: rows labelled outside that range are
            # nn.CrossEntropyLoss's ignored rows (ignore_index = -100) and the mean is over the rest
            acc = torch.zeros(2, dtype=torch.float32, device=pts.device)
            probs, u = ops.softmax_head(hs[3], HEAD_CE, labels=labels.reshape(-1).contiguous(),
                                        out_dtype=prec.act_dtype, cols=cols, want_dz=True,
                                        dz_gain=HEAD_GAIN if prec.scaled else 1.0, loss_sum=acc[0:1],
                                        valid_count=acc[1:2])
            ctx.ce_count = acc[1].clamp_min(1.0)
            out = acc[0] / ctx.ce_count
            extra = probs.view(B, N, cols) if prec.scaled else probs.view(B, N, k_out).transpose(1, 2)
            keep = u
        elif head == "lsm":
            lp, _ = ops.softmax_head(hs[3], HEAD_LSM, out_dtype=prec.act_dtype, cols=cols)
            out = lp.view(B, N, cols) if prec.scaled else lp.view(B, N, k_out).transpose(1, 2)
            extra, keep = None, lp
            ctx.box = GradBox()
        elif head == "ce+lsm":
            # one pass over the labelled clouds (the first labels.shape[0] of the batch: CE + softmax,
            # utils/trainer.py:898-901) and the unlabelled ones (log_softmax, :913-914) of an iteration
            Bg = labels.shape[0]
            Pg = Bg * N
            acc = torch.zeros(2, dtype=torch.float32, device=pts.device)
            dt = prec.act_dtype
            # joint = [softmax(labelled) ; log_softmax(unlabelled)], what the discriminator reads;
            # dzbuf = [gain * (softmax - onehot) ; filled by the backward], the dz of fc4
            joint = torch.empty((P, cols), dtype=dt, device=pts.device)
            dzbuf = torch.empty((P, cols), dtype=dt, device=pts.device)
            if Pg > 0:
                ops.softmax_head(hs[3][:Pg], HEAD_CE, labels=labels.reshape(-1).contiguous(), out_dtype=dt,
                                 cols=cols, dz_gain=HEAD_GAIN if prec.scaled else 1.0, loss_sum=acc[0:1],
                                 valid_count=acc[1:2], probs_out=joint[:Pg], dz_out=dzbuf[:Pg])
            if P > Pg:
                ops.softmax_head(hs[3][Pg:], HEAD_LSM, out_dtype=dt, cols=cols, probs_out=joint[Pg:])
            ctx.ce_count = acc[1].clamp_min(1.0)
            ctx.split = Bg
            out = acc[0] / ctx.ce_count
            if prec.scaled:
                extra = joint[:Pg].view(Bg, N, cols)
                lp_out = joint[Pg:].view(B - Bg, N, cols)
            else:
                extra = joint[:Pg].view(Bg, N, k_out).transpose(1, 2)
                lp_out = joint[Pg:].view(B - Bg, N, k_out).transpose(1, 2)
            ctx.box = GradBox()
            ctx.has_keep = True
            bit_list = xbits + hbits[:3]
            ctx.bit_slots = [i for i, t in enumerate(bit_list) if t is not None]
            ctx.save_for_backward(pts2, cls2, g, idx, *xs, *hs[:3], *params, joint, dzbuf,
                                  *[bit_list[i] for i in ctx.bit_slots])
            if debug is not None:
                debug.update(x=xs, h=hs[:3], g=g, idx=idx, cbias=cbias, logits=logits)
            ctx.mark_non_differentiable(extra)
            return out, extra, lp_out, g
        else:
            raise ValueError("head must be 'logits', 'ce', 'lsm' or 'ce+lsm'")
        ctx.has_keep = keep is not None
        # 1-bit activation masks [x > 0] of x1..x5 / h1..h3 where the layer could emit them
        bit_list = xbits + hbits[:3]
        ctx.bit_slots = [i for i, t in enumerate(bit_list) if t is not None]
        ctx.save_for_backward(pts2, cls2, g, idx, *xs, *hs[:3], *params,
                              *([keep] if keep is not None else []),
                              *[bit_list[i] for i in ctx.bit_slots])
        if debug is not None:
            debug.update(x=xs, h=hs[:3], g=g, idx=idx, cbias=cbias, logits=logits)
        if head == "ce":
            ctx.mark_non_differentiable(extra)
            return out, extra, g
        return out, g

    @staticmethod
    def _head_dz(ctx, prec, keep, d_out, B, N, k_out, dev):
        """(dz of fc4 [P, pad] in the chain's dtype, scale2 | None) from the head's gradient."""
        P = B * N
        if ctx.head == "ce":
            if d_out is None:
                d_out = torch.zeros((), dtype=torch.float32, device=dev)
            d_out = d_out.float().reshape(())
            cnt = ctx.ce_count                                     # rows that are not ignored
            if prec.scaled:
                # dz_true = u * d_out / (GAIN * cnt): the kept u is the scaled dz with S = GAIN * cnt / d_out
                inv = d_out / (HEAD_GAIN * cnt)
                safe = torch.where(inv == 0, torch.ones_like(inv), inv)
                return keep, torch.stack([1.0 / safe, inv])
            return keep * (d_out / cnt), None
        if ctx.head == "ce+lsm":
            return SegFunction._dual_head_dz(ctx, prec, keep, d_out, B, N, k_out, dev)
        # "lsm": d_out is the gradient of the packed log-softmax map
        if prec.scaled:
            scale2 = ctx.box.scale2
            ctx.box.scale2 = None
            if d_out is None:
                d_out = torch.zeros((P, keep.shape[1]), dtype=prec.act_dtype, device=dev)
            d2 = d_out.reshape(P, keep.shape[1])
            if scale2 is None:
                # gradient from a consumer that does not speak the GradBox protocol: unscaled
                d2 = d2.float().contiguous()
                scale2 = ops.amax_scale(d2)
                return ops.logsoftmax_bwd(keep, d2, k_out, scale=scale2[0:1], out_dtype=prec.act_dtype,
                                          cols=keep.shape[1]), scale2
            if d2.stride(1) != 1:
                d2 = d2.contiguous()
            return ops.logsoftmax_bwd(keep, d2, k_out, out_dtype=prec.act_dtype, cols=keep.shape[1]), scale2
        if d_out is None:
            d_out = torch.zeros((B, k_out, N), dtype=torch.float32, device=dev)
        d2 = d_out.transpose(1, 2).reshape(P, k_out)
        if d2.stride(1) != 1 or d2.dtype != torch.float32:
            d2 = d2.contiguous().float()
        return ops.logsoftmax_bwd(keep, d2, k_out), None

    @staticmethod
    def _dual_head_dz(ctx, prec, keep, d_out, B, N, k_out, dev):
        """dz of fc4 for the "ce+lsm" head: rows of the labelled clouds carry the CE gradient the
        forward stored, rows of the unlabelled ones the log_softmax backward of the discriminator's
        gradient.  Both halves share ONE gradient scale, the CE half's S = GAIN * count / dloss (the
        adversarial rows are rescaled from the scale the discriminator's backward left in the box)."""
        joint, dzbuf = keep
        d_loss, d_lp = d_out
        Pg = ctx.split * N
        P = B * N
        if d_loss is None:
            d_loss = torch.zeros((), dtype=torch.float32, device=dev)
        d_loss = d_loss.float().reshape(())
        cnt = ctx.ce_count
        lp = joint[Pg:]
        if prec.scaled:
            inv = d_loss / (HEAD_GAIN * cnt)
            safe = torch.where(inv == 0, torch.ones_like(inv), inv)
            scale2 = torch.stack([1.0 / safe, inv])
            box2 = ctx.box.scale2
            ctx.box.scale2 = None
            if P > Pg:
                if d_lp is None:
                    dzbuf[Pg:].zero_()
                else:
                    d2 = d_lp.reshape(P - Pg, joint.shape[1])
                    if d2.stride(1) != 1:
                        d2 = d2.contiguous()
                    # incoming rows carry the discriminator's scale S_D: bring them to S
                    ratio = scale2[0:1] * box2[1:2] if box2 is not None else scale2[0:1].clone()
                    ops.logsoftmax_bwd(lp, d2, k_out, scale=ratio, cols=joint.shape[1], out=dzbuf[Pg:])
            return dzbuf, scale2
        dz_g = dzbuf[:Pg] * (d_loss / cnt)
        if P == Pg:
            return dz_g, None
        if d_lp is None:
            dz_n = torch.zeros((P - Pg, k_out), dtype=torch.float32, device=dev)
        else:
            d2 = d_lp.transpose(1, 2).reshape(P - Pg, k_out)
            if d2.stride(1) != 1 or d2.dtype != torch.float32:
                d2 = d2.contiguous().float()
            dz_n = ops.logsoftmax_bwd(lp, d2, k_out)
        return torch.cat([dz_g, dz_n], 0), None

    @staticmethod
    def backward(ctx, dlogits, *rest):
        prec = ctx.prec
        B, N = ctx.shape
        P = B * N
        dg_ext = rest[-1]
        sv = ctx.saved_tensors
        pts2, cls2, g, idx = sv[0:4]
        xs, hs = list(sv[4:9]), list(sv[9:12])
        nb = len(ctx.bit_slots)
        bit_list = [None] * 8
        for slot, t in zip(ctx.bit_slots, sv[len(sv) - nb:]):
            bit_list[slot] = t
        xbits, hbits = bit_list[:5], bit_list[5:]
        sv = sv[:len(sv) - nb]
        if ctx.head == "ce+lsm":
            keep = (sv[-2], sv[-1])
            params = sv[12:-2]
            dlogits = (dlogits, rest[1])              # (d loss, d log_softmax map); rest[0]: softmax, no grad
        else:
            keep = sv[-1] if ctx.has_keep else None
            params = sv[12:-1] if ctx.has_keep else sv[12:]
        p = dict(zip(PARAM_NAMES, params))
        need = dict(zip(PARAM_NAMES, ctx.needs_input_grad[6:]))
        W = {n: p[n + ".weight"].reshape(p[n + ".weight"].shape[0], -1) for n in _TRUNK + _HEAD}
        dev = pts2.device
        k_out = W["fc4"].shape[0]
        grads = {}

        if ctx.head != "logits":
            dz, scale2 = SegFunction._head_dz(ctx, prec, keep, dlogits, B, N, k_out, dev)
            inv = scale2[1:2] if scale2 is not None else None
            dl_cm = None
        else:
            if dlogits is None:
                dlogits = torch.zeros((B, N, k_out), dtype=torch.float32, device=dev)
            dl_cm = dlogits.transpose(1, 2)                           # B x k x N
        if dl_cm is None:
            pass
        elif dlogits.stride(2) != 1 and dl_cm.is_contiguous() and ops.is_channel_major(dl_cm):
            # the trainer's CE / softmax backward hand the gradient over channel-major
            # (B x k x N contiguous): one kernel transposes, scales, converts and pads it
            scale2 = ops.amax_scale(dl_cm.reshape(B * k_out, N)) if prec.scaled else None
            inv = scale2[1:2] if prec.scaled else None
            if prec.scaled:
                dz = ops.convert_cm(dl_cm, prec.act_dtype, cols_pad=(k_out + 63) // 64 * 64,
                                    scale=scale2[0:1])
            else:
                dz = ops.convert_cm(dl_cm, torch.float32)
        else:
            dl = dlogits.reshape(P, k_out)
            if dl.stride(1) != 1 or dl.dtype != torch.float32:
                dl = dl.contiguous().float()
            scale2 = ops.amax_scale(dl) if prec.scaled else None
            inv = scale2[1:2] if prec.scaled else None
            dz = prepare_dz(prec, dl, scale2)

        # every dW / dbias accumulator of this pass comes out of one zero-filled buffer
        pool = ZeroPool(ZeroPool.size_for([(64, 128), (128, 256), (256, 256), (256, _C1), (B, 256),
                                           (2048, 512), (512, 128), (128, 128), (128, 128), (128, 64),
                                           (64, 64)]), dev)
        # ---- head: fc4, fc3, fc2 ---------------------------------------------------
        # A level = the dgrad through one stored activation + the weight gradients of the layers it
        # feeds.  Where pcadv_backlevel takes the shape (and _LEVELS asks for it) both come out of one
        # pass over dz; otherwise wgrad and dgrad are separate launches that each read dz.
        inv_s = scale2[1:2] if scale2 is not None else None
        head = (("fc4", ACT_NONE), ("fc3", ACT_RELU), ("fc2", ACT_RELU))
        for li, (name, _) in enumerate(head):
            xin = hs[2 - li]
            wt = dgrad_weight(prec, [W[name]], W[name].shape[1], [dz.shape[1]])
            if name in _LEVELS and need[name + ".weight"] and need[name + ".bias"] and \
                    ops.backlevel_eligible(prec, [dz.shape[1]], xin, hbits[2 - li]):
                dw = pool.take(dz.shape[1], xin.shape[1])
                db = pool.take(dz.shape[1])
                dz = ops.backlevel([dz], wt, xin, mask_bits=hbits[2 - li], dws=[dw], dbiases=[db], scale=inv_s)
                n_, k_ = W[name].shape
                grads[name + ".weight"], grads[name + ".bias"] = dw[:n_, :k_], db[:n_]
                continue
            dw, db = layer_wgrad(prec, dz, [xin], W[name].shape, need[name + ".weight"],
                                 need[name + ".bias"], scale2, pool=pool)
            grads[name + ".weight"], grads[name + ".bias"] = dw, db
            dz, _, _ = ops.linear([dz], wt, mask=xin, mask_act=ACT_RELU, out_dtype=prec.act_dtype,
                                  engine=prec.engine, mask_bits=hbits[2 - li])
        dz_fc1 = dz                                                   # [P, 256], scaled

        # ---- fc1: per-point part (x1..x5) and per-cloud part (g, cls, bias) -----------
        need_w1, need_b1 = need["fc1.weight"], need["fc1.bias"]
        dw1 = pool.take(256, _C1) if need_w1 else None
        db1 = pool.take(256) if need_b1 else None
        dcb = pool.take(B, 256)                                        # d(cbias), scaled
        # d(cbias) = the per-cloud column sums of dz_fc1.  Where the dense part of dz5 (below) can take them as a
        # by-product of its pass over dz_fc1 it runs first, and fc1's weight gradient needs no separate
        # column-sum pass over dz_fc1 (group_colsum16: 1 GB per step at cfg5).
        s5 = _SLICES[4]
        wt5 = dgrad_weight(prec, [W["fc1"][:, s5[0]:s5[1]]], 512, [256])
        dz5_dense = None
        inplace5 = ops.maxpool_inplace_eligible(512, N, 2048)
        want_dcb = dcb
        if _DCB_FROM_DGRAD and inplace5 and ops.group_sum_eligible(prec, [dz_fc1], wt5, 512, xbits[4], N):
            dz5_dense, _, _ = ops.linear([dz_fc1], wt5, mask=xs[4], mask_act=ACT_RELU, out_dtype=prec.act_dtype,
                                         engine=prec.engine, mask_bits=xbits[4], rows_per_group=N,
                                         seg0_group_sum=dcb)
            want_dcb = None
        # trunk levels that go through pcadv_backlevel also form their slice of fc1's weight gradient
        # (dz_fc1^T x_k); fc1's own wgrad launch then covers the remaining x segments only
        fused = [False] * 5
        if need_w1:
            for li in range(1, 5):                                     # level li: x_li feeds conv_{li+1} and fc1
                name = _TRUNK[li]
                fused[li - 1] = ("l%d" % li) in _LEVELS and need[name + ".weight"] and need[name + ".bias"] and \
                    ops.backlevel_eligible(prec, [W[name].shape[0], dz_fc1.shape[1]], xs[li - 1], xbits[li - 1])
        rest = [i for i in range(5) if not fused[i]]
        if need_w1 and len(rest) == 5:
            ops.wgrad(dz_fc1, xs, dw=dw1[:, :_G0], dgroup_bias=want_dcb, rows_per_group=N, scale=inv,
                      engine=prec.engine)
        else:
            # runs of adjacent unfused segments share one launch (their dw columns are contiguous)
            first = True
            i = 0
            while need_w1 and i < len(rest):
                j = i
                while j + 1 < len(rest) and rest[j + 1] == rest[j] + 1:
                    j += 1
                c0, c1 = _SLICES[rest[i]][0], _SLICES[rest[j]][1]
                ops.wgrad(dz_fc1, [xs[q] for q in rest[i:j + 1]], dw=dw1[:, c0:c1],
                          dgroup_bias=want_dcb if first else None, rows_per_group=N, scale=inv, engine=prec.engine)
                first = False
                i = j + 1
            if first and want_dcb is not None:                         # nothing left for fc1's own launch
                ops.wgrad(dz_fc1, [], dgroup_bias=dcb, rows_per_group=N, scale=inv, engine=prec.engine)
        if need_w1 or need_b1:
            ops.wgrad(dcb, [g, cls2] if need_w1 else [], dw=dw1[:, _G0:_C1] if need_w1 else None,
                      dbias=db1, scale=inv)
        grads["fc1.weight"], grads["fc1.bias"] = dw1, db1

        # ---- through the max: dg = dcbias @ fc1.W[:, 960:3008] (+ external grad) -------
        wg_t = W["fc1"][:, _G0:_G1].t().contiguous()                  # [2048, 256]
        dg, _, _ = ops.linear([dcb], wg_t, engine=ENGINE_SIMT)        # [B, 2048], scaled
        if dg_ext is not None:
            dg = dg + (dg_ext.float() * scale2[0] if prec.scaled else dg_ext.float())
        dw6 = pool.take(2048, 512) if need["conv6.weight"] else None
        db6 = pool.take(2048) if need["conv6.bias"] else None
        # ---- trunk: dz_k = relu'(x_k) * ([dz_{k+1} | dz_fc1] @ [W_{k+1}; fc1.W[:, slice_k]]) --
        wt = wt5
        w6 = compute_weight(prec, W["conv6"], [512], 2048)
        if inplace5:
            # dense part first; the max-pool's sparse part (argmax rows only) is added in place
            if dz5_dense is not None:
                dz = dz5_dense
            else:
                dz, _, _ = ops.linear([dz_fc1], wt, mask=xs[4], mask_act=ACT_RELU,
                                      out_dtype=prec.act_dtype, engine=prec.engine, mask_bits=xbits[4])
            ops.maxpool_bwd(dg, g, idx, xs[4], w6, N, act=ACT_RELU, dw=dw6, dbias=db6,
                            dz_inout=dz, prev_act=ACT_RELU, scale=inv)
        else:
            dx5_sparse = torch.zeros((P, 512), dtype=torch.float32, device=dev)
            ops.maxpool_bwd(dg, g, idx, xs[4], w6, N, act=ACT_RELU, dw=dw6, dbias=db6,
                            dx_acc=dx5_sparse, scale=inv)
            dz, _, _ = ops.linear([dz_fc1], wt, addend=dx5_sparse, mask=xs[4], mask_act=ACT_RELU,
                                  out_dtype=prec.act_dtype, engine=prec.engine)
            del dx5_sparse
        grads["conv6.weight"], grads["conv6.bias"] = dw6, db6
        for li in range(4, 0, -1):                                    # conv5 .. conv2
            name = _TRUNK[li]
            xin = xs[li - 1]
            sl = _SLICES[li - 1]
            wt = dgrad_weight(prec, [W[name], W["fc1"][:, sl[0]:sl[1]]], W[name].shape[1],
                              [dz.shape[1], 256])
            if fused[li - 1]:
                dw = pool.take(dz.shape[1], xin.shape[1])
                db = pool.take(dz.shape[1])
                dz = ops.backlevel([dz, dz_fc1], wt, xin, mask_bits=xbits[li - 1],
                                   dws=[dw, dw1[:, sl[0]:sl[1]]], dbiases=[db, None], scale=inv)
                n_, k_ = W[name].shape
                grads[name + ".weight"], grads[name + ".bias"] = dw[:n_, :k_], db[:n_]
                continue
            dw, db = layer_wgrad(prec, dz, [xin], W[name].shape, need[name + ".weight"],
                                 need[name + ".bias"], scale2, pool=pool)
            grads[name + ".weight"], grads[name + ".bias"] = dw, db
            dz, _, _ = ops.linear([dz, dz_fc1], wt, mask=xin, mask_act=ACT_RELU,
                                  out_dtype=prec.act_dtype, engine=prec.engine, mask_bits=xbits[li - 1])
        dw, db = layer_wgrad(prec, dz, [pts2], W["conv1"].shape, need["conv1.weight"],
                             need["conv1.bias"], scale2, pool=pool)
        grads["conv1.weight"], grads["conv1.bias"] = dw, db

        dpts = None
        if ctx.needs_input_grad[4]:
            wt = W["conv1"].t().contiguous()                          # [3, 64]
            dpts, _, _ = ops.linear([dz], wt, out_scale=inv, engine=ENGINE_SIMT)
            dpts = dpts.view(B, N, 3)
        dcls = None
        if ctx.needs_input_grad[5]:
            wc_t = W["fc1"][:, _G1:_C1].t().contiguous()              # [16, 256]
            dcls, _, _ = ops.linear([dcb], wc_t, out_scale=inv, engine=ENGINE_SIMT)
            dcls = dcls.view(B, 1, dcls.shape[-1])

        out = []
        for n_ in PARAM_NAMES:
            gr = grads.get(n_)
            if gr is not None:
                gr = gr.reshape(p[n_].shape)
            out.append(gr)
        return (None, None, None, None, dpts, dcls, *out)
