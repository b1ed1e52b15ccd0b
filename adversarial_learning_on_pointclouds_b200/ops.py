"""Tensor-level wrappers over the libpcadv C ABI.

torch is used here for device memory (torch.empty / torch.zeros), the current
CUDA stream and dtype bookkeeping only; every arithmetic op is a libpcadv call.
All matrices are point-major ``[rows, channels]`` with unit stride on channels.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, ENGINE_SIMT, ENGINE_TC, F16, F32, BF16, HEAD_CE, HEAD_LSM)

_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}


class Precision:
    """Arithmetic mode of the per-point layers.

    fp32 : fp32 storage, fp32 FFMA accumulation (verification mode, 1e-5).
    fp16 : fp16 storage, tcgen05 kind::f16 with fp32 accumulation; gradients are
           carried with a power-of-two dynamic scale (default fast mode, 1e-3).
    bf16 : bf16 storage, same kernels (wide-range mode, ~1e-2).
    """

    def __init__(self, name):
        if name not in ("fp32", "fp16", "bf16"):
            raise ValueError("precision must be fp32, fp16 or bf16, got %r" % (name,))
        self.name = name
        self.act_dtype = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}[name]
        self.scaled = name != "fp32"
        want_tc = name != "fp32" and os.environ.get("PCADV_ENGINE", "tc") != "simt"
        self.engine = ENGINE_TC if want_tc else ENGINE_SIMT

    def __repr__(self):
        return "Precision(%s)" % self.name


_DEFAULT = [Precision(os.environ.get("PCADV_PRECISION", "fp16"))]


def set_default_precision(name):
    _DEFAULT[0] = Precision(name)


def default_precision():
    return _DEFAULT[0]


class KernelTimer:
    """Context manager that brackets every libpcadv call made inside it with CUDA
    events on the current stream (the stream the kernels are launched on) and
    keeps them per call signature, e.g. ``linear:tc:k512:n2048:colmax``.
    ``summary()`` synchronises and returns {signature: (calls, total_ms)}."""

    def __init__(self):
        self.records = {}

    def __enter__(self):
        global _TIMER
        self._prev, _TIMER = _TIMER, self
        return self

    def __exit__(self, *exc):
        global _TIMER
        _TIMER = self._prev

    def add(self, tag, start, end, rows=0):
        self.records.setdefault(tag, []).append((start, end, rows))

    def summary(self):
        """{signature: (calls, total ms, total rows processed)}"""
        torch.cuda.synchronize()
        return {tag: (len(ev), sum(a.elapsed_time(b) for a, b, _ in ev), sum(r for _, _, r in ev))
                for tag, ev in self.records.items()}


_TIMER = None


def _call(tag, fn, *args, rows=0):
    timer = _TIMER
    if timer is None:
        _lib.check(fn(*args))
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    _lib.check(fn(*args))
    end.record()
    timer.add(tag, start, end, rows)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _mat(t):
    """(ptr, ld, dtype code) of a 2-D tensor with unit column stride."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError("expected a [rows, cols] tensor with unit column stride, got %s / %s"
                         % (tuple(t.shape), t.stride()))
    if not t.is_cuda:
        raise RuntimeError("libpcadv ops need CUDA tensors (there is no CPU path)")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return C.c_void_p(t.data_ptr()), int(ld), _DT[t.dtype]


def _f32(t, n=None):
    if t is None:
        return C.c_void_p(0)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError("expected a contiguous fp32 tensor")
    if n is not None and t.numel() != n:
        raise ValueError("expected %d elements, got %d" % (n, t.numel()))
    return C.c_void_p(t.data_ptr())


def tc_eligible(segs, w, n):
    """Shapes the tensor-core engine takes: 16-bit operands, every K segment a
    multiple of 64 with 16-byte aligned rows."""
    for s in segs:
        if s.dtype not in (torch.float16, torch.bfloat16) or s.shape[1] % 64 or s.stride(0) % 8 \
                or s.data_ptr() % 16:
            return False
    if w.dtype != segs[0].dtype or w.stride(0) % 8 or w.data_ptr() % 16:
        return False
    return segs[0].shape[0] >= 128


def linear(segs, w, *, bias=None, group_bias=None, rows_per_group=0, addend=None, act=ACT_NONE,
           slope=0.0, mask=None, mask_act=ACT_NONE, mask_slope=0.0, out_dtype=torch.float32,
           out_scale=None, want_out=True, colmax=False, rowmax=False, engine=ENGINE_SIMT, n=None,
           bits_out=None, mask_bits=None, seg0_group_sum=None):
    """See ``pcadv_linear`` in include/pcadv.h.  Returns (out | None, colmax_key |
    None, rowmax_key | None).  ``bits_out``: a ``new_bits(rows, n)`` tensor that receives the
    sign bits of the output; ``mask_bits``: such a tensor used instead of ``mask`` (both only
    where ``bits_eligible`` holds).  ``seg0_group_sum``: fp32 [rows / rows_per_group, segs[0] width] that receives
    (+=) the per-cloud column sums of the first segment as a by-product (``group_sum_eligible`` shapes only)."""
    a = _lib.LinearArgs()
    rows = segs[0].shape[0]
    n = int(n if n is not None else w.shape[0])
    a.rows, a.n, a.num_seg = rows, n, len(segs)
    ktot = 0
    for i, s in enumerate(segs):
        if s.shape[0] != rows:
            raise ValueError("segment %d has %d rows, expected %d" % (i, s.shape[0], rows))
        p, ld, dt = _mat(s)
        a.seg[i].ptr, a.seg[i].ld, a.seg[i].k, a.seg[i].dtype = p, ld, s.shape[1], dt
        ktot += s.shape[1]
    if w.shape[1] != ktot or w.shape[0] < n:
        raise ValueError("weight %s does not match n=%d, ktot=%d" % (tuple(w.shape), n, ktot))
    a.w, a.ldw, a.w_dtype = _mat(w)
    if engine == ENGINE_TC and not tc_eligible(segs, w, n):
        engine = ENGINE_SIMT
    a.engine = engine
    a.bias = _f32(bias, n) if bias is not None else None
    dev = segs[0].device
    if group_bias is not None:
        if rows_per_group <= 0 or rows % rows_per_group:
            raise ValueError("rows_per_group must divide rows")
        a.group_bias = _f32(group_bias, (rows // rows_per_group) * n)
    a.rows_per_group = int(rows_per_group)
    if addend is not None:
        p, ld, dt = _mat(addend)
        if dt != F32:
            raise ValueError("addend must be fp32")
        a.addend, a.ld_addend = p, ld
    a.act, a.slope = act, float(slope)
    use_bits = mask_bits is not None and engine == ENGINE_TC and out_dtype != torch.float32 and \
        addend is None and out_scale is None and not rowmax and n % 64 == 0 and act == ACT_NONE and \
        bias is None and group_bias is None
    if use_bits:
        a.mask_bits, a.ld_mask_bits = C.c_void_p(mask_bits.data_ptr()), mask_bits.stride(0)
        a.mask_act, a.mask_slope = mask_act, float(mask_slope)
    elif mask is not None:
        p, ld, dt = _mat(mask)
        a.mask, a.ld_mask, a.mask_dtype = p, ld, dt
        a.mask_act, a.mask_slope = mask_act, float(mask_slope)
    elif mask_bits is not None:
        raise ValueError("mask_bits needs the tensor-core engine and a 16-bit output (pass mask too)")
    if bits_out is not None:
        if out_dtype == torch.float32 or not want_out:
            raise ValueError("bits_out needs a 16-bit output")
        a.bits_out, a.ld_bits_out = C.c_void_p(bits_out.data_ptr()), bits_out.stride(0)
    a.out_scale = _f32(out_scale) if out_scale is not None else None
    if seg0_group_sum is not None:
        if not use_bits or rows_per_group <= 0:
            raise ValueError("seg0_group_sum needs the mask_bits dgrad shape and rows_per_group")
        a.seg0_group_sum = _f32(seg0_group_sum, (rows // rows_per_group) * segs[0].shape[1])
    out = ckey = rkey = None
    a.out_dtype = _DT[out_dtype]
    if want_out:
        out = torch.empty((rows, n), dtype=out_dtype, device=dev)
        a.out, a.ld_out = C.c_void_p(out.data_ptr()), n
    if colmax:
        if rows_per_group <= 0 or rows % rows_per_group:
            raise ValueError("colmax needs rows_per_group dividing rows")
        ckey = torch.zeros((rows // rows_per_group, n), dtype=torch.int64, device=dev)
        a.colmax_key = C.c_void_p(ckey.data_ptr())
    if rowmax:
        rkey = torch.zeros((rows,), dtype=torch.int64, device=dev)
        a.rowmax_key = C.c_void_p(rkey.data_ptr())
    tag = "linear:%s:k%d:n%d%s%s%s%s" % ("tc" if engine == ENGINE_TC else "simt", ktot, n,
                                       ":colmax" if colmax else "", ":rowmax" if rowmax else "",
                                       ":maskbits" if use_bits else (":mask" if mask is not None else ""),
                                       ":addend" if addend is not None else "")
    if rows > 0:                                   # an empty batch launches nothing
        _call(tag, _lib.lib().pcadv_linear, C.byref(a), _stream(), rows=rows)
    return out, ckey, rkey


def group_sum_eligible(prec, segs, w, n, mask_bits, rows_per_group):
    """Shapes for which ``linear(..., seg0_group_sum=...)`` works: the tensor-core mask-bits dgrad over 16-bit
    operands, clouds that are whole 128-row tiles, a first segment of at most 256 channels."""
    return (prec.engine == ENGINE_TC and prec.act_dtype != torch.float32 and mask_bits is not None and n % 64 == 0
            and rows_per_group > 0 and rows_per_group % 128 == 0 and segs[0].shape[0] % rows_per_group == 0
            and segs[0].shape[1] <= 256 and tc_eligible(segs, w, n))


def chain_eligible(x, widths, last_f32=False, allow_serial=False):
    """Shapes ``chain`` takes: a 16-bit TMA-compatible input and 2..4 layers whose widths (input
    included) are multiples of 64 in [64, 256] (``widths`` lists the padded widths; with
    ``last_f32`` the last layer is at most 64 wide and leaves as fp32 rows)."""
    if x.dtype not in (torch.float16, torch.bfloat16) or x.dim() != 2 or x.shape[0] < 128:
        return False
    if x.stride(1) != 1 or x.stride(0) % 8 or x.data_ptr() % 16:
        return False
    if not 2 <= len(widths) - 1 <= 4:
        return False
    if not all(w % 64 == 0 and 64 <= w <= 256 for w in widths):
        return False
    if last_f32 and widths[-1] != 64:
        return False
    # shared-memory budget of tc_chain.cu: >= 2 ring stages + a [128 x max_n] tile per epilogue half.
    # Chains that only fit with ONE tile (a 256-wide stored output) run in the kernel's serial mode,
    # which measured slower than separate launches (0.53 vs 0.41 ms for fc2 -> fc3 -> fc4 at 2^20
    # points), so the models do not ask for it.
    max_n = max(widths[1:-1] + ([] if last_f32 else widths[-1:]))
    if last_f32:
        max_n = max(max_n, 128)
    tiles = 1 if allow_serial else 2
    smem = 1024 + 2 * (16384 + max(widths[1:]) * 128) + tiles * 128 * max_n * 2 + 4096 + 8192 + 256
    return smem <= 232448


def chain(x, layers, *, rowmax=False, want_bits=True, last_f32=False):
    """See ``pcadv_chain``.  layers: [(w16 [n, k], bias | None, act, slope)] (weights already in the
    16-bit compute dtype).  Returns (outputs [list of [rows, n] 16-bit tensors; the last one None
    with ``rowmax``, fp32 [rows, n] with ``last_f32``], sign-bit maps [list, None where not
    produced], rowmax key | None)."""
    a = _lib.ChainArgs()
    rows, k0 = x.shape
    a.rows, a.x, a.ldx, a.k0, a.dtype, a.num_layers = rows, C.c_void_p(x.data_ptr()), x.stride(0), k0, _DT[x.dtype], len(layers)
    outs, bits = [], []
    dev = x.device
    for l, (w, b, act, slope) in enumerate(layers):
        n = w.shape[0]
        L = a.layer[l]
        L.w, L.ldw, L.n, L.act, L.slope = C.c_void_p(w.data_ptr()), w.stride(0), n, act, float(slope)
        L.bias = _f32(b, n) if b is not None else None
        last = l == len(layers) - 1
        if last and rowmax:
            outs.append(None); bits.append(None)
            continue
        if last and last_f32:
            o = torch.empty((rows, n), dtype=torch.float32, device=dev)
            a.out_f32, a.n_f32 = C.c_void_p(o.data_ptr()), n
            L.n = (n + 63) // 64 * 64
            outs.append(o); bits.append(None)
            continue
        o = torch.empty((rows, n), dtype=x.dtype, device=dev)
        L.out, L.ld_out = C.c_void_p(o.data_ptr()), n
        outs.append(o)
        if want_bits and act != ACT_NONE:
            bt = new_bits(rows, n, dev)
            L.bits_out, L.ld_bits = C.c_void_p(bt.data_ptr()), bt.stride(0)
            bits.append(bt)
        else:
            bits.append(None)
    rkey = None
    if rowmax:
        rkey = torch.zeros((rows,), dtype=torch.int64, device=dev)
        a.rowmax_key = C.c_void_p(rkey.data_ptr())
    if rows > 0:
        _call("chain:k%d:%s%s" % (k0, "-".join(str(w.shape[0]) for w, _, _, _ in layers), ":rowmax" if rowmax else ""),
              _lib.lib().pcadv_chain, C.byref(a), _stream(), rows=rows)
    return outs, bits, rkey


def bits_eligible(prec, segs, w, n):
    """True when a layer can emit / consume the 1-bit activation mask: tensor-core engine,
    16-bit storage, n a multiple of 64."""
    if prec.engine != ENGINE_TC or prec.act_dtype == torch.float32 or n % 64:
        return False
    if len(segs) == 1 and segs[0].shape[1] <= 4 and segs[0].dtype == torch.float32:
        # the CUDA-core first-layer kernel (Conv1d(3, 64)) writes the map too
        return n <= 256 and segs[0].shape[0] >= 1024
    return tc_eligible(segs, w, n)


def new_bits(rows, n, device):
    """Sign-bit map of a [rows, n] activation: uint32 words, column c in word c // 32 (bit layout
    in include/pcadv.h)."""
    return torch.empty((rows, n // 32), dtype=torch.int32, device=device)


def wgrad(dz, segs, *, dw=None, dbias=None, dgroup_bias=None, rows_per_group=0, scale=None,
          engine=ENGINE_SIMT, n=None):
    """See ``pcadv_wgrad``.  ``dw`` / ``dbias`` / ``dgroup_bias`` are fp32 tensors that
    are accumulated into; ``dw`` may be a column-sliced view (unit column stride)."""
    a = _lib.WgradArgs()
    rows = dz.shape[0]
    n = int(n if n is not None else dz.shape[1])
    a.rows, a.n, a.num_seg = rows, n, len(segs)
    a.dz, a.ld_dz, a.dz_dtype = _mat(dz)
    ktot = 0
    for i, s in enumerate(segs):
        p, ld, dt = _mat(s)
        if s.shape[0] != rows:
            raise ValueError("segment %d has %d rows, expected %d" % (i, s.shape[0], rows))
        a.seg[i].ptr, a.seg[i].ld, a.seg[i].k, a.seg[i].dtype = p, ld, s.shape[1], dt
        ktot += s.shape[1]
    if dw is not None:
        p, ld, dt = _mat(dw)
        if dt != F32 or dw.shape[0] < n or dw.shape[1] != ktot:
            raise ValueError("dw must be fp32 [>=%d, %d], got %s" % (n, ktot, tuple(dw.shape)))
        a.dw, a.ld_dw = p, ld
    a.dbias = _f32(dbias) if dbias is not None else None
    if dgroup_bias is not None:
        a.dgroup_bias = _f32(dgroup_bias, (rows // rows_per_group) * n)
    a.rows_per_group = int(rows_per_group)
    a.scale = _f32(scale) if scale is not None else None
    if engine == ENGINE_TC:
        ok = dz.dtype in (torch.float16, torch.bfloat16) and n % 64 == 0 and rows >= 64 and \
            all(s.dtype == dz.dtype and s.shape[1] % 64 == 0 and s.stride(0) % 8 == 0 for s in segs) \
            and dz.stride(0) % 8 == 0 and dw is not None
        if not ok:
            engine = ENGINE_SIMT
    a.engine = engine
    tag = "wgrad:%s:n%d:k%d%s%s" % ("tc" if engine == ENGINE_TC else "simt", n, ktot,
                                    ":dbias" if dbias is not None else "",
                                    ":dgroup" if dgroup_bias is not None else "")
    if rows > 0:
        _call(tag, _lib.lib().pcadv_wgrad, C.byref(a), _stream(), rows=rows)


_LEVEL_ON = os.environ.get("PCADV_LEVEL", "1") != "0"      # tuning aid: separate dgrad / wgrad launches


def backlevel_eligible(prec, ks, x, mask_bits, rows_per_group=0, want_group=False):
    """Shapes ``backlevel`` takes: tensor-core engine, 16-bit storage, the widths ``ks`` of the incoming
    gradients and the width of ``x`` multiples of 64, sum(ks) small enough for the weight-gradient
    accumulators to stay in TMEM beside one dgrad accumulator, the sign-bit map of x at hand."""
    if not _LEVEL_ON or prec.engine != ENGINE_TC or prec.act_dtype == torch.float32 or mask_bits is None:
        return False
    rows, n = x.shape
    if rows < 128 or n % 64 or x.dtype != prec.act_dtype or x.stride(0) % 8 or x.data_ptr() % 16:
        return False
    if any(k % 64 for k in ks) or len(ks) > _lib.MAX_SEG:
        return False
    if want_group and (rows_per_group <= 0 or rows_per_group % 128):
        return False
    # TMEM budget of tc_level.cu: (K / 128 weight-gradient tiles + one dgrad accumulator) x 64 columns
    return (1 + (sum(ks) + 127) // 128) * 64 <= 512


def backlevel(segs, w, x, *, mask_bits=None, mask_act=ACT_RELU, mask_slope=0.0, dws=None, dbiases=None,
              dgroups=None, rows_per_group=0, scale=None, out=None, onehot=None):
    """See ``pcadv_backlevel``: dz_out = act'(x) * ([segs] @ w^T) together with the weight gradients
    ``dws[i] [k_i, n] += scale * segs[i]^T @ x``, ``dbiases[i] [k_i] += scale * colsum(segs[i])`` and the
    per-cloud column sums ``dgroups[i] [rows / rows_per_group, k_i]`` in one pass over ``segs``.
    ``onehot`` = (dy [rows] fp32, val [rows] fp32, idx [rows] int32, n_pool, act, slope, scale | None): the single
    segment is the gradient of a max over ``n_pool`` channels, generated inside the kernel (``segs`` is
    ignored).  Returns dz_out ([rows, n], the dtype of x)."""
    a = _lib.BackLevelArgs()
    rows, n = x.shape
    if onehot is not None:
        dy, val, idx, n_pool, oh_act, oh_slope, oh_scale = onehot
        if idx.dtype != torch.int32 or not idx.is_contiguous() or idx.numel() != rows:
            raise ValueError("onehot idx must be a contiguous int32 tensor with one entry per row")
        widths = [int(n_pool)]
        a.onehot_dy, a.onehot_val, a.onehot_idx = _f32(dy, rows), _f32(val, rows), _ptr(idx)
        a.onehot_act, a.onehot_slope = oh_act, float(oh_slope)
        a.onehot_scale = _f32(oh_scale) if oh_scale is not None else None
        a.seg[0].ptr, a.seg[0].ld, a.seg[0].k, a.seg[0].dtype = None, int(n_pool), int(n_pool), _DT[x.dtype]
    else:
        widths = [s.shape[1] for s in segs]
        for i, s in enumerate(segs):
            if s.shape[0] != rows or s.dtype != x.dtype:
                raise ValueError("segment %d: expected [%d, k] %s" % (i, rows, x.dtype))
            p, ld, dt = _mat(s)
            a.seg[i].ptr, a.seg[i].ld, a.seg[i].k, a.seg[i].dtype = p, ld, s.shape[1], dt
    a.rows, a.n, a.num_seg = rows, n, len(widths)
    ktot = sum(widths)
    if tuple(w.shape) != (n, ktot) or w.dtype != x.dtype:
        raise ValueError("weight %s does not match n=%d, ktot=%d" % (tuple(w.shape), n, ktot))
    a.w, a.ldw, _ = _mat(w)
    a.x, a.ldx, _ = _mat(x)
    if mask_bits is not None and mask_act != ACT_NONE:
        a.mask_bits, a.ld_mask_bits = C.c_void_p(mask_bits.data_ptr()), mask_bits.stride(0)
        a.mask_act, a.mask_slope = mask_act, float(mask_slope)
    dz = out if out is not None else torch.empty((rows, n), dtype=x.dtype, device=x.device)
    a.dz_out, a.ld_out, _ = _mat(dz)
    for i, k in enumerate(widths):
        dw = dws[i] if dws else None
        if dw is not None:
            p, ld, dt = _mat(dw)
            if dt != F32 or dw.shape[0] < k or dw.shape[1] != n:
                raise ValueError("dws[%d] must be fp32 [>=%d, %d], got %s" % (i, k, n, tuple(dw.shape)))
            a.dw[i], a.ld_dw[i] = p, ld
        db = dbiases[i] if dbiases else None
        if db is not None:
            if db.numel() < k:
                raise ValueError("dbiases[%d] needs %d elements" % (i, k))
            a.dbias[i] = _f32(db)
        dgp = dgroups[i] if dgroups else None
        if dgp is not None:
            a.dgroup[i] = _f32(dgp, (rows // rows_per_group) * k)
    a.rows_per_group = int(rows_per_group)
    a.scale = _f32(scale) if scale is not None else None
    if rows > 0:
        _call("backlevel:k%d:n%d%s" % (ktot, n, ":onehot" if onehot is not None else ""), _lib.lib().pcadv_backlevel,
              C.byref(a), _stream(), rows=rows)
    return dz


def max_finalize(key, act=ACT_NONE, slope=0.0, want_idx=True):
    """Unpack packed max keys -> (val fp32, idx int32) with the shape of ``key``."""
    val = torch.empty(key.shape, dtype=torch.float32, device=key.device)
    idx = torch.empty(key.shape, dtype=torch.int32, device=key.device) if want_idx else None
    if key.numel() > 0:
        _call("max_finalize", _lib.lib().pcadv_max_finalize, _ptr(key), key.numel(), act, float(slope),
              _ptr(val), _ptr(idx), _stream())
    return val, idx


def maxpool_inplace_eligible(k, rows_per_group, n):
    """Shapes ``maxpool_bwd(dz_inout=...)`` takes (see include/pcadv.h)."""
    return k % 64 == 0 and k <= 1024 and rows_per_group <= 8192 and n <= 4096


def maxpool_bwd(dg, gval, idx, x, w, rows_per_group, *, act=ACT_NONE, slope=0.0, dw=None,
                dbias=None, dx_acc=None, dz_inout=None, prev_act=ACT_NONE, prev_slope=0.0,
                scale=None):
    """See ``pcadv_maxpool_bwd``.  ``dz_inout`` ([rows, k], any float dtype) receives the
    sparse contribution in place, through the previous layer's activation mask."""
    a = _lib.MaxBwdArgs()
    groups, n = dg.shape
    if groups == 0:
        return
    a.groups, a.n, a.k = groups, n, w.shape[1]
    a.act, a.slope = act, float(slope)
    a.rows_per_group = int(rows_per_group)
    a.dg, a.gval = _f32(dg, groups * n), _f32(gval, groups * n)
    if idx.dtype != torch.int32 or not idx.is_contiguous():
        raise ValueError("idx must be contiguous int32")
    a.idx = _ptr(idx)
    a.x, a.ldx, a.x_dtype = _mat(x)
    a.w, a.ldw, a.w_dtype = _mat(w)
    if dw is not None:
        p, ld, dt = _mat(dw)
        a.dw, a.ld_dw = p, ld
    a.dbias = _f32(dbias) if dbias is not None else None
    if dx_acc is not None:
        p, ld, dt = _mat(dx_acc)
        if dt != F32:
            raise ValueError("dx_acc must be fp32")
        a.dx_acc, a.ld_dx = p, ld
    if dz_inout is not None:
        a.dz_inout, a.ld_dz, a.dz_dtype = _mat(dz_inout)
        a.prev_act, a.prev_slope = prev_act, float(prev_slope)
        ws_bytes = query_workspace(_lib.WS_MAXPOOL_BWD_INPLACE, groups, rows_per_group, n)
        ws = torch.empty((ws_bytes // 4,), dtype=torch.int32, device=dg.device)
        a.workspace = _ptr(ws)
    a.scale = _f32(scale) if scale is not None else None
    _call("maxpool_bwd:n%d:k%d" % (n, a.k), _lib.lib().pcadv_maxpool_bwd, C.byref(a), _stream())


def rowmax_bwd(dy, val, idx, n, *, act=ACT_NONE, slope=0.0, scale=None, out_dtype=torch.float32):
    rows = dy.numel()
    dz = torch.empty((rows, n), dtype=out_dtype, device=dy.device)
    if rows == 0:
        return dz
    _call("rowmax_bwd:n%d" % n, _lib.lib().pcadv_rowmax_bwd, _f32(dy), _f32(val), _ptr(idx), rows, n,
          act, float(slope), _f32(scale) if scale is not None else None, _ptr(dz), n, _DT[out_dtype],
          _stream())
    return dz


def rowmax_dgrad(dy, val, idx, w, yprev, *, act=ACT_NONE, slope=0.0, scale=None, prev_act=ACT_NONE,
                 prev_slope=0.0, out_dtype=torch.float32):
    """dz of the layer BEFORE a max over channels, straight from (dy, argmax): see
    ``pcadv_rowmax_dgrad``.  w: the pooled layer's [n, k] weight, yprev: its [rows, k] input."""
    rows, k = yprev.shape
    wp, ldw, wdt = _mat(w)
    yp, ldy, ydt = _mat(yprev)
    dz = torch.empty((rows, k), dtype=out_dtype, device=yprev.device)
    if rows == 0:
        return dz
    _call("rowmax_dgrad:k%d" % k, _lib.lib().pcadv_rowmax_dgrad, _f32(dy), _f32(val), _ptr(idx), rows, k,
          act, float(slope), _f32(scale) if scale is not None else None, wp, ldw, wdt, yp, ldy, ydt,
          prev_act, float(prev_slope), _ptr(dz), k, _DT[out_dtype], _stream())
    return dz


def rowmax_wgrad(dy, val, idx, yprev, n, *, act=ACT_NONE, slope=0.0, dw=None, dbias=None):
    """fp32 (dw [n, k], dbias [n]) of a layer followed by a max over channels, accumulated
    into: see ``pcadv_rowmax_wgrad``."""
    rows, k = yprev.shape
    if rows == 0:
        return
    yp, ldy, ydt = _mat(yprev)
    dwp, ld_dw = (C.c_void_p(0), 0)
    if dw is not None:
        dwp, ld_dw, dt = _mat(dw)
        if dt != F32:
            raise ValueError("dw must be fp32")
    _call("rowmax_wgrad:n%d:k%d" % (n, k), _lib.lib().pcadv_rowmax_wgrad, _f32(dy), _f32(val), _ptr(idx),
          rows, n, k, act, float(slope), yp, ldy, ydt, dwp, ld_dw,
          _f32(dbias) if dbias is not None else None, _stream())


def amax_scale(xs, target=256.0):
    """Device-side power-of-two gradient scale over one or several fp32 matrices:
    returns a 2-float tensor [S, 1/S] with S = 2^floor(log2(target / max|x|))."""
    if isinstance(xs, torch.Tensor):
        xs = [xs]
    dev = xs[0].device
    ws = torch.zeros(1, dtype=torch.int32, device=dev)
    s2 = torch.empty(2, dtype=torch.float32, device=dev)
    if all(x.numel() == 0 for x in xs):
        s2.fill_(1.0)
        return s2
    for x in xs:
        if x.numel() == 0:
            continue
        p, ld, dt = _mat(x)
        if dt != F32:
            raise ValueError("amax_scale expects fp32")
        _call("amax_scale", _lib.lib().pcadv_amax_scale, p, x.shape[0], x.shape[1], ld, float(target),
              _ptr(ws), _ptr(s2), _stream())
    return s2


def convert(src, out_dtype, cols_pad=None, scale=None, mask=None, mask_act=ACT_NONE, mask_slope=0.0):
    """dst = convert(src * scale * act'(mask)), zero-padded to ``cols_pad`` columns."""
    p, ld, dt = _mat(src)
    rows, cols = src.shape
    cols_pad = int(cols_pad or cols)
    dst = torch.empty((rows, cols_pad), dtype=out_dtype, device=src.device)
    if rows == 0:
        return dst
    mp, mld, mdt = _mat(mask) if mask is not None else (C.c_void_p(0), 0, F32)
    _call("convert:c%d" % cols_pad, _lib.lib().pcadv_convert, p, dt, ld, rows, cols, _ptr(dst),
          _DT[out_dtype], cols_pad, cols_pad, _f32(scale) if scale is not None else None, mp, mld, mdt,
          mask_act, float(mask_slope), _stream())
    return dst


def is_channel_major(x_bcn):
    """True for a B x C x N fp32 tensor whose point axis has unit stride (contiguous
    B x C x N, or any view with the same inner layout)."""
    return x_bcn.dim() == 3 and x_bcn.dtype == torch.float32 and x_bcn.is_cuda and \
        (x_bcn.shape[2] == 1 or x_bcn.stride(2) == 1) and x_bcn.stride(1) >= x_bcn.shape[2]


def convert_cm(x_bcn, out_dtype, cols_pad=None, scale=None):
    """Channel-major B x C x N fp32 -> point-major [B*N, cols_pad] (see pcadv_convert_cm)."""
    B, Cn, N = x_bcn.shape
    cols_pad = int(cols_pad or Cn)
    dst = torch.empty((B * N, cols_pad), dtype=out_dtype, device=x_bcn.device)
    if B * N == 0:
        return dst
    _call("convert_cm:c%d" % cols_pad, _lib.lib().pcadv_convert_cm, _ptr(x_bcn), x_bcn.stride(0),
          x_bcn.stride(1), B, N, Cn, _ptr(dst), _DT[out_dtype], cols_pad, cols_pad,
          _f32(scale) if scale is not None else None, _stream())
    return dst


def softmax_head(logits, mode, *, labels=None, out_dtype=torch.float32, cols=None, want_probs=True,
                 want_dz=False, dz_gain=1.0, loss_sum=None, valid_count=None, probs_out=None, dz_out=None):
    """See ``pcadv_softmax_head``.  logits: fp32 [rows, n].  Returns (probs | None, dz | None): point-
    major [rows, cols] matrices of ``out_dtype`` (softmax, or log_softmax in HEAD_LSM mode, and
    dz_gain * (softmax - onehot)); ``loss_sum`` (fp32 scalar tensor) accumulates the CE sum and
    ``valid_count`` the number of rows whose label lies in [0, n) (other rows are ignored rows:
    no loss term, zero dz).  ``probs_out`` / ``dz_out``: caller-owned [rows, cols] destinations (row
    slices of a larger matrix) instead of fresh tensors."""
    a = _lib.HeadArgs()
    p, ld, dt = _mat(logits)
    if dt != F32:
        raise ValueError("softmax_head expects fp32 logits")
    rows, n = logits.shape
    cols = int(cols or n)
    a.rows, a.n, a.mode, a.logits, a.ld = rows, n, mode, p, ld
    if labels is not None:
        if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.numel() != rows:
            raise ValueError("labels must be a contiguous int64 tensor with one entry per row")
        a.labels = _ptr(labels)
    probs = dz = None

    def dest(given):
        if given is None:
            return torch.empty((rows, cols), dtype=out_dtype, device=logits.device)
        if tuple(given.shape) != (rows, cols) or given.dtype != out_dtype or (cols > 1 and given.stride(1) != 1):
            raise ValueError("softmax_head destination must be [%d, %d] %s with unit column stride"
                             % (rows, cols, out_dtype))
        return given

    if want_probs or probs_out is not None:
        probs = dest(probs_out)
        a.probs, a.ld_probs, a.probs_dtype, a.probs_cols = _ptr(probs), probs.stride(0) if rows > 1 else cols, \
            _DT[out_dtype], cols
    if want_dz or dz_out is not None:
        dz = dest(dz_out)
        a.dz, a.ld_dz, a.dz_dtype, a.dz_cols = _ptr(dz), dz.stride(0) if rows > 1 else cols, _DT[out_dtype], cols
    a.dz_gain = float(dz_gain)
    a.loss_sum = _f32(loss_sum) if loss_sum is not None else None
    a.valid_count = _f32(valid_count) if valid_count is not None else None
    if rows == 0:
        return probs, dz
    _call("softmax_head:%s" % ("ce" if mode == _lib.HEAD_CE else "lsm"), _lib.lib().pcadv_softmax_head,
          C.byref(a), _stream(), rows=rows)
    return probs, dz


def logsoftmax_bwd(lp, dy, n, *, scale=None, out_dtype=torch.float32, cols=None, out=None):
    """dz = scale * (dy - exp(lp) * rowsum(dy)) over the first ``n`` columns (see
    ``pcadv_logsoftmax_bwd``); returns [rows, cols] of ``out_dtype`` (zero beyond n)."""
    lpp, ld_lp, lpd = _mat(lp)
    dyp, ld_dy, dyd = _mat(dy)
    rows = lp.shape[0]
    cols = int(cols or n)
    if out is not None:
        if tuple(out.shape) != (rows, cols) or (cols > 1 and out.stride(1) != 1):
            raise ValueError("logsoftmax_bwd: out must be [%d, %d] with unit column stride" % (rows, cols))
        dz, out_dtype = out, out.dtype
    else:
        dz = torch.empty((rows, cols), dtype=out_dtype, device=lp.device)
    if rows == 0:
        return dz
    _call("logsoftmax_bwd", _lib.lib().pcadv_logsoftmax_bwd, lpp, lpd, ld_lp, dyp, dyd, ld_dy, rows, int(n),
          _f32(scale) if scale is not None else None, _ptr(dz), _DT[out_dtype],
          dz.stride(0) if rows > 1 else cols, cols, _stream(), rows=rows)
    return dz


def lse_ratio(x, dy=None):
    """StackDiscNet's z / (z + 1), z = logsumexp over the columns of x [rows, S] (fp32): see
    ``pcadv_lse_ratio``.  Without ``dy``: returns y [rows]; with ``dy`` [rows]: returns dx [rows, S]."""
    p, ld, dt = _mat(x)
    if dt != F32:
        raise ValueError("lse_ratio expects fp32")
    rows, S = x.shape
    if dy is None:
        y = torch.empty((rows,), dtype=torch.float32, device=x.device)
        if rows:
            _call("lse_ratio", _lib.lib().pcadv_lse_ratio, p, ld, rows, S, None, _ptr(y), None, 0, _stream(), rows=rows)
        return y
    dx = torch.empty((rows, S), dtype=torch.float32, device=x.device)
    if rows:
        _call("lse_ratio_bwd", _lib.lib().pcadv_lse_ratio, p, ld, rows, S, _f32(dy, rows), None, _ptr(dx), S, _stream(),
              rows=rows)
    return dx


def round_residual(src, scale, out_dtype):
    """Column sums of what converting ``src * scale`` (fp32 [rows, cols <= 64]) to ``out_dtype`` drops:
    see ``pcadv_round_residual``.  Returns fp32 [cols]."""
    p, ld, dt = _mat(src)
    if dt != F32:
        raise ValueError("round_residual expects fp32")
    rows, cols = src.shape
    out = torch.zeros((cols,), dtype=torch.float32, device=src.device)
    if rows > 0:
        _call("round_residual", _lib.lib().pcadv_round_residual, p, ld, rows, cols,
              _f32(scale) if scale is not None else None, _DT[out_dtype], _ptr(out), _stream())
    return out


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise ValueError("expected a contiguous fp32 CUDA tensor")
    return C.c_void_p(t.data_ptr())


def bmm(x, trans, transpose_t=False):
    """y[b] = x[b] @ T[b] (or @ T[b]^T) for x [B, N, k], T [B, k, k] fp32: see ``pcadv_bmm``."""
    B, N, k = x.shape
    y = torch.empty_like(x)
    if x.numel() == 0:
        return y
    _call("bmm:k%d" % k, _lib.lib().pcadv_bmm, _f32c(x), _f32c(trans), _f32c(y), B, N, k,
          1 if transpose_t else 0, _stream())
    return y


def bmm_tgrad(x, dy):
    """dT[b] = x[b]^T @ dy[b] for x, dy [B, N, k] fp32: see ``pcadv_bmm_tgrad``."""
    B, N, k = x.shape
    dT = torch.zeros((B, k, k), dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return dT
    _call("bmm_tgrad:k%d" % k, _lib.lib().pcadv_bmm_tgrad, _f32c(x), _f32c(dy), _f32c(dT), B, N, k, _stream())
    return dT


def ortho_reg(trans):
    """(diff [B, d, d] = T T^T - I, norms [B] = ||diff||_F): see ``pcadv_ortho_reg``."""
    B, d, _ = trans.shape
    diff = torch.empty_like(trans)
    norms = torch.empty((B,), dtype=torch.float32, device=trans.device)
    _call("ortho_reg:d%d" % d, _lib.lib().pcadv_ortho_reg, _f32c(trans), B, d, _f32c(diff), _f32c(norms), _stream())
    return diff, norms


def ortho_reg_bwd(diff, trans, norms, dloss):
    """dT of mean_b ||T T^T - I||_F: see ``pcadv_ortho_reg_bwd``.  dloss: 0-d / 1-element fp32 tensor."""
    B, d, _ = trans.shape
    dT = torch.empty_like(trans)
    _call("ortho_reg_bwd:d%d" % d, _lib.lib().pcadv_ortho_reg_bwd, _f32c(diff), _f32c(trans), _f32c(norms),
          _f32c(dloss.reshape(1)), B, d, _f32c(dT), _stream())
    return dT


def query_workspace(op, groups=0, rows_per_group=0, n=0):
    """Bytes of caller-owned scratch an entry point wants: see ``pcadv_query_workspace``."""
    v = int(_lib.load().pcadv_query_workspace(int(op), int(groups), int(rows_per_group), int(n)))
    if v < 0:
        raise _lib.PcadvError(_lib.load().pcadv_last_error().decode())
    return v


def jitter(pts, sigma=0.01, clip=0.05, seed=0, offset=0, out=None):
    """pts + clip(sigma * N(0, 1), -clip, clip) on the device (dataset/modelNetData.py:80-91): see
    ``pcadv_jitter``.  pts: contiguous fp32 CUDA tensor (any shape); ``out`` may be pts itself."""
    if pts.dtype != torch.float32 or not pts.is_contiguous() or not pts.is_cuda:
        raise ValueError("jitter expects a contiguous fp32 CUDA tensor")
    if clip <= 0:
        raise ValueError("clip must be positive (modelNetData.py:88)")
    out = torch.empty_like(pts) if out is None else out
    if pts.numel():
        _call("jitter", _lib.lib().pcadv_jitter, _ptr(pts), _ptr(out), pts.numel(), float(sigma), float(clip),
              int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), _stream())
    return out


def bn_shape_ok(C):
    g = C // 8
    return C > 0 and C % 8 == 0 and g <= 256 and 256 % g == 0


def bn_stats(x, eps=1e-5, momentum=0.1, running_mean=None, running_var=None):
    """(mean [C], rstd [C]) of the rows of x [rows, C]; updates the running statistics when given:
    see ``pcadv_bn_stats``."""
    p, ld, dt = _mat(x)
    rows, Cn = x.shape
    ws = torch.zeros(2 * Cn, dtype=torch.float32, device=x.device)
    mean = torch.empty(Cn, dtype=torch.float32, device=x.device)
    rstd = torch.empty(Cn, dtype=torch.float32, device=x.device)
    _call("bn_stats:c%d" % Cn, _lib.lib().pcadv_bn_stats, p, dt, ld, rows, Cn, float(eps), float(momentum), _ptr(ws),
          _ptr(mean), _ptr(rstd), _f32(running_mean) if running_mean is not None else None,
          _f32(running_var) if running_var is not None else None, _stream(), rows=rows)
    return mean, rstd


def bn_apply(x, mean, rstd, gamma=None, beta=None, act=ACT_NONE, out_dtype=None):
    p, ld, dt = _mat(x)
    rows, Cn = x.shape
    y = torch.empty((rows, Cn), dtype=out_dtype or x.dtype, device=x.device)
    if rows:
        _call("bn_apply:c%d" % Cn, _lib.lib().pcadv_bn_apply, p, dt, ld, rows, Cn, _f32(mean, Cn), _f32(rstd, Cn),
              _f32(gamma, Cn) if gamma is not None else None, _f32(beta, Cn) if beta is not None else None, act,
              _ptr(y), _DT[y.dtype], Cn, _stream(), rows=rows)
    return y


def bn_bwd(x, dy, mean, rstd, gamma=None, y=None, want_dx=True):
    """(dx | None, dgamma [C], dbeta [C]): see ``pcadv_bn_bwd``; ``y`` = the ReLU'd output when the
    layer has the fused ReLU."""
    p, ld, dt = _mat(x)
    gp, gld, gdt = _mat(dy)
    rows, Cn = x.shape
    yp, yld, ydt = _mat(y) if y is not None else (C.c_void_p(0), 0, F32)
    dgamma = torch.zeros(Cn, dtype=torch.float32, device=x.device)
    dbeta = torch.zeros(Cn, dtype=torch.float32, device=x.device)
    dx = torch.empty((rows, Cn), dtype=dy.dtype, device=x.device) if want_dx else None
    _call("bn_bwd:c%d" % Cn, _lib.lib().pcadv_bn_bwd, p, dt, ld, gp, gdt, gld, yp, ydt, yld, rows, Cn, _f32(mean, Cn),
          _f32(rstd, Cn), _f32(gamma, Cn) if gamma is not None else None, _ptr(dgamma), _ptr(dbeta), _ptr(dx),
          _DT[dx.dtype] if dx is not None else F32, Cn, _stream(), rows=rows)
    return dx, dgamma, dbeta


def part_counts(labels, logits=None, pred=None, want_pred=False):
    """Per-cloud part statistics: see ``pcadv_part_counts``.

    labels: int64 [B, N].  Give either ``logits`` -- a [B, N, C] fp32 tensor with unit stride on C and
    contiguous points (the storage behind the generator's B x C x N view) -- or ``pred`` int64 [B, N].
    Returns (counts int32 [B, 3, C] = inter / pred / gt, correct int32 [B], pred int64 [B, N] | None);
    with ``pred`` given, C is 64 (every label value the kernel supports)."""
    if (logits is None) == (pred is None):
        raise ValueError("give either logits or pred")
    if labels.dtype != torch.int64 or not labels.is_cuda or labels.dim() != 2:
        raise ValueError("labels must be an int64 CUDA tensor [B, N]")
    labels = labels.contiguous()
    B, N = labels.shape
    if logits is not None:
        if (logits.dtype != torch.float32 or not logits.is_cuda or logits.dim() != 3
                or tuple(logits.shape[:2]) != (B, N)):
            raise ValueError("logits must be an fp32 CUDA tensor [B, N, C] matching labels")
        Cn = logits.shape[2]
        if B * N > 0 and (logits.stride(2) != 1 or logits.stride(0) != N * logits.stride(1)):
            raise ValueError("logits need unit stride on C and evenly strided points, got %s" % (logits.stride(),))
        ld = logits.stride(1) if N > 1 or B > 1 else Cn
        lp, pp = _ptr(logits), _ptr(None)
    else:
        if pred.dtype != torch.int64 or not pred.is_cuda or tuple(pred.shape) != (B, N):
            raise ValueError("pred must be an int64 CUDA tensor [B, N]")
        pred = pred.contiguous()
        Cn, ld = 64, 0
        lp, pp = _ptr(None), _ptr(pred)
    counts = torch.zeros((B, 3, Cn), dtype=torch.int32, device=labels.device)
    correct = torch.zeros((B,), dtype=torch.int32, device=labels.device)
    out = torch.empty((B, N), dtype=torch.int64, device=labels.device) if want_pred else None
    if B * N > 0:
        _call("part_counts", _lib.lib().pcadv_part_counts, lp, ld, pp, _ptr(labels), B, N, Cn, _ptr(counts),
              _ptr(correct), _ptr(out), _stream())
    return counts, correct, out


def part_iou(counts, onehot, part_begin):
    """(iou float64 [B], category int32 [B]) from ``part_counts`` counters: see ``pcadv_part_iou``.
    onehot: fp32 [B, ncat] (unit column stride); part_begin: int32 device tensor [ncat + 1]."""
    B, _, Cn = counts.shape
    if onehot.dtype != torch.float32 or onehot.dim() != 2 or onehot.shape[0] != B or (
            onehot.shape[1] > 1 and onehot.stride(1) != 1):
        raise ValueError("onehot must be fp32 [B, ncat] with unit column stride")
    ncat = onehot.shape[1]
    if part_begin.dtype != torch.int32 or part_begin.numel() != ncat + 1 or not part_begin.is_contiguous():
        raise ValueError("part_begin must be a contiguous int32 tensor of ncat + 1 entries")
    iou = torch.empty((B,), dtype=torch.float64, device=counts.device)
    cat = torch.empty((B,), dtype=torch.int32, device=counts.device)
    if B > 0:
        _call("part_iou", _lib.lib().pcadv_part_iou, _ptr(counts.contiguous()), _ptr(onehot),
              onehot.stride(0) if B > 1 else ncat, ncat, _ptr(part_begin), B, Cn, _ptr(iou), _ptr(cat), _stream())
    return iou, cat


def transpose(src, out_dtype=None):
    """dst[c, r] = src[r, c] (weight matrices only)."""
    p, ld, dt = _mat(src)
    out_dtype = out_dtype or src.dtype
    rows, cols = src.shape
    dst = torch.empty((cols, rows), dtype=out_dtype, device=src.device)
    _lib.check(_lib.lib().pcadv_transpose(p, dt, ld, rows, cols, _ptr(dst), _DT[out_dtype], rows,
                                          _stream()))
    return dst


__all__ = ["KernelTimer", "Precision", "set_default_precision", "default_precision", "linear", "wgrad",
           "max_finalize", "maxpool_bwd", "rowmax_bwd", "amax_scale", "convert", "transpose",
           "softmax_head", "logsoftmax_bwd", "HEAD_CE", "HEAD_LSM", "part_counts", "part_iou",
           "ACT_NONE", "ACT_RELU", "ACT_LEAKY", "ENGINE_SIMT", "ENGINE_TC"]
