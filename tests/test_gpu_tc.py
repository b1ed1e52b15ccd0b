"""GPU: the tensor-core engine (tcgen05 / TMEM / TMA kernels) against fp64 torch
on the same 16-bit operands.  Operands are exactly representable in the storage
dtype, so the only error is fp32 accumulation order: tolerance 1e-5 relative on
the fp32 accumulator, 1e-3 after rounding to a 16-bit output."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from adversarial_learning_on_pointclouds_b200 import ops
from adversarial_learning_on_pointclouds_b200.ops import (ACT_LEAKY, ACT_NONE, ACT_RELU, ENGINE_TC)
from helpers import rel_err

DEV = "cuda"
DT = [torch.float16, torch.bfloat16]


def _rand(shape, seed, dtype, scale=1.0):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).to(DEV).to(dtype)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,ks,n", [(128, [64], 64), (1000, [64], 128), (5000, [128], 256),
                                       (4096, [128], 512), (3000, [64, 128, 128, 128, 512], 256),
                                       (2500, [128], 50), (777, [512, 256], 128), (300, [64], 16)])
def test_tc_linear_plain(dtype, rows, ks, n):
    segs = [_rand((rows, k), 10 + i, dtype) for i, k in enumerate(ks)]
    w = _rand((n, sum(ks)), 3, dtype, 0.1)
    out, _, _ = ops.linear(segs, w, out_dtype=torch.float32, engine=ENGINE_TC)
    ref = torch.cat(segs, 1).double() @ w.double().t()
    err = rel_err(out, ref)
    print(dtype, rows, ks, n, "rel err", err)
    assert err < 1e-5


@pytest.mark.parametrize("dtype", DT)
def test_tc_linear_full_epilogue(dtype):
    rows, ks, n, rpg = 5000, [64, 128], 256, 2500
    segs = [_rand((rows, k), 10 + i, dtype) for i, k in enumerate(ks)]
    w = _rand((n, sum(ks)), 3, dtype, 0.1)
    bias = _rand((n,), 4, torch.float32)
    gb = _rand((rows // rpg, n), 5, torch.float32)
    add = _rand((rows, n), 6, torch.float32)
    mask = _rand((rows, n), 7, dtype)
    sc = torch.tensor([0.5], device=DEV)
    out, _, rkey = ops.linear(segs, w, bias=bias, group_bias=gb, rows_per_group=rpg, addend=add,
                              act=ACT_LEAKY, slope=0.2, mask=mask, mask_act=ACT_RELU, out_scale=sc,
                              out_dtype=dtype, rowmax=True, engine=ENGINE_TC)
    pre = torch.cat(segs, 1).double() @ w.double().t() + bias.double() + \
        gb.double().repeat_interleave(rpg, 0) + add.double()
    ref = F.leaky_relu(pre, 0.2) * (mask > 0) * 0.5
    assert out.dtype == dtype
    assert rel_err(out, ref) < (1e-3 if dtype == torch.float16 else 6e-3)
    rval, ridx = ops.max_finalize(rkey, ACT_RELU)
    rv, ri = F.relu(pre).max(1)
    assert rel_err(rval, rv) < 1e-5
    assert ((ridx.long() == ri) | (rv == 0)).double().mean().item() > 0.995


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,N,k,n", [(3, 500, 128, 256), (2, 2500, 512, 2048), (5, 100, 64, 1024),
                                     (1, 4096, 128, 512)])
def test_tc_colmax_over_points(dtype, B, N, k, n):
    x = _rand((B * N, k), 1, dtype)
    w = _rand((n, k), 2, dtype, 0.1)
    b = _rand((n,), 3, torch.float32)
    _, key, _ = ops.linear([x], w, bias=b, want_out=False, colmax=True, rows_per_group=N,
                           engine=ENGINE_TC)
    g, idx = ops.max_finalize(key, ACT_RELU)
    y = F.relu(x.double() @ w.double().t() + b.double()).view(B, N, n)
    gv, gi = y.max(1)
    assert rel_err(g, gv) < 1e-5
    at = torch.gather(y, 1, idx.long().unsqueeze(1)).squeeze(1)
    assert ((gv - at).abs() <= 1e-5 * gv.abs().clamp_min(1e-3)).all()      # argmax within rounding
    assert ((idx.long() == gi) | (gv == 0)).double().mean().item() > 0.99


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,ks,n", [(1000, [64], 64), (5000, [128], 256), (40000, [128], 512),
                                       (3000, [64, 128, 128, 128, 512], 256), (2500, [128], 64),
                                       (100000, [512], 128)])
def test_tc_wgrad(dtype, rows, ks, n):
    dz = _rand((rows, n), 1, dtype)
    segs = [_rand((rows, k), 10 + i, dtype) for i, k in enumerate(ks)]
    sc = torch.tensor([0.5], device=DEV)
    dw = torch.zeros((n, sum(ks)), device=DEV)
    db = torch.zeros((n,), device=DEV)
    ops.wgrad(dz, segs, dw=dw, dbias=db, scale=sc, engine=ENGINE_TC)
    x = torch.cat(segs, 1).double()
    e1, e2 = rel_err(dw, 0.5 * dz.double().t() @ x), rel_err(db, 0.5 * dz.double().sum(0))
    print(dtype, rows, ks, n, "dw", e1, "db", e2)
    assert e1 < 1e-5 and e2 < 1e-5


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,k,n,act", [(1000, 64, 64, ACT_RELU), (4133, 128, 512, ACT_RELU),
                                          (2500, 256, 256, ACT_LEAKY), (129, 64, 128, ACT_RELU)])
def test_tc_bit_masks(dtype, rows, k, n, act):
    """bits_out is exactly [stored output > 0]; a dgrad that reads mask_bits equals the one that
    reads the 16-bit mask."""
    x = _rand((rows, k), 21, dtype)
    w = _rand((n, k), 22, dtype, 0.1)
    bias = _rand((n,), 23, torch.float32)
    bits = ops.new_bits(rows, n, DEV)
    bits.fill_(-1)
    y, _, _ = ops.linear([x], w, bias=bias, act=act, slope=0.2, out_dtype=dtype, engine=ENGINE_TC,
                         bits_out=bits)
    # word w covers columns 32 w .. 32 w + 31: column 2 k at bit k, column 2 k + 1 at bit 16 + k
    j = torch.arange(32, device=DEV, dtype=torch.int64)
    shifts = (j >> 1) + 16 * (j & 1)
    got = ((bits.long().unsqueeze(2) >> shifts) & 1).reshape(rows, n).bool()
    assert torch.equal(got, y.float() > 0)
    dz = _rand((rows, n), 24, dtype)
    wt = _rand((k if k % 64 == 0 else 64, n), 25, dtype, 0.1)
    # dgrad of a following layer: output width n' = wt rows must match the mask width -> use n x n
    w2 = _rand((n, n), 26, dtype, 0.1)
    a, _, _ = ops.linear([dz], w2, mask=y, mask_act=act, mask_slope=0.2, out_dtype=dtype, engine=ENGINE_TC)
    b, _, _ = ops.linear([dz], w2, mask=y, mask_act=act, mask_slope=0.2, out_dtype=dtype, engine=ENGINE_TC,
                         mask_bits=bits)
    assert torch.equal(a, b)


@pytest.mark.parametrize("rows", [129, 389, 1000, 4133])
def test_caller_buffers_are_not_overrun(rows):
    """Guard words around the caller-provided buffers (sign-bit map, dW, dbias, packed head
    outputs) stay untouched for ragged row counts."""
    dtype = torch.float16
    k, n = 64, 128
    x = _rand((rows, k), 31, dtype)
    w = _rand((n, k), 32, dtype, 0.1)
    guard = 64
    raw = torch.full((rows * (n // 32) + 2 * guard,), 0x5A5A5A5A, dtype=torch.int32, device=DEV)
    bits = raw[guard:guard + rows * (n // 32)].view(rows, n // 32)
    ops.linear([x], w, bias=_rand((n,), 33, torch.float32), act=ACT_RELU, out_dtype=dtype, engine=ENGINE_TC,
               bits_out=bits)
    assert (raw[:guard] == 0x5A5A5A5A).all() and (raw[guard + rows * (n // 32):] == 0x5A5A5A5A).all()
    # wgrad into a guarded dW / dbias
    dz = _rand((rows, n), 34, dtype)
    rawf = torch.full((n * k + n + 2 * guard,), 7.25, dtype=torch.float32, device=DEV)
    dw = rawf[guard:guard + n * k].view(n, k)
    db = rawf[guard + n * k:guard + n * k + n]
    dw.zero_(); db.zero_()
    ops.wgrad(dz, [x], dw=dw, dbias=db, engine=ENGINE_TC)
    assert (rawf[:guard] == 7.25).all() and (rawf[guard + n * k + n:] == 7.25).all()
    assert rel_err(dw, dz.double().t() @ x.double()) < 1e-5
    assert rel_err(db, dz.double().sum(0)) < 1e-5
    # packed head outputs are allocated by ops.softmax_head itself; check the values at the ragged end
    logits = _rand((rows, 50), 35, torch.float32) * 3
    lp, _ = ops.softmax_head(logits, ops.HEAD_LSM, out_dtype=dtype, cols=64)
    assert rel_err(lp[:, :50], torch.log_softmax(logits.double(), 1)) < 2e-3
    assert lp[:, 50:].abs().max().item() == 0


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,ks,rpg", [(4096, [64, 128, 128, 128, 512], 1024), (19984, [512], 0),
                                         (70001, [64, 128, 128, 128, 512], 0), (8192, [512, 512], 2048)])
def test_tc_wgrad_cta_pair(dtype, rows, ks, rpg):
    """The cta_group::2 weight-gradient kernel (dz of 256 channels, K-concat of >= 512): dW and the
    fused per-cloud column sums against fp64."""
    n = 256
    dz = _rand((rows, n), 51, dtype)
    segs = [_rand((rows, k), 52 + i, dtype) for i, k in enumerate(ks)]
    dw = torch.zeros((n, sum(ks)), device=DEV)
    dgb = torch.zeros((rows // rpg, n), device=DEV) if rpg else None
    sc = torch.tensor([0.25], device=DEV)
    ops.wgrad(dz, segs, dw=dw, dgroup_bias=dgb, rows_per_group=rpg, scale=sc, engine=ENGINE_TC)
    ref = 0.25 * dz.double().t() @ torch.cat(segs, 1).double()
    assert rel_err(dw, ref) < 1e-5
    if rpg:
        assert rel_err(dgb, dz.double().view(rows // rpg, rpg, n).sum(1)) < 1e-6


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,widths,rowmax", [(1000, [64, 128, 128, 128], False), (4133, [64, 64, 64, 64, 128], True),
                                                (129, [256, 128, 64], False), (40000, [128, 128, 128, 64, 64], False),
                                                (5000, [128, 256, 256], False), (4133, [64, 256, 128, 128], True)])
def test_tc_chain(dtype, rows, widths, rowmax):
    """pcadv_chain (layers multiplied out of shared memory) against the same layers run one by one
    through pcadv_linear: identical stored activations, sign bits and row-max."""
    x = _rand((rows, widths[0]), 61, dtype)
    layers = []
    for l in range(len(widths) - 1):
        w = _rand((widths[l + 1], widths[l]), 62 + l, dtype, 1.5 / widths[l] ** 0.5)
        b = _rand((widths[l + 1],), 70 + l, torch.float32, 0.1)
        layers.append((w, b, ACT_RELU, 0.0))
    assert ops.chain_eligible(x, widths, allow_serial=True)
    outs, bits, rkey = ops.chain(x, layers, rowmax=rowmax)
    cur = x
    for l, (w, b, act, slope) in enumerate(layers):
        last = l == len(layers) - 1
        if last and rowmax:
            _, _, rk = ops.linear([cur], w, bias=b, want_out=False, rowmax=True, engine=ENGINE_TC)
            v1, i1 = ops.max_finalize(rkey, ACT_RELU)
            v2, i2 = ops.max_finalize(rk, ACT_RELU)
            assert torch.equal(v1, v2) and torch.equal(i1, i2)
            assert outs[l] is None
            break
        bt = ops.new_bits(rows, widths[l + 1], DEV)
        ref, _, _ = ops.linear([cur], w, bias=b, act=act, out_dtype=dtype, engine=ENGINE_TC, bits_out=bt)
        assert torch.equal(outs[l], ref), l
        assert torch.equal(bits[l], bt), l
        cur = ref


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows", [4096, 4133, 129])
def test_tc_chain_wide_tail_with_fp32_logits(dtype, rows):
    """The head tail fc2 -> fc3 -> fc4 (256 -> 256 -> 128 -> 50 fp32) as one chained launch (one tile
    in flight, both epilogue halves split its steps) against the layers run one by one."""
    widths = [256, 256, 128, 50]
    x = _rand((rows, 256), 81, dtype).relu_()
    layers = []
    for l in range(3):
        w = _rand((widths[l + 1], widths[l]), 82 + l, dtype, 1.5 / widths[l] ** 0.5)
        b = _rand((widths[l + 1],), 90 + l, torch.float32, 0.1)
        layers.append((w, b, ACT_RELU if l < 2 else ACT_NONE, 0.0))
    assert ops.chain_eligible(x, [256, 256, 128, 64], last_f32=True, allow_serial=True)
    assert not ops.chain_eligible(x, [256, 256, 128, 64], last_f32=True)     # the models skip the serial mode
    outs, bits, _ = ops.chain(x, layers, last_f32=True)
    cur = x
    for l, (w, b, act, slope) in enumerate(layers):
        if l == 2:
            ref, _, _ = ops.linear([cur], w, bias=b, out_dtype=torch.float32, engine=ENGINE_TC)
            assert outs[l].dtype == torch.float32 and tuple(outs[l].shape) == (rows, 50)
            assert torch.equal(outs[l], ref)
            break
        bt = ops.new_bits(rows, widths[l + 1], DEV)
        ref, _, _ = ops.linear([cur], w, bias=b, act=act, out_dtype=dtype, engine=ENGINE_TC, bits_out=bt)
        assert torch.equal(outs[l], ref), l
        assert torch.equal(bits[l], bt), l
        cur = ref


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,ks,n,rpg,act", [
    (1000, [128, 256], 128, 0, ACT_RELU),        # trunk level 2 / 3: one dgrad accumulator beside 3 weight tiles
    (4133, [128, 256], 64, 0, ACT_RELU),         # level 1
    (2560, [512, 256], 128, 0, ACT_RELU),        # level 4: two 64-column slices
    (4096, [256], 512, 1024, ACT_RELU),          # level 5: four slices, per-cloud column sums
    (3000, [256], 256, 0, ACT_LEAKY),            # fc2 level
    (129, [128], 256, 0, ACT_RELU),              # fc3 level, ragged rows
    (5000, [64], 128, 0, ACT_RELU),              # fc4 level: K = 64 zero-padded to one 128-channel tile
    (40000, [512, 256, 128], 64, 0, ACT_NONE),   # K = 896: seven weight tiles + one accumulator = all of TMEM
])
def test_tc_backlevel(dtype, rows, ks, n, rpg, act):
    """pcadv_backlevel = the dgrad GEMM of a level + the weight gradients of the layers fed by x +
    bias / per-cloud sums, against fp64 torch on the same 16-bit operands."""
    segs = [_rand((rows, k), 40 + i, dtype) for i, k in enumerate(ks)]
    w = _rand((n, sum(ks)), 50, dtype, 0.1)
    xpre = _rand((rows, n), 51, dtype)
    x = F.leaky_relu(xpre, 0.2) if act == ACT_LEAKY else (xpre.relu() if act == ACT_RELU else xpre)
    bits = None
    if act != ACT_NONE:
        # sign bits in the layout pcadv_linear writes: column 2 k at bit k, 2 k + 1 at bit 16 + k of word c // 32
        pos = (x.float() > 0).reshape(rows, n // 32, 32).long()
        j = torch.arange(32, device=DEV)
        shifts = (j >> 1) + 16 * (j & 1)
        words = (pos << shifts).sum(2)
        bits = (words - ((words >> 31) << 32)).to(torch.int32).contiguous()
    dws = [torch.zeros((k, n), device=DEV) for k in ks]
    # column sums are taken for at most 8 chunks (512 channels) per launch
    dbs, used = [], 0
    for k in ks:
        dbs.append(torch.zeros((k,), device=DEV) if used + k // 64 <= 8 else None)
        used += k // 64 if dbs[-1] is not None else 0
    groups = rows // rpg if rpg else 0
    dgs = [torch.zeros((groups, k), device=DEV) if rpg and dbs[i] is not None else None for i, k in enumerate(ks)]
    sc = torch.tensor([0.25], device=DEV)
    dz = ops.backlevel(segs, w, x, mask_bits=bits, mask_act=act, mask_slope=0.2, dws=dws, dbiases=dbs,
                       dgroups=dgs if rpg else None, rows_per_group=rpg, scale=sc)
    torch.cuda.synchronize()
    cat = torch.cat(segs, 1).double()
    ref = cat @ w.double().t()
    if act == ACT_RELU:
        ref = ref * (x > 0)
    elif act == ACT_LEAKY:
        ref = torch.where(x > 0, ref, ref * 0.2)
    tol = 1e-3 if dtype == torch.float16 else 6e-3
    assert dz.dtype == dtype and rel_err(dz, ref) < tol
    for i, s in enumerate(segs):
        assert rel_err(dws[i], 0.25 * (s.double().t() @ x.double())) < 1e-5, i
        if dbs[i] is not None:
            assert rel_err(dbs[i], 0.25 * s.double().sum(0)) < 1e-5, i
        if rpg and dgs[i] is not None:
            assert rel_err(dgs[i], s.double().reshape(groups, rpg, -1).sum(1)) < 1e-5, i
    # second call accumulates
    ops.backlevel(segs, w, x, mask_bits=bits, mask_act=act, mask_slope=0.2, dws=dws, dbiases=dbs, scale=sc)
    assert rel_err(dws[0], 0.5 * (segs[0].double().t() @ x.double())) < 1e-5


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,n_pool,k,act", [(5000, 128, 64, ACT_RELU), (129, 128, 64, ACT_LEAKY),
                                               (40000, 64, 128, ACT_NONE), (4133, 256, 64, ACT_RELU)])
def test_tc_backlevel_onehot(dtype, rows, n_pool, k, act):
    """The one-hot mode of pcadv_backlevel (backward of "layer + activation + max over channels",
    models/discriminator.py:67-72) against the gather kernels it replaces and fp64 torch."""
    g = torch.Generator().manual_seed(77)
    dy = torch.randn(rows, generator=g).to(DEV)
    val = (torch.rand(rows, generator=g) - 0.3).to(DEV)          # some pooled values <= 0
    idx = torch.randint(0, n_pool, (rows,), generator=g, dtype=torch.int32).to(DEV)
    y = _rand((rows, k), 61, dtype).relu()
    w = _rand((n_pool, k), 62, dtype, 0.1)                         # the pooled layer's weight
    wt = w.t().contiguous()
    pos = (y.float() > 0).reshape(rows, k // 32, 32).long()
    j = torch.arange(32, device=DEV)
    words = (pos << ((j >> 1) + 16 * (j & 1))).sum(2)
    bits = (words - ((words >> 31) << 32)).to(torch.int32).contiguous()
    S = torch.tensor([8.0], device=DEV)
    inv = torch.tensor([0.125], device=DEV)
    dw = torch.zeros((n_pool, k), device=DEV)
    db = torch.zeros((n_pool,), device=DEV)
    dz = ops.backlevel(None, wt, y, mask_bits=bits, mask_act=ACT_RELU, dws=[dw], dbiases=[db], scale=inv,
                       onehot=(dy, val, idx, n_pool, act, 0.2, S))
    torch.cuda.synchronize()
    d = torch.ones_like(val) if act == ACT_NONE else torch.where(val > 0, torch.ones_like(val),
                                                                 torch.full_like(val, 0.2 if act == ACT_LEAKY else 0.0))
    s = (dy * d).double()
    onehot = torch.zeros((rows, n_pool), dtype=torch.float64, device=DEV)
    onehot[torch.arange(rows, device=DEV), idx.long()] = s
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2               # s itself is rounded to 16 bits
    assert rel_err(dw, onehot.t() @ y.double()) < tol
    assert rel_err(db, onehot.sum(0)) < 1e-5                        # the bias sums the unrounded values
    ref_dz = (onehot * 8.0) @ w.double() * (y > 0)
    assert rel_err(dz, ref_dz) < tol
    # the gather kernels it stands in for
    dw2 = torch.zeros_like(dw)
    db2 = torch.zeros_like(db)
    ops.rowmax_wgrad(dy, val, idx, y, n_pool, act=act, slope=0.2, dw=dw2, dbias=db2)
    dz2 = ops.rowmax_dgrad(dy, val, idx, w, y, act=act, slope=0.2, scale=S, prev_act=ACT_RELU, out_dtype=dtype)
    assert rel_err(dw, dw2) < tol and rel_err(db, db2) < 1e-5 and rel_err(dz, dz2) < tol


@pytest.mark.parametrize("rows", [129, 389, 1000, 4133])
@pytest.mark.parametrize("ks,n", [([64], 128), ([128, 256], 64), ([128], 256)])
def test_backlevel_does_not_overrun_caller_buffers(rows, ks, n):
    """Guard words around dz_out, dW and dbias of pcadv_backlevel stay untouched for ragged row
    counts (compute-sanitizer is closed on the pool: this is the bounds check)."""
    dtype = torch.float16
    segs = [_rand((rows, k), 70 + i, dtype) for i, k in enumerate(ks)]
    w = _rand((n, sum(ks)), 80, dtype, 0.1)
    x = _rand((rows, n), 81, dtype).relu()
    bits = torch.full((rows, n // 32), -1, dtype=torch.int32, device=DEV)      # every sign bit set
    guard = 256
    raw16 = torch.full((rows * n + 2 * guard,), 3.5, dtype=dtype, device=DEV)
    out = raw16[guard:guard + rows * n].view(rows, n)
    sizes = [k * n for k in ks] + [ks[0]]
    rawf = torch.full((sum(sizes) + 2 * guard,), 7.25, dtype=torch.float32, device=DEV)
    off, dws = guard, []
    for k in ks:
        dws.append(rawf[off:off + k * n].view(k, n))
        off += k * n
    db = rawf[off:off + ks[0]]
    for t in dws + [db]:
        t.zero_()
    ops.backlevel(segs, w, x, mask_bits=bits, dws=dws, dbiases=[db] + [None] * (len(ks) - 1), out=out)
    torch.cuda.synchronize()
    assert (raw16[:guard] == 3.5).all() and (raw16[guard + rows * n:] == 3.5).all()
    assert (rawf[:guard] == 7.25).all() and (rawf[off + ks[0]:] == 7.25).all()
    assert rel_err(out, torch.cat(segs, 1).double() @ w.double().t()) < 1e-3
    assert rel_err(dws[-1], segs[-1].double().t() @ x.double()) < 1e-5
    assert rel_err(db, segs[0].double().sum(0)) < 1e-5


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,ks,n,rpg", [(4096, [256], 512, 1024), (2560, [128, 64], 128, 640), (1024, [64], 64, 128)])
def test_tc_linear_group_sum_byproduct(dtype, rows, ks, n, rpg):
    """seg0_group_sum: the per-cloud column sums of the first segment come out of the mask-bits dgrad
    launch (exact fp32 sums of the 16-bit values), the dgrad result itself is unchanged."""
    segs = [_rand((rows, k), 90 + i, dtype) for i, k in enumerate(ks)]
    w = _rand((n, sum(ks)), 93, dtype, 0.1)
    y = _rand((rows, n), 94, dtype).relu()
    bits = ops.new_bits(rows, n, DEV)
    pos = (y.float() > 0).reshape(rows, n // 32, 32).long()
    j = torch.arange(32, device=DEV)
    words = (pos << ((j >> 1) + 16 * (j & 1))).sum(2)
    bits.copy_((words - ((words >> 31) << 32)).to(torch.int32))
    ref, _, _ = ops.linear(segs, w, mask=y, mask_act=ACT_RELU, out_dtype=dtype, engine=ENGINE_TC, mask_bits=bits)
    gs = torch.zeros((rows // rpg, ks[0]), device=DEV)
    out, _, _ = ops.linear(segs, w, mask=y, mask_act=ACT_RELU, out_dtype=dtype, engine=ENGINE_TC, mask_bits=bits,
                           rows_per_group=rpg, seg0_group_sum=gs)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    assert rel_err(gs, segs[0].double().reshape(rows // rpg, rpg, -1).sum(1)) < 1e-6
    # accumulates
    ops.linear(segs, w, mask=y, mask_act=ACT_RELU, out_dtype=dtype, engine=ENGINE_TC, mask_bits=bits,
               rows_per_group=rpg, seg0_group_sum=gs)
    assert rel_err(gs, 2 * segs[0].double().reshape(rows // rpg, rpg, -1).sum(1)) < 1e-6
