"""Branch-conditioned parity harness ("Parity method").

ReLU masks and max-pool argmaxes are the only non-smooth points of the
network.  A gradient comparison between two correct implementations fails at
the 1e-4 level as soon as ONE pre-activation that is within rounding distance
of zero lands on the other side (measured on the reference itself: 1 vs 8 CPU
threads differ by 3.7e-4 in conv3.bias.grad at B=4, N=1024).  So gradients are
compared against the oracle evaluated with the CUDA path's branch decisions,
and every decision that differs from the oracle's own is shown to sit within
rounding distance of its boundary.
"""
import torch
import torch.nn.functional as F

from oracle import pointnet_oracle as PO, steps
from helpers import rel_err


def to_bcn(x_pm, B, N):
    """point-major [B*N, C] -> reference layout B x C x N (CPU)."""
    return x_pm.detach().float().cpu().view(B, N, -1).transpose(1, 2)


def seg_branches(debug, B, N):
    br = {}
    for i, x in enumerate(debug["x"]):
        br["act:conv%d" % (i + 1)] = to_bcn(x, B, N) > 0
    for i, h in enumerate(debug["h"]):
        br["act:fc%d" % (i + 1)] = h.detach().float().cpu().view(B, N, -1) > 0
    br["argmax:conv6"] = debug["idx"].detach().cpu().long()
    return br


# unit roundoff of the storage / operand format of each precision mode
U = {"fp32": 2.0 ** -24, "fp16": 2.0 ** -11, "bf16": 2.0 ** -8}


def flip_bounds(mode):
    """How far from its boundary a decision may sit and still legitimately land on the other side,
    in units of the layer's RMS: (|pre-activation| bound, top-2 gap bound).

    A pre-activation is a K-term dot product of operands rounded to the format (unit roundoff u) on
    top of inputs that already carry a few u of relative error, so its error is a few u of the
    layer's RMS magnitude, independent of K (the terms' errors add like a random walk, as the terms
    themselves do).  Measured on the reference with emulated fp16 / bf16 operands and storage
    (B = 4, N = 2048, PointNetSeg): flipped ReLU decisions sit within 2.0 u (fp16) / 2.7 u (bf16) of
    zero, flipped argmaxes within 5.1 u / 5.5 u of the maximum.  The gate is 8 u and 16 u.  In the fp32
    mode the difference to the oracle is the summation order of K <= 960 terms: sqrt(K) u, gate 256 u
    / 512 u (1.5e-5 / 3e-5 of the RMS)."""
    if mode == "fp32":
        return 256 * U[mode], 512 * U[mode]
    return 8 * U[mode], 16 * U[mode]


def check_branch_boundaries(branch, record, tol, arg_tol=None, stats=None):
    """Every branch decision that differs from the oracle's own must be within ``tol`` (ReLU sign;
    relative to the layer's RMS magnitude) / ``arg_tol`` (argmax: gap to the maximum) of its
    boundary.  Returns the number of differing decisions; ``stats`` (dict) receives the worst
    observed margins, in units of the layer RMS."""
    arg_tol = tol if arg_tol is None else arg_tol
    n_diff = 0
    worst_act = worst_arg = 0.0
    for key, mine in branch.items():
        kind, name = key.split(":", 1)
        if kind == "act":
            pre = record["pre:" + name]
            theirs = pre > 0
            diff = mine.reshape(pre.shape) != theirs
            if diff.any():
                scale = pre.float().pow(2).mean().sqrt().item()
                worst = pre[diff].abs().max().item()
                worst_act = max(worst_act, worst / max(scale, 1e-30))
                assert worst <= tol * max(scale, 1e-30), (key, worst, scale, int(diff.sum()))
                n_diff += int(diff.sum())
        elif kind in ("argmax", "rowargmax"):
            x = record[("maxin:" if kind == "argmax" else "chanmaxin:") + name]   # B x C x N (post-activation)
            dim = 2 if kind == "argmax" else 1
            theirs = record[kind + ":" + name]
            diff = mine != theirs
            if diff.any():
                top = x.max(dim)[0]
                at_mine = torch.gather(x, dim, mine.unsqueeze(dim)).squeeze(dim)
                gap = (top - at_mine)[diff]
                scale = x.float().pow(2).mean().sqrt().item()
                worst_arg = max(worst_arg, gap.max().item() / max(scale, 1e-30))
                assert gap.max().item() <= arg_tol * max(scale, 1e-30), (key, gap.max().item(), scale)
                n_diff += int(diff.sum())
    if stats is not None:
        stats["worst_act_margin"] = max(stats.get("worst_act_margin", 0.0), worst_act)
        stats["worst_argmax_margin"] = max(stats.get("worst_argmax_margin", 0.0), worst_arg)
    return n_diff


# ---------------------------------------------------------------------------------------------
# generic Functions (models/_mlp.py): the DEBUG_TAPE of a pass -> the oracle's ``branch`` dict
# ---------------------------------------------------------------------------------------------
def _to_layout(y, layout, B, C):
    """point-major / per-cloud rows [rows, >= C] of the CUDA path -> the oracle's tensor layout."""
    y = y.detach().float().cpu()[:, :C].contiguous()
    if layout == "bcn":
        return y.view(B, -1, C).transpose(1, 2)
    if layout == "bnc":
        return y.view(B, -1, C)
    if layout == "bc":
        return y.view(B, C)
    if layout == "bc1":
        return y.view(B, C, 1)
    raise ValueError(layout)


def tape_branches(tape, plan, record, B):
    """``plan``: one entry per PointMLPFunction call of the pass, in call order:
    dict(layers=[(oracle layer name, layout) | None, ...] for the stored layers of the call,
    reduce=(oracle layer name, "points" | "channels") | None).  ``record``: the oracle's own record
    of the same pass (shapes, and the reduce layer's pre-activations away from the argmax).  Returns
    the oracle ``branch`` dict with the CUDA path's decisions."""
    assert len(tape) == len(plan), (len(tape), len(plan))
    br = {}
    for rec, pl in zip(tape, plan):
        layers = pl.get("layers", [])
        for i, ent in enumerate(layers):
            if ent is None:
                continue
            name, layout = ent
            pre = record["pre:" + name]
            C = pre.shape[2] if layout == "bnc" else pre.shape[1]
            br["act:" + name] = _to_layout(rec["ys"][i], layout, B, C) > 0
        red = pl.get("reduce")
        if red is not None:
            name, kind = red
            idx = rec["red_idx"].detach().cpu().long()
            val = rec["red_val"].detach().float().cpu()
            if kind == "points":                                  # max over the cloud's points: idx [B, C]
                br["argmax:" + name] = idx
                if ("pre:" + name) in record:                     # the pooled layer has an activation
                    m = (record["pre:" + name] > 0).clone()
                    m.scatter_(2, idx.unsqueeze(2), (val > 0).unsqueeze(2))
                    br["act:" + name] = m
            else:                                                 # max over channels: idx per row -> B x N
                pre = record["pre:" + name]
                idx = idx.view(B, -1)
                br["rowargmax:" + name] = idx
                m = (pre > 0).clone()
                m.scatter_(1, idx.unsqueeze(1), (val.view(B, -1) > 0).unsqueeze(1))
                br["act:" + name] = m
    return br


def grad_report(named_cuda, oracle_params, extra=()):
    """Relative L2 error per tensor and over all tensors together."""
    pairs = [(k, v.grad, oracle_params[k].grad) for k, v in named_cuda if oracle_params[k].grad is not None
             or v.grad is not None]
    pairs += list(extra)
    errs, num, den = {}, 0.0, 0.0
    for k, a, b in pairs:
        assert a is not None and b is not None, k
        errs[k] = rel_err(a, b)
        num += (a.detach().double().cpu() - b.detach().double()).pow(2).sum().item()
        den += b.detach().double().pow(2).sum().item()
    return errs, (num / max(den, 1e-300)) ** 0.5


def _to(obj, dev):
    if torch.is_tensor(obj):
        return obj.to(dev)
    if isinstance(obj, dict):
        return {k: _to(v, dev) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to(v, dev) for v in obj)
    return obj


class tf32_eager:
    """Context: stock PyTorch eager on the GPU with TF32 tensor cores (cuDNN / cuBLAS) -- the
    arithmetic BASELINE.json's north_star names for the 1e-3 gate -- used as the yardstick: the SAME
    oracle, the SAME pinned decisions, only the matmul operands rounded to a 10-bit mantissa."""

    def __enter__(self):
        self.old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.old


def gate(ours, yard, tol, what):
    """The 16-bit gate: within ``tol``, or -- where the network itself amplifies operand rounding
    beyond it -- within 1.25 x what stock TF32 eager does under the same conditioning."""
    assert ours <= max(tol, 1.25 * yard), (what, ours, yard, tol)


def branch_parity(mode, tol, cuda_run, oracle_run, plan, named_cuda, sd, B, extra_grads=None):
    """The branch-conditioned comparison for a model built on the generic Functions.

    cuda_run()                              -> (list of output tensors, loss)
    oracle_run(params, branch, record, dev) -> (list of output tensors, loss), on device ``dev``
    The CUDA forward runs under a DEBUG_TAPE; the oracle runs twice on the CPU: with its own decisions
    (forward parity; every decision of the CUDA path that differs must be within ``flip_bounds`` of
    its boundary) and with the CUDA path's decisions (loss and gradient parity at ``tol``).  In the
    16-bit modes the oracle also runs a third time, by stock TF32 eager on the GPU with the same
    pinned decisions: the yardstick of ``gate``."""
    from adversarial_learning_on_pointclouds_b200.models import _mlp
    _mlp.DEBUG_TAPE = tape = []
    try:
        outs, loss = cuda_run()
    finally:
        _mlp.DEBUG_TAPE = None
    loss.backward()
    rec = {}
    with torch.no_grad():
        o_outs, _ = oracle_run(sd, None, rec, "cpu")
    rep = {"fwd": [rel_err(a, b) for a, b in zip(outs, o_outs)]}
    branch = tape_branches(tape, plan, rec, B)
    act_tol, arg_tol = flip_bounds(mode)
    stats = {}
    rep["n_branch_diff"] = check_branch_boundaries(branch, rec, act_tol, arg_tol, stats)
    rep["worst_act_margin_u"] = stats["worst_act_margin"] / U[mode]
    rep["worst_argmax_margin_u"] = stats["worst_argmax_margin"] / U[mode]
    p = steps.leaf_params(sd)
    b_outs, b_loss = oracle_run(p, branch, None, "cpu")
    b_loss.backward()
    rep["loss"] = abs(loss.item() - b_loss.item()) / max(abs(b_loss.item()), 1e-30)
    extra = extra_grads(b_outs) if extra_grads is not None else ()
    rep["grads"], rep["grad_total"] = grad_report(named_cuda, p, extra)
    # yardstick: the same oracle, same decisions, stock TF32 eager on the GPU
    rep["yard_fwd"], rep["yard_grads"], rep["yard_total"] = [0.0] * len(outs), {}, 0.0
    if mode != "fp32":
        q = steps.leaf_params(_to(sd, "cuda"))
        with tf32_eager():
            with torch.no_grad():
                y_own, _ = oracle_run(_to(sd, "cuda"), None, None, "cuda")
            y_outs, y_loss = oracle_run(q, _to(branch, "cuda"), None, "cuda")
            y_loss.backward()
        rep["yard_fwd"] = [rel_err(a, b) for a, b in zip(y_own, o_outs)]
        rep["yard_grads"], rep["yard_total"] = grad_report([(k, q[k]) for k, _ in named_cuda], p)
    for e, y in zip(rep["fwd"], rep["yard_fwd"]):
        gate(e, y, tol, "forward")
    gate(rep["loss"], 0.0, tol, "loss")
    gate(rep["grad_total"], rep["yard_total"], tol, "all gradients")
    for k, e in rep["grads"].items():
        gate(e, rep["yard_grads"].get(k, 0.0), 4 * tol, k)
    return rep


def seg_parity(model, pts, cls, seg, tol, glob_weight=0.5, mode=None):
    """Run PointNetSeg on the GPU and the oracle on the CPU with the same
    parameters and inputs; compare logits, global feature, loss and every
    parameter gradient.  Returns a report dict."""
    dev = pts.device
    B, N, _ = pts.shape
    model.zero_grad()
    dbg = {}
    model._debug = dbg
    pred, glob = model(pts, cls)
    model._debug = None
    loss = F.cross_entropy(pred, seg) + glob_weight * glob.square().mean()
    loss.backward()
    assert tuple(pred.shape) == (B, model.output_dim, N)
    assert tuple(pred.stride()) == (N * model.output_dim, 1, model.output_dim)
    assert tuple(glob.shape) == (B, 2048, 1)

    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    branch = seg_branches(dbg, B, N)
    # 1. oracle with its own decisions: forward parity + boundary check
    rec = {}
    with torch.no_grad():
        o_pred, o_glob = PO.pointnet_seg_forward(sd, pts.cpu(), cls.cpu(), record=rec)
    act_tol, arg_tol = flip_bounds(mode) if mode is not None else (max(64 * tol, 1e-4),) * 2
    stats = {}
    n_diff = check_branch_boundaries(branch, rec, act_tol, arg_tol, stats)
    rep = dict(n_branch_diff=n_diff,
               pred=rel_err(pred, o_pred), glob=rel_err(glob, o_glob))
    if mode is not None:
        rep["worst_act_margin_u"] = stats["worst_act_margin"] / U[mode]
        rep["worst_argmax_margin_u"] = stats["worst_argmax_margin"] / U[mode]
    # 2. oracle with the CUDA path's decisions: gradient parity
    p = steps.leaf_params(sd)
    b_pred, b_glob = PO.pointnet_seg_forward(p, pts.cpu(), cls.cpu(), branch=branch)
    b_loss = F.cross_entropy(b_pred, seg.cpu()) + glob_weight * b_glob.square().mean()
    b_loss.backward()
    rep["loss"] = abs(loss.item() - b_loss.item()) / abs(b_loss.item())
    rep["grads"] = {k: rel_err(v.grad, p[k].grad) for k, v in model.named_parameters()}
    num = sum((v.grad.double().cpu() - p[k].grad.double()).pow(2).sum()
              for k, v in model.named_parameters())
    den = sum(p[k].grad.double().pow(2).sum() for k, _ in model.named_parameters())
    rep["grad_total"] = (num / den).sqrt().item()
    rep["idx_equal"] = bool((branch["argmax:conv6"] == rec["argmax:conv6"]).all())
    assert rep["pred"] <= tol and rep["glob"] <= tol and rep["loss"] <= tol, rep
    assert rep["grad_total"] <= tol, rep
    assert max(rep["grads"].values()) <= 4 * tol, rep
    return rep
