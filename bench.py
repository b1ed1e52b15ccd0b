#!/usr/bin/env python3
"""Benchmark of the hot path: one adversarial PointNet segmentation train step
(utils/trainer.py:873-966 of the reference: PointNetSeg(50) generator +
PointwiseDiscNet discriminator, CE + BCE-with-logits, Adam on both).

    python bench.py --gpus N --steps K --warmup W          # the CUDA path
    python bench.py --impl reference ...                   # the CPU oracle port

Default workload = BASELINE.json configs[4] ("cfg5"): B = 256 labelled + 256
unlabelled clouds per GPU, N = 4096 points, weak scaling over GPUs.  Prints one
JSON line (see the contract in the task description / DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "adversarial_seg_train_step_throughput"
UNIT = "clouds/s"
WORKLOADS = {
    # name: (B labelled, B unlabelled, N points)
    "cfg5": (256, 256, 4096),
    "cfg3": (16, 16, 2048),
    "cfg3x2": (32, 32, 2048),
}
# algorithmic (useful) FLOPs of the fused conv6 + ReLU + max-over-points kernel: 2 * K * Cout per
# point, K = 512, Cout = 2048 (2 097 152 FLOP / point)
CONV6_FLOP_PER_POINT = 2 * 512 * 2048
CONV6_TAG = "linear:tc:k512:n2048:colmax"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons back to back (one query is ~30-50 ms) while running."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        sm, reasons, smax = [], set(), None
        for s in self.samples:
            if len(s) < 7:
                continue
            try:
                sm.append(float(s[0])); smax = float(s[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm))


def synthetic_inputs(B, N, seed):
    """The synthetic batch of SURVEY.md 8c, drawn in this order from one generator: pts U[-1,1)^3,
    y in [0,40), seg in [0,50), shape in [0,16).  (Kept here so that the CUDA arm imports nothing
    from oracle/; tests/test_oracle_golden.py checks it against the oracle's generator.)"""
    import torch
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(B, N, 3, generator=g) * 2 - 1
    y = torch.randint(0, 40, (B,), generator=g)
    seg = torch.randint(0, 50, (B, N), generator=g)
    shp = torch.randint(0, 16, (B,), generator=g)
    cls = torch.nn.functional.one_hot(shp, 16).to(torch.float32).view(B, 1, 16)
    return pts, y, seg, cls


def synthetic_batches(Bg, Bn, N, rank):
    """Synthetic ShapeNet-part-shaped clouds: seeds 1234 / 4321 (+rank)."""
    pts, _, seg, cls = synthetic_inputs(Bg, N, 1234 + 1000 * rank)
    pts2, _, _, cls2 = synthetic_inputs(Bn, N, 4321 + 1000 * rank)
    return (pts, cls, seg), (pts2, cls2)


# --------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the step: the oracle port (oracle/steps.py,
    pinned to the reference's golden vectors) on the host cores, on a bounded sample."""
    import torch
    from adversarial_learning_on_pointclouds_b200 import models as M
    from adversarial_learning_on_pointclouds_b200.utils import init_net
    from oracle import steps
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bg, Bn, N = WORKLOADS[args.workload]
    sb = max(1, min(Bg, args.cpu_sample_clouds))
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    opt = torch.optim.Adam(list(gp.values()), lr=1e-4, betas=(0.9, 0.999))
    optD = torch.optim.Adam(list(dp.values()), lr=1e-5, betas=(0.9, 0.999))
    bg, bn = synthetic_batches(sb, sb, N, 0)

    def step():
        opt.zero_grad(); optD.zero_grad()
        steps.adversarial_seg_step(gp, dp, bg, bn)
        opt.step(); optD.step()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = 2 * sb / dt
    sample = "%d+%d clouds of N=%d per step (bounded sample of %s), %d steps" % (sb, sb, N, args.workload,
                                                                                args.steps)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(name, world):
    Bg, Bn, N = WORKLOADS[name]
    return {"workload": "%s: adversarial PointNetSeg(50) + PointwiseDiscNet step "
                        "(utils/trainer.py:873-966), %d labelled + %d unlabelled clouds per GPU, "
                        "N=%d points, 50 parts" % (name, Bg, Bn, N),
            "clouds_per_gpu_per_step": Bg + Bn, "points_per_cloud": N, "parallelism": "dp%d" % world,
            "l2_policy": "working set per step (several GB of activations) exceeds the 126 MB L2; "
                         "no explicit flush"}


def cpu_baseline_sample(args, N):
    """Bounded CPU timing of the oracle port on this box's host cores (rank 0, N=1)."""
    import torch
    from adversarial_learning_on_pointclouds_b200 import models as M
    from adversarial_learning_on_pointclouds_b200.utils import init_net
    from oracle import steps
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sb = args.cpu_sample_clouds
    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier")
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier")
    gp, dp = steps.leaf_params(g.state_dict()), steps.leaf_params(d.state_dict())
    bg, bn = synthetic_batches(sb, sb, N, 0)
    times = []
    for i in range(3):
        for p_ in list(gp.values()) + list(dp.values()):
            p_.grad = None
        t0 = time.perf_counter()
        steps.adversarial_seg_step(gp, dp, bg, bn)
        times.append(time.perf_counter() - t0)
    dt = min(times[1:])
    return {"value": 2 * sb / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d+%d clouds of N=%d (bounded sample), best of 2 after 1 warm-up, fwd+bwd "
                      "without the Adam update" % (sb, sb, N)}


# --------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import adversarial_learning_on_pointclouds_b200 as pkg
    from adversarial_learning_on_pointclouds_b200 import models as M, ops, Precision
    from adversarial_learning_on_pointclouds_b200.utils import init_net
    from adversarial_learning_on_pointclouds_b200.trainer import (adversarial_seg_step,
                                                                  adversarial_seg_step_fused,
                                                                  GraphedAdversarialSegStep)
    from adversarial_learning_on_pointclouds_b200.parallel import DistributedOptimizer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Bg, Bn, N = WORKLOADS[args.workload]
    prec = Precision(args.precision)

    torch.manual_seed(0)
    g = init_net(M.PointNetSeg(50), "cpu", "xavier").to(dev)
    d = init_net(M.PointwiseDiscNet(N, 50), "cpu", "xavier").to(dev)
    g.precision = d.precision = prec
    use_graph = not args.no_graph
    opt = torch.optim.Adam(g.parameters(), lr=1e-4, betas=(0.9, 0.999), fused=True, capturable=use_graph)
    optD = torch.optim.Adam(d.parameters(), lr=1e-5, betas=(0.9, 0.999), fused=True, capturable=use_graph)
    if world > 1:
        opt, optD = DistributedOptimizer(opt), DistributedOptimizer(optD)
    targs = argparse.Namespace(device=dev, lambda_seg=1.0, lambda_adv=1e-3)
    gan_loss, seg_loss = torch.nn.BCEWithLogitsLoss(), torch.nn.CrossEntropyLoss()

    host_gt, host_nogt = synthetic_batches(Bg, Bn, N, rank)
    host_gt = tuple(t.pin_memory() for t in host_gt)
    host_nogt = tuple(t.pin_memory() for t in host_nogt)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_gt + host_nogt)
    dev_gt = tuple(t.to(dev) for t in host_gt)
    dev_nogt = tuple(t.to(dev) for t in host_nogt)
    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()

    fused = not args.no_fused
    step_fn = adversarial_seg_step_fused if fused else adversarial_seg_step

    def step(bg, bn):
        return step_fn(g, d, gan_loss, seg_loss, opt, optD, bg, bn, targs,
                       device_labels=args.device_labels)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        step(dev_gt, dev_nogt)

    # per-kernel times (eager pass of the same step; events on the launching stream)
    launches0 = pkg._lib.launch_count()
    with ops.KernelTimer() as kt:
        ms_eager = timed(lambda: step(dev_gt, dev_nogt), args.steps)
    launches = pkg._lib.launch_count() - launches0
    ksum = kt.summary()

    gstep, graph_note = None, "eager launches"
    if use_graph:
        try:
            gstep = GraphedAdversarialSegStep(g, d, gan_loss, seg_loss, opt, optD, targs, dev_gt,
                                              dev_nogt, warmup=1, device_labels=args.device_labels,
                                              fused=fused)
            graph_note = "whole step replayed from one CUDA graph (%d libpcadv launches per step)" \
                % gstep.launches_per_step
            launches = gstep.launches_per_step * args.steps
        except Exception as exc:                      # keep the eager path if capture is refused
            gstep, graph_note = None, "eager launches (graph capture failed: %s)" % str(exc)[:120]
            torch.cuda.synchronize()

    if gstep is not None:
        def resident_step():
            gstep()                                    # inputs already in the static HBM buffers

        gstep.prefetch(host_gt, host_nogt)

        def e2e_step():
            # this step's batch was put on the wire (pinned host -> device staging, copy stream)
            # while the previous step computed; the next batch's copy starts right after this
            # step is launched, so every timed step still carries one full host -> device copy
            losses = gstep.step_prefetched()
            gstep.prefetch(host_gt, host_nogt)
            loss_host.copy_(losses, non_blocking=True)
            torch.cuda.current_stream().synchronize()
    else:
        def resident_step():
            step(dev_gt, dev_nogt)

        def e2e_step():
            bg = tuple(t.to(dev, non_blocking=True) for t in host_gt)
            bn = tuple(t.to(dev, non_blocking=True) for t in host_nogt)
            losses = step(bg, bn)
            loss_host.copy_(torch.stack(losses), non_blocking=True)
            torch.cuda.current_stream().synchronize()

    for _ in range(3):
        resident_step()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(resident_step, args.steps)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    clouds = (Bg + Bn) * world * args.steps
    value = clouds / (ms / 1e3)
    e2e_value = clouds / (ms_e2e / 1e3)
    peaks = _peaks()
    roofline = None
    if CONV6_TAG in ksum:
        calls, tot_ms = ksum[CONV6_TAG]
        per_launch_s = tot_ms / calls / 1e3
        # one launch = one generator pass: clouds-per-pass x N points
        flops = CONV6_FLOP_PER_POINT * float(tot_points_per_launch(ksum, Bg, Bn, N))
        achieved = flops / per_launch_s / 1e12
        roofline = {"kernel": "tc_colmax_kernel (conv6 512->2048 + ReLU + max over points)",
                    "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch at 2^20 points from
                    # profiles/r01c_ncu_full_conv6max_r1c.csv (algorithmic: 1.074e9)
                    "traffic": 1.0898e9 if (Bg + Bn) // 2 * N == (1 << 20) else None,
                    "traffic_unit": "bytes per launch (ncu --set full, round 1)",
                    "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                    "peak_burst": peaks.get("bf16_burst"),
                    "frac_of_burst": (achieved / peaks["bf16_burst"]) if peaks.get("bf16_burst") else None,
                    "launch_ms": per_launch_s * 1e3, "share_of_step": tot_ms / ms_eager,
                    "timing": "CUDA events around the launch on the launching stream, in an eager "
                              "pass of the same step (%d steps, %.2f ms/step eager)" % (args.steps,
                                                                                       ms_eager / args.steps)}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp16": "fp16", "bf16": "bf16", "fp32": "f32"}[args.precision],
        "data": "synthetic", "config": workload_config(args.workload, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 12, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "launch_mode": graph_note,
        "step_api": ("fused loss heads (PointNetSeg.forward_ce / forward_logsoftmax, "
                     "trainer.adversarial_seg_step_fused)" if fused else
                     "reference-shaped modules + torch losses (trainer.adversarial_seg_step)"),
        "clocks": sampler.summary(),
        "roofline": roofline,
        "kernel_rooflines": kernel_rooflines(ksum, (Bg + Bn) / 2.0 * N, peaks, args.steps),
        "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in
                               sorted(ksum.items(), key=lambda kv: -kv[1][1])[:args.top_kernels]},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample(args, N)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: destroy_process_group() can block for minutes when a
        # captured CUDA graph still holds the communicator's kernels (seen on 2 x B200), and
        # nothing is left to flush.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def kernel_rooflines(ksum, points, peaks, steps, top=14):
    """Achieved HBM GB/s and tensor TFLOP/s of the per-point kernels against the measured peaks,
    from their call signatures (algorithmic bytes / FLOPs per point: DESIGN.md 4) and their mean
    launch time (CUDA events on the launching stream).  ``points`` = rows of one launch."""
    import re
    out = []
    for tag, (calls, tot_ms) in ksum.items():
        m = re.match(r"(linear|wgrad):tc:(?:k(\d+):n(\d+)|n(\d+):k(\d+))(.*)", tag)
        by = fl = None
        if m:
            kind, flags = m.group(1), m.group(6)
            k = int(m.group(2) or m.group(5))
            n = int(m.group(3) or m.group(4))
            fl = 2.0 * k * n
            if kind == "wgrad":
                by = 2.0 * (k + n)                                  # dz + x, 16-bit
            elif "colmax" in flags:
                by = 2.0 * k                                        # output never stored
            elif "rowmax" in flags:
                by = 2.0 * k + 8
            else:
                out_b = 4.0 * n if n == 50 else 2.0 * n             # fp32 logits, else 16-bit
                by = 2.0 * k + out_b + (n / 8.0 if n % 64 == 0 else 0.0)
                if ":mask" in flags and "maskbits" not in flags:
                    by += 2.0 * n
        elif tag == "linear:simt:k3:n64":
            by, fl = 12.0 + 128.0, 2.0 * 3 * 64
        elif tag.startswith("softmax_head"):
            by, fl = 200.0 + (256.0 if tag.endswith("ce") else 128.0) + 8.0, 0.0
        elif tag == "logsoftmax_bwd":
            by, fl = 3 * 128.0, 0.0
        if by is None:
            continue
        per_call_s = tot_ms / calls / 1e3
        gbs = by * points / per_call_s / 1e9
        tfs = fl * points / per_call_s / 1e12
        out.append({"kernel": tag, "calls_per_step": calls / steps, "ms_per_call": round(tot_ms / calls, 4),
                    "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / peaks["hbm"], 3),
                    "tflops": round(tfs, 1), "tensor_frac": round(tfs / peaks["bf16_sustained"], 3),
                    "ms_per_step": round(tot_ms / steps, 4)})
    out.sort(key=lambda d: -d["ms_per_step"])
    return out[:top]


def tot_points_per_launch(ksum, Bg, Bn, N):
    """Points one conv6 launch processes: the two generator passes of a step have Bg and Bn
    clouds; the average is what the per-launch mean duration corresponds to."""
    return (Bg + Bn) / 2.0 * N


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--device-labels", action="store_true",
                    help="draw the smoothed GAN labels on the device instead of the CPU")
    ap.add_argument("--cpu-sample-clouds", type=int, default=8,
                    help="clouds per batch of the bounded CPU sample (8 + 8 clouds of N points: about 1 s per step on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--top-kernels", type=int, default=12)
    ap.add_argument("--no-fused", action="store_true",
                    help="run the loop body on the reference-shaped forward() + torch softmax / CE "
                         "instead of the fused loss heads")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel eagerly")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
