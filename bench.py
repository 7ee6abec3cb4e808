#!/usr/bin/env python
"""bench.py — SimpleTrainModule (FC-DenseNet67, 4 classes) training throughput on synthetic Duckietown-shaped data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): batch 64 per GPU, 3x120x160 fp32 images + int64 labels, bf16 activations,
fwd + class-weighted CE on probabilities + bwd + AdamW, Dropout2d active (in-kernel RNG), BatchNorm batch statistics.
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train images/sec (whole box)"
KCLASS = ["dense_fwd", "gemm_px", "dgrad_dense", "dgrad_td", "wgrad", "head", "dz_prepare", "channel_stats", "first_fwd",
          "tu_col2im", "pack_weights", "optimizer", "misc", "dgrad_bulk", "wgrad_dense"]
NK = len(KCLASS)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--height", type=int, default=120)
    ap.add_argument("--width", type=int, default=160)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler(threading.Thread):
    """nvidia-smi SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons}


def algorithmic_bytes(B, H, W):
    """per-step algorithmic HBM bytes of each tensor-core kernel class: every operand read
    once, every result written once, bf16 activations.  Returns {class: bytes}."""
    skips = (128, 208, 288, 368, 448)
    lv = [(H >> l, W >> l) for l in range(6)]
    px = [B * h * w for h, w in lv]
    dense = []  # (level, cin, rmw_from)
    cin = 48
    for l in range(5):
        dense += [(l, cin + 16 * j, 0) for j in range(5)]
        cin = skips[l]
    dense += [(5, 448 + 16 * j, 448 if j == 4 else 0) for j in range(5)]
    for l in range(5):
        dense += [(l, skips[l] + 80 + 16 * j, (skips[l] + 80) if (j == 4 and l != 0) else 0) for j in range(5)]
    out = {k: 0 for k in KCLASS}
    for l, c, rmw in dense:
        out["dense_fwd"] += px[l] * (c + 16) * 2
        out["dgrad_dense"] += px[l] * (2 * c + 2 * c + 2 * (c - rmw) + 32)   # x, dX write, dX read, dz16
        out["wgrad"] += px[l] * (2 * c + 32)
    for l in range(5):
        C = skips[l]
        out["gemm_px"] += px[l] * C * 2 + px[l + 1] * C * 2                   # TD forward
        out["dgrad_td"] += px[l] * (2 * C + 4 * C + 2 * C)                    # x, dX rmw, dz
        out["wgrad"] += px[l] * (2 * C + 2 * C)
        out["gemm_px"] += px[l + 1] * (80 * 2 + 720 * 2)                      # TU forward column matrix
        out["gemm_px"] += px[l + 1] * (720 * 2 + 80 * 2)                      # TU input gradient
        out["wgrad"] += px[l + 1] * (80 * 2 + 720 * 2)
    out["wgrad"] += px[0] * (288 * 2 + 32) + px[0] * (64 + 96)                # classifier, first conv
    out["head"] = px[0] * (288 * 2 + 288 * 2 + 32 + 1)
    return out


def conv_flops_per_image(H, W):
    """forward conv FLOPs (2*MAC) per image, SURVEY.md §6: 15.906 G at 120x160"""
    skips = (128, 208, 288, 368, 448)
    lv = [(H >> l) * (W >> l) for l in range(6)]
    f = lv[0] * 2 * 27 * 48
    cin = 48
    for l in range(5):
        f += sum(lv[l] * 2 * 9 * (cin + 16 * j) * 16 for j in range(5))
        f += lv[l] * 2 * skips[l] * skips[l]
        cin = skips[l]
    f += sum(lv[5] * 2 * 9 * (448 + 16 * j) * 16 for j in range(5))
    for l in range(5):
        f += lv[l + 1] * 2 * 9 * 80 * 80
        f += sum(lv[l] * 2 * 9 * (skips[l] + 80 + 16 * j) * 16 for j in range(5))
    f += lv[0] * 2 * 288 * 4
    return f


def synthetic_batch(B, H, W, num_cls, seed):
    """Duckietown-shaped synthetic data: x ~ N(0,1) like post-Normalize images, y a piecewise-constant
    4-class map (background, right-lane trapezoid, left-lane trapezoid, obstacle blob) so classes are spatially coherent."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, H, W, generator=g)
    y = torch.zeros(B, H, W, dtype=torch.int64)
    ys = torch.arange(H).view(H, 1).float()
    xs = torch.arange(W).view(1, W).float()
    r = torch.rand(B, 6, generator=g)
    for b in range(B):
        top = H * (0.45 + 0.1 * float(r[b, 0]))
        cx = W * (0.35 + 0.3 * float(r[b, 1]))
        half = W * (0.03 + 0.22 * ((ys - top) / max(H - top, 1)).clamp(0, 1))
        low = ys >= top
        y[b][low & (xs >= cx) & (xs < cx + 2 * half)] = 1
        if num_cls > 2:
            y[b][low & (xs < cx) & (xs >= cx - 2 * half)] = 2
        if num_cls > 3 and float(r[b, 2]) > 0.3:
            oy, ox = H * (0.5 + 0.4 * float(r[b, 3])), W * float(r[b, 4])
            y[b][((ys - oy) ** 2 + (xs - ox) ** 2) < (0.06 * H) ** 2] = 3
    return x, y


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    """The reference's own CPU path.  Its model code cannot travel to the GPU box (/root/reference is absent there) and
    pytorch_lightning 1.2.1 is not installable, so this times the oracle port (plain PyTorch fp32, same ops as the
    reference: nn.functional conv2d / batch_norm-equivalent / conv_transpose2d / max_pool2d, torch.optim.AdamW) on all
    host cores, on a bounded sample of the workload."""
    if rank != 0:
        return
    import torch
    from oracle import fcdensenet as net
    from oracle import steps
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    total = args.steps + args.warmup
    bs = max(1, min(8, int(150.0 / (max(total, 1) * 0.65))))
    sd = net.init_state_dict(num_cls=4, seed=42)
    x, y = steps.synthetic_batch(bs, args.height, args.width, 4, seed=42)
    tr = steps.Trainer(sd, num_cls=4)
    g = torch.Generator().manual_seed(0)
    for _ in range(args.warmup):
        tr.step(x, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr.step(x, y)
    dt = time.perf_counter() - t0
    v = bs * args.steps / dt
    sample = f"batch {bs} of the 64-image step, {args.height}x{args.width}, fp32, {args.steps} steps"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "SimpleTrainModule (FC-DenseNet67, 4 classes) fwd+weighted-CE+bwd+AdamW, 3x120x160, "
                                   "batch 64/GPU (README --batch_size=64 --augment config)", "sample": sample},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from sim2real_lane_segment_b200 import _lib, ddp
    from sim2real_lane_segment_b200.trainingModules import SimpleTrainModule

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, H, W, C = args.batch, args.height, args.width, 4
    torch.manual_seed(42)
    model = SimpleTrainModule(lr=1e-3, decay=1e-4, lrRatio=1e3, num_cls=C).to(dev).train()
    (opt,), _ = model.configure_optimizers()
    rt = model.runtime()
    sync = ddp.GradSync(n_buckets=3) if world > 1 else None
    L = _lib.lib()
    L.rl_launch_count.restype = ctypes.c_long

    # synthetic data: a few distinct batches, per-rank seed
    nb = 4
    host = []
    for i in range(nb):
        x, y = synthetic_batch(B, H, W, C, seed=42 + rank * 100 + i)
        host.append((x.pin_memory(), y.pin_memory()))
    devb = [(x.to(dev), y.to(dev)) for x, y in host]

    def step_resident(i):
        x, y = devb[i % nb]
        opt.zero_grad()
        loss = model.training_step((x, y), i)
        loss.backward()
        if world > 1:
            sync.start(rt.grads)
            opt.grad_scale = sync.finish()
        opt.step()
        return loss

    def step_local(i):
        """same step without the gradient exchange (used for the per-kernel event timing on rank 0 only)"""
        x, y = devb[i % nb]
        opt.zero_grad()
        loss = model.training_step((x, y), i)
        loss.backward()
        opt.step()
        return loss

    loss_host = torch.empty(1, pin_memory=True)
    # end to end: pinned HOST batches, each copied to the device inside the timed region (one H2D per step, issued by the
    # package's DevicePrefetcher on a copy stream one step ahead), and the loss read back to the host every step
    from sim2real_lane_segment_b200.prefetch import DevicePrefetcher

    def host_batches():
        i = 0
        while True:
            yield host[i % nb]
            i += 1

    feed = iter(DevicePrefetcher(host_batches(), dev))

    def step_e2e(i):
        xin, yin = next(feed)
        opt.zero_grad()
        loss = model.training_step((xin, yin), i)
        loss.backward()
        if world > 1:
            sync.start(rt.grads)
            opt.grad_scale = sync.finish()
        opt.step()
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = L.rl_launch_count(-1)
    ms = timed(step_resident, args.steps)
    launches = L.rl_launch_count(-1) - n0
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    # ---- per-class CUDA-event timing of our kernels on the launching stream (2 profiled steps)
    roof, shares = None, None
    if rank == 0:
        msb = (ctypes.c_double * NK)()
        lnb = (ctypes.c_long * NK)()
        for eng in rt.engines.values():   # time every kernel alone (the side-stream overlap would smear the class times)
            eng.set_option("overlap", 0)
        L.rl_profile_begin()
        nprof = 2
        for i in range(nprof):
            step_local(i)
        L.rl_profile_end(msb, lnb)
        for eng in rt.engines.values():
            eng.set_option("overlap", 1)
        per = {KCLASS[i]: (msb[i] / nprof, lnb[i] // nprof) for i in range(NK)}
        tot = sum(v[0] for v in per.values())
        shares = {k: round(v[0] / tot, 4) for k, v in per.items() if v[1]}
        algo = algorithmic_bytes(B, H, W)
        top = max((k for k in per if algo.get(k, 0) > 0), key=lambda k: per[k][0])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        t_ms, n_l = per[top]
        achieved = algo[top] / (t_ms / 1e3) / 1e9
        roof = {"kernel": top, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": which,
                "launches_per_step": int(n_l), "avg_launch_ms": round(t_ms / max(n_l, 1), 4),
                "algorithmic_bytes_per_step": int(algo[top]),
                "timing_note": "per-class CUDA-event sums of 2 extra steps run with the side-stream overlap disabled "
                               "(each kernel alone on the GPU); the timed region above runs with overlap enabled",
                "class_ms_per_step": {k: round(v[0], 3) for k, v in per.items() if v[1]},
                "time_shares": shares}
        ncu_traffic = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(ncu_traffic):
            try:
                tj = json.load(open(ncu_traffic))
                roof["traffic"] = tj.get(top)
                roof["traffic_note"] = tj.get(top + "_note", tj.get("note"))
            except Exception:
                pass

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import fcdensenet as net      # cpu_baseline leg: the only place this arm touches oracle/
        from oracle import steps as osteps
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = net.init_state_dict(num_cls=C, seed=42)
        x, y = osteps.synthetic_batch(8, H, W, C, seed=42)
        tr = osteps.Trainer(sd, num_cls=C)
        tr.step(x, y)
        t0 = time.perf_counter()
        nrep = 3
        for _ in range(nrep):
            tr.step(x, y)
        dt = time.perf_counter() - t0
        cpu = {"value": round(8 * nrep / dt, 3), "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"batch 8 of the 64-image step, {H}x{W}, fp32, {nrep} steps after 1 warm-up (oracle/steps.py Trainer)"}

    if rank == 0:
        fl = conv_flops_per_image(H, W)
        line = {"metric": METRIC, "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "SimpleTrainModule (FC-DenseNet67, 4 classes) fwd+weighted-CE+bwd+AdamW, 3x120x160, "
                                       "batch 64/GPU (README --batch_size=64 --augment config)",
                           "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2": "per-step working set (>2 GB of activation arenas) exceeds the 126 MB L2; 4 rotating input batches",
                           "dropout": "Dropout2d(0.2) active, in-kernel counter RNG", "bn": "batch statistics, local per rank"},
                "e2e": {"value": round(e2e, 2), "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W * 4 + B * H * W * 8,
                        "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 4)},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": roof,
                "cpu_baseline": cpu,
                "model_tflops": round(3 * fl * value / 1e12, 1)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()   # ranks > 0 wait here while rank 0 finishes its (collective-free) profiling pass
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
