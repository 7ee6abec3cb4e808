#!/usr/bin/env python
"""bench.py — the reference's headline workloads on synthetic Duckietown-shaped data (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload train|infer|mme]

  train (default; BASELINE.json configs[1], configs[2] under torchrun): SimpleTrainModule, batch 64 per GPU, 3x120x160
        fp32 images + int64 labels, bf16 activations, fwd + class-weighted CE on probabilities + bwd + AdamW, Dropout2d
        active (in-kernel RNG), BatchNorm batch statistics.  At N = 1 the whole step is replayed as one CUDA graph.
  infer (configs[3]): TrainingBase.evaluate_batch on 256 x 3 x 480 x 640 per GPU (masks + confusion matrix), eval BatchNorm
  mme   (configs[4]): MMETrainingModule, two streams of 64 images per GPU: SGD-nesterov entropy step + AdamW labelled step
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

KCLASS = ["dense_fwd", "gemm_px", "dgrad_dense", "dgrad_td", "wgrad", "head", "dz_prepare", "channel_stats", "first_fwd",
          "tu_col2im", "pack_weights", "optimizer", "misc", "dgrad_bulk", "wgrad_dense"]
NK = len(KCLASS)
WORKLOADS = {
    "train": ("train images/sec (whole box)",
              "SimpleTrainModule (FC-DenseNet67, 4 classes) fwd+weighted-CE+bwd+AdamW, 3x120x160, batch 64/GPU "
              "(README --batch_size=64 --augment config)"),
    "infer": ("inference images/sec (whole box)",
              "test.py loop: TrainingBase.evaluate_batch (FC-DenseNet67 eval mode, 4 classes) on 256x3x480x640 per GPU, "
              "masks + 4x4 confusion matrix + loss"),
    "mme": ("train images/sec (whole box)",
            "MMETrainingModule two-stream step: 64 unlabelled (entropy, grad-reverse, SGD-nesterov) + 64 labelled "
            "(weighted CE, AdamW) images per GPU and step, 3x120x160"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="train workload: eager launches instead of one CUDA graph per step")
    a = ap.parse_args()
    big = a.workload == "infer"
    a.batch = a.batch or (256 if big else 64)
    a.height = a.height or (480 if big else 120)
    a.width = a.width or (640 if big else 160)
    return a


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


def _dense_layers(H, W, B):
    """(level, cin, rmw_from, C0 of the block, position in block) of the 55 DenseLayers + per-level pixel counts"""
    skips = (128, 208, 288, 368, 448)
    px = [B * (H >> l) * (W >> l) for l in range(6)]
    dense, cin = [], 48
    for l in range(5):
        dense += [(l, cin + 16 * j, 0, cin, j) for j in range(5)]
        cin = skips[l]
    dense += [(5, 448 + 16 * j, 448 if j == 4 else 0, 448, j) for j in range(5)]
    for l in range(5):
        c0 = skips[l] + 80
        dense += [(l, c0 + 16 * j, c0 if (j == 4 and l != 0) else 0, c0, j) for j in range(5)]
    return dense, px, skips


def algorithmic_bytes(B, H, W, train=True):
    """per-step algorithmic HBM bytes of each kernel class: every operand read once, every result written
    once, bf16 activations.  Backward = the block-fused schedule (chain over a block's own new channels + ONE pass over the
    block input + per-layer weight gradients)."""
    dense, px, skips = _dense_layers(H, W, B)
    out = {k: 0 for k in KCLASS}
    for l, c, rmw, c0, j in dense:
        out["dense_fwd"] += px[l] * (c + 16) * 2
        if train:
            out["dgrad_dense"] += px[l] * (16 * j * 6 + 32) if j else 0          # chain: x, dX read + write of 16j channels, dz16
            out["wgrad_dense"] += px[l] * (2 * c + 32)                          # x, dz16
            if j == 4:                                                           # one bulk pass per block: 5 dz16, x, dX (write or rmw)
                out["dgrad_bulk"] += px[l] * (5 * 32 + 2 * c0 + (2 * c0 if rmw else 4 * c0))
    for l in range(5):
        C = skips[l]
        out["gemm_px"] += px[l] * C * 2 + px[l + 1] * C * 2                      # TD forward
        out["gemm_px"] += px[l + 1] * 80 * 2 + px[l] * 80 * 2                    # TU forward: input slice, output slice
        if train:
            out["dgrad_td"] += px[l] * (2 * C + 4 * C + 2 * C)                   # x, dX rmw, dz
            out["wgrad"] += px[l] * (2 * C + 2 * C)
            out["gemm_px"] += px[l] * 80 * 2 + px[l + 1] * 80 * 2                # TU input gradient: dz parity planes, dX slice
            out["wgrad"] += px[l + 1] * 80 * 2 + px[l] * 80 * 2                  # TU weight gradient: x slice, dz parity planes
    if train:
        out["wgrad"] += px[0] * (288 * 2 + 32) + px[0] * (64 + 96)               # classifier, first conv
        out["head"] = px[0] * (288 * 2 + 288 * 2 + 32 + 1)                       # features, d(features), dl, label
    else:
        out["head"] = px[0] * (288 * 2 + 1 + 1)                                  # features, label, mask
    return out


def class_flops(B, H, W, train=True):
    """tensor-core FLOPs (2*MAC) per step and kernel class (the mask operand doubles wgrad_dense's MMA work)"""
    dense, px, skips = _dense_layers(H, W, B)
    f = {k: 0 for k in KCLASS}
    for l, c, rmw, c0, j in dense:
        f["dense_fwd"] += px[l] * 2 * 9 * c * 16
        if train:
            f["dgrad_dense"] += px[l] * 2 * 9 * 16 * j * 16
            f["dgrad_bulk"] += px[l] * 2 * 9 * c0 * 16
            f["wgrad_dense"] += px[l] * 2 * 9 * (2 * c) * 16
    for l in range(5):
        C = skips[l]
        f["gemm_px"] += px[l] * 2 * C * C + px[l + 1] * 2 * 80 * 720 * (2 if train else 1)
        if train:
            f["dgrad_td"] += px[l] * 2 * C * C
            f["wgrad"] += px[l] * 2 * C * C + px[l + 1] * 2 * 80 * 720
    return f


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


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------ CPU reference (oracle port)
def cpu_reference(workload, H, W, steps, warmup, budget_s):
    """The reference's own CPU path: its model code cannot travel to the GPU box (/root/reference is absent there) and
    pytorch_lightning 1.2.1 is not installable, so this times the oracle port (plain PyTorch fp32, the reference's ops:
    conv2d / batch_norm-equivalent / conv_transpose2d / max_pool2d / cross_entropy, torch.optim.AdamW / SGD) on all host
    cores, on a bounded sample of the workload.  Returns (images/s, cores, sample description, ms per step)."""
    import torch
    from oracle import fcdensenet as net
    from oracle import steps as osteps
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    total = max(steps + warmup, 1)
    sd = net.init_state_dict(num_cls=4, seed=42)
    if workload == "infer":
        bs = 1
        x, y = osteps.synthetic_batch(bs, H, W, 4, seed=42)
        fn = lambda: osteps.evaluate_batch(sd, x, y)
        per_step, sample = bs, f"1 of the 256 images, {H}x{W}, fp32, evaluate_batch (oracle/steps.py)"
    elif workload == "mme":
        bs = max(1, min(4, int(budget_s / (total * 1.4))))
        x, y = osteps.synthetic_batch(bs, H, W, 4, seed=42)
        xu, _ = osteps.synthetic_batch(bs, H, W, 4, seed=43)
        tr = osteps.Trainer(sd, num_cls=4)
        ext = [tr.sd[n] for n in tr.param_names if n.startswith("featureExtractor.")]
        cls = [tr.sd[n] for n in tr.param_names if n.startswith("classifier.")]
        sgd = torch.optim.SGD([{"params": ext, "lr": 1e-3 / 3}, {"params": cls, "lr": 1e-3}], lr=1e-3, momentum=0.9,
                              nesterov=True, weight_decay=1e-4)

        def fn():
            sgd.zero_grad(set_to_none=True)
            loss, _ = osteps.mme_unlabelled_loss(tr.sd, xu)
            loss.backward()
            sgd.step()
            tr.step(x, y)
        per_step = 2 * bs
        sample = f"{bs}+{bs} of the 64+64 images of a step, {H}x{W}, fp32 (oracle/steps.py mme_unlabelled_loss + Trainer)"
    else:
        bs = max(1, min(8, int(budget_s / (total * 0.65))))
        x, y = osteps.synthetic_batch(bs, H, W, 4, seed=42)
        tr = osteps.Trainer(sd, num_cls=4)
        fn = lambda: tr.step(x, y)
        per_step, sample = bs, f"batch {bs} of the 64-image step, {H}x{W}, fp32 (oracle/steps.py Trainer)"
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    return per_step * steps / dt, cores, f"{sample}, {steps} steps after {warmup} warm-up", 1e3 * dt / max(steps, 1)


def run_reference(args, rank):
    if rank != 0:
        return
    metric, wl = WORKLOADS[args.workload]
    v, cores, sample, ms = cpu_reference(args.workload, args.height, args.width, args.steps, args.warmup, 150.0)
    line = {"impl": "reference", "metric": metric, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": {"workload": wl, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
class MultiPrefetcher:
    """sim2real_lane_segment_b200.prefetch.DevicePrefetcher for batches of any number of tensors (the MME batch: labelled
    x, labels, unlabelled x): pinned host batch -> device on a copy stream, issued one step ahead"""

    def __init__(self, batches, device):
        import torch
        self.batches, self.device = batches, device
        self.copy_stream = torch.cuda.Stream(device)

    def __iter__(self):
        import torch
        it = iter(self.batches)
        bufs, copied, released = [None, None], [torch.cuda.Event(), torch.cuda.Event()], [None, None]

        def stage(slot, batch):
            if bufs[slot] is None:
                bufs[slot] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch)
            with torch.cuda.stream(self.copy_stream):
                if released[slot] is not None:
                    self.copy_stream.wait_event(released[slot])
                for d, s in zip(bufs[slot], batch):
                    d.copy_(s, non_blocking=True)
                copied[slot].record(self.copy_stream)
        stage(0, next(it))
        slot = 0
        while True:
            torch.cuda.current_stream(self.device).wait_event(copied[slot])
            out = bufs[slot]
            stage(slot ^ 1, next(it))
            yield out
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            released[slot] = ev
            slot ^= 1


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from sim2real_lane_segment_b200 import _lib, ddp
    from sim2real_lane_segment_b200.graphs import GraphedTrainStep
    from sim2real_lane_segment_b200.trainingModules import MMETrainingModule, SimpleTrainModule

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version / debug lines to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    B, H, W, C = args.batch, args.height, args.width, 4
    wk = args.workload
    metric, wl_name = WORKLOADS[wk]
    torch.manual_seed(42)
    L = _lib.lib()
    L.rl_launch_count.restype = ctypes.c_long
    nb = 4 if wk != "infer" else 2
    imgs_per_step = B * (2 if wk == "mme" else 1)

    # ---- synthetic data: a few distinct batches, per-rank seed: pinned on the host + resident copies
    host = []
    for i in range(nb):
        x, y = synthetic_batch(B, H, W, C, seed=42 + rank * 100 + i)
        item = [x.pin_memory(), y.pin_memory()]
        if wk == "mme":
            item.append(synthetic_batch(B, H, W, C, seed=4242 + rank * 100 + i)[0].pin_memory())
        host.append(tuple(item))
    devb = [tuple(t.to(dev) for t in item) for item in host]
    graph_note, collective, graphed = None, None, None
    result_host = torch.empty(1, pin_memory=True)

    if wk == "train":
        model = SimpleTrainModule(lr=1e-3, decay=1e-4, lrRatio=1e3, num_cls=C).to(dev).train()
        (opt,), _ = model.configure_optimizers()
        rt = model.runtime()
        if world > 1:
            rt.grad_sync = ddp.GradSync()      # all-reduce buckets are fired from inside the backward pass
            collective = ddp.GradSync.collective

        def eager_step(batch, i):
            opt.zero_grad()
            loss = model.training_step((batch[0], batch[1]), i)
            loss.backward()
            if rt.grad_sync is not None:
                opt.grad_scale = rt.grad_sync.finish()
            opt.step()
            return loss
        step_fn = eager_step
        if not args.no_graph:
            ok = True
            try:
                graphed = GraphedTrainStep(model, opt, *devb[0], warmup=3)
            except Exception as e:                    # (a capture the NCCL build refuses: stay eager on every rank)
                ok, graphed = False, None
                sys.stderr.write(f"rank {rank}: CUDA-graph capture failed ({type(e).__name__}: {str(e)[:200]}); eager steps\n")
            if world > 1:
                flag = torch.tensor([1 if ok else 0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag) == 0:
                    graphed = None
            if graphed is not None:
                graph_note = ("whole step (zero_grad, forward, head+loss, backward, AdamW) replayed as ONE CUDA graph" if world == 1 else
                              "step replayed as FOUR CUDA graphs (forward + up-path backward | deep down path | shallow down path | "
                              "AdamW) with the NCCL all-reduce of each finished parameter range launched between them")
                step_fn = lambda batch, i: graphed(batch[0], batch[1])
        prof_step = eager_step
        h2d, d2h = B * 3 * H * W * 4 + B * H * W * 8, 4
        fetch = lambda out: result_host.copy_(out.detach().reshape(1), non_blocking=True)
    elif wk == "mme":
        model = MMETrainingModule(lr=1e-3, decay=1e-4, lrRatio=1e3, num_cls=C).to(dev).train()
        (optG, optF), _ = model.configure_optimizers()
        rt = model.runtime()
        if world > 1:
            rt.grad_sync = ddp.GradSync()
            collective = ddp.GradSync.collective + "; twice per step (one exchange per optimizer)"

        def step_fn(batch, i):
            out = None
            for idx, o in enumerate((optG, optF)):          # Lightning: one training_step per optimizer and batch
                o.zero_grad()
                out = model.training_step((batch[0], batch[2], batch[1], None), i, optimizer_idx=idx)
                out.backward()
                if rt.grad_sync is not None:
                    o.grad_scale = rt.grad_sync.finish()
                o.step()
            return out
        prof_step = step_fn
        h2d, d2h = 2 * B * 3 * H * W * 4 + B * H * W * 8, 4
        fetch = lambda out: result_host.copy_(out.detach().reshape(1), non_blocking=True)
    else:
        model = SimpleTrainModule(num_cls=C).to(dev).eval()
        rt = model.runtime()
        cm_total = torch.zeros(C, C, dtype=torch.long, device=dev)

        def step_fn(batch, i):
            out = model.evaluate_batch((batch[0], batch[1]), return_predictions=True, return_probs=False)
            cm_total.add_(out["confusion_matrix"])
            return out
        prof_step = step_fn
        h2d, d2h = B * 3 * H * W * 4 + B * H * W * 8, B * H * W + C * C * 8 + 4
        mask_host = torch.empty(B, H, W, dtype=torch.uint8).pin_memory()
        cm_host = torch.empty(C, C, dtype=torch.long).pin_memory()

        def fetch(out):
            mask_host.copy_(out["labels_hat"], non_blocking=True)
            cm_host.copy_(out["confusion_matrix"], non_blocking=True)
            result_host.copy_((out["loss"] / out["weight"]).reshape(1), non_blocking=True)

    step_resident = lambda i: step_fn(devb[i % nb], i)

    def host_batches():
        i = 0
        while True:
            yield host[i % nb]
            i += 1
    feed = iter(MultiPrefetcher(host_batches(), dev))

    def step_e2e(i):
        """pinned HOST batch -> device (the copy is issued one step ahead on a copy stream, inside the timed region) ->
        step through the public module API -> result (loss / masks + confusion matrix) back to the host"""
        out = step_fn(next(feed), i)
        fetch(out)
        return out

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

    W_ = max(args.warmup, 3)
    for i in range(W_):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = L.rl_launch_count(-1)
    ms = timed(step_resident, args.steps)
    launches = L.rl_launch_count(-1) - n0
    if graphed is not None:       # replays launch what was captured; the library's host-side counter only saw the capture
        launches = graphed.captured_launches * args.steps
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    value = world * imgs_per_step * args.steps / (ms / 1e3)
    e2e = world * imgs_per_step * args.steps / (ms_e2e / 1e3)

    # ---- per-class CUDA-event timing of our kernels on the launching stream (2 profiled eager steps, overlap off)
    roof, tc_util = None, None
    if rank == 0:
        msb = (ctypes.c_double * NK)()
        lnb = (ctypes.c_long * NK)()
        sync_saved, rt.grad_sync = rt.grad_sync, None   # rank-0-only pass: no collectives
        prof_step(devb[0], 0)                           # (re)creates the eager plan state after graph replays
        for eng in rt.engines.values():                 # time every kernel alone (side-stream overlap would smear the class times)
            eng.set_option("overlap", 0)
        torch.cuda.synchronize()
        L.rl_profile_begin()
        nprof = 2
        for i in range(nprof):
            prof_step(devb[i % nb], i)
        L.rl_profile_end(msb, lnb)
        for eng in rt.engines.values():
            eng.set_option("overlap", 1)
        rt.grad_sync = sync_saved
        per = {KCLASS[i]: (msb[i] / nprof, lnb[i] // nprof) for i in range(NK)}
        tot = sum(v[0] for v in per.values())
        shares = {k: round(v[0] / tot, 4) for k, v in per.items() if v[1]}
        mult = 2 if wk == "mme" else 1
        algo = {k: v * mult for k, v in algorithmic_bytes(B, H, W, train=wk != "infer").items()}
        flops = {k: v * mult for k, v in class_flops(B, H, W, train=wk != "infer").items()}
        top = max((k for k in per if algo.get(k, 0) > 0), key=lambda k: per[k][0])
        pk = peaks()
        peak, which = (pk["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in pk else (6650.0, "fallback")
        tcp, tcw = ((pk["bf16_tflops_sustained"], "measured sustained cuBLAS bf16 (MEASURED_PEAKS.json bf16_tflops_sustained)")
                    if "bf16_tflops_sustained" in pk else (1400.0, "fallback"))
        t_ms, n_l = per[top]
        achieved = algo[top] / (t_ms / 1e3) / 1e9
        roof = {"kernel": top, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": which,
                "launches_per_step": int(n_l), "avg_launch_ms": round(t_ms / max(n_l, 1), 4),
                "algorithmic_bytes_per_step": int(algo[top]),
                "timing_note": "per-class CUDA-event sums of 2 extra eager steps run with the side-stream overlap disabled "
                               "(each kernel alone on the GPU); the timed region above runs with overlap enabled",
                "class_ms_per_step": {k: round(v[0], 3) for k, v in per.items() if v[1]},
                "class_frac_of_hbm_peak": {k: round(algo[k] / (per[k][0] / 1e3) / 1e9 / peak, 3)
                                           for k in per if per[k][1] and algo.get(k, 0) > 0},
                "time_shares": shares,
                "step_algorithmic_gb": round(sum(algo.values()) / 1e9, 2),
                "step_frac_of_hbm_peak": round(sum(algo.values()) / (ms / args.steps / 1e3) / 1e9 / peak, 3)}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            roof["traffic"] = tj.get(top)
            roof["traffic_note"] = tj.get(top + "_note")
        except Exception:
            pass
        # conv tensor-core utilisation (BASELINE.json metric): model FLOPs of each tensor-core kernel class / its time,
        # against the measured cuBLAS bf16 throughput
        tc_util = {"peak_tflops": tcp, "peak_source": tcw, "per_class": {}}
        for k in per:
            if per[k][1] and flops.get(k, 0) > 0:
                tf = flops[k] / (per[k][0] / 1e3) / 1e12
                tc_util["per_class"][k] = {"tflops": round(tf, 1), "frac": round(tf / tcp, 3)}
        conv_t = sum(per[k][0] for k in per if flops.get(k, 0) > 0)
        tf_all = sum(flops.values()) / (conv_t / 1e3) / 1e12 if conv_t > 0 else 0.0
        tc_util["all_conv_kernels"] = {"tflops": round(tf_all, 1), "frac": round(tf_all / tcp, 3)}

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference(wk, H, W, 3, 1, 20.0)     # cpu_baseline leg: the only place this arm touches oracle/
        cpu = {"value": round(v, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        fl = conv_flops_per_image(H, W)
        cfg = {"workload": wl_name, "global_batch": imgs_per_step * world, "parallelism": f"dp{world}",
               "l2": f"per-step working set (GBs of activation arenas) exceeds the 126 MB L2; {nb} rotating input batches",
               "bn": "batch statistics, local per rank" if wk != "infer" else "running statistics (eval mode)"}
        if wk != "infer":
            cfg["dropout"] = "Dropout2d(0.2) active, in-kernel counter RNG (device-side step counter)"
        if graph_note:
            cfg["cuda_graph"] = graph_note
        if collective:
            cfg["collective"] = collective
        line = {"metric": metric, "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": W_, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
                "e2e": {"value": round(e2e, 2), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": round(ms_e2e / args.steps, 4)},
                "gpu_launches": int(launches),
                "clocks": sampler.summary(),
                "roofline": roof,
                "conv_tc_util": tc_util,
                "cpu_baseline": cpu,
                "model_tflops": round((3 if wk != "infer" else 1) * fl * value / 1e12, 1)}
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
