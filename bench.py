#!/usr/bin/env python
"""Headline benchmark: 7-camera frames/s -> 3-D pose (BASELINE.json metric).

A step = one pass of the hot path over a batch of synthetic frames:
    256 frames x 7 cameras of 256x256 uint8 (BASELINE.json configs[1] shape, 8-stack hourglass,
    bf16 tensor-core convs, 19 maps/image) -> arg-max -> 19->38 packing -> DLT -> bundle adjustment
    -> DLT  (the 2D->3D tail of configs[2]), i.e. everything Core.pose2d_estimation +
    calibrate_calc + save run in the reference, minus file I/O.

  value : frames/s with the images already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public pipeline call with PINNED HOST images: every timed step issues
          one H2D copy of a full step's images (double-buffered on a copy stream: the copy of step k+1
          overlaps the compute of step k) and the D2H of the 3-D joints + cameras
  roofline : dominant kernels = conv_chain_kernel + conv_gemm_kernel (tcgen05 conv chains and implicit-GEMM
          convs), achieved = algorithmic conv FLOPs of the step / summed device time of their launches (CUDA
          events on the launching stream, taken during the timed region), peak = MEASURED_PEAKS.json
          bf16_tflops_sustained (fallback 1400 TF/s "of fallback"); traffic = DRAM bytes per conv launch
          from the committed ncu metrics pass (profiles/conv_gemm_traffic.json)
  cpu_baseline : the CPU oracle (PyTorch fp32 hourglass + numpy DLT + SciPy BA, all host threads)
          on a bounded sample

N > 1 (torchrun): frames shard across ranks (weak scaling: 256 frames per rank), the bundle
adjustment all-reduces its 42x42 reduced camera system per iteration, one all-gather of the 3-D
joints at the end.

`--impl reference` times the CPU oracle alone (the reference's df2d/pyba are not installable
here: un-vendored dependencies, no network) and prints the same JSON line with impl=reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_RANK = 256
IN_H = IN_W = 256
NUM_STACKS = 8
CAMS = 7
GFLOP_PER_IMAGE = 54.974742528  # 8-stack, 256x256, K = 19


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_RANK, help="frames per rank and step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="ncu helper: warm up, then run --steps resident steps inside the NVTX range 'df3d_step' and exit "
                         "(no JSON line; numbers taken under a profiler are never bench values)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_sustained": p["bf16_tflops_sustained"], "tf_burst": p["bf16_tflops"],
                "which": "measured"}
    return {"hbm_gbs": 6650.0, "tf_sustained": 1400.0, "tf_burst": 1590.0, "which": "fallback"}


def synthetic_images(n, h, w, seed, device):
    """Sum of 19 Gaussian blobs (sigma 6 px) + N(0, 0.05) noise, uint8 gray."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, h, w), dtype=torch.uint8, device=device)
    ys = torch.arange(h, dtype=torch.float32, device=device).view(1, 1, h, 1)
    xs = torch.arange(w, dtype=torch.float32, device=device).view(1, 1, 1, w)
    for i in range(0, n, 64):
        m = min(64, n - i)
        cy = torch.rand((m, 19, 1, 1), generator=g, device=device) * h
        cx = torch.rand((m, 19, 1, 1), generator=g, device=device) * w
        img = torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / 72.0).sum(1)
        img = img + 0.05 * torch.randn((m, h, w), generator=g, device=device)
        out[i:i + m] = (img.clamp_(0, 1) * 255).round().to(torch.uint8)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline of the b200 arm, and the whole --impl reference arm)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_frames_per_s(sample_frames, frames_3d, seed=0):
    """Times the oracle on `sample_frames` frames (x7 images) for the hourglass and on `frames_3d`
    frames for the 3-D half; returns frames/s = 1 / (t_2d per frame + t_3d per frame)."""
    from oracle import argmax as oargmax
    from oracle import geometry as g
    from oracle import hourglass as ohg
    from oracle import pack as opack
    from oracle import procrustes as oproc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = ohg.make_model(NUM_STACKS, seed=seed)
    imgs = ohg.to_uint8(ohg.synthetic_images(CAMS * sample_frames, IN_H, IN_W, seed=seed + 1))
    flip = np.zeros((CAMS, sample_frames), dtype=bool)
    flip[4:] = True
    x = ohg.preprocess_u8(imgs, flip=flip.reshape(-1))
    t0 = time.perf_counter()
    with torch.no_grad():
        heat = torch.cat([model(x[i:i + 8])[-1] for i in range(0, x.shape[0], 8)])   # batch 8 like cli.py:141-145
    idx, conf = oargmax.heatmap_argmax(heat.numpy())
    t_2d = time.perf_counter() - t0

    # 3-D half on synthetic geometry of `frames_3d` frames (template skeleton projected with calib)
    G = os.path.join(ROOT, "tests", "golden")
    calib = dict(np.load(os.path.join(G, "calib.npz")))
    tmpl = np.load(os.path.join(G, "template.npz"))["points3d"]
    rng = np.random.default_rng(seed)
    T = frames_3d
    X = tmpl[rng.integers(0, tmpl.shape[0], size=T)] + rng.normal(scale=0.05, size=(T, 38, 3))
    p19 = np.zeros((CAMS, T, 19, 2))
    for c in range(CAMS):
        half = slice(0, 19) if c < 3 else slice(19, 38)
        uv = g.project(X[:, half].reshape(-1, 3), calib["R"][c], calib["tvec"][c], calib["intr"][c]).reshape(T, 19, 2)
        col = np.clip(np.round(uv[..., 0] / 960 * 64), 1, 63) / 64       # 64x64 heat-map grid
        row = np.clip(np.round(uv[..., 1] / 480 * 64), 1, 63) / 64
        p19[c, ..., 0], p19[c, ..., 1] = row, (1 - col if c > 3 else col)
    t0 = time.perf_counter()
    p38 = opack.pack_points2d(p19, range(7))
    out = g.calibrate_and_triangulate(p38, calib, image_shape=(960, 480))
    oproc.procrustes_separate(out["points3d_wo_procrustes"], tmpl)
    t_3d = time.perf_counter() - t0
    fps = 1.0 / (t_2d / sample_frames + t_3d / T)
    return fps, {"t_hourglass_s": t_2d, "hourglass_frames": sample_frames, "t_3d_s": t_3d, "frames_3d": T, "cores": cores,
                 "torch_threads": torch.get_num_threads()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    detail = None
    for i in range(args.warmup + args.steps):
        fps, detail = cpu_oracle_frames_per_s(sample_frames=1, frames_3d=32, seed=i)
        if i >= args.warmup:
            vals.append(fps)
        if i == 0 and detail["t_hourglass_s"] > 40:      # keep the whole run within minutes on slow hosts
            args.warmup, args.steps = 0, 1
            vals = [fps]
            break
    v = float(np.mean(vals))
    sample = "1 frame (7 images 256x256, 8-stack fp32 PyTorch-CPU, batch<=8) + 3-D half (DLT, SciPy TRF BA, procrustes) on 32 frames per step"
    print(json.dumps({
        "impl": "reference", "metric": "7-cam frames/sec -> 3D pose", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1000.0 / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, note="CPU oracle port of df2d+pyba (reference deps not installable offline)"),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": detail["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, note=None):
    cfg = {"workload": f"configs[1] shape ({args.frames} frames x 7 cams, 256x256 u8, 8-stack hourglass, 19 maps) "
                       "+ configs[2] 2D->3D tail (arg-max, pack, DLT, LM bundle adjust, DLT)",
           "frames_per_gpu": args.frames, "images_per_step_per_gpu": args.frames * CAMS,
           "cache": "inputs+activations (>1 GB per chunk) larger than the 126 MB L2; no explicit flush",
           "parallelism": f"frames sharded x{args.gpus}"}
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    from deepfly3d_b200.inference import random_state_dict
    from deepfly3d_b200.pipeline import Pose3DPipeline, gather_frames

    T = args.frames
    n_img = CAMS * T
    pipe = Pose3DPipeline(random_state_dict(NUM_STACKS, seed=0), IN_H, IN_W, n_img, image_shape=[IN_W, IN_H],
                          device=dev, ba_max_iters=10)
    images = synthetic_images(n_img, IN_H, IN_W, seed=1 + rank, device=dev)      # resident in HBM
    host_images = torch.empty((n_img, IN_H, IN_W), dtype=torch.uint8).pin_memory()
    host_images.copy_(images)
    host_x3d = torch.empty((T * world, 38, 3), dtype=torch.float64).pin_memory()
    host_cam = torch.empty((CAMS, 6), dtype=torch.float64).pin_memory()
    # e2e: two device staging buffers and a copy stream -- the host->device copy of step k+1 runs while step k
    # computes (every timed step still issues exactly one copy of a full step's inputs inside the timed region)
    stages = [torch.empty_like(images), torch.empty_like(images)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"k": 0, "primed": False}

    def h2d(slot):
        copy_stream.wait_event(consumed[slot])          # the step that last read this buffer is done with it
        with torch.cuda.stream(copy_stream):
            stages[slot].copy_(host_images, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_resident():
        out = pipe.run(images, T, group=group)
        return gather_frames(out["points3d_wo_procrustes"], group), out

    def step_e2e():
        cur = e2e_state["k"] & 1
        main = torch.cuda.current_stream()
        if not e2e_state["primed"]:                                                  # very first step: its own copy
            consumed[0].record(main)
            consumed[1].record(main)
            h2d(cur)
            e2e_state["primed"] = True
        h2d(cur ^ 1)                                                                 # H2D of the next step's inputs
        main.wait_event(ready[cur])
        out = pipe.run(stages[cur], T, group=group)
        consumed[cur].record(main)
        x3d = gather_frames(out["points3d_wo_procrustes"], group)
        host_x3d.copy_(x3d, non_blocking=True)                                       # D2H of the result
        host_cam.copy_(out["cam_rt"], non_blocking=True)
        e2e_state["k"] += 1
        return x3d, out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None, timing=False):
        barrier()
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        conv = {"conv_ms": 0.0, "conv_flop": 0.0, "conv_launches": 0.0, "other_ms": 0.0, "conv3x3_ms": 0.0, "conv3x3_flop": 0.0}
        ev0.record()
        for _ in range(steps):
            fn()
            if timing:          # reading the per-launch events synchronises; they sit inside the timed region
                t = pipe.engine.read_timing()
                for k in conv:
                    conv[k] += t[k]
        ev1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks, conv

    for _ in range(max(args.warmup, 3)):
        step_resident()
    if args.profile:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("df3d_step")
        for _ in range(args.steps):
            step_resident()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return
    ms_res, clocks, _ = timed(step_resident, args.steps, ClockSampler(local) if rank == 0 else None)
    # separate pass with per-launch events for the roofline of the dominant kernel
    pipe.engine.set_timing(True)
    step_resident()
    _, _, conv = timed(step_resident, args.steps, None, timing=True)
    pipe.engine.set_timing(False)
    for _ in range(2):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)

    frames_total = T * world * args.steps
    value = frames_total / (ms_res / 1e3)
    e2e = frames_total / (ms_e2e / 1e3)
    pk = peaks()
    achieved_tf = conv["conv_flop"] / (conv["conv_ms"] / 1e3) / 1e12 if conv["conv_ms"] > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "conv_gemm_traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    line = {
        "metric": "7-cam frames/sec -> 3D pose", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_res / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(host_images.numel()) * world,
                "d2h_bytes_per_step": int(host_x3d.numel() * 8 + host_cam.numel() * 8)},
        "gpu_launches": int(pipe.launches(n_img)) * args.steps,
        "roofline": {"bound": "tensor", "kernel": "conv_chain_kernel + conv_gemm_kernel<BN> (tcgen05 implicit-GEMM convs and conv chains)",
                     "achieved": achieved_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": achieved_tf / pk["tf_sustained"], "peak_source": f"bf16_tflops_sustained, of {pk['which']}",
                     "launches_per_step": conv["conv_launches"] / args.steps,
                     "flop_per_launch": conv["conv_flop"] / max(conv["conv_launches"], 1),
                     "ms_per_launch": conv["conv_ms"] / max(conv["conv_launches"], 1),
                     "share_of_step": conv["conv_ms"] / max(conv["conv_ms"] + conv["other_ms"], 1e-9),
                     "conv3x3_tflops": conv["conv3x3_flop"] / max(conv["conv3x3_ms"], 1e-9) / 1e9,
                     "end_to_end_frac": value / world * CAMS * GFLOP_PER_IMAGE / 1e3 / pk["tf_sustained"],
                     "traffic": traffic},
        "clocks": clocks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                fps, d = cpu_oracle_frames_per_s(sample_frames=1, frames_3d=32)
                line["cpu_baseline"] = {
                    "value": fps, "unit": "frames/s", "cores": d["cores"], "kind": "port",
                    "sample": f"1 frame (7 images, 8-stack 256x256 fp32 PyTorch-CPU): {d['t_hourglass_s']:.1f} s; "
                              f"3-D half on 32 frames: {d['t_3d_s']:.2f} s"}
            except Exception as e:  # the oracle is test infrastructure; never let it break the GPU line
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
