#!/usr/bin/env python
"""Headline benchmark: 7-camera frames/s -> 3-D pose (BASELINE.json metric).

Workloads (BASELINE.json `configs`); a step = one pass of the hot path over one batch of frames:

  --config 2 (default)  configs[2], the end-to-end configuration: 1 000 frames x 7 cameras of 256x256 uint8 per GPU
                        -> 8-stack hourglass (bf16 tensor-core convs, 19 maps / image) -> arg-max -> 19->38 packing
                        -> DLT -> bundle adjustment over ALL frames -> DLT -> procrustes, i.e. everything
                        Core.pose2d_estimation + calibrate_calc + save run in the reference, minus file I/O.
                        N > 1: weak scaling, 1 000 frames per rank.
  --config 1            configs[1]: 256 frames x 7 cameras, hourglass + arg-max only.
  --config 4            configs[3]: 100 000 frames in total, contiguous blocks of 100 000 / N frames per rank (strong
                        scaling), bundle adjustment on a strided subset of <= 1 000 frames, one all-gather of points3d.

  value : frames/s with the images already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public pipeline call with PINNED HOST images: every timed step issues one H2D copy
          of a full step's images (double-buffered on a copy stream: the copy of step k+1 overlaps the compute of
          step k) and the D2H of the registered 3-D joints, the raw 3-D joints and the cameras
  e2e_files : (N = 1) the drop-in surface itself: Core(folder of 480x960 JPEG files) -> pose2d_estimation ->
          calibrate_calc -> save, frames/s from files on disk to the result pickle
  roofline : dominant kernels = conv_chain_kernel + conv_gemm_kernel (tcgen05 conv chains and implicit-GEMM convs),
          achieved = algorithmic conv FLOPs of the launches that ran / summed device time of those launches (CUDA
          events on the launching stream, inside a timed pass), peak = MEASURED_PEAKS.json bf16_tflops_sustained
          (fallback 1400 TF/s "of fallback"); traffic = DRAM bytes per conv launch from the committed ncu metrics
          pass (profiles/conv_gemm_traffic.json)
  cpu_baseline : the CPU oracle (PyTorch fp32 hourglass at batch 8 + numpy DLT + SciPy BA + procrustes, all host
          threads) on a bounded sample, per-stage seconds in `sample`

N > 1 (torchrun): frames shard across ranks; the packed 2-D points are all-gathered, the bundle adjustment runs
replicated (bit-reproducible kernels: identical cameras on every rank, no collective inside the solver), one
all-gather of the 3-D joints at the end; procrustes (global medians) on the gathered array.

`--impl reference` times the CPU oracle alone (the reference's df2d/pyba are not installable here: un-vendored
dependencies, no network) and prints the same JSON line with impl=reference.
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

IN_H = IN_W = 256
NUM_STACKS = 8
CAMS = 7
GFLOP_PER_IMAGE = 54.974742528  # 8-stack, 256x256, K = 19
CONFIG_FRAMES = {1: 256, 2: 1000, 4: 100000}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 4], help="BASELINE.json configs[] workload (see the docstring)")
    ap.add_argument("--frames", type=int, default=None, help="override the frames per rank and step (config 4: in total)")
    ap.add_argument("--ref-frames", type=int, default=8, help="frames of the CPU hourglass sample (x7 images, batch 8)")
    ap.add_argument("--ba-solver", default="exact", choices=["exact", "lsmr"],
                    help="regularised Gauss-Newton step of the bundle adjustment: exact Schur solve (default here: the one that "
                         "scales to the 8-GPU weak-scaling problem) or SciPy's truncated LSMR (the default of the drop-in Core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the Core-from-JPEG-folder end-to-end leg")
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


def config3_joint_pixels(T, seed, device, h=IN_H, w=IN_W):
    """jittered template skeleton projected by the perturbed packaged cameras -> joint
    positions in an h x w image frame, (7, T, 19, 2) (x, y).  Product-side data only (no oracle import)."""
    from deepfly3d_b200.camera_network import rodrigues_vec
    from deepfly3d_b200.pipeline import load_default_calib
    from deepfly3d_b200.procrustes import read_template_pose3d

    rng = np.random.default_rng(seed)
    calib, tmpl = load_default_calib(), read_template_pose3d()

    def rot(r):
        th = np.linalg.norm(r)
        k = r / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) * np.cos(th) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * K

    X = tmpl[rng.integers(0, tmpl.shape[0], size=T)] + rng.normal(scale=0.05, size=(T, 38, 3))
    out = np.zeros((CAMS, T, 19, 2), dtype=np.float32)
    for c in range(CAMS):
        R = rot(rodrigues_vec(calib["R"][c]) + rng.normal(scale=0.01, size=3))
        t = calib["tvec"][c] + rng.normal(scale=0.5, size=3)
        half = X[:, :19] if c < 4 else X[:, 19:]
        Xc = half.reshape(-1, 3) @ R.T + t
        K = calib["intr"][c]
        u = (K[0, 0] * Xc[:, 0] / Xc[:, 2] + K[0, 2]) / 960.0 * w
        v = (K[1, 1] * Xc[:, 1] / Xc[:, 2] + K[1, 2]) / 480.0 * h
        out[c, ..., 0], out[c, ..., 1] = u.reshape(T, 19), v.reshape(T, 19)
    return torch.as_tensor(out, device=device)


def synthetic_images(T, h, w, seed, device):
    """(7*T, h, w) uint8, camera-major: 19 Gaussian blobs (sigma 6 px) at the projected joints of the config-3
    skeleton + N(0, 0.05) noise."""
    centres = config3_joint_pixels(T, seed, device, h, w).reshape(CAMS * T, 19, 2)
    n = CAMS * T
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, h, w), dtype=torch.uint8, device=device)
    ys = torch.arange(h, dtype=torch.float32, device=device).view(1, 1, h, 1)
    xs = torch.arange(w, dtype=torch.float32, device=device).view(1, 1, 1, w)
    for i in range(0, n, 64):
        m = min(64, n - i)
        cx = centres[i:i + m, :, 0].reshape(m, 19, 1, 1)
        cy = centres[i:i + m, :, 1].reshape(m, 19, 1, 1)
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
def cpu_oracle_stages(hourglass_frames, frames_3d, seed=0):
    """Times the oracle stage by stage: the hourglass + arg-max on `hourglass_frames` frames (x7 images, batch 8 like
    df3d/cli.py:141-145), and packing + SciPy-TRF bundle adjustment + DLT + procrustes on `frames_3d` frames of the
    config-3 geometry.  Returns frames/s = 1 / (t_2d per frame + t_3d per frame) and the per-stage seconds."""
    from oracle import argmax as oargmax
    from oracle import geometry as g
    from oracle import hourglass as ohg
    from oracle import pack as opack
    from oracle import procrustes as oproc
    from oracle import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = ohg.make_model(NUM_STACKS, seed=seed)
    imgs = ohg.to_uint8(ohg.synthetic_images(CAMS * hourglass_frames, IN_H, IN_W, seed=seed + 1))
    flip = np.zeros((CAMS, hourglass_frames), dtype=bool)
    flip[4:] = True
    x = ohg.preprocess_u8(imgs, flip=flip.reshape(-1))
    t0 = time.perf_counter()
    with torch.no_grad():
        heat = torch.cat([model(x[i:i + 8])[-1] for i in range(0, x.shape[0], 8)])
    t_hg = time.perf_counter() - t0
    t0 = time.perf_counter()
    oargmax.heatmap_argmax(heat.numpy())
    t_am = time.perf_counter() - t0

    T = frames_3d
    calib, p38, _, _ = synth.config3_points2d(T, seed=2 + seed)
    tmpl = synth.load_template()
    p19 = np.zeros((CAMS, T, 19, 2))
    t0 = time.perf_counter()
    opack.pack_points2d(p19, range(7))
    t_pack = time.perf_counter() - t0
    pts_xy = g.to_pixels_xy(p38, [960, 480])
    t0 = time.perf_counter()
    R, t = g.bundle_adjust(calib["R"], calib["tvec"], calib["intr"], pts_xy)          # includes its initial DLT
    t_ba = time.perf_counter() - t0
    t0 = time.perf_counter()
    X = g.triangulate_dlt(g.projection_matrices(R, t, calib["intr"]), pts_xy)
    t_dlt = time.perf_counter() - t0
    t0 = time.perf_counter()
    oproc.procrustes_separate(X, tmpl)
    t_proc = time.perf_counter() - t0
    t_2d, t_3d = t_hg + t_am, t_pack + t_ba + t_dlt + t_proc
    fps = 1.0 / (t_2d / hourglass_frames + t_3d / T)
    return fps, {"hourglass_s": t_hg, "argmax_s": t_am, "hourglass_frames": hourglass_frames, "pack_s": t_pack, "ba_s": t_ba,
                 "dlt_s": t_dlt, "procrustes_s": t_proc, "frames_3d": T, "cores": cores, "torch_threads": torch.get_num_threads()}


def sample_text(d):
    return (f"{d['hourglass_frames']} frames ({7 * d['hourglass_frames']} images 256x256, 8-stack fp32 PyTorch-CPU, batch 8): hourglass "
            f"{d['hourglass_s']:.1f} s + arg-max {d['argmax_s']:.3f} s; 3-D half on {d['frames_3d']} frames of the config-3 geometry: pack "
            f"{d['pack_s']:.3f} s, SciPy TRF bundle adjustment (incl. its DLT) {d['ba_s']:.2f} s, DLT {d['dlt_s']:.2f} s, procrustes "
            f"{d['procrustes_s']:.3f} s; frames/s = 1 / (t_2d per frame + t_3d per frame); {d['torch_threads']} torch threads")


def frames_for(args, world):
    total = args.frames if args.frames is not None else CONFIG_FRAMES[args.config]
    if args.config == 4:
        return (total + world - 1) // world, "strong"
    return total, "weak"


def workload_config(args, world, T, note=None):
    if args.config == 1:
        wl = f"configs[1]: {T} frames x 7 cams, 256x256 u8, 8-stack hourglass, 19 maps -> arg-max only"
    elif args.config == 2:
        wl = (f"configs[2]: {T} frames x 7 cams per GPU, 256x256 u8, 8-stack hourglass (19 maps) -> arg-max -> pack -> DLT -> "
              f"bundle adjustment over all {T * world} frames -> DLT -> procrustes")
    else:
        wl = (f"configs[3]: {T * world} frames x 7 cams in contiguous blocks of {T} per GPU, 256x256 u8, 8-stack hourglass -> arg-max -> "
              "pack -> bundle adjustment on a strided subset of <= 1000 frames -> DLT -> one all-gather of points3d -> procrustes")
    cfg = {"workload": wl, "frames_per_gpu": T, "images_per_step_per_gpu": T * CAMS,
           "weights": "random-init (seeded; no pretrained weights offline)",
           "cache": "inputs+activations (>1 GB per 1792-image chunk) larger than the 126 MB L2; no explicit flush",
           "parallelism": f"frames sharded x{world}"}
    if note:
        cfg["note"] = note
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    T, scaling = frames_for(args, world)
    frames_3d = min(1000, T * world) if args.config != 1 else 64
    vals, detail = [], None
    t_start = time.perf_counter()
    n_total = args.warmup + args.steps
    warm = args.warmup
    for i in range(n_total):
        fps, detail = cpu_oracle_stages(args.ref_frames, frames_3d, seed=i)
        if i >= warm:
            vals.append(fps)
        elapsed = time.perf_counter() - t_start
        if elapsed / (i + 1) * (i + 2) > 240 and i + 1 < n_total:       # keep the whole run within a few minutes
            if not vals:
                vals, warm = [fps], i
            break
    v = float(np.mean(vals))
    print(json.dumps({
        "impl": "reference", "metric": "7-cam frames/sec -> 3D pose", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": len(vals), "warmup": warm, "ms_per_step": 1000.0 * T * world / v,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world, T, note="CPU oracle port of df2d+pyba (the reference's dependencies are not installable "
                                                        "offline); every step times a bounded sample, ms_per_step = the workload's frames / value"),
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": detail["cores"], "kind": "port", "sample": sample_text(detail),
                         "stages": detail},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
def core_from_files(dev, frames=256, ba_solver="exact"):
    """The drop-in surface end to end: writes `frames` x 7 synthetic 480x960 JPEG files, then times
    Core(folder) -> pose2d_estimation -> calibrate_calc -> save (threaded libjpeg decode, copy stream, device resize to
    256x512, 8-stack hourglass, BA, DLT, procrustes, pickle)."""
    import shutil
    import tempfile
    from concurrent.futures import ThreadPoolExecutor

    import cv2

    from deepfly3d_b200 import inference
    from deepfly3d_b200.core import Core

    tmp = tempfile.mkdtemp(prefix="df3d_bench_")
    folder = os.path.join(tmp, "sample", "test")
    os.makedirs(folder)
    try:
        small = synthetic_images(frames, 240, 480, seed=7, device=dev).cpu().numpy()       # (7*frames, 240, 480)

        def write(i):
            c, t = divmod(i, frames)
            cv2.imwrite(os.path.join(folder, f"camera_{c}_img_{t}.jpg"), cv2.resize(small[i], (960, 480)), [cv2.IMWRITE_JPEG_QUALITY, 90])

        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
            list(pool.map(write, range(CAMS * frames)))
        sd = inference.random_state_dict(NUM_STACKS, seed=0)
        out = {}
        variants = [("host_libjpeg", False, None), ("nvjpeg", True, None), ("nvjpeg_gpu_hybrid", "gpu_hybrid", None)]
        if os.environ.get("DF3D_BENCH_BLOCKS"):                          # tuning: extra block sizes, host decode
            variants += [(f"host_libjpeg_block{b}", False, int(b)) for b in os.environ["DF3D_BENCH_BLOCKS"].split(",")]
            variants += [(f"nvjpeg_block{b}", True, int(b)) for b in os.environ["DF3D_BENCH_BLOCKS"].split(",")]
        for label, gpu_decode, block in variants:
            try:
                times = []
                for rep in range(2):                                    # first pass builds the engine (untimed)
                    for f in os.listdir(tmp):
                        if f.endswith("_df3d"):
                            shutil.rmtree(os.path.join(tmp, f))
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    core = Core(folder, num_images_max=0, camera_ordering=[0, 1, 2, 3, 4, 5, 6], state_dict=sd, gpu_decode=gpu_decode, block_frames=block,
                                ba_solver=ba_solver)
                    t1 = time.perf_counter()
                    core.pose2d_estimation()
                    t2 = time.perf_counter()
                    core.calibrate_calc(0, frames)
                    core.save()
                    torch.cuda.synchronize()
                    t3 = time.perf_counter()
                    times.append((t3 - t0, t1 - t0, t2 - t1, t3 - t2, dict(core.ingest_stats)))
                tot, t_open, t_2d, t_3d, st = times[-1]
                out[label] = {"frames_per_s": frames / tot, "pose2d_frames_per_s": frames / t_2d, "open_s": t_open, "pose2d_s": t_2d,
                              "calibrate_save_s": t_3d, "decode_wait_s": st.get("decode_wait_s"), "workers": st.get("workers"),
                              "blocks": st.get("blocks"), "block_frames": st.get("block_frames"), "decode": st.get("decode")}
            except Exception as e:  # nvJPEG may be absent on a box: report, do not fail the bench
                out[label] = {"error": str(e)[:200]}
        inference.drop_engine()
        return {"value": out.get("host_libjpeg", {}).get("frames_per_s"), "unit": "frames/s", "frames": frames,
                "input": "7 x frames JPEG files 480x960 (quality 90) -> 256x512 network input, 8-stack hourglass, 64x128 heat-maps",
                "path": "Core(folder).pose2d_estimation() + calibrate_calc() + save()", "ba_solver": ba_solver, "variants": out}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


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

    from deepfly3d_b200 import ops
    from deepfly3d_b200.inference import random_state_dict
    from deepfly3d_b200.pipeline import Pose3DPipeline, gather_frames

    T, scaling = frames_for(args, world)
    n_img = CAMS * T
    # configs[3]: BA on a strided subset of <= 1 000 frames; configs[2]: over all frames (DF3D_BENCH_BA_MAX_FRAMES only
    # serves the multi-GPU attribution runs of tools/gpu_n8.sh and is named in the line when set)
    ba_cap = 1000 if args.config == 4 else None
    if os.environ.get("DF3D_BENCH_BA_MAX_FRAMES"):
        ba_cap = int(os.environ["DF3D_BENCH_BA_MAX_FRAMES"])
    pipe = Pose3DPipeline(random_state_dict(NUM_STACKS, seed=0), IN_H, IN_W, n_img, image_shape=[IN_W, IN_H],
                          device=dev, ba_max_iters=10, ba_max_frames=ba_cap, ba_solver=args.ba_solver)
    images = synthetic_images(T, IN_H, IN_W, seed=1 + rank, device=dev)      # resident in HBM
    host_images = torch.empty((n_img, IN_H, IN_W), dtype=torch.uint8).pin_memory()
    host_images.copy_(images)
    full3d = args.config != 1
    host_x3d = torch.empty((T * world, 38, 3), dtype=torch.float64).pin_memory()
    host_x3d_raw = torch.empty((T * world, 38, 3), dtype=torch.float64).pin_memory()
    host_cam = torch.empty((CAMS, 6), dtype=torch.float64).pin_memory()
    host_idx = torch.empty((n_img, 19), dtype=torch.int32).pin_memory()
    host_conf = torch.empty((n_img, 19), dtype=torch.float32).pin_memory()
    proc_ws = ops.procrustes_workspace(T * world, dev) if full3d else None
    # e2e: two device staging buffers and a copy stream -- the host->device copy of step k+1 runs while step k
    # computes (every timed step still issues exactly one copy of a full step's inputs inside the timed region)
    # (a step of more than 16 GB of images -- config 4 on few GPUs -- is staged single-buffered)
    n_stage = 2 if images.numel() <= (16 << 30) else 1
    stages = [torch.empty_like(images) for _ in range(n_stage)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"k": 0, "primed": False}

    def h2d(slot):
        copy_stream.wait_event(consumed[slot])          # the step that last read this buffer is done with it
        with torch.cuda.stream(copy_stream):
            stages[slot].copy_(host_images, non_blocking=True)
            ready[slot].record(copy_stream)

    def compute(imgs):
        if not full3d:                                  # configs[1]: hourglass + arg-max
            idx, conf = pipe.engine.forward(imgs, flip=pipe.flip_flags(T))
            return {"idx": idx, "conf": conf}
        out = pipe.run(imgs, T, group=group)
        out["x3d_raw"] = gather_frames(out["points3d_wo_procrustes"], group)
        out["x3d"] = ops.procrustes(out["x3d_raw"], workspace=proc_ws)       # global medians: on the gathered array
        return out

    def step_resident():
        return compute(images)

    def step_e2e():
        cur = e2e_state["k"] % n_stage
        main = torch.cuda.current_stream()
        if not e2e_state["primed"]:                                                  # very first step: its own copy
            consumed[0].record(main)
            consumed[1].record(main)
            h2d(cur)
            e2e_state["primed"] = True
        if n_stage == 2:
            h2d(cur ^ 1)                                                             # H2D of the next step's inputs
        main.wait_event(ready[cur])
        out = compute(stages[cur])
        consumed[cur].record(main)
        if n_stage == 1:
            h2d(0)                                                                   # next step's copy, behind this step's compute
        if full3d:                                                                   # D2H of the results
            host_x3d.copy_(out["x3d"], non_blocking=True)
            host_x3d_raw.copy_(out["x3d_raw"], non_blocking=True)
            host_cam.copy_(out["cam_rt"], non_blocking=True)
        else:
            host_idx.copy_(out["idx"], non_blocking=True)
            host_conf.copy_(out["conf"], non_blocking=True)
        e2e_state["k"] += 1
        return out

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

    warm = max(args.warmup, 3)
    for _ in range(warm):
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
    t_steps = max(1, min(args.steps, 2 if T > 2000 else args.steps))
    pipe.engine.set_timing(True)
    step_resident()
    _, _, conv = timed(step_resident, t_steps, None, timing=True)
    pipe.engine.set_timing(False)
    for _ in range(2):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    rep = ops.ba_report(step_resident()["ba_report"]) if full3d else None
    ba_ms = None
    if full3d and rank == 0 and world == 1:      # device time of the 2D->3D tail on this step's data, both solvers
        ba_ms = {}
        out = step_resident()
        for solver in ("exact", "lsmr"):
            cam = pipe.cam_rt0.clone()
            P0, _ = ops.projection_matrices(cam, pipe.intr4)
            X0 = ops.triangulate_dlt(P0, out["pts_xy"])
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            ops.bundle_adjust(cam, pipe.intr4, out["pts_xy"], X0, max_iters=pipe.ba_max_iters, ftol=pipe.ba_ftol, solver=solver)
            ev[1].record()
            torch.cuda.synchronize()
            ba_ms[solver] = ev[0].elapsed_time(ev[1])

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
    d2h = int(host_x3d.numel() * 8 * 2 + host_cam.numel() * 8) if full3d else int(host_idx.numel() * 4 + host_conf.numel() * 4)
    ba_frames = next(iter(pipe._ba_ws))[1] if full3d and pipe._ba_ws else min(T * world, pipe.ba_max_frames or T * world)
    sharded_ba = bool(full3d and world > 1 and args.ba_solver == "exact" and ops.ba_sharded_plan(CAMS, ba_frames, 38, world) is not None)
    launches = int(pipe.launches(n_img, sharded_ba=sharded_ba) + 4) if full3d else int(pipe.engine.launches(n_img))
    line = {
        "metric": "7-cam frames/sec -> 3D pose", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_res / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world, T),
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(host_images.numel()) * world,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches * args.steps,
        "roofline": {"bound": "tensor", "kernel": "conv_chain_kernel + conv_gemm_kernel<BN> (tcgen05 implicit-GEMM convs and conv chains)",
                     "achieved": achieved_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": achieved_tf / pk["tf_sustained"], "peak_source": f"bf16_tflops_sustained, of {pk['which']}",
                     "launches_per_step": conv["conv_launches"] / t_steps,
                     "flop_per_launch": conv["conv_flop"] / max(conv["conv_launches"], 1),
                     "ms_per_launch": conv["conv_ms"] / max(conv["conv_launches"], 1),
                     "share_of_step": conv["conv_ms"] / max(conv["conv_ms"] + conv["other_ms"], 1e-9),
                     "conv3x3_tflops": conv["conv3x3_flop"] / max(conv["conv3x3_ms"], 1e-9) / 1e9,
                     "end_to_end_frac": value / world * CAMS * GFLOP_PER_IMAGE / 1e3 / pk["tf_sustained"],
                     "traffic": traffic},
        "clocks": clocks,
    }
    if rep is not None:
        line["bundle_adjust"] = {"frames": ba_frames, "sharded": sharded_ba, "observations": rep["n_obs"],
                                 "evaluations": rep["iters"], "accepted": rep["accepted"], "status": rep["status"],
                                 "solver": args.ba_solver, "device_ms_by_solver": ba_ms}
    if rank == 0:
        if world == 1 and not args.no_files:
            try:
                del stages
                torch.cuda.empty_cache()
                import contextlib

                with contextlib.redirect_stdout(sys.stderr):         # Core prints like the reference; stdout carries ONE line
                    line["e2e_files"] = core_from_files(dev, ba_solver=args.ba_solver)
            except Exception as e:
                line["e2e_files"] = {"value": None, "error": str(e)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            try:
                fps, d = cpu_oracle_stages(args.ref_frames, min(1000, T) if full3d else 64)
                line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": d["cores"], "kind": "port", "sample": sample_text(d),
                                        "stages": d}
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
