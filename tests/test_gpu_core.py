"""The df3d.core.Core surface end to end on the GPU (mirrors the reference's test_pose_estimation
and test_calibration, tests/test_df3d.py:150-244, on the committed fixtures)."""
import os
import pickle
import shutil

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import argmax as oargmax
from oracle import hourglass as ohg
from oracle import pack as opack

HERE = os.path.dirname(os.path.abspath(__file__))
IMAGES = os.path.join(HERE, "golden", "images")


@pytest.fixture()
def working(tmp_path, lib_built):
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test needs a CUDA device")
    d = tmp_path / "sample" / "test"          # 'sample/test' also exercises the default-ordering regex
    d.mkdir(parents=True)
    for f in os.listdir(IMAGES):
        shutil.copy(os.path.join(IMAGES, f), d / f)
    return str(d)


def test_core_loads_folder(working):
    from deepfly3d_b200.core import Core

    core = Core(input_folder=working, output_folder=working + "_out", num_images_max=0, camera_ordering=None)
    assert core.num_images == 3 and core.max_img_id == 2
    assert core.image_shape == [960, 480]
    assert np.all(core.camera_ordering == np.arange(7))
    assert core.save_path.endswith("df3d_result_{}.pkl".format(working.replace("/", "_")))
    with pytest.raises(FileNotFoundError):
        os.makedirs(working + "_empty")
        Core(input_folder=working + "_empty")


def test_pose_estimation_matches_oracle(working):
    """Images -> hourglass -> arg-max -> packing through Core, against the CPU oracle fed the same
    JPEGs and the same (seeded) weights; tolerance of the reference test: atol 0.02 on points2d."""
    import cv2

    from deepfly3d_b200.core import Core

    model = ohg.make_model(2, seed=0)
    core = Core(input_folder=working, num_images_max=0, camera_ordering=[0, 1, 2, 3, 4, 5, 6], state_dict=model.state_dict())
    core.pose2d_estimation()
    assert core.points2d.shape == (7, 3, 38, 2) and core.conf.shape == (7, 3, 19, 1)

    imgs = np.stack([[cv2.resize(cv2.imread(os.path.join(working, f"camera_{c}_img_{t}.jpg"), cv2.IMREAD_GRAYSCALE),
                                 (512, 256), interpolation=cv2.INTER_LINEAR) for t in range(3)] for c in range(7)])
    flip = np.zeros((7, 3), dtype=bool)
    flip[4:] = True
    with torch.no_grad():
        x = ohg.preprocess_u8(torch.as_tensor(imgs.reshape(21, 256, 512)), flip=flip.reshape(-1))
        heat = model(x, emulate_bf16=True)[-1]
    idx, conf = oargmax.heatmap_argmax(heat.numpy())
    ref = opack.pack_points2d(opack.indices_to_points2d(idx.reshape(7, 3, 19), (64, 128)), range(7))
    # the packing structure is exact: camera 3 dropped, camera 2 joints 15.., camera 4 joints 34.. blanked,
    # unseen halves zero / (0, 1) after the un-flip (core.py:187-199)
    blank = np.zeros((7, 3, 38), dtype=bool)
    blank[3] = True
    blank[:3, :, 19:] = True
    blank[4:, :, :19] = True
    blank[2, :, 15:] = True
    blank[4, :, 34:] = True
    assert np.array_equal(core.points2d[blank], ref[blank])
    assert np.all(core.points2d[:4][blank[:4]] == 0) and np.all(core.points2d[4:][blank[4:]] == [0.0, 1.0])
    # coordinates: the seeded random network has nearly flat score maps, so a bf16-level difference in
    # the accumulation order moves some arg-maxes (tests/test_gpu_hourglass.py measures exactly that);
    # the reference's own tolerance (atol 0.02, test_df3d.py:171) must hold for the bulk of the joints
    close = np.abs(core.points2d - ref).max(axis=-1) <= 0.02
    assert close.mean() > 0.8, f"only {close.mean():.3f} of the joints within atol 0.02 of the oracle"
    rngv = float(heat.max() - heat.min())
    assert np.abs(core.conf.reshape(21, 19) - conf).max() < 0.02 * rngv
    core.save()
    with open(core.save_path, "rb") as f:
        saved = pickle.load(f)
    assert set(saved) == {"points2d", "camera_ordering", "heatmap_confidence"}      # 2-D only: no calibration yet


def test_calibration_matches_golden(working, golden):
    """Reference test_calibration: golden 2-D in, BA + DLT + procrustes, compare with golden 3-D.
    North-star tolerance 1e-3 mm on the joints (the reference's SciPy path gets 1e-5)."""
    from deepfly3d_b200.core import Core

    core = Core(input_folder=working, num_images_max=0, camera_ordering=[0, 1, 2, 3, 4, 5, 6])
    core.points2d = golden["result_2d"]["points2d"]
    core.conf = golden["result_2d"]["heatmap_confidence"]
    core.calibrate_calc(0, 100)
    core.save()
    with open(core.save_path, "rb") as f:
        saved = pickle.load(f)
    r3 = golden["result_3d"]
    assert set(saved) == {0, 1, 2, 3, 4, 5, 6, "points3d", "points2d", "points3d_wo_procrustes", "camera_ordering",
                          "heatmap_confidence"}
    np.testing.assert_allclose(saved["points3d_wo_procrustes"], r3["points3d_wo_procrustes"], atol=1e-3)
    np.testing.assert_allclose(saved["points3d"], r3["points3d"], atol=1e-3)
    for cam in range(7):
        np.testing.assert_allclose(saved[cam]["R"], r3["R"][cam], atol=1e-3)
        np.testing.assert_allclose(saved[cam]["tvec"], r3["tvec"][cam], atol=5e-3)
        np.testing.assert_array_equal(saved[cam]["intr"], r3["intr"][cam])
        np.testing.assert_array_equal(saved[cam]["distort"], r3["distort"][cam])
    np.testing.assert_array_equal(saved[3]["R"], golden["calib"]["R"][3])              # untouched camera
    # resume from the pickle (core.py:108-126)
    again = Core(input_folder=working, num_images_max=0, camera_ordering=[0, 1, 2, 3, 4, 5, 6])
    assert again.camNet is not None and again.has_calibration
    np.testing.assert_allclose(again.camNet.points3d, saved["points3d_wo_procrustes"], atol=1e-9)


def test_pipeline_matches_stagewise(golden):
    """Pose3DPipeline (what bench.py times) == the stages called one by one."""
    from deepfly3d_b200 import ops
    from deepfly3d_b200.inference import random_state_dict
    from deepfly3d_b200.pipeline import Pose3DPipeline

    T = 4
    pipe = Pose3DPipeline(random_state_dict(2, seed=0), 128, 128, 7 * T, image_shape=[960, 480])
    img = ohg.to_uint8(ohg.synthetic_images(7 * T, 128, 128, seed=5)).cuda()
    out = pipe.run(img, T)
    idx, conf = pipe.engine.forward(img, flip=pipe.flip_flags(T))
    p2d, pxy = ops.pack_points2d(idx, 7, T, (32, 32), range(7), [960, 480])
    torch.cuda.synchronize()
    assert torch.equal(out["idx"], idx) and torch.equal(out["points2d"], p2d) and torch.equal(out["pts_xy"], pxy)
    assert out["points3d_wo_procrustes"].shape == (T, 38, 3) and torch.isfinite(out["points3d_wo_procrustes"]).all()
    assert pipe.launches(7 * T) > 100


def test_inference_folder_streams_blocks_and_loads_checkpoint(working, tmp_path):
    """The streaming loader (blocks of frames, threaded decode, copy stream) gives the same result whatever the
    block size; `weights=` takes a file with the layout of df2d's sh8_deepfly.tar ({'state_dict': {'module.*'}}) and
    `mean=` a mean.pth.tar (reference df3d/config.py:30-39); a second call with OTHER weights must not be served
    by the cached engine of the first."""
    from deepfly3d_b200 import inference

    model_a, model_b = ohg.make_model(2, seed=0), ohg.make_model(2, seed=9)
    wa, wb = tmp_path / "a.tar", tmp_path / "b.tar"
    torch.save({"state_dict": {"module." + k: v for k, v in model_a.state_dict().items()}, "epoch": 1}, wa)
    torch.save({"state_dict": {"module." + k: v for k, v in model_b.state_dict().items()}, "epoch": 1}, wb)
    torch.save({"mean": torch.tensor([0.22, 0.22, 0.22])}, tmp_path / "mean.pth.tar")
    kw = dict(folder=working, camera_ids_to_flip=[4, 5, 6], max_img_id=2, mean=str(tmp_path / "mean.pth.tar"))
    stats = {}
    p_a, c_a = inference.inference_folder(weights=str(wa), stats=stats, **kw)
    assert stats["blocks"] == 1 and p_a.shape == (7, 3, 19, 2) and c_a.shape == (7, 3, 19, 1)
    p_a2, c_a2 = inference.inference_folder(weights=str(wa), block_frames=2, stats=stats, **kw)     # blocks (0,2), (2,3)
    assert stats["blocks"] == 2
    assert np.array_equal(p_a, p_a2) and np.array_equal(c_a, c_a2)
    p_sd, c_sd = inference.inference_folder(state_dict=model_a.state_dict(), **kw)                  # same weights as a dict
    assert np.array_equal(p_a, p_sd) and np.array_equal(c_a, c_sd)
    p_b, c_b = inference.inference_folder(weights=str(wb), **kw)
    assert not np.array_equal(c_a, c_b), "second checkpoint was served by the first one's engine"
    p_m, c_m = inference.inference_folder(weights=str(wa), **{**kw, "mean": 0.5})
    assert not np.array_equal(c_a, c_m), "the mean file was ignored"
    # against the oracle with that mean
    import cv2

    imgs = np.stack([[cv2.resize(cv2.imread(os.path.join(working, f"camera_{c}_img_{t}.jpg"), cv2.IMREAD_GRAYSCALE),
                                 (512, 256), interpolation=cv2.INTER_LINEAR) for t in range(3)] for c in range(7)])
    flip = np.zeros((7, 3), dtype=bool)
    flip[4:] = True
    with torch.no_grad():
        x = ohg.preprocess_u8(torch.as_tensor(imgs.reshape(21, 256, 512)), flip=flip.reshape(-1), mean=0.22)
        heat = model_a(x, emulate_bf16=True)[-1]
    _, conf = oargmax.heatmap_argmax(heat.numpy())
    rngv = float(heat.max() - heat.min())
    assert np.abs(c_a.reshape(21, 19) - conf).max() < 0.01 * rngv
    inference.drop_engine()


def test_inference_folder_device_decode_streams_blocks(working):
    """gpu_decode: files read by the pool, blocks decoded with nvJPEG on a side stream (hardware engines when the box has
    them) overlapping the previous block's hourglass.  Same result whatever the block size; against the host-decode
    path the decoded frames differ by a grey level or two (tests/test_gpu_ingest.py), so the heat-map scores agree
    closely and most arg-maxes exactly."""
    from deepfly3d_b200 import inference

    sd = ohg.make_model(2, seed=0).state_dict()
    kw = dict(folder=working, camera_ids_to_flip=[4, 5, 6], max_img_id=2, state_dict=sd)
    p_h, c_h = inference.inference_folder(**kw)
    stats = {}
    p_1, c_1 = inference.inference_folder(gpu_decode=True, stats=stats, **kw)
    assert stats["blocks"] == 1 and stats["decode"] in ("nvjpeg-hardware", "nvjpeg-default")
    print(f"  device decode backend: {stats['decode']}")
    p_3, c_3 = inference.inference_folder(gpu_decode=True, block_frames=1, stats=stats, **kw)        # three blocks
    assert stats["blocks"] == 3
    assert np.array_equal(p_1, p_3) and np.array_equal(c_1, c_3)
    p_d, c_d = inference.inference_folder(gpu_decode="default", block_frames=2, **kw)
    if stats["decode"] == "nvjpeg-default":
        assert np.array_equal(p_1, p_d) and np.array_equal(c_1, c_d)
    for p, c in ((p_1, c_1), (p_d, c_d)):
        assert np.abs(c - c_h).max() < 0.02 * max(float(np.abs(c_h).max()), 1e-6) + 1e-3
        assert (np.abs(p - p_h).max(axis=-1) <= 0.02).mean() > 0.8
    inference.drop_engine()


def _sharded_worker(rank, world, port, T, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from deepfly3d_b200.inference import random_state_dict
    from deepfly3d_b200.pipeline import Pose3DPipeline, gather_frames, shard_frames

    full = ohg.to_uint8(ohg.synthetic_images(7 * T, 128, 128, seed=5)).view(7, T, 128, 128)
    lo, hi = shard_frames(T, rank, world)
    res = {"range": (lo, hi)}
    for solver in ("lsmr", "exact"):                 # replicated LSMR solve / exact solve with the per-point work sharded
        pipe = Pose3DPipeline(random_state_dict(2, seed=0), 128, 128, 7 * (hi - lo), image_shape=[960, 480], device=f"cuda:{rank}",
                              ba_solver=solver)
        out = pipe.run(full[:, lo:hi].reshape(7 * (hi - lo), 128, 128).cuda(), hi - lo, group=dist.group.WORLD)
        x3d = gather_frames(out["points3d_wo_procrustes"], dist.group.WORLD)
        torch.cuda.synchronize()
        res[solver] = {"x3d": x3d.cpu(), "cam": out["cam_rt"].cpu(), "idx": out["idx"].cpu()}
        del pipe
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_sharded_two_gpus_equals_single_gpu(tmp_path, lib_built):
    """Frame-sharded run over NCCL (2 ranks: all-gather of the 2-D points, bundle adjustment -- replicated with the LSMR
    solver, per-point work sharded + per-block partials all-gathered with the exact one --, local DLT, all-gather of
    the 3-D joints) == the single-GPU run on the concatenated frames, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import socket

    import torch.multiprocessing as mp

    from deepfly3d_b200.inference import random_state_dict
    from deepfly3d_b200.pipeline import Pose3DPipeline

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    T = 26                                           # 26 x 38 = 988 points = 8 blocks of 128: splits over 2 ranks
    from deepfly3d_b200 import ops

    assert ops.ba_sharded_plan(7, T, 38, 2) is not None
    mp.spawn(_sharded_worker, args=(2, port, T, str(tmp_path)), nprocs=2, join=True)
    full = ohg.to_uint8(ohg.synthetic_images(7 * T, 128, 128, seed=5))
    for solver in ("lsmr", "exact"):
        pipe = Pose3DPipeline(random_state_dict(2, seed=0), 128, 128, 7 * T, image_shape=[960, 480], ba_solver=solver)
        ref = pipe.run(full.cuda(), T)
        torch.cuda.synchronize()
        for r in range(2):
            d = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
            lo, hi = d["range"]
            assert torch.equal(d[solver]["idx"], ref["idx"].cpu().view(7, T, -1)[:, lo:hi].reshape(7 * (hi - lo), -1))
            assert torch.equal(d[solver]["cam"], ref["cam_rt"].cpu()), f"{solver}: multi-GPU bundle adjustment differs from the single-GPU solve"
            assert torch.equal(d[solver]["x3d"], ref["points3d_wo_procrustes"].cpu())
        del pipe


def test_core_streams_videos_without_expanding_them(tmp_path, lib_built):
    """config 5 ingest: Core(stream_videos=True) decodes camera_N.mp4 straight into the pipeline (no JPEG files are
    written) and gives what the same decoded frames give through the engine directly."""
    import cv2

    from deepfly3d_b200 import inference
    from deepfly3d_b200.core import Core
    from deepfly3d_b200.hourglass import HourglassEngine
    from deepfly3d_b200.skeleton import HEATMAP_SHAPE

    d = tmp_path / "sample" / "test"
    d.mkdir(parents=True)
    T = 3
    for c in range(7):
        vw = cv2.VideoWriter(str(d / f"camera_{c}.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 100.0, (960, 480))
        if not vw.isOpened():
            pytest.skip("no mp4 encoder in this OpenCV build")
        for t in range(T):
            vw.write(cv2.imread(os.path.join(IMAGES, f"camera_{c}_img_{t}.jpg")))
        vw.release()
    model = ohg.make_model(2, seed=0)
    core = Core(str(d), num_images_max=0, camera_ordering=[0, 1, 2, 3, 4, 5, 6], state_dict=model.state_dict(), stream_videos=True)
    assert core.num_images == T and core.image_shape == [960, 480] and abs(core.fps - 100.0) < 1e-6
    core.pose2d_estimation()
    assert not [f for f in os.listdir(d) if f.endswith(".jpg")]
    frames = np.zeros((7, T, 480, 960), dtype=np.uint8)
    for c in range(7):
        cap = cv2.VideoCapture(str(d / f"camera_{c}.mp4"))
        for t in range(T):
            frames[c, t] = cv2.cvtColor(cap.read()[1], cv2.COLOR_BGR2GRAY)
    from deepfly3d_b200 import ops

    dev = ops.resize_gray_u8(torch.as_tensor(frames.reshape(7 * T, 480, 960)).cuda(), (256, 512))
    eng = HourglassEngine(model.state_dict(), 256, 512, max_batch=7 * T)
    flip = torch.zeros((7, T), dtype=torch.uint8)
    flip[4:] = 1
    idx, conf = eng.forward(dev, flip=flip.reshape(-1).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(core.conf.reshape(7 * T, 19), conf.cpu().numpy())
    ref = opack.pack_points2d(opack.indices_to_points2d(idx.cpu().numpy().reshape(7, T, 19), HEATMAP_SHAPE), range(7))
    assert np.array_equal(core.points2d, ref)
    eng.close()
    inference.drop_engine()
