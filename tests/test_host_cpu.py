"""Host-side pieces of the product that need no GPU: the product's procrustes registration against the
reference's golden 3-D result, the checkpoint / mean-file loaders (the layout of df2d's sh8_deepfly.tar and
mean.pth.tar, reference df3d/config.py:30-39), and the block plan of the streaming image loader."""
import os

import numpy as np
import pytest
import torch


def test_product_procrustes_matches_golden(golden):
    """deepfly3d_b200.procrustes (the file Core.save runs) on golden points3d_wo_procrustes == golden points3d,
    at the 1e-12 the oracle copy reaches (reference df3d/procrustes.py:51, tests/test_df3d.py:227-232)."""
    from deepfly3d_b200.procrustes import procrustes_seperate, read_template_pose3d

    r3 = golden["result_3d"]
    np.testing.assert_array_equal(read_template_pose3d(), golden["template"]["points3d"])   # packaged template
    out = procrustes_seperate(r3["points3d_wo_procrustes"])
    np.testing.assert_allclose(out, r3["points3d"], atol=1e-12, rtol=0)
    # and against the oracle's independent restatement on a second input (longer, jittered)
    from oracle import procrustes as oproc

    rng = np.random.default_rng(0)
    X = np.repeat(r3["points3d_wo_procrustes"], 7, axis=0) + rng.normal(scale=0.03, size=(105, 38, 3))
    np.testing.assert_allclose(procrustes_seperate(X), oproc.procrustes_separate(X, golden["template"]["points3d"]),
                               atol=1e-12, rtol=0)


def test_checkpoint_layout_roundtrip(tmp_path, lib_built):
    """A file shaped like sh8_deepfly.tar -- {'state_dict': {'module.<key>': tensor}, 'epoch': ...} -- loads by key
    into the flat blob the C ABI takes, identical to flattening the model's own state dict."""
    from deepfly3d_b200 import hourglass, inference
    from oracle import hourglass as ohg

    model = ohg.make_model(2, seed=4)
    sd = model.state_dict()
    path = tmp_path / "sh2_synth.tar"
    torch.save({"epoch": 7, "state_dict": {"module." + k: v.clone() for k, v in sd.items()}, "best_acc": 0.5}, path)
    loaded = inference.load_state_dict(str(path))
    assert all(k.startswith("module.") for k in loaded)
    blob_a, S, K = hourglass.flatten_state_dict(loaded)
    blob_b, _, _ = hourglass.flatten_state_dict(sd)
    assert (S, K) == (2, 19) and np.array_equal(blob_a, blob_b)
    # a bare state dict (no wrapper) and $DF3D_B200_WEIGHTS work too
    torch.save(sd, tmp_path / "bare.tar")
    os.environ["DF3D_B200_WEIGHTS"] = str(tmp_path / "bare.tar")
    try:
        blob_c, _, _ = hourglass.flatten_state_dict(inference.load_state_dict(None))
    finally:
        del os.environ["DF3D_B200_WEIGHTS"]
    assert np.array_equal(blob_c, blob_b)
    with pytest.raises(ValueError):
        hourglass.flatten_state_dict({"conv1.weight": torch.zeros(1)})        # not an hourglass checkpoint
    # the cache key follows the file, not an object id (id() of a freed dict is recycled)
    k1 = inference._weights_key(None, str(path))
    os.utime(path, ns=(1, 1))
    assert inference._weights_key(None, str(path)) != k1
    assert inference._weights_key(None, None) == ("random", 0)


def test_mean_file_layouts(tmp_path):
    from deepfly3d_b200.inference import load_mean

    assert load_mean(None) == (0.5, 0.5, 0.5)
    assert load_mean(0.25) == (0.25, 0.25, 0.25)
    assert load_mean([0.1, 0.2, 0.3]) == (0.1, 0.2, 0.3)
    torch.save({"mean": torch.tensor([0.22, 0.22, 0.22]), "std": torch.ones(3)}, tmp_path / "mean.pth.tar")
    m = load_mean(str(tmp_path / "mean.pth.tar"))
    assert np.allclose(m, 0.22)
    torch.save(torch.tensor([0.4, 0.5, 0.6]), tmp_path / "bare.pth.tar")
    assert np.allclose(load_mean(str(tmp_path / "bare.pth.tar")), [0.4, 0.5, 0.6])
    with pytest.raises(ValueError):
        load_mean([1.0, 2.0])


def test_block_plan_bounds_memory():
    """The streaming loader cuts a recording into frame blocks sized for one engine launch sequence: device and
    pinned host memory are two blocks whatever T is (100 000 frames)."""
    from deepfly3d_b200.inference import block_frames_for, plan_blocks

    assert block_frames_for(256, 256, 100000) == 256           # 1 792 images
    assert block_frames_for(256, 512, 100000) == 128           # reference input size: 896 images
    assert block_frames_for(256, 512, 15) == 15                # small folder: one block
    assert block_frames_for(256, 512, 100000, batch_size=7 * 400) == 400
    blocks = plan_blocks(100000, 128)
    assert blocks[0] == (0, 128) and blocks[-1] == (99968, 100000) and len(blocks) == 782
    assert all(b[1] - b[0] <= 128 for b in blocks) and sum(b[1] - b[0] for b in blocks) == 100000
    assert plan_blocks(0, 8) == [] and plan_blocks(5, 8) == [(0, 5)]
    # two pinned + two device staging blocks of 480 x 960 frames
    assert 4 * 7 * 128 * 480 * 960 < 2 * 2**30
