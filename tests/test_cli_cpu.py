"""The df3d-cli mirror (deepfly3d_b200/cli.py) parses the reference's command lines (df3d/cli.py:62-166) and keeps
its folder logic (cli.py:170-273, 329-354).  No GPU: Core is replaced by a recorder."""
import os

import pytest


def test_flags_match_the_reference(tmp_path):
    from deepfly3d_b200 import cli

    a = cli.parse_cli_args([str(tmp_path)])
    assert a.output_folder == str(tmp_path) + "_df3d" and a.order == [0, 1, 2, 3, 4, 5, 6]        # cli.py:159-165, 116-122
    assert (a.num_images_max, a.batch_size, a.skip_estimation, a.pin_memory_disabled) == (0, 8, False, False)
    a = cli.parse_cli_args([str(tmp_path), "-n", "100", "--camera-ids", "6", "5", "4", "3", "2", "1", "0", "--skip-pose-estimation",
                            "--batch-size", "12", "--pin-memory-disabled", "-x", "-vv", "--output-folder", str(tmp_path / "o"),
                            "--output-fps", "30", "-r"])
    assert a.order == [6, 5, 4, 3, 2, 1, 0] and a.num_images_max == 100 and a.skip_estimation and a.delete_images
    assert a.batch_size == 12 and a.pin_memory_disabled and a.verbose2 and a.recursive and a.output_fps == 30.0
    assert a.output_folder == str(tmp_path / "o")


def test_recursive_and_from_file_run_every_folder(tmp_path, monkeypatch):
    from deepfly3d_b200 import cli, core

    calls = []

    class FakeCore:
        def __init__(self, inp, out, n, order, weights=None, mean=None):
            if "bad" in inp:
                raise ValueError("boom")
            self.max_img_id = 2
            calls.append(["init", inp, out, n, list(order)])

        def pose2d_estimation(self, bs, pin):
            calls[-1].append(("pose2d", bs, pin))

        def calibrate_calc(self, lo, hi):
            calls[-1].append(("calib", lo, hi))

        def save(self):
            calls[-1].append("save")

        def delete_images(self):
            calls[-1].append("delete")

    monkeypatch.setattr(core, "Core", FakeCore)
    for d in ("a/images", "b/deep/images", "bad/images", "c/other"):
        os.makedirs(tmp_path / d)
    (tmp_path / "a/images/images").mkdir()                     # a match is not descended into
    assert sorted(cli.find_subfolders(str(tmp_path), "images")) == sorted(str(tmp_path / d) for d in ("a/images", "b/deep/images", "bad/images"))
    rc = cli.main([str(tmp_path), "-r", "-x"])
    assert rc == 1                                             # one folder failed, the others ran (cli.py:252-273)
    assert len(calls) == 2
    assert calls[0][5:] == [("pose2d", 8, False), "save", ("calib", 0, 2), "save", "delete"]       # cli.py:296-300, 323-324
    assert all(c[2] == c[1] + "_df3d" for c in calls)
    calls.clear()
    lst = tmp_path / "folders.txt"
    lst.write_text(f"{tmp_path / 'a/images'}\n\n{tmp_path / 'a/images'}\n{tmp_path / 'c/other'}\n")
    assert cli.main([str(lst), "-f", "--skip-pose-estimation"]) == 0
    assert calls == []                                         # skip without a video flag: "Nothing to do" (cli.py:282-289)
    assert cli.main([str(lst), "-f", "-n", "5", "--order", "6", "5", "4", "3", "2", "1", "0"]) == 0
    assert [c[1] for c in calls] == [str(tmp_path / "a/images"), str(tmp_path / "c/other")]         # duplicates and blanks dropped
    assert calls[0][3:5] == [5, [6, 5, 4, 3, 2, 1, 0]]
    lst.write_text(f"{tmp_path / 'missing'}\n")
    assert cli.main([str(lst), "-f"]) == 1
    assert cli.main([str(tmp_path), "-f", "-r"]) == 1
    with pytest.raises(NotImplementedError):
        cli.main([str(tmp_path / "a/images"), "--video-2d"])
