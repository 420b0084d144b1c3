"""The bench contract (JSON line keys) on the committed bench lines of the last GPU session, and the reference arm
(CPU oracle port) end to end -- the driver runs `bench.py --impl reference` beside the CUDA arm."""
import glob
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    with open(files[-1]) as f:
        return json.loads(f.read())


def test_committed_bench_line_has_the_contract_keys():
    d = _latest("r02?_bench.json")
    assert BASE_KEYS <= set(d)
    assert d["metric"].startswith("7-cam frames/sec") and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    frames = d["config"]["frames_per_gpu"]
    assert "configs[2]" in d["config"]["workload"] and frames == 1000          # the literal BASELINE.json configuration
    assert e["h2d_bytes_per_step"] == 7 * frames * 256 * 256 * d["n_gpus"] and e["d2h_bytes_per_step"] > 0
    assert d["e2e_files"]["value"] > 0 and "Core(folder)" in d["e2e_files"]["path"]   # files on disk -> result pickle
    assert d["bundle_adjust"]["frames"] == frames * d["n_gpus"]
    r = d["roofline"]
    assert abs(r["achieved"] - r["flop_per_launch"] * r["launches_per_step"] / (r["ms_per_launch"] * r["launches_per_step"] / 1e3) / 1e12) < 1e-6 * r["achieved"]
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    for bad in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"):
        assert bad not in d["clocks"]["reasons"]


def test_committed_reference_line():
    d = _latest("r02?_bench_reference.json")
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]


def test_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-frames", "1", "--frames", "64"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    st = d["cpu_baseline"]["stages"]                                  # per-stage seconds
    assert {"hourglass_s", "argmax_s", "pack_s", "ba_s", "dlt_s", "procrustes_s"} <= set(st) and st["frames_3d"] == 64
    assert abs(d["ms_per_step"] - 1000.0 * 64 / d["value"]) < 1e-6    # per STEP of the workload, not per frame
