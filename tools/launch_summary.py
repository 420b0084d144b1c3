"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`, see tools/gpu_round.sh).
    python tools/launch_summary.py gpurun_out/launches.csv r01c > profiles/r01c_launches_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, tag):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    tot = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"].replace("df3d::", "")).replace("void ", "")
        ms = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r["Metric Unit"]]
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + ms)
    total = sum(t for _, t in tot.values())
    print(f"# {tag}: ncu launch list of ONE full-size bench step (bench.py --profile, NVTX range df3d_step): 256 frames x 7 cams (bench.py --profile --frames 256),")
    print("# 8-stack 256x256, conv-chain plan (DF3D_HG_FUSE=2).  ncu --metrics gpu__time_duration.sum --clock-control none")
    print(f"# {sum(n for n, _ in tot.values())} launches, {total:.3f} ms summed (cold-cache, serialised: compare SHARES with bench.py's roofline.share_of_step)")
    print(f"{'kernel':40s} {'launches':>8s} {'ms':>10s} {'share':>8s}")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:40]:40s} {n:8d} {t:10.3f} {100 * t / total:7.2f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "rNN")
