"""DRAM bytes per conv launch of one bench step, from an ncu metrics pass
    ncu --nvtx --nvtx-include "df3d_step/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/traffic.csv python bench.py --profile --steps 1 --warmup 3
    python tools/traffic_summary.py gpurun_out/traffic.csv > profiles/conv_gemm_traffic.json
bench.py reports `dram_bytes_per_launch` as roofline.traffic (same averaging as roofline.achieved: all tcgen05 conv launches)."""
import csv
import json
import re
import sys


def main(path):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if l.startswith('"')]))
    per = {}
    for r in rows:
        k = per.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"].replace("df3d::", "")).replace("void ", "")})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1)
        k[r["Metric Name"]] = v * scale
    conv = [k for k in per.values() if k["name"].startswith(("conv_chain_kernel", "conv_gemm_kernel"))]
    tot = sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in conv)
    ms = sum(k.get("gpu__time_duration.sum", 0) for k in conv)
    classes = {}
    for k in conv:
        c = classes.setdefault(k["name"], [0, 0.0, 0.0])
        c[0] += 1
        c[1] += k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0)
        c[2] += k.get("gpu__time_duration.sum", 0)
    out = {
        "what": "dram__bytes_read.sum + dram__bytes_write.sum of every tcgen05 conv launch of ONE bench step (1792 images), ncu metrics pass",
        "conv_launches": len(conv),
        "dram_bytes_per_step_conv": tot,
        "dram_bytes_per_launch": tot / max(len(conv), 1),
        "ms_summed_under_ncu": ms,
        "per_kernel": {n: {"launches": c[0], "dram_bytes": c[1], "ms": c[2], "gbps_under_ncu": c[1] / c[2] / 1e6 if c[2] else 0} for n, c in classes.items()},
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
