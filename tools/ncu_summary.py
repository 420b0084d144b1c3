"""Summarise an .ncu-rep (read here, without a GPU) into the per-launch numbers the roofline uses.
    python tools/ncu_summary.py gpurun_out/prof_conv.ncu-rep > profiles/rNN_conv_ncu.txt"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for h, i in list(idx.items()):       # some metrics carry a unit prefix (FBSP.TriageCompute.dram__throughput...)
        idx.setdefault(h.split(".", 2)[-1] if h.count(".") > 2 and not h.startswith(("dram", "lts", "sm", "gpu", "launch")) else h, i)
    print(f"# {path}: {len(data)} launches (ncu --set full --clock-control none; cold-cache, serialised replays)")
    print(f"{'#':>3s} {'kernel':44s} " + " ".join(f"{n:>12s}" for _, n in COLS))
    for k, r in enumerate(data):
        name = r[idx["Kernel Name"]].replace("df3d::", "")[:44]
        cells = []
        for m, _ in COLS:
            i = idx.get(m)
            if i is None:
                cells.append("-")
                continue
            v, u = r[i], units[i]
            try:
                f = float(v.replace(",", ""))
                cells.append(f"{f:.3f}{u[:5]}" if u not in ("%", "") else f"{f:.2f}{u}")
            except ValueError:
                cells.append(v[:12])
        print(f"{k:3d} {name:44s} " + " ".join(f"{c:>12s}" for c in cells))


if __name__ == "__main__":
    main(sys.argv[1])
