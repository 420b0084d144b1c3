"""Per-instruction warp-state samples of one launch of an .ncu-rep (read here, without a GPU).
    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep <launch index> [top N]
Prints the totals per stall reason, the totals per code region (role branches of the chain kernel are told apart by
the SASS address ranges between the role's first/last tcgen05 instruction), and the N instructions with the most samples."""
import csv
import io
import subprocess
import sys


def main(path, launch, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1",
                          "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = raw.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    reasons = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
    tot = {k: 0 for k in reasons}
    all_samples = 0
    for r in rows:
        for k in reasons:
            tot[k] += int(r[k] or 0)
        all_samples += int(r["# Samples"] or 0)
    print(f"{lines[0][:120]}")
    print(f"samples {all_samples}, instructions {len(rows)}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print(f"  {k[6:]:20s} {v:8d} {100.0 * v / max(all_samples, 1):6.1f}%")
    print("top instructions by samples: idx, samples, executed, top reasons, SASS")
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        rs = sorted(((int(r[k] or 0), k[6:]) for k in reasons), reverse=True)[:3]
        print(f"  {i:5d} {int(r['# Samples']):6d} {int(r['Instructions Executed'] or 0):9d}  " +
              " ".join(f"{n}:{c}" for c, n in rs if c) + "   " + r["Source"].strip()[:90])
    return rows


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40)
