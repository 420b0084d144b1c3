"""In-tree build of libdf3d_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m deepfly3d_b200.build          # incremental
    python -m deepfly3d_b200.build --force

The shared library is written next to this file so that it travels with the repo snapshot to
the GPU box (it is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdf3d_b200.so")
STAMP = os.path.join(HERE, ".libdf3d_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
]


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith(".cu"):
                out.append(os.path.join(root, f))
    return sorted(out)


def _fingerprint():
    h = hashlib.sha256()
    files = [os.path.join(HERE, "..", "include", "df3d_b200.h")]
    for root, _, fs in os.walk(CSRC):
        files += [os.path.join(root, f) for f in fs if f.endswith((".cu", ".cuh", ".h"))]
    for p in sorted(files):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == fp:
                return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + sources() + ["-ldl"]  # dlopen: nvJPEG is bound at run time (jpeg.cu)
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libdf3d_b200.so")
    if verbose:
        print(res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
