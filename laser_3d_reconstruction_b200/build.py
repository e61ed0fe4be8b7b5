"""Build libl3d.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m laser_3d_reconstruction_b200.build [--force]

The shared object lands next to this file (git-ignored; it travels to the GPU box with gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libl3d.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--fmad=false", "-Xptxas", "-v",
]


def sources():
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "l3d.h"))
    return deps


def build(force: bool = False, verbose: bool = False) -> str:
    deps = sources()
    if (not force and os.path.exists(OUT)
            and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps)):
        return OUT
    cmd = [NVCC] + FLAGS + ["-o", OUT, os.path.join(CSRC, "l3d_all.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libl3d.so")
    with open(os.path.join(HERE, "libl3d.ptxas.log"), "w") as f:
        f.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
