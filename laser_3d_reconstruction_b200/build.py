"""Build libl3d.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m laser_3d_reconstruction_b200.build [--force]

Every csrc/*.cu is its own translation unit (compiled in parallel into build/, no relocatable device code:
kernels only call device functions of their own file), linked into one shared object that lands next to
this file (git-ignored; it travels to the GPU box with gpurun).
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libl3d.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--fmad=false", "-Xptxas", "-v",
]


def units():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def headers():
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    hs.append(os.path.join(HERE, "..", "include", "l3d.h"))
    return hs


def sources():
    return units() + headers()


def _compile(src, obj):
    res = subprocess.run([NVCC] + FLAGS + ["-c", "-o", obj, src], capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in headers())
    todo, objs = [], []
    for src in units():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            todo.append((src, obj))
    if not todo and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(o) for o in objs):
        return OUT
    logs = {}
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for src, rc, log in ex.map(lambda t: _compile(*t), todo):
            logs[os.path.basename(src)] = log
            if verbose or rc != 0:
                sys.stderr.write(log)
            if rc != 0:
                raise RuntimeError("nvcc failed on %s" % src)
    res = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libl3d.so")
    for name, log in logs.items():
        with open(os.path.join(OBJ, name + ".ptxas.log"), "w") as f:
            f.write(log)
    with open(os.path.join(HERE, "libl3d.ptxas.log"), "w") as f:
        for name in sorted(os.listdir(OBJ)):
            if name.endswith(".ptxas.log"):
                f.write("==== %s\n" % name[:-10])
                f.write(open(os.path.join(OBJ, name)).read())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
