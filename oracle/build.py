"""Build the oracle's C restatement (test infrastructure only) into oracle/_ref/liborc.so.

The reference is pure Python (no C/C++ sources to compile), so oracle/_ref holds only this
restatement's shared object; it is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("orc_sgbm.c", "orc_image.c", "orc_wls.c", "orc_bm.c")]
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "liborc.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = SRC + [os.path.join(HERE, "csrc", "l3d_oracle.h")]
    if (not force and os.path.exists(OUT)
            and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in deps)):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=gnu11",
           "-I", os.path.join(HERE, "csrc"), "-o", OUT] + SRC + ["-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
