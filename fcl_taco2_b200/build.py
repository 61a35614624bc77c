"""Build recipe of libfcl_taco2.so: every kernel compiled for sm_100a with nvcc, in-tree."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libfcl_taco2.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(HERE), "include", "fcl_taco2.h")]
    return any(os.path.getmtime(d) > t for d in deps)


PROF_LIB = os.path.join(HERE, "lib", "libfcl_taco2_prof.so")


def build(force: bool = False, verbose: bool = False, prof: bool = False) -> str:
    """`prof=True`: the profiling build (-DFCL_DEC_PROF: counter-mode probes and what-if switches in the pair decoders,
    tools/decoder_prof.py) as lib/libfcl_taco2_prof.so; the product library carries none of it."""
    if prof:
        out, extra = PROF_LIB, ["-DFCL_DEC_PROF"]
    else:
        if not force and not needs_build():
            return LIB
        out, extra = LIB, []
    os.makedirs(os.path.dirname(out), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out
