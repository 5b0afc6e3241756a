"""Builds libfnp_sm100.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m findnpropagate_b200.build [--force] [--verbose]

-fmad=false: every fused multiply-add in the library is written explicitly (__fmaf_rn), so
the results are bit-reproducible against the reference kernels' compiled arithmetic.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libfnp_sm100.so")
SOURCES = ["fnp_ops.cu", "fnp_seeker.cu", "fnp_host.cpp"]
HEADERS = ["fnp_common.cuh", "fnp_sweep.cuh", os.path.join("..", "..", "include", "fnp.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, defs=(), out=None):
    """defs / out: A/B builds with other compile-time constants (-DNAME=value) into another file; a process picks
    one with FNP_LIB_PATH (see _lib.py)."""
    if out is None and not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defs] + ["-o", out or SO] + [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libfnp_sm100.so")
    return out or SO


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defs=defs, out=outs[0] if outs else None))
