"""TEST INFRASTRUCTURE ONLY -- builds the *reference's own* native ops, unmodified.

Compiles the six source files of `pcdet/ops/roiaware_pool3d` and `pcdet/ops/iou3d_nms`
straight from where they lie under /root/reference (nothing is copied into this repo)
into two pybind11/ATen extension modules under `oracle/_ref/`:

    oracle/_ref/roiaware_pool3d_cuda.so   (points_in_boxes_gpu / points_in_boxes_cpu / pooling)
    oracle/_ref/iou3d_nms_cuda.so         (boxes_iou_bev_gpu/cpu, boxes_overlap_bev_gpu,
                                           nms_gpu, nms_normal_gpu ...)

This is our own recipe (plain nvcc + g++ command lines, run in parallel), not the
reference's setup.py.  The resulting `.so` files are git-ignored but travel to the GPU
box with the gpurun snapshot, where the `-m gpu` tests use them as the bit-exact checker
and `bench.py --impl reference` uses `points_in_boxes_cpu` as the timed CPU baseline.
/root/reference does not exist on the GPU box: only the prebuilt `.so` files are used
there.

Usage:  python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("FNP_REFERENCE_ROOT", "/root/reference")

MODULES = {
    "roiaware_pool3d_cuda": [
        "pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp",
        "pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu",
    ],
    "iou3d_nms_cuda": [
        "pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp",
        "pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp",
        "pcdet/ops/iou3d_nms/src/iou3d_nms.cpp",
        "pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu",
    ],
}


def _flags(name):
    import torch
    from torch.utils import cpp_extension as ce

    inc = []
    for p in ce.include_paths(device_type="cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    common = [
        "-DTORCH_EXTENSION_NAME=%s" % name,
        "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
        "-std=c++17", "-O3", "-w",
    ]
    cxx = ["g++", "-fPIC"] + common + inc
    # same code generation the reference would get from torch's BuildExtension on this
    # box (TORCH_CUDA_ARCH_LIST=10.0a): -O3 default fp model (fmad on, prec-div on)
    nv = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-D__CUDA_NO_HALF_OPERATORS__",
          "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"] + common + inc
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = ["g++", "-shared", "-L" + libdir, "-Wl,-rpath," + libdir,
            "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
            "-ltorch", "-ltorch_python", "-lcudart"]
    return cxx, nv, link


def _compile(job):
    cmd, obj = job
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference build failed:\n%s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return obj


def build(force=False):
    if not os.path.isdir(REF):
        # GPU box: the prebuilt files are all there is.
        missing = [m for m in MODULES if not os.path.exists(os.path.join(OUT, m + ".so"))]
        if missing:
            raise RuntimeError("reference tree %s absent and oracle/_ref lacks %s" % (REF, missing))
        return
    os.makedirs(OUT, exist_ok=True)
    jobs, links = [], []
    for name, srcs in MODULES.items():
        so = os.path.join(OUT, name + ".so")
        srcs = [os.path.join(REF, s) for s in srcs]
        if (not force and os.path.exists(so)
                and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs)):
            continue
        cxx, nv, link = _flags(name)
        objs = []
        for s in srcs:
            obj = os.path.join(OUT, "%s__%s.o" % (name, os.path.basename(s).replace(".", "_")))
            objs.append(obj)
            jobs.append(((nv if s.endswith(".cu") else cxx) + ["-c", s, "-o", obj], obj))
        links.append((link, objs, so))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(6, len(jobs))) as ex:
            list(ex.map(_compile, jobs))
    for link, objs, so in links:
        _compile((link[:2] + objs + link[2:] + ["-o", so], so))
        for o in objs:
            os.remove(o)


def load(name):
    """Import one of the prebuilt reference modules (pybind11) from oracle/_ref."""
    import importlib.util
    import torch  # noqa: F401  (the module links against libtorch)

    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        raise ImportError("oracle/_ref/%s.so not built -- run python oracle/build_ref.py" % name)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print("built:", sorted(os.listdir(OUT)))
