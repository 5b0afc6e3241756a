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

It also installs the ten Python files of the reference that `FrustumProposerOG` needs at import
time (the head, its op wrappers and utils), byte for byte, under `oracle/_ref/pysrc/pcdet/...`
-- the equivalent of `pip install --target` for the slice of the package on this path
(the full package needs spconv / SharedArray / kornia and is not installable here).  Like the
`.so` files they are git-ignored build outputs that travel with the gpurun snapshot, so that
`tests/test_reference_gpu.py` and `tools/ref_gpu_bench.py` can run the reference's own
`get_proposals` on the B200.  They are never imported by the product.

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


# reference Python files FrustumProposerOG imports (tools/ref_seeker.py lists why each one is needed)
PY_FILES = [
    "pcdet/models/dense_heads/frustum_proposals_v1.py",
    "pcdet/models/dense_heads/frustum_proposals_v1_kitti.py",
    "pcdet/models/dense_heads/target_assigner/hungarian_assigner.py",
    "pcdet/models/model_utils/centernet_utils.py",
    "pcdet/models/model_utils/model_nms_utils.py",
    "pcdet/models/preprocessed_detector.py",
    "pcdet/utils/box_utils.py",
    "pcdet/utils/common_utils.py",
    "pcdet/utils/loss_utils.py",
    "pcdet/utils/calibration_kitti.py",
    "pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py",
    "pcdet/ops/iou3d_nms/iou3d_nms_utils.py",
]
PYSRC = os.path.join(OUT, "pysrc")


def install_py(force=False):
    """Copies PY_FILES (unmodified) to oracle/_ref/pysrc/; a no-op without the reference tree."""
    import shutil
    if not os.path.isdir(REF):
        return
    for rel in PY_FILES:
        src, dst = os.path.join(REF, rel), os.path.join(PYSRC, rel)
        if not os.path.exists(src):
            continue
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)


def py_root():
    """Root of the reference's Python files: the reference tree itself where it exists, else the installed copy."""
    if os.path.isdir(os.path.join(REF, "pcdet")):
        return REF
    if os.path.isdir(os.path.join(PYSRC, "pcdet")):
        return PYSRC
    raise ImportError("neither %s nor oracle/_ref/pysrc holds the reference's Python files" % REF)


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
    install_py(force)
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
