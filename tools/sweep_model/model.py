"""Driver of the host model of the depth-sweep scoring kernel (sweep_model.cpp).

TEST INFRASTRUCTURE: builds libsweep_model.so with g++ from the same header the device kernels
include (findnpropagate_b200/csrc/fnp_sweep.cuh) and runs it on the frustums the CPU oracle
produces, so the range logic of FNP_SCORE_SWEEP is checked against brute-force counting without a
GPU.  Used by tests/test_sweep_model_cpu.py and, as a script, to print workload statistics:

    python tools/sweep_model/model.py cfg2 [n_frames]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libsweep_model.so")
SRC = os.path.join(HERE, "sweep_model.cpp")
HDR = os.path.join(ROOT, "findnpropagate_b200", "csrc", "fnp_sweep.cuh")


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.sweep_model_counts.restype = C.c_int
    lib.sweep_model_counts.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _round_down_f32(t):
    """largest float32 <= t (t float64 array)"""
    f = t.astype(np.float32)
    up = f.astype(np.float64) > t
    return np.where(up, np.nextafter(f, np.float32(-np.inf)), f).astype(np.float32)


def prep_boxes(boxes, cosf, sinf):
    """(n,7) boxes -> (n,8) [cx,cy,cz,hz,cosa,sina,tx,ty], the arithmetic of prep_box()
    (findnpropagate_b200/csrc/fnp_common.cuh); cosf/sinf: CUDA-exact single-precision functions."""
    b = np.asarray(boxes, np.float32).reshape(-1, 7)
    out = np.empty((b.shape[0], 8), np.float32)
    out[:, 0:3] = b[:, 0:3]
    out[:, 3] = _round_down_f32(b[:, 5].astype(np.float64) * 0.5)
    ang = -b[:, 6]
    uniq = {float(a): (cosf(np.float32(a)), sinf(np.float32(a))) for a in np.unique(ang)}   # scalar functions
    out[:, 4] = [uniq[float(a)][0] for a in ang]
    out[:, 5] = [uniq[float(a)][1] for a in ang]
    for col, dim in ((6, 3), (7, 4)):
        t = b[:, dim].astype(np.float64) * 0.5 + np.float64(np.float32(1e-5))
        f = _round_down_f32(t)
        eq = f.astype(np.float64) == t
        out[:, col] = np.where(eq, np.nextafter(f, np.float32(-np.inf)), f)
    return out


def run_frustum(lib, xyz, prep, hidx, J, M, split=2048, cols_out=None, dev_out=None):
    """-> counts_sweep, counts_brute (nv), stats dict"""
    xyz = np.ascontiguousarray(xyz[:, :3], np.float32)
    prep = np.ascontiguousarray(prep, np.float32)
    hidx = np.ascontiguousarray(hidx, np.int32)
    nv = hidx.shape[0]
    cs, cb = np.zeros(max(nv, 1), np.int32), np.zeros(max(nv, 1), np.int32)
    st = np.zeros(8, np.int64)
    maxabs = float(np.abs(xyz).max()) if xyz.shape[0] else 0.0
    maxabs_z = float(np.abs(xyz[:, 2]).max()) if xyz.shape[0] else 0.0
    rc = lib.sweep_model_counts(xyz.ctypes.data, xyz.shape[0], prep.ctypes.data, hidx.ctypes.data, nv, J, M,
                                C.c_float(maxabs), C.c_float(maxabs_z), split, cs.ctypes.data, cb.ctypes.data, st.ctypes.data,
                                cols_out.ctypes.data if cols_out is not None else None,
                                dev_out.ctypes.data if dev_out is not None else None)
    assert rc == 0
    return cs[:nv], cb[:nv], dict(exact_tests=int(st[0]), adds=int(st[1]), pairs=int(st[2]), pseudo_axes=int(st[3]),
                                  columns=int(st[4]), pairs_definite=int(st[5]), pairs_uncertain=int(st[6]))


def run_frame(lib, cfg_name, seed=0, split=2048, params_override=None):
    """Oracle frame -> per-frustum (sweep, brute, oracle) counts and summed stats."""
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle as O
    import seeker_oracle as SO
    from findnpropagate_b200 import synth
    cfg = synth.CONFIGS[cfg_name]
    params = synth.seeker_params(cfg)
    if params_override:
        params.update(params_override)
    f = synth.make_frame(seed, cfg)
    o = SO.seek_frame(f.points, f.lidar2image, f.camera2lidar, f.camera_intrinsics,
                      (f.det_boxes, f.det_labels, f.det_scores, f.det_cam_idx), params, keep_intermediates=True)
    M = max(int(params["num_mags"]), 1)
    J = int(params["num_rotations"]) * int(params["num_sizes"])
    tot = dict(exact_tests=0, adds=0, pairs=0, pseudo_axes=0, columns=0, pairs_definite=0, pairs_uncertain=0,
               brute_tests=0, frustums=0)
    res = []
    for r in o["frustums"]:
        if "xyz" not in r or not r["valid"].any():
            continue
        hidx = np.nonzero(r["valid"])[0].astype(np.int32)
        prep = prep_boxes(r["hyp_boxes"][hidx], O.cosf, O.sinf)
        cs, cb, st = run_frustum(lib, r["xyz"], prep, hidx, J, M, split)
        res.append((cs, cb, r["counts"][hidx]))
        for k in st:
            tot[k] += st[k]
        tot["brute_tests"] += r["xyz"].shape[0] * hidx.shape[0]
        tot["frustums"] += 1
    return res, tot


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = build()
    for seed in range(n):
        res, tot = run_frame(lib, name, seed)
        bad = sum(int((a != b).sum()) + int((b != c).sum()) for a, b, c in res)
        print(name, "frame", seed, "mismatching counts:", bad, tot,
              "exact tests per (point, column): %.3f" % (tot["exact_tests"] / max(tot["pairs"], 1)),
              "shared adds per pair: %.3f" % (tot["adds"] / max(tot["pairs"], 1)),
              "brute tests per pair: %.1f" % (tot["brute_tests"] / max(tot["pairs"], 1)))
