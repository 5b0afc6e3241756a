"""CPU check of the depth-sweep scoring logic (FNP_SCORE_SWEEP).

tools/sweep_model/sweep_model.cpp compiles the very functions the device kernels call
(findnpropagate_b200/csrc/fnp_sweep.cuh: sweep_col_build, sweep_point, in_box) for the host and
runs them next to a brute-force count with in_box(); both must give the same integers, and on
oracle frames they must equal the oracle's counts (reference predicate,
roiaware_pool3d_kernel.cu:16-36).  The GPU parity tests (tests/test_seeker_gpu.py) check the
kernels themselves; this one keeps the range logic honest where there is no GPU.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "sweep_model"))
import model as SM  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    return SM.build()


@pytest.mark.parametrize("cfg,seed,override", [
    ("tiny", 0, None), ("tiny", 1, None), ("tiny", 2, dict(num_mags=24)), ("cfg1", 0, None),
    ("cfg1", 1, dict(num_mags=48, num_rotations=7)), ("cfg1", 2, dict(num_mags=33, num_sizes=2)),
])
def test_sweep_equals_brute_force_and_oracle_on_frames(lib, cfg, seed, override):
    res, tot = SM.run_frame(lib, cfg, seed, split=512, params_override=override)
    assert tot["frustums"] > 0
    for cs, cb, co in res:
        assert np.array_equal(cb, co), "brute force with in_box() differs from the oracle"
        assert np.array_equal(cs, cb), "sweep differs from brute force"
    if (override or {}).get("num_mags", 0) >= 24:
        assert tot["exact_tests"] < 0.2 * tot["brute_tests"]      # the sweep does take the short cut


def _random_case(rng, J, M, scale, step, jitter, n_pts, valid_p):
    """A frustum made to hurt: centres on a line of step `step` with `jitter` of deviation, an
    offset of `scale` metres, points sprinkled on the faces of the boxes (+- a few ulp)."""
    prep, hidx = [], []
    yaw = rng.uniform(-np.pi, np.pi, J)
    dims = rng.uniform(0.3, 12.0, (J, 3))
    origin = rng.uniform(-1, 1, 3) * scale
    direction = rng.normal(size=3)
    direction /= np.linalg.norm(direction)
    if rng.random() < 0.3:
        direction[rng.integers(0, 3)] = 0.0       # no travel along one world axis
    for m in range(M):
        for j in range(J):
            if rng.random() > valid_p:
                continue
            c = origin + direction * step * m + rng.normal(size=3) * jitter
            t = dims[j] * 0.5
            prep.append([c[0], c[1], c[2], t[2], np.cos(-yaw[j]), np.sin(-yaw[j]), t[0], t[1]])
            hidx.append(m * J + j)
    prep = np.asarray(prep, np.float32).reshape(-1, 8)
    hidx = np.asarray(hidx, np.int32)
    # all hypotheses of a column must share rotation and size exactly (fp32)
    for j in range(J):
        sel = (hidx % J) == j
        if sel.any():
            prep[sel, 3:] = prep[sel][0, 3:]
    pts = []
    for _ in range(n_pts):
        if len(prep) and rng.random() < 0.8:
            b = prep[rng.integers(0, len(prep))]
            l = rng.uniform(-1.2, 1.2, 3) * np.array([b[6], b[7], b[3]])
            if rng.random() < 0.7:                  # right on a face, a few ulp either side
                k = rng.integers(0, 3)
                l[k] = np.sign(l[k] + 1e-30) * np.array([b[6], b[7], b[3]])[k] * (1 + rng.integers(-4, 5) * 6e-8)
            ca, sa = np.float64(b[4]), np.float64(b[5])
            # local = Rot(-yaw)(p - c)  =>  p = c + Rot(yaw) local;  cosa = cos(-yaw), sina = sin(-yaw)
            x = b[0] + l[0] * ca + l[1] * sa
            y = b[1] - l[0] * sa + l[1] * ca
            pts.append([x, y, b[2] + l[2]])
        else:
            pts.append(origin + rng.normal(size=3) * 5)
    return np.asarray(pts, np.float32).reshape(-1, 3), prep, hidx


@pytest.mark.parametrize("seed", range(12))
def test_sweep_equals_brute_force_on_adversarial_columns(lib, seed):
    rng = np.random.default_rng(100 + seed)
    J = int(rng.integers(1, 9))
    M = int(rng.choice([1, 2, 5, 16, 64, 100]))
    scale = float(rng.choice([1.0, 50.0, 200.0, 1000.0]))
    step = float(rng.choice([0.0, 1e-7, 1e-5, 1e-3, 2e-2, 0.5]))
    jitter = float(rng.choice([0.0, 1e-6, 1e-4, 1e-2, 0.5]))
    pts, prep, hidx = _random_case(rng, J, M, scale, step, jitter, 3000, float(rng.choice([1.0, 0.7, 0.2])))
    if hidx.shape[0] == 0:
        pytest.skip("no valid hypothesis drawn")
    cs, cb, st = SM.run_frustum(lib, pts, prep, hidx, J, M, split=int(rng.choice([64, 1000, 4096])))
    assert np.array_equal(cs, cb), (J, M, scale, step, jitter, st)
    assert cb.sum() > 0
