// Host model of the depth-sweep scoring kernel (FNP_SCORE_SWEEP): runs the SAME functions the
// device kernels call (findnpropagate_b200/csrc/fnp_sweep.cuh: sweep_col_build, sweep_solve, sweep_emit, sweep_exact_step_col,
// in_box) serially on the CPU, next to the brute-force count with in_box(), so that the range
// logic can be checked without a GPU (tests/test_sweep_model_cpu.py).  Test infrastructure; not
// part of the product library.
//
//   g++ -O2 -ffp-contract=off -shared -fPIC -o libsweep_model.so sweep_model.cpp
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

struct float4 { float x, y, z, w; };
#define FNP_SWEEP_MODEL 1
static long long g_exact_tests = 0;   // incremented by sweep_exact_step_col under FNP_SWEEP_MODEL
#include "../../findnpropagate_b200/csrc/fnp_sweep.cuh"

using namespace fnp;

// pts (n,3) xyz; prep (nv,8) compacted hypotheses [cx,cy,cz,hz,cosa,sina,tx,ty]; hidx (nv) original
// index h = m*J + j, ascending; maxabs_pts = max |coordinate| of the frustum's point AABB.
// counts_sweep / counts_brute (nv).  stats (8): [0] exact tests taken by the sweep, [1] shared-memory
// adds, [2] (point, column) pairs, [3] constant-axis columns*axes, [4] columns, [5] pairs with a
// definite range.  split: points per split (the splits must add up like on the device).
extern "C" int sweep_model_counts(const float *pts, int n, const float *prep, const int *hidx, int nv, int J, int M,
                                  float maxabs_pts, float maxabs_pts_z, int split, int *counts_sweep, int *counts_brute, long long *stats,
                                  float *cols_out /* (J, FNP_SWEEP_COL_FLOATS) or NULL */, float *dev_out /* (J,3) or NULL */)
{
    const int H = J * M;
    if (H > 32767 || M > 255) return -1;
    memset(stats, 0, 8 * sizeof(long long));
    // ---- sweep_prep_kernel
    std::vector<int> first(J, 0x7fffffff), last(J, -1), r0(J, 0);
    std::vector<float> c0(3 * J, 0.f), c1(3 * J, 0.f), dev(3 * J, 0.f), den(3 * J, 0.f);
    float maxabs = maxabs_pts, maxabs_z = maxabs_pts_z;
    for (int r = 0; r < nv; r++) {
        const int h = hidx[r], m = h / J, j = h - m * J;
        if (m < first[j]) first[j] = m;
        if (m > last[j]) last[j] = m;
        for (int k = 0; k < 3; k++) maxabs = fmaxf(maxabs, fabsf(prep[r * 8 + k]));
        maxabs_z = fmaxf(maxabs_z, fabsf(prep[r * 8 + 2]));
    }
    for (int r = 0; r < nv; r++) {
        const int h = hidx[r], m = h / J, j = h - m * J;
        float C[3];
        sweep_axes(load_prep(prep, r), C);
        if (m == first[j]) { for (int k = 0; k < 3; k++) c0[3 * j + k] = C[k]; r0[j] = r; }
        if (m == last[j]) for (int k = 0; k < 3; k++) c1[3 * j + k] = C[k];
    }
    for (int i = 0; i < 3 * J; i++) {
        const int j = i / 3, span = last[j] - first[j];
        c1[i] = span > 0 ? f_div(f_sub(c1[i], c0[i]), (float)span) : 0.f;
    }
    for (int r = 0; r < nv; r++) {
        const int h = hidx[r], m = h / J, j = h - m * J;
        float C[3];
        sweep_axes(load_prep(prep, r), C);
        const float dm = (float)(m - first[j]);
        for (int k = 0; k < 3; k++) {
            const float line = f_fma(c1[3 * j + k], dm, c0[3 * j + k]);
            const float d = f_sub(C[k], line);
            if (d > 0.f) dev[3 * j + k] = fmaxf(dev[3 * j + k], d);
            if (d < 0.f) den[3 * j + k] = fmaxf(den[3 * j + k], -d);
        }
    }
    const float eps = sweep_eps(maxabs), eps_z = sweep_eps_z(maxabs_z);
    std::vector<SweepCol> cols(J);
    for (int j = 0; j < J; j++) {
        if (last[j] >= first[j]) {
            cols[j] = sweep_col_build(first[j], last[j], &c0[3 * j], &c1[3 * j], &dev[3 * j], &den[3 * j], load_prep(prep, r0[j]), eps, eps_z);
            stats[4]++;
            for (int k = 0; k < 3; k++) stats[3] += (cols[j].pseudo_mask >> k) & 1;
        } else {
            memset(&cols[j], 0, sizeof(SweepCol));
            cols[j].m0 = 0; cols[j].m1 = -1;
        }
    }
    if (cols_out) memcpy(cols_out, cols.data(), sizeof(SweepCol) * J);
    if (dev_out) memcpy(dev_out, dev.data(), sizeof(float) * 3 * J);
    // ---- sweep_score_kernel, one item per split
    std::vector<short> slot(H, -1);
    for (int r = 0; r < nv; r++) slot[hidx[r]] = (short)r;
    memset(counts_sweep, 0, sizeof(int) * nv);
    long long n_add = 0;
    g_exact_tests = 0;
    for (int p0 = 0; p0 < n; p0 += split) {
        const int np = (n - p0 < split) ? n - p0 : split;
        std::vector<int> diff(H, 0);
        for (int j = 0; j < J; j++) {
            const SweepCol &c = cols[j];
            if (c.m1 < c.m0) continue;
            int *d = diff.data() + j * M + c.m0;
            const short *sl = slot.data() + c.m0 * J + j;
            int base = 0;
            const int D = c.m1 - c.m0;
            auto add = [&](int *q, int v) { *q += v; n_add++; };
            for (int i = 0; i < np; i++) {
                const float *p = pts + (size_t)(p0 + i) * 3;
                const SweepRanges r = sweep_solve(c, p[0], p[1], p[2]);
                // the device kernel's bookkeeping (sweep_emit): unconditional adds, lanes without one aim at a scratch word
                const SweepEmit e = sweep_emit(r, D);
                if (e.add_lo) add(d + r.a, 1);
                if (e.add_hi) add(d + r.e + 1, -1);
                base += e.from_zero ? 1 : 0;
                // the device queues (point, column, packed steps) and drains the queue densely; the
                // arithmetic per queued step is sweep_packed_step + sweep_exact_step_col, as here
                const int cnt = e.uncertain ? sweep_packed_count(e.packed) : 0;
                for (int k = 0; k < cnt; k++)
                    sweep_exact_step_col(p[0], p[1], p[2], sweep_packed_step(e.packed, k), D, d, sl, J, prep, c.cosa, c.sina, c.tx, c.ty, add);
                stats[5] += (r.a <= r.e);
                stats[6] += e.uncertain ? 1 : 0;
                stats[2]++;
            }
            d[0] += base;
        }
        for (int j = 0; j < J; j++) {
            int run = 0;
            for (int m = 0; m < M; m++) { run += diff[j * M + m]; diff[j * M + m] = run; }
        }
        for (int h = 0; h < H; h++)
            if (slot[h] >= 0) counts_sweep[slot[h]] += diff[(h % J) * M + h / J];
    }
    stats[1] = n_add;
    // ---- brute force with the exact predicate
    for (int r = 0; r < nv; r++) {
        const BoxPrep b = load_prep(prep, r);
        int c = 0;
        for (int i = 0; i < n; i++) c += in_box(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], b) ? 1 : 0;
        counts_brute[r] = c;
    }
    stats[0] = g_exact_tests;
    return 0;
}
