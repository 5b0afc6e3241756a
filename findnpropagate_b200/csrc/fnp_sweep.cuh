// Depth-sweep scoring (stage 2b, FNP_SCORE_SWEEP): the arithmetic shared by the device kernels
// in fnp_seeker.cu and by the host model tools/sweep_model.cu, which runs the same functions on
// the CPU to check the range logic without a GPU.  Every operation is a single IEEE fp32
// operation on both sides (the library is built with -fmad=false; the host model with
// -ffp-contract=off), so the two agree bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/fnp.h"

#if defined(__CUDACC__)
#define FNP_HD __host__ __device__ __forceinline__
#else
#define FNP_HD inline
#endif

namespace fnp {

#if defined(__CUDA_ARCH__)
FNP_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
FNP_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
FNP_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
FNP_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
FNP_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
FNP_HD int f2i_up(float a) { return __float2int_ru(a); }
FNP_HD int f2i_down(float a) { return __float2int_rd(a); }
FNP_HD float4 ld4(const float4 *p) { return __ldg(p); }
#else
FNP_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
FNP_HD float f_sub(float a, float b) { volatile float r = a - b; return r; }
FNP_HD float f_mul(float a, float b) { volatile float r = a * b; return r; }
FNP_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
FNP_HD float f_div(float a, float b) { volatile float r = a / b; return r; }
FNP_HD int f2i_sat(float a) { return a >= 2147483648.f ? 0x7fffffff : a <= -2147483648.f ? (int)0x80000000 : (int)a; }
FNP_HD int f2i_up(float a) { return f2i_sat(ceilf(a)); }      // saturating like cvt.rpi.s32.f32
FNP_HD int f2i_down(float a) { return f2i_sat(floorf(a)); }
FNP_HD float4 ld4(const float4 *p) { return *p; }
#endif

// ---------------------------------------------------------------------------------------
// In-box predicate.  Reference arithmetic (roiaware_pool3d_kernel.cu:16-36 as compiled for
// sm_100a): z test and x/y tests are evaluated in fp64,
//     in_z  = !((double)|z-cz| > (double)dz * 0.5)
//     in_xy = (double)|lx| < (double)dx*0.5 + (double)1e-5f   (same for y)
// with lx = fma(sx, cosa, rn(sy * -sina)), ly = fma(sy, cosa, rn(sx * sina)),
// cosa = cosf(-rz), sina = sinf(-rz).  The fp64 compares are hoisted exactly into fp32
// thresholds per box: |l| < t  <=>  |l| <= pred(t), pred(t) = largest float strictly below t.
// ---------------------------------------------------------------------------------------
struct BoxPrep {
    float cx, cy, cz, hz;      // centre, half height threshold
    float cosa, sina, tx, ty;  // rotation by -heading, strict-less thresholds as <=
};

FNP_HD bool in_box(float x, float y, float z, const BoxPrep &p)
{
    const float sz = f_sub(z, p.cz);
    const float sx = f_sub(x, p.cx);
    const float sy = f_sub(y, p.cy);
    const float lx = f_fma(sx, p.cosa, f_mul(sy, -p.sina));
    const float ly = f_fma(sy, p.cosa, f_mul(sx, p.sina));
    return !(fabsf(sz) > p.hz) && (fabsf(lx) <= p.tx) && (fabsf(ly) <= p.ty);
}

FNP_HD BoxPrep load_prep(const float *hyp_prep, size_t idx)
{
    const float4 *src = reinterpret_cast<const float4 *>(hyp_prep + idx * 8);
    const float4 a = ld4(src), c = ld4(src + 1);
    BoxPrep p;
    p.cx = a.x; p.cy = a.y; p.cz = a.z; p.hz = a.w;
    p.cosa = c.x; p.sina = c.y; p.tx = c.z; p.ty = c.w;
    return p;
}

// The H = M*J hypotheses of a frustum form J columns (one per yaw x size entry of the prior
// table); the M hypotheses of a column share cosa, sina, tx, ty, hz and differ only in their
// centre, which advances along the centre line (plus the slowly varying front shift).  In the
// column's rotated frame,  lx = U - Cu(m),  ly = V - Cv(m),  sz = z - Cz(m)  with
//     U = x cosa - y sina,  V = y cosa + x sina,   Cu(m) = cx_m cosa - cy_m sina, ...
// so a point is inside hypothesis m iff Cu(m), Cv(m), Cz(m) fall into three intervals around
// (U, V, z).  Each C.(m) is a straight line in m up to a measured deviation:
//     |C(m) - (C0 + s (m - m0))| <= delta      for every VALID m of the column,
// with (C0, s) through the first and last valid depth step m0, m1 and delta taken over the actual
// centres (sweep_prep_kernel).  With eps bounding the fp32 rounding of both the exact predicate
// and this solve, and dl = delta + eps,
//     |P - C0 - s dm| <= t - dl   =>  inside   on that axis (definitely),
//     |P - C0 - s dm| >  t + dl   =>  outside  on that axis (definitely),
// which are two nested ranges of dm = m - m0 per axis; intersected over the three axes they give
// a DEFINITE range [a, e] and a POSSIBLE range [A, B] containing it.  The definite range goes into
// the column's difference array (+1 at a, -1 at e+1), the at most few depth steps of
// [A, B] \ [a, e] take the exact predicate in_box() against the hypothesis itself, and a prefix
// sum over m yields the counts.  The counts are the same integers the direct kernel produces: the
// only approximate quantity, the range ends, is used with a margin that covers its error, and
// everything inside the margin is decided by the exact predicate.
//
// A column in which t - dl < 0 has no definite range and degrades to exact tests of the possible
// range, so the result is correct for any geometry and fast for the seeker's.
struct SweepCol {
    float cosa, sina, tx, ty;   // rotation and x/y thresholds shared by the column's hypotheses (one 16-byte load:
                                // with the centre + hz of a hypothesis they are its whole BoxPrep)
    int m0, m1;          // first / last valid depth step of the column (m0 > m1: none)
    int pseudo_mask;     // bit k: axis k does not travel; it carries a pseudo slope (statistics only)
    float eps;           // rounding bound used for the x/y axes (statistics only)
    float c0[3];         // Cu, Cv, Cz at m0
    float inv_s[3];      // 1 / slope per depth step
    float w_in[3];       // (t - dl) |inv_s|  (depth steps; -inf when t < dl)
    float w_p[3];        // (t + dl) |inv_s|
};
static_assert(sizeof(SweepCol) == FNP_SWEEP_COL_FLOATS * 4, "SweepCol layout is part of the ABI workspace size");

FNP_HD void sweep_axes(const BoxPrep &p, float C[3])
{
    C[0] = f_fma(p.cx, p.cosa, f_mul(p.cy, -p.sina));
    C[1] = f_fma(p.cy, p.cosa, f_mul(p.cx, p.sina));
    C[2] = p.cz;
}

// eps: bound of the accumulated fp32 rounding of the exact predicate and of the range solve.
// Operands reach 2 maxabs (x - cx); both sides together stay below 12 ulp of that magnitude
// (DESIGN.md section 4); eps = 2 maxabs 2^-18 = 32 ulp leaves a factor > 2.5.
FNP_HD float sweep_eps(float maxabs)
{
    const float scale = fmaxf(f_mul(2.f, maxabs), 64.f);
    return f_mul(scale, 3.814697265625e-06f);
}
// The z axis is not rotated: its predicate |rn(z - cz)| <= hz and its solve only see z-magnitudes
// (a few metres), so its rounding bound is taken from the largest |z|, |cz| alone: < 4 ulp of
// 2 maxabs_z on both sides together, eps_z = 2 maxabs_z 2^-18 leaves a factor 4.  This matters
// because the z slope is the smallest (~1 mm per depth step): eps / |slope| is the width, in
// depth steps, of the band that takes exact predicates.
FNP_HD float sweep_eps_z(float maxabs_z)
{
    const float scale = fmaxf(f_mul(2.f, maxabs_z), 1.f);
    return f_mul(scale, 3.814697265625e-06f);
}

// Column parameters from the line fit: c0 = C at the first valid step, slope per step, dev_pos /
// dev_neg = max of C(m) - line(m) and of line(m) - C(m) over the valid steps (both >= 0), p = any
// hypothesis of the column (rotation, size).  The line is moved to the middle of the band the
// centres occupy (the front shift bends every column the same way, so the deviations are mostly
// one-sided and this halves delta).
// An axis whose travel |s| (m1 - m0) is below 4 eps gets the pseudo slope eps / (D + 1) instead
// (the line then strays from the fitted one by at most travel + eps, which is added to dl), so
// that the range solve needs no special case for constant axes.
FNP_HD SweepCol sweep_col_build(int m0, int m1, const float c0[3], const float slope[3], const float dev_pos[3],
                                const float dev_neg[3], const BoxPrep &p, float eps_xy, float eps_z)
{
    const float INF = INFINITY;
    SweepCol c;
    c.m0 = m0; c.m1 = m1;
    c.pseudo_mask = 0;
    c.eps = eps_xy;
    c.cosa = p.cosa; c.sina = p.sina; c.tx = p.tx; c.ty = p.ty;
    const float t[3] = {p.tx, p.ty, p.hz};
    const float span = (float)(m1 - m0);
    for (int k = 0; k < 3; k++) {
        const float eps = k == 2 ? eps_z : eps_xy;
        float s = slope[k];
        float dl = f_add(f_mul(0.5f, f_add(dev_pos[k], dev_neg[k])), eps);
        c.c0[k] = f_add(c0[k], f_mul(0.5f, f_sub(dev_pos[k], dev_neg[k])));
        const float travel = f_mul(fabsf(s), span);
        if (!(travel > f_mul(4.f, eps))) {
            c.pseudo_mask |= 1 << k;
            s = f_div(eps, f_add(span, 1.f));
            dl = f_add(f_add(dl, travel), eps);
        }
        const float inv = f_div(1.f, s);
        c.inv_s[k] = inv;
        const float win = f_sub(t[k], dl);
        c.w_in[k] = win >= 0.f ? f_mul(win, fabsf(inv)) : -INF;
        c.w_p[k] = f_mul(f_add(t[k], dl), fabsf(inv));
    }
    return c;
}

// Ranges of dm = m - m0 for one point against one column: definite [a, e], possible [A, B]
// (empty when the first bound exceeds the second); [a, e] lies inside [A, B], both inside [0, D].
struct SweepRanges {
    int a, e, A, B;
};

FNP_HD SweepRanges sweep_solve(const SweepCol &c, const float x, const float y, const float z)
{
    const float Df = (float)(c.m1 - c.m0);
    const float P[3] = {f_fma(x, c.cosa, f_mul(y, -c.sina)), f_fma(y, c.cosa, f_mul(x, c.sina)), z};
    float lo = 0.f, hi = Df, plo = 0.f, phi = Df;      // already clamped to the column's depth steps
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float q = f_mul(f_sub(P[k], c.c0[k]), c.inv_s[k]);
        lo = fmaxf(lo, f_sub(q, c.w_in[k])); hi = fminf(hi, f_add(q, c.w_in[k]));
        plo = fmaxf(plo, f_sub(q, c.w_p[k])); phi = fminf(phi, f_add(q, c.w_p[k]));
    }
    SweepRanges r;
    r.a = f2i_up(lo);      // float -> int saturates: an empty range stays empty
    r.e = f2i_down(hi);
    r.A = f2i_up(plo);
    r.B = f2i_down(phi);
    return r;
}

// ---------------------------------------------------------------------------------------
// The bookkeeping after sweep_solve, branch-free (sweep_score_kernel since round 2, session 3).
//   definite range [a, e] (if a <= e):  +1 at a unless a == 0 (those are summed per warp), -1 at e + 1 unless e == D;
//   uncertain steps: [A, a2 - 1] and [e2 + 1, B] with (a2, e2) = (a, e) when there is a definite range and
//   (B + 1, B) otherwise -- one formula for both cases; packed as four bytes A | a2 << 8 | e2 << 16 | B << 24
//   (all in [0, 255]: the sweep mode needs M <= 255, and a2 <= B + 1 <= D + 1 <= M).
// ---------------------------------------------------------------------------------------
struct SweepEmit {
    bool add_lo, add_hi, from_zero, uncertain;   // +1 at a; -1 at e + 1; range starts at step 0; has uncertain steps
    unsigned packed;                              // valid when uncertain
};

FNP_HD SweepEmit sweep_emit(const SweepRanges &r, const int D)
{
    SweepEmit o;
    const bool def = r.a <= r.e;
    o.add_lo = def && r.a != 0;
    o.add_hi = def && r.e < D;
    o.from_zero = def && r.a == 0;
    const unsigned a2 = def ? (unsigned)r.a : (unsigned)r.B + 1u, e2 = def ? (unsigned)r.e : (unsigned)r.B;
    const unsigned n_unc = (a2 - (unsigned)r.A) + ((unsigned)r.B - e2);
    o.uncertain = (r.A <= r.B) && n_unc != 0u;
    o.packed = (unsigned)r.A | (a2 << 8) | (e2 << 16) | ((unsigned)r.B << 24);
    return o;
}
FNP_HD int sweep_packed_count(unsigned w) { return (int)(((w >> 8) & 0xffu) - (w & 0xffu)) + (int)((w >> 24) - ((w >> 16) & 0xffu)); }
// k-th uncertain step (0 <= k < count) of a packed word
FNP_HD int sweep_packed_step(unsigned w, int k)
{
    const int A = (int)(w & 0xffu), n1 = (int)((w >> 8) & 0xffu) - A;
    return k < n1 ? A + k : (int)((w >> 16) & 0xffu) + 1 + (k - n1);
}

// Exact predicate of point (x, y, z) against the hypothesis at depth step m0 + dm of column j, the column's shared
// parameters supplied by the caller (cosa, sina, tx, ty are equal, bit for
// bit, for all hypotheses of a column: prep_box computes them from the same row of the prior table), so that
// only the 16 bytes {cx, cy, cz, hz} of the hypothesis are loaded.
// `slot_col` = slot table of the column (slot_col[dm * J], -1: not a valid hypothesis), diff = the
// column's difference array indexed by dm, D = m1 - m0.  add(ptr, v): *ptr += v.
template <class Add>
FNP_HD void sweep_exact_step_col(const float x, const float y, const float z, const int dm, const int D, int *diff,
                                 const short *slot_col, const int J, const float *prep_f, const float cosa,
                                 const float sina, const float tx, const float ty, Add add)
{
    const int r = slot_col[dm * J];
    if (r < 0) return;
#ifdef FNP_SWEEP_MODEL
    g_exact_tests++;
#endif
    const float4 a = ld4(reinterpret_cast<const float4 *>(prep_f + (size_t)r * 8));
    BoxPrep p;
    p.cx = a.x; p.cy = a.y; p.cz = a.z; p.hz = a.w;
    p.cosa = cosa; p.sina = sina; p.tx = tx; p.ty = ty;
    if (in_box(x, y, z, p)) {
        add(diff + dm, 1);
        if (dm < D) add(diff + dm + 1, -1);
    }
}

}  // namespace fnp
