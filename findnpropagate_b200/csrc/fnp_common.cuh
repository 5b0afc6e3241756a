// Shared device helpers of libfnp_sm100.so.  sm_100a only; compiled with -fmad=false so
// that every fused multiply-add in this library is one that is written out explicitly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fnp.h"
#include "fnp_sweep.cuh"   // BoxPrep, in_box (shared with the host model of the sweep)

#define FNP_LAUNCH_CHECK()                         \
    do {                                           \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

namespace fnp {

__host__ __device__ inline int divup(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float strict_lt_threshold(float dim)
{
    const double t = (double)dim * 0.5 + (double)1e-5f;
    if (!(t > 0.0)) return (t != t) ? __int_as_float(0x7fc00000) : -1.0f;
    float f = __double2float_rd(t);
    if ((double)f == t) f = __int_as_float(__float_as_int(f) - 1);  // t > 0 so f > 0 here
    return f;
}

__device__ __forceinline__ BoxPrep prep_box(const float *__restrict__ b)
{
    BoxPrep p;
    p.cx = b[0];
    p.cy = b[1];
    p.cz = b[2];
    p.hz = __double2float_rd((double)b[5] * 0.5);
    const float ang = -b[6];
    p.cosa = cosf(ang);
    p.sina = sinf(ang);
    p.tx = strict_lt_threshold(b[3]);
    p.ty = strict_lt_threshold(b[4]);
    return p;
}

// cnt += p as ONE predicated IADD (the C form compiles to add + predicated move)
__device__ __forceinline__ void count_if(int &cnt, bool p)
{
    asm("{\n .reg .pred q;\n setp.ne.s32 q, %1, 0;\n @q add.s32 %0, %0, 1;\n}" : "+r"(cnt) : "r"((int)p));
}

// ---------------------------------------------------------------------------------------
// Small math with a fixed evaluation order (twin: oracle/fnp_oracle.c)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float dot3(const float *__restrict__ a, float x, float y, float z)
{
    return __fmaf_rn(a[2], z, __fmaf_rn(a[1], y, __fmul_rn(a[0], x)));
}

// 3-term dot product in the order torch's batched (L,3,3)@(L,3,1) matmul uses on B200
// (measured, tools/probe_gpu.py): rn(fma(a1,y, rn(a0*x)) + rn(a2*z)).
__device__ __forceinline__ float dot3_bmm(const float *__restrict__ a, float x, float y, float z)
{
    return __fadd_rn(__fmaf_rn(a[1], y, __fmul_rn(a[0], x)), __fmul_rn(a[2], z));
}

__device__ __forceinline__ float norm3(float x, float y, float z)
{
    return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
}

// fma-only exp for x <= 0 (softmax of negative distances); <= 1 ulp.
__device__ __forceinline__ float fnp_exp(float x)
{
    if (!(x > -87.0f)) return (x != x) ? x : 0.0f;
    if (x > 88.0f) return __int_as_float(0x7f800000);
    const float n = rintf(__fmul_rn(x, 1.44269502162933349609375f));
    float r = __fmaf_rn(n, -0.693145751953125f, x);
    r = __fmaf_rn(n, -1.428606765330187045e-06f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float r2 = __fmul_rn(r, r);
    const float e = __fadd_rn(__fmaf_rn(p, r2, r), 1.0f);
    const int ni = (int)n;
    return __fmul_rn(e, __int_as_float((ni + 127) << 23));
}

// axis-aligned BEV IoU (iou_normal): Sa + Sb is one fma, x -+ dx/2 are exact-half fmas
__device__ __forceinline__ float iou_normal(const float *__restrict__ a, const float *__restrict__ b)
{
    // a, b: x, y, (z), dx, dy at [0], [1], [3], [4] (iou3d_nms_kernel.cu:327-338)
    const float left = fmaxf(__fmaf_rn(a[3], -0.5f, a[0]), __fmaf_rn(b[3], -0.5f, b[0]));
    const float right = fminf(__fmaf_rn(a[3], 0.5f, a[0]), __fmaf_rn(b[3], 0.5f, b[0]));
    const float top = fmaxf(__fmaf_rn(a[4], -0.5f, a[1]), __fmaf_rn(b[4], -0.5f, b[1]));
    const float bottom = fminf(__fmaf_rn(a[4], 0.5f, a[1]), __fmaf_rn(b[4], 0.5f, b[1]));
    const float width = fmaxf(__fsub_rn(right, left), 0.f), height = fmaxf(__fsub_rn(bottom, top), 0.f);
    const float inter = __fmul_rn(width, height);
    const float sasb = __fmaf_rn(b[3], b[4], __fmul_rn(a[3], a[4]));
    return __fdiv_rn(inter, fmaxf(__fsub_rn(sasb, inter), 1e-8f));
}

// LiDAR -> image (frustum_proposals_v1.py:1431-1475 without augmentation).  L = rows 0..2 of
// lidar2image, 12 floats row-major [r0c0 r0c1 r0c2 r0c3 | r1.. | r2..].
__device__ __forceinline__ bool project(const float *__restrict__ L, float x, float y, float z,
                                        float img_w, float img_h, float &u, float &v, float &d)
{
    const float wx = __fadd_rn(dot3(L + 0, x, y, z), L[3]);
    const float wy = __fadd_rn(dot3(L + 4, x, y, z), L[7]);
    const float wz = __fadd_rn(dot3(L + 8, x, y, z), L[11]);
    d = fminf(fmaxf(wz, 1e-5f), 1e5f);
    u = __fdiv_rn(wx, d);
    v = __fdiv_rn(wy, d);
    return (v < img_h) & (v >= 0.f) & (u < img_w) & (u >= 0.f);
}

// image (u,v,d) -> LiDAR (frustum_proposals_v1.py:1509-1545): combine (9) then trans (3).
__device__ __forceinline__ void unproject(const float *__restrict__ C, const float *__restrict__ t,
                                          float u, float v, float d, float &x, float &y, float &z)
{
    const float px = __fmul_rn(u, d), py = __fmul_rn(v, d);
    x = __fadd_rn(dot3_bmm(C + 0, px, py, d), t[0]);
    y = __fadd_rn(dot3_bmm(C + 3, px, py, d), t[1]);
    z = __fadd_rn(dot3_bmm(C + 6, px, py, d), t[2]);
}

// ---------------------------------------------------------------------------------------
// KITTI calibration (pcdet/utils/calibration_kitti.py:128-216, CalibrationTorch) for FNP_VARIANT_KITTI.
// K = the 48 floats of the frame (include/fnp.h): M1 (4,3) | P2T (4,3) | cu cv fu fv tx ty | Minv (4,4).
// [x y z 1] @ M (4,C), column c: torch's matmul accumulates k = 0..3 as one fma chain on B200 for >= 33 rows,
// and rounds every product on its own below that (tools/probe_kitti.py); `chain` selects the order.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4h(const float *__restrict__ M, const int c, const int C, const float x, const float y,
                                       const float z, const bool chain = true)
{
    if (chain) {
        float acc = __fmul_rn(x, M[c]);
        acc = __fmaf_rn(y, M[C + c], acc);
        acc = __fmaf_rn(z, M[2 * C + c], acc);
        return __fadd_rn(acc, M[3 * C + c]);                      // fma(1, m, acc)
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, M[c]), __fmul_rn(y, M[C + c])), __fmul_rn(z, M[2 * C + c])), M[3 * C + c]);
}

// lidar_to_img (:196-203): rect = [p 1] @ M1, hom = [rect 1] @ P2T, (u, v) = hom.xy / rect.z, depth = hom.z - P2T[3][2].
// No clamp and no on-image test in this head (frustum_proposals_v1_kitti.py:693-700).
__device__ __forceinline__ void project_kitti(const float *__restrict__ K, const float x, const float y, const float z,
                                              float &u, float &v, float &d)
{
    const float r0 = dot4h(K, 0, 3, x, y, z), r1 = dot4h(K, 1, 3, x, y, z), r2 = dot4h(K, 2, 3, x, y, z);
    const float h0 = dot4h(K + 12, 0, 3, r0, r1, r2), h1 = dot4h(K + 12, 1, 3, r0, r1, r2), h2 = dot4h(K + 12, 2, 3, r0, r1, r2);
    u = __fdiv_rn(h0, r2);
    v = __fdiv_rn(h1, r2);
    d = __fsub_rn(h2, K[12 + 9 + 2]);
}

// img_to_rect (:205-216) then rect_to_lidar (:151-169): x = ((u - cu) d) / fu + tx, y likewise, [x y d 1] @ Minv
__device__ __forceinline__ void unproject_kitti(const float *__restrict__ K, const float u, const float v, const float d,
                                                float &x, float &y, float &z, const bool chain = true)
{
    const float xr = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(u, K[24]), d), K[26]), K[28]);
    const float yr = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(v, K[25]), d), K[27]), K[29]);
    x = dot4h(K + 32, 0, 4, xr, yr, d, chain);
    y = dot4h(K + 32, 1, 4, xr, yr, d, chain);
    z = dot4h(K + 32, 2, 4, xr, yr, d, chain);
}

// ---------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS: UBLKCP) helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion signalled on `bar` (bytes % 16 == 0, 16 B aligned)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace fnp
