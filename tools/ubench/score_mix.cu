// Micro-benchmark (run under gpurun): the ceiling of the scoring loop's instruction mix on B200.
// Each thread keeps K=4 hypotheses in registers and tests them against point pairs read with
// broadcast LDS.128 from a shared-memory tile that never changes (no TMA, no barriers, no
// global traffic): what the in-box predicate costs when nothing but the SM pipes is in the way.
//   V0: 7 packed FP + 6 FSETP + 2 @p IADD per two tests (the shipped count_pair)
//   V1: 7 packed FP + 4 FSETP + 2 FSET.BF + 1 FADD2 (float-pair accumulator)
//   V2: V0 and V1 alternating by hypothesis (pipe balance)
// Prints point-box tests per clock per SM and the implied chip-wide tests/s at the max clock.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Hyp { unsigned long long cx2, cy2, cz2, cosa2, nsina2, sina2; float hz, tx, ty; };

__device__ __forceinline__ unsigned long long dup2(float v)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void pair_v0(int &cnt, unsigned long long xx, unsigned long long yy, unsigned long long zz, const Hyp &h)
{
    asm("{\n .reg .b64 sx, sy, sz, m1, m2, lx, ly;\n .reg .f32 a0, a1, b0, b1, c0, c1;\n .reg .pred p, q;\n"
        " sub.rn.f32x2 sx, %1, %4;\n sub.rn.f32x2 sy, %2, %5;\n sub.rn.f32x2 sz, %3, %6;\n"
        " mul.rn.f32x2 m1, sy, %8;\n mul.rn.f32x2 m2, sx, %9;\n"
        " fma.rn.f32x2 lx, sx, %7, m1;\n fma.rn.f32x2 ly, sy, %7, m2;\n"
        " mov.b64 {a0, a1}, lx;\n mov.b64 {b0, b1}, ly;\n mov.b64 {c0, c1}, sz;\n"
        " abs.f32 a0, a0;\n abs.f32 a1, a1;\n abs.f32 b0, b0;\n abs.f32 b1, b1;\n abs.f32 c0, c0;\n abs.f32 c1, c1;\n"
        " setp.leu.f32 p, c0, %10;\n setp.le.and.f32 p, a0, %11, p;\n setp.le.and.f32 p, b0, %12, p;\n"
        " setp.leu.f32 q, c1, %10;\n setp.le.and.f32 q, a1, %11, q;\n setp.le.and.f32 q, b1, %12, q;\n"
        " @p add.s32 %0, %0, 1;\n @q add.s32 %0, %0, 1;\n}\n"
        : "+r"(cnt)
        : "l"(xx), "l"(yy), "l"(zz), "l"(h.cx2), "l"(h.cy2), "l"(h.cz2), "l"(h.cosa2), "l"(h.nsina2), "l"(h.sina2),
          "f"(h.hz), "f"(h.tx), "f"(h.ty));
}

__device__ __forceinline__ void pair_v1(unsigned long long &acc, unsigned long long xx, unsigned long long yy, unsigned long long zz, const Hyp &h)
{
    asm("{\n .reg .b64 sx, sy, sz, m1, m2, lx, ly, d;\n .reg .f32 a0, a1, b0, b1, c0, c1, f0, f1;\n .reg .pred p, q;\n"
        " sub.rn.f32x2 sx, %1, %4;\n sub.rn.f32x2 sy, %2, %5;\n sub.rn.f32x2 sz, %3, %6;\n"
        " mul.rn.f32x2 m1, sy, %8;\n mul.rn.f32x2 m2, sx, %9;\n"
        " fma.rn.f32x2 lx, sx, %7, m1;\n fma.rn.f32x2 ly, sy, %7, m2;\n"
        " mov.b64 {a0, a1}, lx;\n mov.b64 {b0, b1}, ly;\n mov.b64 {c0, c1}, sz;\n"
        " abs.f32 a0, a0;\n abs.f32 a1, a1;\n abs.f32 b0, b0;\n abs.f32 b1, b1;\n abs.f32 c0, c0;\n abs.f32 c1, c1;\n"
        " setp.leu.f32 p, c0, %10;\n setp.le.and.f32 p, a0, %11, p;\n set.le.and.f32.f32 f0, b0, %12, p;\n"
        " setp.leu.f32 q, c1, %10;\n setp.le.and.f32 q, a1, %11, q;\n set.le.and.f32.f32 f1, b1, %12, q;\n"
        " mov.b64 d, {f0, f1};\n add.rn.f32x2 %0, %0, d;\n}\n"
        : "+l"(acc)
        : "l"(xx), "l"(yy), "l"(zz), "l"(h.cx2), "l"(h.cy2), "l"(h.cz2), "l"(h.cosa2), "l"(h.nsina2), "l"(h.sina2),
          "f"(h.hz), "f"(h.tx), "f"(h.ty));
}

constexpr int K = 4, TILE_PAIRS = 256;

template <int V>
__global__ void __launch_bounds__(128) kern(float *out, int iters, float seed, const float *__restrict__ prm)
{
    __shared__ __align__(16) float4 tile[TILE_PAIRS][2];
    for (int i = threadIdx.x; i < TILE_PAIRS; i += 128) {
        tile[i][0] = make_float4(seed * i, seed * i + 1.f, 0.5f * i, 0.25f * i);
        tile[i][1] = make_float4(0.1f * i, 0.2f * i, 1.f, 1.f);
    }
    __syncthreads();
    Hyp hp[K];
    int cnt[K];
    unsigned long long acc[K];
    for (int k = 0; k < K; k++) {
        // every hypothesis parameter is a run-time value in its own register, as in the product kernel
        const float *q = prm + ((threadIdx.x + 128 * k) & 1023) * 8;
        hp[k].cx2 = dup2(q[0]); hp[k].cy2 = dup2(q[1]); hp[k].cz2 = dup2(q[2]);
        hp[k].cosa2 = dup2(q[4]); hp[k].nsina2 = dup2(-q[5]); hp[k].sina2 = dup2(q[5]);
        hp[k].hz = q[3]; hp[k].tx = q[6]; hp[k].ty = q[7];
        cnt[k] = 0; acc[k] = 0ull;
    }
    const ulonglong2 *tp = reinterpret_cast<const ulonglong2 *>(tile);
    for (int it = 0; it < iters; it++) {
        for (int i = 0; i + 2 <= TILE_PAIRS; i += 2) {
            const ulonglong2 xy0 = tp[2 * i], zd0 = tp[2 * i + 1];
            const ulonglong2 xy1 = tp[2 * i + 2], zd1 = tp[2 * i + 3];
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (V == 0 || (V == 2 && (k & 1) == 0)) {
                    pair_v0(cnt[k], xy0.x, xy0.y, zd0.x, hp[k]);
                    pair_v0(cnt[k], xy1.x, xy1.y, zd1.x, hp[k]);
                } else {
                    pair_v1(acc[k], xy0.x, xy0.y, zd0.x, hp[k]);
                    pair_v1(acc[k], xy1.x, xy1.y, zd1.x, hp[k]);
                }
            }
        }
    }
    float s = 0.f;
    for (int k = 0; k < K; k++) s += cnt[k] + __uint_as_float((unsigned)acc[k]) + __uint_as_float((unsigned)(acc[k] >> 32));
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int V>
void run(const char *name, int ctas_per_sm)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out, *prm;
    const int blocks = sms * ctas_per_sm, iters = 400;
    cudaMalloc(&out, blocks * 128 * sizeof(float));
    cudaMalloc(&prm, 1024 * 8 * sizeof(float));
    {
        float h[1024 * 8];
        for (int i = 0; i < 1024; i++) {
            h[i * 8 + 0] = 0.37f * i; h[i * 8 + 1] = 0.11f * i; h[i * 8 + 2] = 0.01f * i; h[i * 8 + 3] = 2.f + 0.001f * i;
            h[i * 8 + 4] = 0.8f; h[i * 8 + 5] = 0.6f; h[i * 8 + 6] = 40.f + 0.01f * i; h[i * 8 + 7] = 20.f + 0.01f * i;
        }
        cudaMemcpy(prm, h, sizeof(h), cudaMemcpyHostToDevice);
    }
    kern<V><<<blocks, 128>>>(out, 4, 0.37f, prm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<V><<<blocks, 128>>>(out, iters, 0.37f, prm);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tests = (double)blocks * 128 * K * iters * TILE_PAIRS * 2;
    const double clk = ms * 1e-3 * khz * 1e3;
    printf("%-34s ctas/SM %d  %8.3f ms  %7.2f tests/clk/SM  %.3e tests/s (clock %d MHz assumed)\n", name, ctas_per_sm, ms,
           tests / clk / sms, tests / (ms * 1e-3), khz / 1000);
    cudaFree(out);
    cudaFree(prm);
}

int main()
{
    for (int c = 4; c <= 8; c += 2) {
        run<0>("V0 setp x6 + @p iadd x2", c);
        run<1>("V1 setp x4 + set x2 + add.f32x2", c);
        run<2>("V2 alternate V0/V1 by hypothesis", c);
    }
    return 0;
}
