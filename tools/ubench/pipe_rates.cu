// Micro-benchmark (run under gpurun): issue/pipe rates on B200 of the instructions the
// scoring loop is made of -- FADD/FMUL/FFMA, their packed f32x2 forms (FADD2/FMUL2/FFMA2),
// FSETP and predicated IADD/FADD.  Prints warp-instructions per clock per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 64
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a0, float b0)
{
    float a[8], c[8];
    uint64_t p[8];
    int cnt[8];
    for (int i = 0; i < 8; i++) { a[i] = a0 + i + threadIdx.x; c[i] = b0 * i; cnt[i] = 0;
        p[i] = ((uint64_t)__float_as_uint(a[i]) << 32) | __float_as_uint(c[i]); }
    uint64_t pb = ((uint64_t)__float_as_uint(b0) << 32) | __float_as_uint(a0);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < REP; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b0), "f"(c[i]));
                if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb));
                if (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b0));
                if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (MODE == 4) asm volatile("{.reg .pred q; setp.le.f32 q, %1, %2; @q add.s32 %0, %0, 1;}" : "+r"(cnt[i]) : "f"(a[i]), "f"(c[(i + r) & 7]));
                if (MODE == 5) asm volatile("{.reg .pred q; setp.le.f32 q, %1, %2; @q add.f32 %0, %0, %3;}" : "+f"(a[i]) : "f"(c[i]), "f"(c[(i + r) & 7]), "f"(b0));
                if (MODE == 6) asm volatile("{.reg .pred q; setp.le.f32 q, %1, %2; setp.le.and.f32 q, %2, %3, q; setp.leu.and.f32 q, %1, %3, q; @q add.s32 %0, %0, 1;}" : "+r"(cnt[i]) : "f"(a[i]), "f"(c[(i + r) & 7]), "f"(c[(i + r + 1) & 7]));
                if (MODE == 7) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
                if (MODE == 8) { // mix: 1 packed fma + 2 setp-chain(3)+iadd  (per 2 tests)
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb));
                    asm volatile("{.reg .pred q; setp.le.f32 q, %1, %2; @q add.s32 %0, %0, 1;}" : "+r"(cnt[i]) : "f"(a[i]), "f"(c[(i + r) & 7]));
                }
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + c[i] + cnt[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int instr_per_slot)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float *out;
    const int blocks = sms * 4, iters = 200;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    k<MODE><<<blocks, 256>>>(out, 10, 1.0f, 1.0f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 1.0f, 1.0000001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double winstr = (double)blocks * 8 /*warps*/ * iters * REP * 8 * instr_per_slot;
    double clk = ms * 1e-3 * khz * 1e3;
    printf("%-44s %8.3f ms  %6.3f warp-instr/clk/SMSP (at max clock %d MHz)\n", name, ms, winstr / clk / sms / 4, khz / 1000);
    cudaFree(out);
}

int main()
{
    run<0>("FFMA (3-reg)", 1);
    run<1>("FFMA2 (f32x2)", 1);
    run<2>("FADD", 1);
    run<3>("FADD2 (f32x2)", 1);
    run<7>("FMUL2 (f32x2)", 1);
    run<4>("FSETP + @p IADD", 2);
    run<5>("FSETP + @p FADD", 2);
    run<6>("3x FSETP chain + @p IADD", 4);
    run<8>("FFMA2 + FSETP + @p IADD", 3);
    return 0;
}
