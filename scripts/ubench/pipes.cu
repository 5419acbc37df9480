// Micro-benchmarks of the sm_100a issue / FMA / XU pipes used to size the force kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Each kernel runs ITER iterations of a fixed instruction mix on independent register chains,
// 1024 threads per SM (8 warps per scheduler); reports warp-instructions per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

template <int NF, int NP, int NM, int NA>
__global__ void __launch_bounds__(256) mix(float* out, float seed)
{
    // NF scalar FFMA (3 distinct regs), NP packed FFMA2, NM MUFU (half rsqrt half ex2), NA alu (FMNMX)
    float a[8], b[8];
    unsigned long long p[8], q[8];
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = seed + k + threadIdx.x; b[k] = seed * 0.5f + k;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[k]) : "f"(a[k]), "f"(b[k]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(q[k]) : "f"(b[k]), "f"(a[k]));
        m[k] = 1.0f + 0.001f * (k + threadIdx.x);
    }
    float c0 = seed * 1.0001f, c1 = 0.999f;
    unsigned long long pc;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c1), "f"(c1));
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < NF; ++k)
            asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[k & 7]) : "f"(b[(k + 1) & 7]), "f"(b[(k + 3) & 7]));
#pragma unroll
        for (int k = 0; k < NP; ++k)
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[k & 7]) : "l"(q[(k + 1) & 7]), "l"(q[(k + 3) & 7]));
#pragma unroll
        for (int k = 0; k < NM; ++k) {
            if (k & 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[k & 7]));
            else asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[k & 7]));
        }
#pragma unroll
        for (int k = 0; k < NA; ++k)
            asm volatile("max.f32 %0, %0, %1;" : "+f"(b[k & 7]) : "f"(c0));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[k]));
        s += a[k] + b[k] + lo + hi + m[k];
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(q[k]));
        s += lo + hi;
    }
    if (s == 12345.678f) out[0] = s;
}

template <int NF, int NP, int NM, int NA>
void run(const char* name, int sms, double ghz)
{
    float* out;
    cudaMalloc(&out, 4);
    const int blocks = sms * 4;
    mix<NF, NP, NM, NA><<<blocks, 256>>>(out, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) mix<NF, NP, NM, NA><<<blocks, 256>>>(out, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 / 5 * ghz * 1e9;               // SM clocks per launch
    const double warps_per_sm = 4.0 * 8;                        // 4 blocks x 8 warps
    const double per_iter = clk / ITER;                         // clocks per loop iteration (all warps of an SM)
    const double inst = (NF + NP + NM + NA) * warps_per_sm;
    printf("%-34s NF=%2d NP=%2d NM=%2d NA=%2d  clk/iter/SM %8.2f  warp-inst/clk/SM %5.2f  "
           "fma-lane-ops/clk/SM %6.1f  mufu/clk/SM %5.2f\n", name, NF, NP, NM, NA, per_iter,
           inst / per_iter, (NF + 2.0 * NP) * warps_per_sm * 32 / per_iter, NM * warps_per_sm * 32 / per_iter);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s  SMs %d  clock %.3f GHz (max; results assume the GPU runs at it)\n", pr.name, pr.multiProcessorCount, ghz);
    const int sms = pr.multiProcessorCount;
    run<16, 0, 0, 0>("scalar FFMA 3-reg", sms, ghz);
    run<0, 16, 0, 0>("packed FFMA2 3-reg", sms, ghz);
    run<0, 0, 8, 0>("MUFU only", sms, ghz);
    run<0, 0, 0, 16>("FMNMX only", sms, ghz);
    run<16, 0, 0, 16>("FFMA + FMNMX 1:1", sms, ghz);
    run<0, 16, 0, 16>("FFMA2 + FMNMX 1:1", sms, ghz);
    run<8, 0, 2, 0>("FFMA 8 : MUFU 2", sms, ghz);
    run<12, 0, 2, 0>("FFMA 12 : MUFU 2", sms, ghz);
    run<14, 0, 2, 0>("FFMA 14 : MUFU 2", sms, ghz);
    run<16, 0, 2, 0>("FFMA 16 : MUFU 2", sms, ghz);
    run<0, 4, 2, 0>("FFMA2 4 : MUFU 2", sms, ghz);
    run<0, 6, 2, 0>("FFMA2 6 : MUFU 2", sms, ghz);
    run<0, 7, 2, 0>("FFMA2 7 : MUFU 2", sms, ghz);
    run<0, 8, 2, 0>("FFMA2 8 : MUFU 2", sms, ghz);
    run<0, 12, 4, 0>("FFMA2 12 : MUFU 4", sms, ghz);
    run<0, 12, 4, 4>("FFMA2 12 : MUFU 4 : ALU 4", sms, ghz);
    run<0, 14, 3, 4>("FFMA2 14 : MUFU 3 : ALU 4", sms, ghz);
    run<0, 6, 1, 0>("FFMA2 6 : MUFU 1", sms, ghz);
    run<0, 7, 1, 0>("FFMA2 7 : MUFU 1", sms, ghz);
    run<0, 8, 1, 0>("FFMA2 8 : MUFU 1", sms, ghz);
    return 0;
}
