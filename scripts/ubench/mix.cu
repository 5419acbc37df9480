// What an instruction mix like the pair law's can reach on sm_100a when nothing but the pipes limits it:
// independent register chains, operands that hit the reuse cache, W warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix mix.cu && ./mix
// Per loop iteration and thread: NM MUFU (rsqrt / ex2 alternating), NP packed FFMA2 (a = a * b + c, b and c
// shared), NA ALU-pipe operations (FMNMX).  Reports the busy fraction of the three pipes assuming 8 / 2 / 2
// cycles per warp instruction and scheduler (measured: mufu.cu, pipes.cu).
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048

template <int NM, int NP, int NA>
__global__ void __launch_bounds__(128) mix(float* out, float seed)
{
    unsigned long long p[8], pb, pc;
    float m[8], a[8];
    const float c0 = seed * 1.0001f;
    {
        const float b0 = 1.0000001f, b1 = 1e-7f;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b0));
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pc) : "f"(b1));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float v = seed + k + threadIdx.x;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(p[k]) : "f"(v));
        m[k] = 1.0f + 0.001f * (k + threadIdx.x);
        a[k] = v;
    }
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        // interleave the three kinds round-robin so that the stream looks like a scheduled kernel
        constexpr int N = (NM > NP ? (NM > NA ? NM : NA) : (NP > NA ? NP : NA));
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (k * NM / N != (k + 1) * NM / N || (NM == N)) {
                const int j = (k * NM / N) & 7;
                if (j & 1) asm volatile("{.reg .f32 t; neg.f32 t, %0; ex2.approx.ftz.f32 %0, t;}" : "+f"(m[j]));
                else asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[j]));
            }
            if (k * NP / N != (k + 1) * NP / N || (NP == N))
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[(k * NP / N) & 7]) : "l"(pb), "l"(pc));
            if (k * NA / N != (k + 1) * NA / N || (NA == N))
                asm volatile("max.f32 %0, %0, %1;" : "+f"(a[(k * NA / N) & 7]) : "f"(c0));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[k]));
        s += lo + hi + m[k] + a[k];
    }
    if (s == 12345.678f) out[0] = s;
}

template <int NM, int NP, int NA>
void run(const char* name, int sms, double ghz, int warps_per_sched)
{
    float* out;
    cudaMalloc(&out, 4);
    const int blocks = sms * warps_per_sched;                   // 128-thread blocks: 1 warp per scheduler each
    mix<NM, NP, NA><<<blocks, 128>>>(out, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) mix<NM, NP, NA><<<blocks, 128>>>(out, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 / 5 * ghz * 1e9;
    const double per = clk / ((double)ITER * warps_per_sched);  // cycles per iteration of one warp, per scheduler
    printf("%-30s W=%d  %7.1f clk/iter/sched   XU %5.1f %%  FMA %5.1f %%  ALU %5.1f %%  issue %5.1f %%\n", name,
           warps_per_sched, per, 100 * 8.0 * NM / per, 100 * 2.0 * NP / per, 100 * 2.0 * NA / per,
           100.0 * (NM + NP + NA) / per);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s  SMs %d  clock %.3f GHz\n", pr.name, pr.multiProcessorCount, ghz);
    const int sms = pr.multiProcessorCount;
    for (int w = 4; w <= 8; w += 4) {
        run<8, 0, 0>("MUFU only", sms, ghz, w);
        run<0, 16, 0>("FFMA2 only", sms, ghz, w);
        run<0, 0, 16>("FMNMX only", sms, ghz, w);
        run<4, 13, 0>("far law 4 : 13 : 0", sms, ghz, w);
        run<4, 14, 2>("far law + overhead 4 : 14 : 2", sms, ghz, w);
        run<10, 26, 0>("general 10 : 26 : 0", sms, ghz, w);
        run<10, 26, 24>("general law 10 : 26 : 24", sms, ghz, w);
        run<10, 26, 12>("general 10 : 26 : 12", sms, ghz, w);
        run<8, 26, 24>("general, 4 MUFU 8 : 26 : 24", sms, ghz, w);
        run<10, 20, 24>("general 10 : 20 : 24", sms, ghz, w);
    }
    return 0;
}
