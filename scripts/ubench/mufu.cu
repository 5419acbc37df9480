// Throughput of the individual MUFU (XU pipe) operations on sm_100a, and of the general pair law's mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
// 1024 threads per SM (8 warps per scheduler), 8 independent operations per iteration.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

// OP: 0 rsqrt, 1 ex2, 2 rcp, 3 sqrt, 4 lg2, 5 sin, 6 the general law's five (rsqrt sqrt rcp ex2 ex2),
// 7 five full-rate candidates (rsqrt rsqrt rsqrt ex2 ex2)
template <int OP>
__device__ __forceinline__ void op(float& y, float x, int k)
{
    const int o = (OP == 6) ? (k % 5 == 0 ? 0 : k % 5 == 1 ? 3 : k % 5 == 2 ? 2 : 1)
                : (OP == 7) ? (k % 5 < 3 ? 0 : 1) : OP;
    // dependent chains (y = op(y)) that stay finite: the loop cannot be hoisted
    if (o == 0) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(y));
    if (o == 1) asm volatile("{.reg .f32 t; neg.f32 t, %0; ex2.approx.ftz.f32 %0, t;}" : "+f"(y));
    if (o == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(y));
    if (o == 3) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(y));
    if (o == 4) asm volatile("{.reg .f32 t; lg2.approx.ftz.f32 t, %0; abs.f32 %0, t;}" : "+f"(y));
    if (o == 5) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(y));
}

template <int OP>
__global__ void __launch_bounds__(256) mufu(float* out, float seed)
{
    float x[10], y[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) { x[k] = 0.f; y[k] = seed + 0.01f * k + 1e-4f * threadIdx.x; }
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 10; ++k) op<OP>(y[k], x[k], k);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) s += y[k];
    if (s == 12345.678f) out[0] = s;
}

template <int OP>
void run(const char* name, int sms, double ghz)
{
    float* out;
    cudaMalloc(&out, 4);
    const int blocks = sms * 4;
    mufu<OP><<<blocks, 256>>>(out, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) mufu<OP><<<blocks, 256>>>(out, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 / 5 * ghz * 1e9;
    const double lane_ops = 10.0 * ITER * 32 * 32;              // per SM: 32 warps x 32 lanes
    printf("%-44s %6.2f results/clk/SM\n", name, lane_ops / clk);
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
    run<0>("rsqrt.approx.ftz.f32", sms, ghz);
    run<1>("ex2.approx.ftz.f32", sms, ghz);
    run<2>("rcp.approx.ftz.f32", sms, ghz);
    run<3>("sqrt.approx.ftz.f32", sms, ghz);
    run<4>("lg2.approx.ftz.f32", sms, ghz);
    run<5>("sin.approx.ftz.f32", sms, ghz);
    run<6>("general law mix (rsqrt sqrt rcp ex2 ex2)", sms, ghz);
    run<7>("rsqrt rsqrt rsqrt ex2 ex2", sms, ghz);
    return 0;
}
