// Does the XU pipe spend time on predicated-off lanes?  MUFU throughput with part of each warp masked.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_pred mufu_pred.cu && ./mufu_pred
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

// PATTERN: 0 all lanes, 1 lanes 0-15, 2 even lanes, 3 lanes 0-7, 4 lane 0 only, 5 no lane (warp-uniformly off)
template <int PATTERN>
__global__ void __launch_bounds__(256) mufu(float* out, float seed, int dummy)
{
    const int lane = threadIdx.x & 31;
    bool on = true;
    if (PATTERN == 1) on = lane < 16;
    if (PATTERN == 2) on = (lane & 1) == 0;
    if (PATTERN == 3) on = lane < 8;
    if (PATTERN == 4) on = lane == 0;
    if (PATTERN == 5) on = lane == dummy;          // dummy = 99: never
    float y[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) y[k] = seed + 0.01f * k + 1e-4f * threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < 10; ++k)
            asm volatile("{.reg .pred p; setp.ne.s32 p, %1, 0; @p rsqrt.approx.ftz.f32 %0, %0;}"
                         : "+f"(y[k]) : "r"((int)on));
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k) s += y[k];
    if (s == 12345.678f) out[0] = s;
}

template <int PATTERN>
void run(const char* name, int sms, double ghz)
{
    float* out;
    cudaMalloc(&out, 4);
    const int blocks = sms * 4;
    mufu<PATTERN><<<blocks, 256>>>(out, 1.0f, 99);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) mufu<PATTERN><<<blocks, 256>>>(out, 1.0f, 99);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double clk = ms * 1e-3 / 5 * ghz * 1e9;
    const double warp_inst = 10.0 * ITER * 32;                  // per SM
    printf("%-40s %6.3f warp-MUFU/clk/SM  (%5.2f clk per warp instruction per scheduler)\n", name,
           warp_inst / clk, 4.0 * clk / warp_inst);
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
    run<0>("all 32 lanes", sms, ghz);
    run<1>("lanes 0-15", sms, ghz);
    run<2>("even lanes", sms, ghz);
    run<3>("lanes 0-7", sms, ghz);
    run<4>("lane 0 only", sms, ghz);
    run<5>("no lane (predicate false everywhere)", sms, ghz);
    return 0;
}
