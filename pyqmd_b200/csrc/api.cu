// api.cu -- library plumbing, the FP32 peak microbenchmark and the reference-shaped host-buffer
// entry points of libpyqmd_b200.so (see include/pyqmd_b200.h).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace pyqmd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

// ---- FP32 peak microbenchmark --------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_chain(float* out, int iters)
{
    float a[8];
    const float b = 1.0000001f, c = 1e-7f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) ffma2_chain(float2* out, int iters)
{
    unsigned long long a[8], b, c;
    {
        float2 bf = make_float2(1.0000001f, 0.9999999f), cf = make_float2(1e-7f, 2e-7f);
        b = *reinterpret_cast<unsigned long long*>(&bf);
        c = *reinterpret_cast<unsigned long long*>(&cf);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float2 t = make_float2((float)(threadIdx.x + k), (float)k);
        a[k] = *reinterpret_cast<unsigned long long*>(&t);
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[k]) : "l"(b), "l"(c));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s ^= a[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = *reinterpret_cast<float2*>(&s);
}

// ---- cached scratch for the host-buffer entry points --------------------------------------------
struct HostPathScratch {
    std::mutex mu;
    void* dev = nullptr;
    size_t dev_bytes = 0;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaStream_t stream = nullptr;

    int ensure(size_t dbytes, size_t hbytes)
    {
        if (!stream) PYQMD_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        if (dbytes > dev_bytes) {
            if (dev) cudaFree(dev);
            dev = nullptr; dev_bytes = 0;
            PYQMD_CUDA_CHECK(cudaMalloc(&dev, dbytes));
            dev_bytes = dbytes;
        }
        if (hbytes > pinned_bytes) {
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr; pinned_bytes = 0;
            PYQMD_CUDA_CHECK(cudaMallocHost(&pinned, hbytes));
            pinned_bytes = hbytes;
        }
        return PYQMD_OK;
    }
};
static HostPathScratch g_scratch;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// cloud_host.cu: sort once -> n_steps of the symmetric scheme -> un-sort, host arrays in and out
int cloud_host_steps(float* h_pos, float* h_vel, const uint8_t* h_is_proton, float* h_force, int64_t n,
                     float strong, float coulomb, float pauli, float dt, int32_t n_steps);

// Runs n_steps Jacobi steps on nucleus-relative FP32 state staged in pinned memory.
// h_pos/h_vel: float2[n] in pinned memory (in/out); h_isp: uint8[n]; centre (optional):
// float[2] used for every step instead of the mean position.
static int run_host_steps(HostPathScratch& S, int64_t n, float strong, float coulomb, float pauli,
                          float dt, int n_steps, const float* centre_override)
{
    const size_t b_pos = align_up(sizeof(float) * 2 * n, 256);
    const size_t b_isp = align_up((size_t)n, 256);
    unsigned char* hp = reinterpret_cast<unsigned char*>(S.pinned);
    unsigned char* dp = reinterpret_cast<unsigned char*>(S.dev);
    // device layout: pos | vel | isp | misc(256)
    float* d_pos = reinterpret_cast<float*>(dp);
    float* d_vel = reinterpret_cast<float*>(dp + b_pos);
    uint8_t* d_isp = dp + 2 * b_pos;
    unsigned char* d_misc = dp + 2 * b_pos + b_isp;
    cudaStream_t st = S.stream;
    PYQMD_CUDA_CHECK(cudaMemcpyAsync(dp, hp, 2 * b_pos + b_isp, cudaMemcpyHostToDevice, st));

    int rc = PYQMD_OK;
    if (n <= 1024) {
        // one nucleus, one block: the ensemble kernel with decay disabled
        int64_t* d_off = reinterpret_cast<int64_t*>(d_misc);
        int32_t* d_cnt = reinterpret_cast<int32_t*>(d_misc + 8);
        float* d_ctr = reinterpret_cast<float*>(d_misc + 16);
        struct { int64_t off; int32_t cnt; int32_t pad; float c[2]; } hm = {0, (int32_t)n, 0, {0.f, 0.f}};
        if (centre_override) { hm.c[0] = centre_override[0]; hm.c[1] = centre_override[1]; }
        PYQMD_CUDA_CHECK(cudaMemcpyAsync(d_misc, &hm, sizeof hm, cudaMemcpyHostToDevice, st));
        pyqmd_ensemble e;
        memset(&e, 0, sizeof e);
        e.pos = d_pos; e.vel = d_vel; e.is_proton = d_isp;
        e.offset = d_off; e.count = d_cnt;
        e.n_nuclei = 1; e.n_list = 1; e.cap = (int32_t)n;
        e.centre = centre_override ? d_ctr : nullptr;
        e.strong = strong; e.coulomb = coulomb; e.pauli = pauli; e.dt_phys = dt;
        rc = pyqmd_ensemble_step(&e, n_steps, st);
        if (rc != PYQMD_OK) return rc;
    } else {
        set_error("run_host_steps handles n <= 1024 only");
        return PYQMD_ERR_INVALID;
    }
    PYQMD_CUDA_CHECK(cudaMemcpyAsync(hp, dp, 2 * b_pos, cudaMemcpyDeviceToHost, st));
    PYQMD_CUDA_CHECK(cudaStreamSynchronize(st));
    return PYQMD_OK;
}

static int prepare_scratch(HostPathScratch& S, int64_t n)
{
    const size_t b_pos = align_up(sizeof(float) * 2 * n, 256);
    const size_t b_isp = align_up((size_t)n, 256);
    // n > 1024 goes through cloud_host_steps, which owns its device buffers: pinned staging only
    return S.ensure(n <= 1024 ? 2 * b_pos + b_isp + 256 : 0, 2 * b_pos + b_isp);
}

// ---- host-buffer ensemble pipeline: three in-order lanes linked by per-chunk events ------------------
struct HostPipeline {
    std::mutex mu;
    cudaStream_t up = nullptr, run = nullptr, down = nullptr;
    cudaEvent_t fork = nullptr, join[3] = {nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> ev_up, ev_run;

    int ensure(int n_chunks)
    {
        if (!up) {
            PYQMD_CUDA_CHECK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
            PYQMD_CUDA_CHECK(cudaStreamCreateWithFlags(&run, cudaStreamNonBlocking));
            PYQMD_CUDA_CHECK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
            PYQMD_CUDA_CHECK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
            for (auto& j : join) PYQMD_CUDA_CHECK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
        }
        while ((int)ev_up.size() < n_chunks) {
            cudaEvent_t a, b;
            PYQMD_CUDA_CHECK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            PYQMD_CUDA_CHECK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            ev_up.push_back(a);
            ev_run.push_back(b);
        }
        return PYQMD_OK;
    }
};
static HostPipeline g_pipe;

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int pyqmd_ensemble_step_host(const pyqmd_ensemble* e, float* h_pos, float* h_vel,
                                        uint8_t* h_is_proton, int32_t* h_count, int32_t* h_zn,
                                        const pyqmd_host_chunk* chunks, int32_t n_chunks,
                                        int32_t n_steps, void* stream)
{
    PYQMD_REQUIRE(e != nullptr && chunks != nullptr && n_chunks >= 0 && n_steps >= 0, "arguments");
    PYQMD_REQUIRE(h_pos && h_vel, "host arrays");
    PYQMD_REQUIRE(h_is_proton || !e->decay_enabled, "h_is_proton may only be NULL when decay is disabled");
    PYQMD_REQUIRE(e->pos && e->vel && e->is_proton && e->offset && e->count, "device staging arrays");
    if (e->decay_enabled) PYQMD_REQUIRE(h_count && h_zn && e->zn, "count / zn arrays (decay enabled)");
    if (n_chunks == 0 || n_steps == 0) return PYQMD_OK;
    std::lock_guard<std::mutex> lock(g_pipe.mu);
    int rc = g_pipe.ensure(n_chunks);
    if (rc != PYQMD_OK) return rc;
    cudaStream_t user = (cudaStream_t)stream;
    PYQMD_CUDA_CHECK(cudaEventRecord(g_pipe.fork, user));
    PYQMD_CUDA_CHECK(cudaStreamWaitEvent(g_pipe.up, g_pipe.fork, 0));
    PYQMD_CUDA_CHECK(cudaStreamWaitEvent(g_pipe.run, g_pipe.fork, 0));
    PYQMD_CUDA_CHECK(cudaStreamWaitEvent(g_pipe.down, g_pipe.fork, 0));
    for (int k = 0; k < n_chunks; ++k) {
        const pyqmd_host_chunk& c = chunks[k];
        PYQMD_REQUIRE(c.slot0 >= 0 && c.slot0 <= c.slot1 && c.nuc0 >= 0 && c.nuc0 <= c.nuc1 &&
                      c.nuc1 <= e->n_nuclei && c.n_launch >= 0 && c.n_launch <= PYQMD_MAX_CHUNK_LAUNCHES,
                      "chunk descriptor");
        const size_t ns = (size_t)(c.slot1 - c.slot0), nn = (size_t)(c.nuc1 - c.nuc0);
        PYQMD_CUDA_CHECK(cudaMemcpyAsync(e->pos + 2 * c.slot0, h_pos + 2 * c.slot0, 8 * ns,
                                         cudaMemcpyHostToDevice, g_pipe.up));
        PYQMD_CUDA_CHECK(cudaMemcpyAsync(e->vel + 2 * c.slot0, h_vel + 2 * c.slot0, 8 * ns,
                                         cudaMemcpyHostToDevice, g_pipe.up));
        if (h_is_proton)      // types cannot change without decay: the caller may upload them once
            PYQMD_CUDA_CHECK(cudaMemcpyAsync(e->is_proton + c.slot0, h_is_proton + c.slot0, ns,
                                             cudaMemcpyHostToDevice, g_pipe.up));
        PYQMD_CUDA_CHECK(cudaEventRecord(g_pipe.ev_up[k], g_pipe.up));
        PYQMD_CUDA_CHECK(cudaStreamWaitEvent(g_pipe.run, g_pipe.ev_up[k], 0));
        for (int l = 0; l < c.n_launch; ++l) {
            pyqmd_ensemble d = *e;
            d.cap = c.cap[l];
            d.list = c.list[l];
            d.n_list = c.n_list[l];
            rc = pyqmd_ensemble_step(&d, n_steps, g_pipe.run);
            if (rc != PYQMD_OK) return rc;
        }
        PYQMD_CUDA_CHECK(cudaEventRecord(g_pipe.ev_run[k], g_pipe.run));
        PYQMD_CUDA_CHECK(cudaStreamWaitEvent(g_pipe.down, g_pipe.ev_run[k], 0));
        PYQMD_CUDA_CHECK(cudaMemcpyAsync(h_pos + 2 * c.slot0, e->pos + 2 * c.slot0, 8 * ns,
                                         cudaMemcpyDeviceToHost, g_pipe.down));
        PYQMD_CUDA_CHECK(cudaMemcpyAsync(h_vel + 2 * c.slot0, e->vel + 2 * c.slot0, 8 * ns,
                                         cudaMemcpyDeviceToHost, g_pipe.down));
        if (e->decay_enabled) {
            PYQMD_CUDA_CHECK(cudaMemcpyAsync(h_is_proton + c.slot0, e->is_proton + c.slot0, ns,
                                             cudaMemcpyDeviceToHost, g_pipe.down));
            PYQMD_CUDA_CHECK(cudaMemcpyAsync(h_count + c.nuc0, e->count + c.nuc0, 4 * nn,
                                             cudaMemcpyDeviceToHost, g_pipe.down));
            PYQMD_CUDA_CHECK(cudaMemcpyAsync(h_zn + c.nuc0, e->zn + c.nuc0, 4 * nn,
                                             cudaMemcpyDeviceToHost, g_pipe.down));
        }
    }
    cudaStream_t lanes[3] = {g_pipe.up, g_pipe.run, g_pipe.down};
    for (int l = 0; l < 3; ++l) {
        PYQMD_CUDA_CHECK(cudaEventRecord(g_pipe.join[l], lanes[l]));
        PYQMD_CUDA_CHECK(cudaStreamWaitEvent(user, g_pipe.join[l], 0));
    }
    return PYQMD_OK;
}

extern "C" int pyqmd_abi_version(void) { return PYQMD_ABI_VERSION; }

extern "C" const char* pyqmd_last_error(void) { return g_err; }

extern "C" int pyqmd_device_props(int device, int64_t out[8])
{
    PYQMD_REQUIRE(out != nullptr, "out is NULL");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        set_error("no CUDA device visible");
        return PYQMD_ERR_NO_DEVICE;
    }
    PYQMD_REQUIRE(device >= 0 && device < count, "device index");
    cudaDeviceProp p;
    PYQMD_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, device);
    out[0] = p.multiProcessorCount;
    out[1] = p.major;
    out[2] = p.minor;
    out[3] = clock_khz;
    out[4] = p.l2CacheSize;
    out[5] = (int64_t)p.sharedMemPerBlockOptin;
    out[6] = (int64_t)(p.totalGlobalMem >> 20);
    out[7] = 0;
    return PYQMD_OK;
}

extern "C" int pyqmd_struct_sizes(int64_t out[4])
{
    static_assert(sizeof(pyqmd_free_particle) == 64 && sizeof(pyqmd_free_frame) == 64, "ABI layout");
    PYQMD_REQUIRE(out != nullptr, "out is NULL");
    out[0] = sizeof(pyqmd_nuclide_entry);
    out[1] = sizeof(pyqmd_decay_event);
    out[2] = sizeof(pyqmd_ensemble);
    out[3] = sizeof(pyqmd_population);
    return PYQMD_OK;
}

extern "C" int pyqmd_fp32_peak(int iters, double* tflops_ffma, double* tflops_ffma2, void* stream)
{
    PYQMD_REQUIRE(iters > 0 && tflops_ffma && tflops_ffma2, "iters > 0, outputs non-NULL");
    int dev = 0, sms = 0;
    PYQMD_CUDA_CHECK(cudaGetDevice(&dev));
    PYQMD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256;
    float2* buf = nullptr;
    PYQMD_CUDA_CHECK(cudaMalloc(&buf, sizeof(float2) * (size_t)blocks * threads));
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    double best1 = 0.0, best2 = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, st);
        ffma_chain<<<blocks, threads, 0, st>>>(reinterpret_cast<float*>(buf), iters);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double f1 = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && f1 > best1) best1 = f1;
        cudaEventRecord(e0, st);
        ffma2_chain<<<blocks, threads, 0, st>>>(buf, iters);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double f2 = 4.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && f2 > best2) best2 = f2;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaError_t err = cudaGetLastError();
    cudaFree(buf);
    if (err != cudaSuccess) {
        set_error("fp32 peak microbenchmark failed: %s", cudaGetErrorString(err));
        return PYQMD_ERR_CUDA;
    }
    *tflops_ffma = best1;
    *tflops_ffma2 = best2;
    return PYQMD_OK;
}

extern "C" int pyqmd_update_forces_and_positions(float* particles, const int32_t* types,
                                                 int32_t num_particles, float center_x,
                                                 float center_y, float strong_strength,
                                                 float coulomb_strength, float pauli_strength,
                                                 float dt)
{
    const int64_t n = num_particles;
    PYQMD_REQUIRE(n >= 0, "num_particles >= 0");
    if (n == 0) return PYQMD_OK;                            // nuclear_forces.py:186-188
    PYQMD_REQUIRE(particles && types, "NULL pointer");
    PYQMD_REQUIRE(n <= 1024, "num_particles <= 1024 for the float4 entry point; use "
                             "pyqmd_update_particles_f64 or pyqmd_cloud_step for larger systems");
    HostPathScratch& S = g_scratch;
    std::lock_guard<std::mutex> lock(S.mu);
    int rc = prepare_scratch(S, n);
    if (rc != PYQMD_OK) return rc;
    const size_t b_pos = align_up(sizeof(float) * 2 * n, 256);
    unsigned char* hp = reinterpret_cast<unsigned char*>(S.pinned);
    float* h_pos = reinterpret_cast<float*>(hp);
    float* h_vel = reinterpret_cast<float*>(hp + b_pos);
    uint8_t* h_isp = hp + 2 * b_pos;
    // nucleus-relative coordinates: x - centre is exact in float64 and loses nothing in FP32
    const double cx = center_x, cy = center_y;
    for (int64_t i = 0; i < n; ++i) {
        h_pos[2 * i] = (float)((double)particles[4 * i] - cx);
        h_pos[2 * i + 1] = (float)((double)particles[4 * i + 1] - cy);
        h_vel[2 * i] = particles[4 * i + 2];
        h_vel[2 * i + 1] = particles[4 * i + 3];
        h_isp[i] = types[i] == 0 ? 1 : 0;                   // 0 = proton, :199
    }
    const float zero_centre[2] = {0.f, 0.f};
    rc = run_host_steps(S, n, strong_strength, coulomb_strength, pauli_strength, dt, 1, zero_centre);
    if (rc != PYQMD_OK) return rc;
    for (int64_t i = 0; i < n; ++i) {
        particles[4 * i] = (float)((double)h_pos[2 * i] + cx);
        particles[4 * i + 1] = (float)((double)h_pos[2 * i + 1] + cy);
        particles[4 * i + 2] = h_vel[2 * i];
        particles[4 * i + 3] = h_vel[2 * i + 1];
    }
    return PYQMD_OK;
}

extern "C" int pyqmd_update_particles_f64(double* x, double* y, double* vx, double* vy,
                                          const uint8_t* is_proton, int64_t n,
                                          double strong_strength, double coulomb_strength,
                                          double pauli_strength, double dt, int32_t n_steps)
{
    PYQMD_REQUIRE(n >= 0 && n_steps >= 0, "n >= 0, n_steps >= 0");
    if (n == 0 || n_steps == 0) return PYQMD_OK;            // nuclear_forces.py:238-239
    PYQMD_REQUIRE(x && y && vx && vy && is_proton, "NULL pointer");
    HostPathScratch& S = g_scratch;
    std::lock_guard<std::mutex> lock(S.mu);
    int rc = prepare_scratch(S, n);
    if (rc != PYQMD_OK) return rc;
    const size_t b_pos = align_up(sizeof(float) * 2 * n, 256);
    unsigned char* hp = reinterpret_cast<unsigned char*>(S.pinned);
    float* h_pos = reinterpret_cast<float*>(hp);
    float* h_vel = reinterpret_cast<float*>(hp + b_pos);
    uint8_t* h_isp = hp + 2 * b_pos;
    // centre of mass in float64 (nuclear_forces.py:242-243); it only serves as the origin of
    // the FP32 working frame, the device recomputes the containment centre every step
    double cx = 0.0, cy = 0.0;
    for (int64_t i = 0; i < n; ++i) { cx += x[i]; cy += y[i]; }
    cx /= (double)n;
    cy /= (double)n;
    for (int64_t i = 0; i < n; ++i) {
        h_pos[2 * i] = (float)(x[i] - cx);
        h_pos[2 * i + 1] = (float)(y[i] - cy);
        h_vel[2 * i] = (float)vx[i];
        h_vel[2 * i + 1] = (float)vy[i];
        h_isp[i] = is_proton[i] ? 1 : 0;
    }
    if (n <= 1024)
        rc = run_host_steps(S, n, (float)strong_strength, (float)coulomb_strength,
                            (float)pauli_strength, (float)dt, n_steps, nullptr);
    else       // one large system: sorted symmetric scheme (cloud_host.cu)
        rc = cloud_host_steps(h_pos, h_vel, h_isp, nullptr, n, (float)strong_strength,
                              (float)coulomb_strength, (float)pauli_strength, (float)dt, n_steps);
    if (rc != PYQMD_OK) return rc;
    for (int64_t i = 0; i < n; ++i) {
        x[i] = (double)h_pos[2 * i] + cx;
        y[i] = (double)h_pos[2 * i + 1] + cy;
        vx[i] = (double)h_vel[2 * i];
        vy[i] = (double)h_vel[2 * i + 1];
    }
    return PYQMD_OK;
}
