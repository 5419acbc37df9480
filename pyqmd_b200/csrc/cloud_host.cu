// cloud_host.cu -- reference-shaped, host-buffer entry point for ONE large system (sm_100a).
//
// NuclearForces.update_particles_gpu (OtsoBear/PyQMD nuclear_forces.py:185-234) packs the particle
// list into host arrays, uploads them, runs its kernel, waits, downloads and writes back -- every
// step.  pyqmd_cloud_step_host is that call for a system of any size: host arrays in, host arrays
// out, blocking.  Between the two copies it does what NucleonCloud does for device-resident state:
// sort once (type bit + 2-D Morton code, stable radix sort) so that whole 256-nucleon tiles take the
// far-field path, run n_steps Jacobi steps of the symmetric scheme (every unordered pair once,
// cloud_sym.cu), un-sort on the way out.  The caller's nucleon order is preserved.
#include <cub/device/device_radix_sort.cuh>

#include <mutex>

#include "cloud.cuh"

namespace pyqmd {

int cloud_prepass(const float* pos, const uint8_t* is_proton, int64_t n, const CloudWorkspace& w,
                  cudaStream_t st);

// bounding box of the whole cloud from the per-tile boxes of the pre-pass
__global__ void __launch_bounds__(256) cloud_bounds_kernel(CloudWorkspace w, int64_t n_tiles,
                                                           float4* __restrict__ out)
{
    __shared__ float4 s[256];
    float4 b = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
    for (int64_t k = threadIdx.x; k < n_tiles; k += 256) {
        const float4 t = w.bbox[k];
        b.x = fminf(b.x, t.x); b.y = fminf(b.y, t.y);
        b.z = fmaxf(b.z, t.z); b.w = fmaxf(b.w, t.w);
    }
    s[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const float4 t = s[threadIdx.x + o];
            float4& m = s[threadIdx.x];
            m.x = fminf(m.x, t.x); m.y = fminf(m.y, t.y);
            m.z = fmaxf(m.z, t.z); m.w = fmaxf(m.w, t.w);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s[0];
}

__global__ void __launch_bounds__(256)
cloud_keys_from_bounds_kernel(const float2* __restrict__ pos, const uint8_t* __restrict__ isp, int64_t n,
                              const float4* __restrict__ bounds, uint64_t* __restrict__ keys,
                              int32_t* __restrict__ idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = *bounds;
    const float extent = fmaxf(b.z - b.x, b.w - b.y) * 1.0001f + 1e-6f;
    keys[i] = cloud_sort_key(pos[i], isp[i] != 0, b.x, b.y, 1.0f / extent);
    idx[i] = (int32_t)i;
}

// sorted[k] = original[perm[k]]
__global__ void __launch_bounds__(256)
cloud_gather_kernel(const int32_t* __restrict__ perm, int64_t n, const float2* __restrict__ pos,
                    const float2* __restrict__ vel, const uint8_t* __restrict__ isp,
                    float2* __restrict__ pos_s, float2* __restrict__ vel_s, uint8_t* __restrict__ isp_s)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int32_t i = perm[k];
    pos_s[k] = pos[i];
    vel_s[k] = vel[i];
    isp_s[k] = isp[i];
}

// original[perm[k]] = sorted[k]
__global__ void __launch_bounds__(256)
cloud_scatter_kernel(const int32_t* __restrict__ perm, int64_t n, const float2* __restrict__ pos_s,
                     const float2* __restrict__ vel_s, const float2* __restrict__ force_s,
                     float2* __restrict__ pos, float2* __restrict__ vel, float2* __restrict__ force)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int32_t i = perm[k];
    pos[i] = pos_s[k];
    vel[i] = vel_s[k];
    if (force) force[i] = force_s[k];
}

struct CloudHostScratch {
    std::mutex mu;
    int device = -1;
    int64_t cap_n = 0;
    cudaStream_t stream = nullptr;
    unsigned char* buf = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    // carved from buf
    float2 *pos, *vel, *pos_s, *pos_s2, *vel_s, *force_s, *force;
    uint8_t *isp, *isp_s;
    uint64_t *keys, *keys_out;
    int32_t *idx, *perm;
    long long* acc;
    float4* bounds;
    void* ws;

    void release()
    {
        if (buf) cudaFree(buf);
        if (cub_tmp) cudaFree(cub_tmp);
        buf = nullptr; cub_tmp = nullptr; cub_bytes = 0; cap_n = 0;
    }

    int ensure(int64_t n)
    {
        int dev = 0;
        PYQMD_CUDA_CHECK(cudaGetDevice(&dev));
        if (dev != device) {                                  // buffers and stream belong to a device
            release();
            if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
            device = dev;
        }
        if (!stream) PYQMD_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        if (n <= cap_n) return PYQMD_OK;
        release();
        auto up = [](size_t v) { return (v + 255) / 256 * 256; };
        const size_t b2 = up(sizeof(float2) * (size_t)n), b1 = up((size_t)n), b8 = up(8 * (size_t)n),
                     b4 = up(4 * (size_t)n), bacc = up(16 * (size_t)n),
                     bws = up((size_t)pyqmd_cloud_workspace_bytes(n));
        const size_t total = 7 * b2 + 2 * b1 + 2 * b8 + 2 * b4 + bacc + 256 + bws;
        PYQMD_CUDA_CHECK(cudaMalloc(&buf, total));
        unsigned char* p = buf;
        auto take = [&](size_t bytes) { unsigned char* r = p; p += bytes; return r; };
        pos = (float2*)take(b2); vel = (float2*)take(b2); pos_s = (float2*)take(b2);
        pos_s2 = (float2*)take(b2); vel_s = (float2*)take(b2); force_s = (float2*)take(b2);
        force = (float2*)take(b2);
        isp = take(b1); isp_s = take(b1);
        keys = (uint64_t*)take(b8); keys_out = (uint64_t*)take(b8);
        idx = (int32_t*)take(b4); perm = (int32_t*)take(b4);
        acc = (long long*)take(bacc);
        bounds = (float4*)take(256);
        ws = take(bws);
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys_out, idx, perm, (int)n, 0, 63, stream);
        PYQMD_CUDA_CHECK(cudaMalloc(&cub_tmp, need));
        cub_bytes = need;
        cap_n = n;
        return PYQMD_OK;
    }
};
static CloudHostScratch g_cloud_host;

// Shared by pyqmd_cloud_step_host and the large-n branch of pyqmd_update_particles_f64.
int cloud_host_steps(float* h_pos, float* h_vel, const uint8_t* h_is_proton, float* h_force, int64_t n,
                     float strong, float coulomb, float pauli, float dt, int32_t n_steps)
{
    PYQMD_REQUIRE(n <= 2147483647LL, "n must fit in int32");
    CloudHostScratch& S = g_cloud_host;
    std::lock_guard<std::mutex> lock(S.mu);
    int rc = S.ensure(n);
    if (rc != PYQMD_OK) return rc;
    cudaStream_t st = S.stream;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    PYQMD_CUDA_CHECK(cudaMemcpyAsync(S.pos, h_pos, sizeof(float2) * n, cudaMemcpyHostToDevice, st));
    PYQMD_CUDA_CHECK(cudaMemcpyAsync(S.vel, h_vel, sizeof(float2) * n, cudaMemcpyHostToDevice, st));
    PYQMD_CUDA_CHECK(cudaMemcpyAsync(S.isp, h_is_proton, (size_t)n, cudaMemcpyHostToDevice, st));
    // sort once: type bit + Morton code of the position inside the cloud's bounding box
    const CloudWorkspace w = carve(S.ws, n);
    rc = cloud_prepass(reinterpret_cast<const float*>(S.pos), S.isp, n, w, st);
    if (rc != PYQMD_OK) return rc;
    cloud_bounds_kernel<<<1, 256, 0, st>>>(w, n_tiles_of(n), S.bounds);
    cloud_keys_from_bounds_kernel<<<blocks, 256, 0, st>>>(S.pos, S.isp, n, S.bounds, S.keys, S.idx);
    size_t tmp = S.cub_bytes;
    PYQMD_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(S.cub_tmp, tmp, S.keys, S.keys_out, S.idx, S.perm,
                                                     (int)n, 0, 63, st));
    cloud_gather_kernel<<<blocks, 256, 0, st>>>(S.perm, n, S.pos, S.vel, S.isp, S.pos_s, S.vel_s,
                                                S.isp_s);
    PYQMD_CUDA_CHECK(cudaMemsetAsync(S.acc, 0, 16 * (size_t)n, st));
    PYQMD_CUDA_CHECK(cudaGetLastError());
    float2* in = S.pos_s;
    float2* out = S.pos_s2;
    for (int s = 0; s < n_steps; ++s) {
        rc = pyqmd_cloud_pair_forces(reinterpret_cast<const float*>(in), S.isp_s, n, 0, 1, strong,
                                     coulomb, pauli, S.acc, S.ws, st);
        if (rc != PYQMD_OK) return rc;
        rc = pyqmd_cloud_integrate(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out),
                                   reinterpret_cast<float*>(S.vel_s),
                                   h_force ? reinterpret_cast<float*>(S.force_s) : nullptr, n, 0, n, dt,
                                   S.acc, S.ws, st);
        if (rc != PYQMD_OK) return rc;
        float2* t = in; in = out; out = t;
    }
    cloud_scatter_kernel<<<blocks, 256, 0, st>>>(S.perm, n, in, S.vel_s, h_force ? S.force_s : nullptr,
                                                 S.pos, S.vel, h_force ? S.force : nullptr);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    PYQMD_CUDA_CHECK(cudaMemcpyAsync(h_pos, S.pos, sizeof(float2) * n, cudaMemcpyDeviceToHost, st));
    PYQMD_CUDA_CHECK(cudaMemcpyAsync(h_vel, S.vel, sizeof(float2) * n, cudaMemcpyDeviceToHost, st));
    if (h_force)
        PYQMD_CUDA_CHECK(cudaMemcpyAsync(h_force, S.force, sizeof(float2) * n, cudaMemcpyDeviceToHost, st));
    PYQMD_CUDA_CHECK(cudaStreamSynchronize(st));
    return PYQMD_OK;
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int pyqmd_cloud_step_host(float* h_pos, float* h_vel, const uint8_t* h_is_proton,
                                     float* h_force, int64_t n, float strong, float coulomb,
                                     float pauli, float dt, int32_t n_steps)
{
    PYQMD_REQUIRE(n >= 0 && n_steps >= 0, "n >= 0, n_steps >= 0");
    if (n == 0 || n_steps == 0) return PYQMD_OK;            // nuclear_forces.py:186-188
    PYQMD_REQUIRE(h_pos && h_vel && h_is_proton, "NULL pointer");
    return cloud_host_steps(h_pos, h_vel, h_is_proton, h_force, n, strong, coulomb, pauli, dt, n_steps);
}
