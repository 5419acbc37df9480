// overlaps.cu -- per-frame overlap projection for ensembles of nuclei (sm_100a).
//
// Replaces NuclearSimulation.resolve_overlaps (OtsoBear/PyQMD nuclear_sim.py:355-379), which the
// app runs once per frame after the sub-steps (:175-176): a *sequential* Gauss-Seidel sweep over
// i < j that pushes any two nucleons closer than 5.0 apart symmetrically, with immediate
// updates.  The order dependence is part of the behaviour, so it is kept: one warp owns one
// nucleus; for row i the 32 lanes test 32 partners j against the current position of i in
// parallel, and the pushes of a chunk are applied one at a time in j order (ballot + first set
// bit), re-testing the remaining lanes against the updated i.  Rows with no overlap cost one
// ballot per 32 partners.  Positions live in shared memory for the whole sweep.
#include <mutex>
#include "common.cuh"
#include "decay_device.cuh"

namespace pyqmd {

constexpr int kOvWarps = 4;            // nuclei per block
constexpr float kMinDist = 5.0f;       // 2 * radius, nuclear_sim.py:357
constexpr uint32_t kOverlapSlot0 = 16; // Philox slots >= 16 are reserved for this kernel

__global__ void __launch_bounds__(kOvWarps * 32)
resolve_overlaps_kernel(const pyqmd_ensemble e, const double* __restrict__ uniforms,
                        const int uniforms_per_nucleus, unsigned long long* __restrict__ n_pushes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float2* sp = reinterpret_cast<float2*>(smem_raw) + (size_t)w * e.cap;
    const int64_t q = (int64_t)blockIdx.x * kOvWarps + w;
    if (q >= e.n_list) return;                               // whole warp exits together
    const int nuc = e.list ? e.list[q] : (int)q;
    const int cnt = e.count[nuc];
    const int64_t off = e.offset[nuc];
    float2* gpos = reinterpret_cast<float2*>(e.pos) + off;
    for (int k = lane; k < cnt; k += 32) sp[k] = gpos[k];
    __syncwarp();

    const uint64_t gid = (uint64_t)(e.id_base + nuc);
    int draws = 0;
    unsigned long long pushes = 0;
    for (int i = 0; i + 1 < cnt; ++i) {                      // :359
        float2 pi = sp[i];
        for (int base = i + 1; base < cnt; base += 32) {     // :360, 32 partners at a time
            const int j = base + lane;
            const bool valid = j < cnt;
            float2 pj = valid ? sp[j] : make_float2(0.f, 0.f);
            int done = -1;                                   // lanes <= done are settled
            while (true) {
                float dx = pj.x - pi.x, dy = pj.y - pi.y;    // :361-362
                const float d2 = dx * dx + dy * dy;          // :363
                const bool ov = valid && lane > done && d2 < kMinDist * kMinDist;   // :365
                const unsigned mask = __ballot_sync(0xffffffffu, ov);
                if (!mask) break;
                const int l = __ffs(mask) - 1;               // first overlapping partner in j order
                float dist = sqrtf(d2);                      // :366
                const bool degenerate = __shfl_sync(0xffffffffu, dist < 0.001f ? 1 : 0, l) != 0;
                if (degenerate) {                            // :367-370, random direction
                    double u;
                    if (uniforms && draws < uniforms_per_nucleus)
                        u = uniforms[(int64_t)nuc * uniforms_per_nucleus + draws];
                    else {
                        uint32_t r[4];
                        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), e.step0,
                                      kOverlapSlot0 + (uint32_t)draws, (uint32_t)e.seed,
                                      (uint32_t)(e.seed >> 32), r);
                        u = u53(r[0], r[1]);
                    }
                    ++draws;
                    const float ang = (float)(6.283185307179586 * u);
                    dx = cosf(ang);
                    dy = sinf(ang);
                    dist = 0.001f;
                } else {
                    dx /= dist;                              // :372-373
                    dy /= dist;
                }
                const float push = (kMinDist - dist) * 0.5f; // :375
                const float nix = pi.x - dx * push, niy = pi.y - dy * push;   // :376-377
                if (lane == l) {
                    pj.x += dx * push;                       // :378-379
                    pj.y += dy * push;
                    sp[j] = pj;
                }
                pi.x = __shfl_sync(0xffffffffu, nix, l);
                pi.y = __shfl_sync(0xffffffffu, niy, l);
                done = l;
                ++pushes;
            }
        }
        if (lane == 0) sp[i] = pi;
        __syncwarp();
    }
    for (int k = lane; k < cnt; k += 32) gpos[k] = sp[k];
    if (n_pushes && lane == 0 && pushes) atomicAdd(n_pushes, pushes);
}

// ---- branch census -----------------------------------------------------------------------------------
// Counts, for every ordered pair of the listed nuclei, which branches of the law
// (nuclear_forces.py:257-291) it takes.  Used for the algorithmic-FLOP accounting of the roofline
// (15 per evaluated pair + 4/7/8 core/attractive/tail + 5 hard core + 3 p-p
// + 6 Pauli); not on the hot path.
// counts: [0] evaluated, [1] skipped (d2 < 0.01), [2] hard, [3] core, [4] attractive, [5] tail,
//         [6] p-p, [7] Pauli
__global__ void __launch_bounds__(256) census_kernel(const pyqmd_ensemble e,
                                                     unsigned long long* __restrict__ counts)
{
    __shared__ unsigned long long sc[8];
    if (threadIdx.x < 8) sc[threadIdx.x] = 0;
    __syncthreads();
    const int64_t q = blockIdx.x;
    const int nuc = e.list ? e.list[q] : (int)q;
    const int cnt = e.count[nuc];
    const float2* pos = reinterpret_cast<const float2*>(e.pos) + e.offset[nuc];
    const uint8_t* isp = e.is_proton + e.offset[nuc];
    unsigned c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t p = threadIdx.x; p < (int64_t)cnt * cnt; p += blockDim.x) {
        const int i = (int)(p / cnt), j = (int)(p % cnt);
        if (i == j) continue;
        const float dx = pos[j].x - pos[i].x, dy = pos[j].y - pos[i].y;
        const float d2 = dx * dx + dy * dy;
        if (d2 < 0.01f) { ++c[1]; continue; }
        ++c[0];
        if (d2 < 4.25f * 4.25f) ++c[2];
        if (d2 < 2.8f * 2.8f) ++c[3];
        else if (d2 < 81.0f) ++c[4];
        else ++c[5];
        if (isp[i] && isp[j]) ++c[6];
        if ((isp[i] != 0) == (isp[j] != 0) && d2 < 64.0f) ++c[7];
    }
    for (int k = 0; k < 8; ++k)
        if (c[k]) atomicAdd(&sc[k], (unsigned long long)c[k]);
    __syncthreads();
    if (threadIdx.x < 8 && sc[threadIdx.x]) atomicAdd(counts + threadIdx.x, sc[threadIdx.x]);
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int pyqmd_ensemble_census(const pyqmd_ensemble* e, unsigned long long* counts,
                                     void* stream)
{
    PYQMD_REQUIRE(e != nullptr && counts != nullptr, "NULL argument");
    PYQMD_REQUIRE(e->pos && e->is_proton && e->offset && e->count, "state arrays");
    const int64_t n_list = e->list ? e->n_list : e->n_nuclei;
    if (n_list == 0) return PYQMD_OK;
    PYQMD_REQUIRE(n_list <= 2147483647LL, "too many nuclei for one launch");
    census_kernel<<<(unsigned)n_list, 256, 0, (cudaStream_t)stream>>>(*e, counts);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}

extern "C" int pyqmd_resolve_overlaps(const pyqmd_ensemble* e, const double* uniforms,
                                      int32_t uniforms_per_nucleus, unsigned long long* n_pushes,
                                      void* stream)
{
    PYQMD_REQUIRE(e != nullptr, "ensemble descriptor is NULL");
    PYQMD_REQUIRE(e->pos && e->offset && e->count, "state arrays");
    PYQMD_REQUIRE(e->cap >= 1 && e->cap <= 4096, "cap must be in [1, 4096]");
    const int64_t n_list = e->list ? e->n_list : e->n_nuclei;
    if (n_list == 0) return PYQMD_OK;
    pyqmd_ensemble d = *e;
    d.n_list = n_list;
    const int64_t grid = (n_list + kOvWarps - 1) / kOvWarps;
    PYQMD_REQUIRE(grid <= 2147483647LL, "too many nuclei for one launch");
    const size_t smem = (size_t)kOvWarps * e->cap * sizeof(float2);
    if (smem > 48 * 1024) {
        // cudaFuncSetAttribute is per device: remember which devices have been configured
        static unsigned char done[64] = {0};
        static std::mutex mu;
        int dev = 0;
        PYQMD_CUDA_CHECK(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 0 || dev >= 64 || !done[dev]) {
            PYQMD_CUDA_CHECK(cudaFuncSetAttribute(resolve_overlaps_kernel,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  200 * 1024));
            if (dev >= 0 && dev < 64) done[dev] = 1;
        }
    }
    resolve_overlaps_kernel<<<(unsigned)grid, kOvWarps * 32, smem, (cudaStream_t)stream>>>(
        d, uniforms, uniforms_per_nucleus, n_pushes);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
