// free_particles.cu -- life cycle of the particles a decay emits (alpha, e-, e+, gamma, n, p), on the
// device (sm_100a).
//
// Reference (OtsoBear/PyQMD): handle_decay appends products(x, y) to self.particles after rewriting
// their speed and lifetime (nuclear_sim.py:294-349); every sub-step update_particle advances and ages
// each free particle and drops the expired ones (nuclear_sim.py:162,178-210).  Here the decay events
// the ensemble kernel logged during a frame (pyqmd_decay_event) are turned into free particles and
// the pool of free particles is advanced by the frame's sub-steps, without the host looking at a
// single count: both kernels read their loop bounds from device counters.
//
// Everything that the reference computes once per frame from (time scale, sub-steps, physics dt) --
// speed scale, aging scale, the time-scaled dt of nucleon-type products, the lifetime of fast-forward
// frames -- is computed on the HOST with the reference's expressions (pyqmd_b200/sim.py:
// frame_constants) and passed in, so the device only does IEEE add / mul / div / sqrt in float64.
#include "common.cuh"

namespace pyqmd {

__device__ __forceinline__ bool is_animated(int ptype)      // nuclear_sim.py:182-183
{
    return ptype == PYQMD_PT_ALPHA || ptype == PYQMD_PT_ELECTRON || ptype == PYQMD_PT_GAMMA ||
           ptype == PYQMD_PT_POSITRON;
}

// n_updates calls of update_particle (nuclear_sim.py:178-210); returns false once the particle expired
__device__ __forceinline__ bool advance(pyqmd_free_particle& p, int n_updates, const pyqmd_free_frame& f)
{
    if (is_animated(p.type)) {
        for (int k = 0; k < n_updates; ++k) {
            // explicit _rn intrinsics: no FMA contraction, every operation rounds like CPython's
            p.x = __dadd_rn(p.x, __dmul_rn(__dmul_rn(p.vx, 1.0 / 240.0), f.speed_scale));   // :194
            p.y = __dadd_rn(p.y, __dmul_rn(__dmul_rn(p.vy, 1.0 / 240.0), f.speed_scale));   // :195
            p.age = __dadd_rn(p.age, __dmul_rn(f.age_dt, f.aging_scale));                   // :201
            if (!(p.age < p.lifetime)) return false;            // :204
        }
    } else {
        for (int k = 0; k < n_updates; ++k) {
            p.x = __dadd_rn(p.x, __dmul_rn(p.vx, f.nucleon_dt));                            // :207-209
            p.y = __dadd_rn(p.y, __dmul_rn(p.vy, f.nucleon_dt));
            p.age = __dadd_rn(p.age, f.age_dt);
        }
    }
    return true;
}

__device__ __forceinline__ void append(pyqmd_free_particle* out, unsigned long long* n_out, int64_t capacity,
                                       unsigned long long* dropped, const pyqmd_free_particle& p)
{
    const unsigned long long slot = atomicAdd(n_out, 1ULL);
    if ((int64_t)slot < capacity) out[slot] = p;
    else if (dropped) atomicAdd(dropped, 1ULL);
}

// pool -> pool: the frame's num_steps updates of every particle that was already free
__global__ void __launch_bounds__(256)
free_advance_kernel(const pyqmd_free_particle* __restrict__ in, const unsigned long long* __restrict__ n_in,
                    pyqmd_free_particle* __restrict__ out, unsigned long long* n_out, int64_t capacity,
                    unsigned long long* dropped, pyqmd_free_frame f)
{
    const int64_t n = min((int64_t)*n_in, capacity);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        pyqmd_free_particle p = in[i];
        if (advance(p, f.num_steps, f)) append(out, n_out, capacity, dropped, p);
    }
}

// event log -> pool: speed / lifetime rewrite of handle_decay (nuclear_sim.py:295-342), then the
// sub-steps that were left in the frame when the particle was emitted
__global__ void __launch_bounds__(256)
free_spawn_kernel(const pyqmd_decay_event* __restrict__ events, const unsigned long long* __restrict__ n_events,
                  int64_t event_capacity, pyqmd_free_particle* __restrict__ out, unsigned long long* n_out,
                  int64_t capacity, unsigned long long* dropped, pyqmd_free_frame f)
{
    const int64_t n = min((int64_t)*n_events, event_capacity);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const pyqmd_decay_event e = events[i];
        if (e.ptype < 0) continue;
        pyqmd_free_particle p;
        p.x = e.x; p.y = e.y; p.vx = e.vx; p.vy = e.vy;
        p.age = 0.0;
        p.nucleus = e.nucleus;
        p.type = e.ptype;
        p.pad = 0;
        const double base = e.ptype == PYQMD_PT_ALPHA ? 30.0                                   // :298-305
                            : (e.ptype == PYQMD_PT_GAMMA ? 60.0
                               : ((e.ptype == PYQMD_PT_ELECTRON || e.ptype == PYQMD_PT_POSITRON) ? 50.0 : 40.0));
        const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(p.vx, p.vx), __dmul_rn(p.vy, p.vy)));   // :308
        if (mag > 0.001) {                                                                     // :309-314
            p.vx = __dmul_rn(__ddiv_rn(p.vx, mag), base);
            p.vy = __dmul_rn(__ddiv_rn(p.vy, mag), base);
        }
        if (f.fast_forward) {
            p.lifetime = f.lifetime_fast;                                                      // :320-338
        } else {                                                                               // :340-341
            const double dflt = e.ptype == PYQMD_PT_ALPHA ? 2.0                                // particles.py:31-38
                                : (e.ptype == PYQMD_PT_GAMMA ? 1.0
                                   : ((e.ptype == PYQMD_PT_ELECTRON || e.ptype == PYQMD_PT_POSITRON) ? 3.0
                                      : INFINITY));
            p.lifetime = fmax(dflt, f.lifetime_floor);
        }
        int remaining = f.num_steps - 1 - (int)(e.step - (int32_t)f.step0);
        remaining = max(0, min(remaining, f.num_steps));
        if (advance(p, remaining, f)) append(out, n_out, capacity, dropped, p);
    }
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int pyqmd_free_particles_frame(const pyqmd_free_particle* pool_in, const unsigned long long* n_in,
                                          pyqmd_free_particle* pool_out, unsigned long long* n_out,
                                          int64_t capacity, const pyqmd_decay_event* events,
                                          unsigned long long* event_count, int64_t event_capacity,
                                          const pyqmd_free_frame* frame, unsigned long long* dropped,
                                          int32_t reset_event_count, void* stream)
{
    PYQMD_REQUIRE(pool_in && n_in && pool_out && n_out && frame, "NULL pointer");
    PYQMD_REQUIRE(pool_in != pool_out && capacity >= 0, "pool_in and pool_out must differ");
    PYQMD_REQUIRE(frame->num_steps >= 0, "num_steps >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    PYQMD_CUDA_CHECK(cudaGetDevice(&dev));
    PYQMD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)(sms * 4);           // loop bounds live on the device: grid-stride
    PYQMD_CUDA_CHECK(cudaMemsetAsync(n_out, 0, sizeof(unsigned long long), st));
    free_advance_kernel<<<grid, 256, 0, st>>>(pool_in, n_in, pool_out, n_out, capacity, dropped, *frame);
    if (events && event_count) {
        free_spawn_kernel<<<grid, 256, 0, st>>>(events, event_count, event_capacity, pool_out, n_out,
                                                capacity, dropped, *frame);
        if (reset_event_count)
            PYQMD_CUDA_CHECK(cudaMemsetAsync(event_count, 0, sizeof(unsigned long long), st));
    }
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
