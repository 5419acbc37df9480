// population.cu -- decay-only Monte Carlo over a population of particle-less nuclei
// (BASELINE config 5: 1e8 C-14 / U-238 nuclei, half-life statistics) for sm_100a.
//
// Per nucleus and sub-step this is the reference's light Nucleus.should_decay
// (OtsoBear/PyQMD decay_chains.py:400-421, same code as particles.py:126-147) followed, on a
// hit, by the (Z, N) / half-life part of handle_decay (nuclear_sim.py:213,288-289,353).
//
// One thread owns one nucleus for all n_steps sub-steps of a launch (state in registers,
// HBM read once; written back only for nuclei that actually decayed -- 20 B per nucleus-launch
// instead of 40).  Draws come from Philox4x32-10 keyed by (seed; global
// nucleus id, step, slot) or, for the bit-exact parity path, from a caller-supplied array.
// Per-step decay counts are reduced warp -> block -> one atomicAdd per block and column.
#include "common.cuh"
#include "decay_device.cuh"

namespace pyqmd {

constexpr int kPopThreads = 256;

__global__ void __launch_bounds__(kPopThreads) population_kernel(const pyqmd_population P,
                                                                 const int n_steps)
{
    __shared__ unsigned int scount[PYQMD_COUNT_COLS];
    const int64_t i = (int64_t)blockIdx.x * kPopThreads + threadIdx.x;
    const bool ok = i < P.n;
    int32_t zn = 0;
    double T = 0.0, p = -1.0;
    if (ok) {
        zn = P.zn[i];
        T = P.half_life[i];
        p = P.p_decay[i];
    }
    const DrawSource draws{P.uniforms, P.seed, P.uniforms_n};
    const uint64_t gid = (uint64_t)(P.id_base + i);
    bool dirty = false;

    for (int s = 0; s < n_steps; ++s) {
        const uint32_t step_abs = P.step0 + (uint32_t)s;
        bool fired = false;
        int mode = PYQMD_DECAY_NONE;
        int watch = -1;
        if (ok && p >= 0.0) {                               // stable: no draw, decay_chains.py:403
            double u0, u1;
            draws.pair(gid, i, step_abs, s, 0, u0, u1);
            fired = u0 < p;                                 // :421
            if (fired) {
                const pyqmd_nuclide_entry* cur = lookup(P.table, zn);
                const int k = pick_option(cur, u1);         // decay_chains.py:218-229
                mode = cur->opt_mode[k];
                if (mode != PYQMD_DECAY_NONE) {             // :231-232
                    for (int wch = 0; wch < P.n_watch; ++wch)
                        if (P.watch_zn[wch] == zn) watch = wch;
                    dirty = true;
                    zn = cur->opt_zn[k];                    // nuclear_sim.py:288-289
                    const pyqmd_nuclide_entry* nxt = lookup(P.table, zn);
                    double u3 = 0.0;
                    if (nxt->kind == PYQMD_HL_BAND) u3 = draws.one(gid, i, step_abs, s, 3);
                    bool used3;
                    daughter_half_life(nxt, u3, P.dt_decay, T, p, used3);   // nuclear_sim.py:353
                }
            }
        }
        if (P.decided && ok) P.decided[(int64_t)s * P.n + i] = fired ? 1 : 0;
        // block-aggregated counters; one barrier per sub-step when nothing in the block decayed
        const bool counted = fired && mode != PYQMD_DECAY_NONE;
        if (__syncthreads_or(counted)) {
            if (threadIdx.x < PYQMD_COUNT_COLS) scount[threadIdx.x] = 0;
            __syncthreads();
            if (counted) {
                atomicAdd(&scount[mode], 1u);
                if (watch >= 0) atomicAdd(&scount[8 + watch], 1u);
            }
            __syncthreads();
            if (threadIdx.x < PYQMD_COUNT_COLS && scount[threadIdx.x] && P.step_counts)
                atomicAdd(P.step_counts + (int64_t)s * PYQMD_COUNT_COLS + threadIdx.x,
                          (unsigned long long)scount[threadIdx.x]);
        }
    }
    if (ok && dirty) {
        P.zn[i] = zn;
        P.half_life[i] = T;
        P.p_decay[i] = p;
    }
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int pyqmd_population_step(const pyqmd_population* p, int32_t n_steps, void* stream)
{
    PYQMD_REQUIRE(p != nullptr, "population descriptor is NULL");
    PYQMD_REQUIRE(n_steps >= 0 && p->n >= 0, "n_steps, n >= 0");
    if (p->n == 0 || n_steps == 0) return PYQMD_OK;
    PYQMD_REQUIRE(p->zn && p->half_life && p->p_decay && p->table, "state arrays / table");
    PYQMD_REQUIRE(p->n_watch >= 0 && p->n_watch <= 8, "n_watch in [0, 8]");
    const int64_t blocks = (p->n + kPopThreads - 1) / kPopThreads;
    PYQMD_REQUIRE(blocks <= 2147483647LL, "population too large for one launch");
    population_kernel<<<(unsigned)blocks, kPopThreads, 0, (cudaStream_t)stream>>>(*p, n_steps);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
