// decay_device.cuh -- device side of the per-nucleus stochastic decay (sm_100a).
//
// Behavioural spec (reference root = OtsoBear/PyQMD):
//   Nucleus.should_decay          particles.py:126-147  (copy at decay_chains.py:400-421)
//   get_decay_product             decay_chains.py:203-245 (branch pick :218-229)
//   get_half_life                 decay_chains.py:247-328 (estimate bands :309-328)
//   Nucleus.adjust_particles      particles.py:149-203
//   create_alpha .. create_proton decay_chains.py:331-371
// The (Z,N) -> {half-life class, options} table is built on the host (pyqmd_b200/nuclides.py)
// so every float64 the reference would compute with libm (probabilities, cumulative branch
// sums) reaches the device bit-for-bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pyqmd_b200.h"

namespace pyqmd {

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based RNG -------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform from two 32-bit words, the same map as CPython's random.random()
// (Modules/_randommodule.c): ((w0 >> 5) * 2^26 + (w1 >> 6)) / 2^53.
__device__ __forceinline__ double u53(uint32_t w0, uint32_t w1)
{
    return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) * (1.0 / 9007199254740992.0);
}

// Draw source for one nucleus-step.  Slot meaning: 0 should_decay (particles.py:147),
// 1 branch (decay_chains.py:221), 2 emission angle (:332..367), 3 half-life estimate (:312-328).
// With `uniforms` != nullptr the draws are read from a caller-supplied [step][nucleus][4] array
// (the bit-exact parity path); otherwise Philox4x32-10 keyed by seed, counters built from GLOBAL ids
// -- independent of how nuclei are sharded over GPUs:
//   slot 0     counter (id >> 1, step, 0): words (0,1) for even ids, (2,3) for odd ids -- ONE call
//              serves the every-step draw of two neighbouring nuclei (the decay-only population
//              kernel keeps two nuclei per thread for exactly this reason)
//   slots 1,2  counter (id, step, 1): words (0,1) / (2,3)     -- only when a nucleus decays
//   slot 3     counter (id, step, 2): words (0,1)             -- only for estimated half-lives
// (oracle twin: orc_philox_uniform, oracle/pyqmd_oracle.c)
struct DrawSource {
    const double* uniforms;   // optional [step_rel][uniforms_n][4]
    uint64_t seed;
    int64_t uniforms_n;       // nuclei per step in the uniforms array

    // the same two draws as 53-bit integers m (u = m / 2^53): `u < p` is exactly `m < ceil(p 2^53)`
    __device__ __forceinline__ void slot0_pair_bits(uint64_t pair_id, uint32_t step_abs, uint64_t& m_even,
                                                    uint64_t& m_odd) const
    {
        uint32_t w[4];
        philox4x32_10((uint32_t)pair_id, (uint32_t)(pair_id >> 32), step_abs, 0u, (uint32_t)seed,
                      (uint32_t)(seed >> 32), w);
        m_even = ((uint64_t)(w[0] >> 5) << 26) | (uint64_t)(w[1] >> 6);
        m_odd = ((uint64_t)(w[2] >> 5) << 26) | (uint64_t)(w[3] >> 6);
    }

    // slot-0 draws of the nuclei 2 * pair_id and 2 * pair_id + 1 from one Philox call
    __device__ __forceinline__ void slot0_pair(uint64_t pair_id, uint32_t step_abs, double& u_even,
                                               double& u_odd) const
    {
        uint32_t w[4];
        philox4x32_10((uint32_t)pair_id, (uint32_t)(pair_id >> 32), step_abs, 0u, (uint32_t)seed,
                      (uint32_t)(seed >> 32), w);
        u_even = u53(w[0], w[1]);
        u_odd = u53(w[2], w[3]);
    }

    // id_global / step_abs feed the Philox counter; id_local / step_rel index `uniforms`.
    __device__ __forceinline__ double one(uint64_t id_global, int64_t id_local, uint32_t step_abs,
                                          uint32_t step_rel, uint32_t slot) const
    {
        if (uniforms) return uniforms[((int64_t)step_rel * uniforms_n + id_local) * 4 + slot];
        if (slot == 0) {
            double a, b;
            slot0_pair(id_global >> 1, step_abs, a, b);
            return (id_global & 1) ? b : a;
        }
        uint32_t w[4];
        philox4x32_10((uint32_t)id_global, (uint32_t)(id_global >> 32), step_abs, slot == 3 ? 2u : 1u,
                      (uint32_t)seed, (uint32_t)(seed >> 32), w);
        return (slot == 2) ? u53(w[2], w[3]) : u53(w[0], w[1]);
    }
};

// ceil(p 2^53) for a per-nucleus probability (estimated half-lives, caller-supplied values)
__device__ __forceinline__ uint64_t decay_threshold(double p)
{
    return (p > 0.0) ? __double2ull_ru(p * 9007199254740992.0) : 0ull;
}

__device__ __forceinline__ const pyqmd_nuclide_entry* lookup(const pyqmd_nuclide_entry* table,
                                                              int32_t zn)
{
    int z = zn >> 16, n = zn & 0xffff;
    z = min(max(z, 0), PYQMD_TABLE_ZDIM - 1);
    n = min(max(n, 0), PYQMD_TABLE_NDIM - 1);
    return table + z * PYQMD_TABLE_NDIM + n;
}

// Decay probability for a half-life that is not in the host table (estimated per nucleus):
// particles.py:134-144 evaluated on the device.  The linear branch is plain IEEE mul/div
// (bit-exact); the direct branch uses CUDA's pow (<= 2 ulp from glibc's -- documented).
__device__ __forceinline__ double decay_probability_device(double T, double dt)
{
    if (isinf(T)) return -1.0;
    double p;
    if (dt > __dmul_rn(T, 0.01))
        p = 1.0 - pow(0.5, __ddiv_rn(dt, T));
    else
        p = __dmul_rn(__ddiv_rn(0.693, T), dt);
    p = (p < 1.0) ? p : 1.0;
    p = (p > 0.0) ? p : 0.0;
    return p;
}

// Branch pick, decay_chains.py:218-229.  Returns the option index.
__device__ __forceinline__ int pick_option(const pyqmd_nuclide_entry* e, double r)
{
    if (e->n_opt <= 1) return 0;
    for (int k = 0; k < e->n_opt; ++k)
        if (r <= e->opt_cum[k]) return k;
    return 0;
}

// Half-life and per-step probability of a (daughter) nuclide; slot-3 draw `u3` used only for
// the estimate bands (decay_chains.py:311-328): 10 ** uniform(a, b) * unit.
__device__ __forceinline__ void daughter_half_life(const pyqmd_nuclide_entry* e, double u3,
                                                   double dt_decay, double& T, double& p,
                                                   bool& used_draw)
{
    used_draw = false;
    if (e->kind == PYQMD_HL_BAND) {
        const double ex = __dadd_rn(e->band_a, __dmul_rn(__dadd_rn(e->band_b, -e->band_a), u3));
        T = __dmul_rn(pow(10.0, ex), e->band_unit);
        p = decay_probability_device(T, dt_decay);
        used_draw = true;
    } else {
        T = e->half_life;
        p = e->p_decay;
    }
}

// Emitted particle type and speed per decay mode (decay_chains.py:331-371);
// returns false when the mode emits nothing here (NONE, fission stub).
__device__ __forceinline__ bool emission_of(int mode, int& ptype, double& speed)
{
    switch (mode) {
        case PYQMD_DECAY_ALPHA:       ptype = PYQMD_PT_ALPHA;    speed = 100.0; return true;
        case PYQMD_DECAY_BETA_MINUS:  ptype = PYQMD_PT_ELECTRON; speed = 150.0; return true;
        case PYQMD_DECAY_BETA_PLUS:   ptype = PYQMD_PT_POSITRON; speed = 150.0; return true;
        case PYQMD_DECAY_GAMMA:       ptype = PYQMD_PT_GAMMA;    speed = 200.0; return true;
        case PYQMD_DECAY_NEUTRON:     ptype = PYQMD_PT_NEUTRON;  speed = 60.0;  return true;
        case PYQMD_DECAY_PROTON:      ptype = PYQMD_PT_PROTON;   speed = 50.0;  return true;
        default: return false;
    }
}

}  // namespace pyqmd
