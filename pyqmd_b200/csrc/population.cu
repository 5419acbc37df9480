// population.cu -- decay-only Monte Carlo over a population of particle-less nuclei
// (BASELINE config 5: 1e8 C-14 / U-238 nuclei, half-life statistics) for sm_100a.
//
// Per nucleus and sub-step this is the reference's light Nucleus.should_decay
// (OtsoBear/PyQMD decay_chains.py:400-421, same code as particles.py:126-147) followed, on a
// hit, by the (Z, N) / half-life part of handle_decay (nuclear_sim.py:213,288-289,353).
//
// One thread owns FOUR neighbouring nuclei (global ids 4k .. 4k+3) for all n_steps sub-steps of a
// launch: a Philox4x32-10 call yields the every-step draws of two nuclei (DrawSource, slot 0), which
// halves the integer work that bounded the round-1 kernel (ncu r01h: ALU 41 %, issue 66 %, HBM 32 %),
// and the two independent calls plus one 16-byte load of the four (Z, N) words give the scheduler
// something to overlap with the dependent table-row loads.  State is
// kept in registers; HBM traffic per nucleus-launch is the 4-byte (Z, N) word: half-life and per-step
// probability of a tabulated nuclide come from its (L2/L1-resident) table row, the per-nucleus side
// arrays are read only for nuclides with an ESTIMATED half-life (10**uniform(a, b), one value per
// nucleus) or when the caller supplied its own per-nucleus values (PYQMD_POP_PER_NUCLEUS_STATE), and
// written only for nuclei that decayed.  Draws: Philox keyed by (seed; global id, step, slot) or, for
// the bit-exact parity path, a caller-supplied array.  Per-step decay counts: shared-memory atomics
// into a block-local table, flushed once per launch.
#include "common.cuh"
#include "decay_device.cuh"

namespace pyqmd {

#ifndef PYQMD_POP_MINBLOCKS
#define PYQMD_POP_MINBLOCKS 2
#endif
constexpr int kPopThreads = 256;
constexpr int kNPT = 4;           // nuclei per thread: global ids 4k .. 4k+3, two Philox calls

// table row of an in-range (Z, N): callers validate on the host, daughters come from the table
__device__ __forceinline__ const pyqmd_nuclide_entry* row_of(const pyqmd_nuclide_entry* table, int32_t zn)
{
    return table + ((zn >> 16) * PYQMD_TABLE_NDIM + (zn & 0xffff));
}

struct DecayOut {
    double p;
    int32_t zn;
    int mode, watch;
};

// The (Z, N) / half-life part of handle_decay (nuclear_sim.py:213,288-289,353) for a nucleus whose
// should_decay fired; writes the new state back at once.  Rare and heavy (float64 pow for estimated
// half-lives): out of line and by value, so that the every-step path keeps its state in few registers.
__device__ __noinline__ DecayOut pop_decay(const pyqmd_population& P, const DrawSource draws, int i,
                                           int32_t zn, double p, uint32_t step_abs, int s)
{
    DecayOut o;
    o.p = p; o.zn = zn; o.mode = PYQMD_DECAY_NONE; o.watch = -1;
    const uint64_t gid = (uint64_t)(P.id_base + i);
    const pyqmd_nuclide_entry* cur = row_of(P.table, zn);
    int k = 0;
    if (cur->n_opt > 1) k = pick_option(cur, draws.one(gid, i, step_abs, s, 1));     // :218-229
    const int mode = cur->opt_mode[k];
    if (mode == PYQMD_DECAY_NONE) return o;                 // :231-232
    for (int wch = 0; wch < P.n_watch; ++wch)
        if (P.watch_zn[wch] == zn) o.watch = wch;
    o.mode = mode;
    o.zn = cur->opt_zn[k];                                  // nuclear_sim.py:288-289
    const pyqmd_nuclide_entry* nxt = lookup(P.table, o.zn);
    const double u3 = (nxt->kind == PYQMD_HL_BAND) ? draws.one(gid, i, step_abs, s, 3) : 0.0;
    bool used3;
    double T;
    daughter_half_life(nxt, u3, P.dt_decay, T, o.p, used3); // nuclear_sim.py:353
    P.zn[i] = o.zn;
    P.half_life[i] = T;
    P.p_decay[i] = o.p;
    return o;
}

constexpr int kSmemSteps = 64;    // sub-steps per launch whose counters fit the block-local table

struct Quad {
    int32_t zn[kNPT];
    bool ok[kNPT];
};

__device__ __forceinline__ Quad load_quad(const pyqmd_population& P, int64_t q, int n, bool aligned)
{
    Quad r;
    const int64_t i0 = kNPT * q - (P.id_base & 3);
    if (aligned && i0 + kNPT <= n) {                        // one 16-byte load
        const int4 v = *reinterpret_cast<const int4*>(P.zn + i0);
        r.zn[0] = v.x; r.zn[1] = v.y; r.zn[2] = v.z; r.zn[3] = v.w;
#pragma unroll
        for (int k = 0; k < kNPT; ++k) r.ok[k] = true;
    } else {
#pragma unroll
        for (int k = 0; k < kNPT; ++k) {
            r.ok[k] = i0 + k >= 0 && i0 + k < n;
            r.zn[k] = r.ok[k] ? P.zn[i0 + k] : 0;
        }
    }
    return r;
}

// Persistent grid (a few blocks per SM), grid-stride over quads, the next quad's (Z, N) words
// prefetched while the current one is processed: the kernel no longer pays one DRAM round trip plus
// one block launch per 256 threads of useful work (ncu r02b: 32 % of the stall samples sat on the
// first load, 9 % on the per-step barrier).  Decay counts go to a block-local table with shared-memory
// atomics and are flushed once at the end.
//
// `random() < p` is decided in integers: the draw is the 53-bit integer m of CPython's m / 2^53, the
// nuclide table carries p_thr = ceil(p 2^53), and m < p_thr is the same predicate exactly -- no
// int -> float64 conversions, no float64 arithmetic on the every-step path (ncu r02d: 87 instructions
// per nucleus-step, a third of them conversions, float64 compares and per-nucleus bookkeeping).
// FAST = Philox draws, no per-nucleus `decided` output, every quad complete and 16-byte aligned: the
// production case, without the predicates of the general one.
template <bool FAST>
__global__ void __launch_bounds__(kPopThreads, PYQMD_POP_MINBLOCKS) population_kernel(const pyqmd_population P,
                                                                    const int n_steps, const int64_t n_quads)
{
    __shared__ unsigned int scount[kSmemSteps][PYQMD_COUNT_COLS];
    const bool local_counts = n_steps <= kSmemSteps;
    if (local_counts) {
        for (int k = threadIdx.x; k < n_steps * PYQMD_COUNT_COLS; k += kPopThreads)
            (&scount[0][0])[k] = 0;
        __syncthreads();
    }
    const int n = (int)P.n;
    const bool aligned = FAST || (P.id_base & 3) == 0;
    const DrawSource draws{P.uniforms, P.seed, P.uniforms_n};
    const int64_t stride = (int64_t)gridDim.x * kPopThreads;
    int64_t q = (int64_t)blockIdx.x * kPopThreads + threadIdx.x;
    Quad cur;
    if (q < n_quads) cur = load_quad(P, q, n, aligned);
    for (; q < n_quads; q += stride) {
        const Quad me = cur;
        // prefetch: the next quad's words travel from DRAM while this one is processed (a second stage
        // for the dependent table-row loads was measured and did not pay: 2.19e11 vs 2.37e11 steps/s)
        if (q + stride < n_quads) cur = load_quad(P, q + stride, n, aligned);
        const int i0 = (int)(kNPT * q - (FAST ? 0 : (P.id_base & 3)));   // local index of the first nucleus
        const uint64_t quad = (uint64_t)(P.id_base >> 2) + (uint64_t)q;  // global quad id
        int32_t zn[kNPT];
        uint64_t thr[kNPT];
        // threshold of the per-step probability: the table row's, unless the half-life is a per-nucleus
        // estimate (the row says so) or the caller supplied its own per-nucleus values
#pragma unroll
        for (int k = 0; k < kNPT; ++k) {
            zn[k] = me.zn[k];
            thr[k] = 0;
            if (FAST || me.ok[k]) {
                thr[k] = row_of(P.table, zn[k])->p_thr;
                if ((P.flags & PYQMD_POP_PER_NUCLEUS_STATE) || thr[k] == PYQMD_THR_PER_NUCLEUS)
                    thr[k] = decay_threshold(P.p_decay[i0 + k]);
            }
        }
        for (int s = 0; s < n_steps; ++s) {
            const uint32_t step_abs = P.step0 + (uint32_t)s;
            uint64_t m[kNPT] = {0, 0, 0, 0};
            if (FAST || !P.uniforms) {
                // unconditional (a stable nucleus ignores its draw): Philox does not wait for the loads
                draws.slot0_pair_bits(2 * quad, step_abs, m[0], m[1]);
                draws.slot0_pair_bits(2 * quad + 1, step_abs, m[2], m[3]);
            }
#pragma unroll
            for (int k = 0; k < kNPT; ++k) {
                // stable (threshold 0): never; otherwise random() < p, decay_chains.py:421
                bool fired;
                if (!FAST && P.uniforms) {
                    // parity path: caller-supplied doubles (not necessarily multiples of 2^-53) against
                    // the nucleus' probability itself (the side array is kept current by pop_decay)
                    const double p = (thr[k] == 0) ? -1.0 : P.p_decay[i0 + k];
                    fired = me.ok[k] && p >= 0.0 && draws.one(0, i0 + k, step_abs, s, 0) < p;
                } else {
                    fired = (FAST || me.ok[k]) && m[k] < thr[k];
                }
                if (!FAST && P.decided && me.ok[k]) P.decided[(int64_t)s * n + i0 + k] = fired ? 1 : 0;
                if (fired) {
                    const DecayOut o = pop_decay(P, draws, i0 + k, zn[k], 0.0, step_abs, s);
                    zn[k] = o.zn;
                    thr[k] = (o.mode == PYQMD_DECAY_NONE) ? thr[k] : decay_threshold(o.p);
                    if (o.mode != PYQMD_DECAY_NONE && P.step_counts) {
                        if (local_counts) {
                            atomicAdd(&scount[s][o.mode], 1u);
                            if (o.watch >= 0) atomicAdd(&scount[s][8 + o.watch], 1u);
                        } else {
                            unsigned long long* row = P.step_counts + (int64_t)s * PYQMD_COUNT_COLS;
                            atomicAdd(row + o.mode, 1ULL);
                            if (o.watch >= 0) atomicAdd(row + 8 + o.watch, 1ULL);
                        }
                    }
                }
            }
        }
    }
    if (local_counts && P.step_counts) {
        __syncthreads();
        for (int k = threadIdx.x; k < n_steps * PYQMD_COUNT_COLS; k += kPopThreads) {
            const unsigned int c = (&scount[0][0])[k];
            if (c) atomicAdd(P.step_counts + k, (unsigned long long)c);
        }
    }
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int pyqmd_population_step(const pyqmd_population* p, int32_t n_steps, void* stream)
{
    PYQMD_REQUIRE(p != nullptr, "population descriptor is NULL");
    PYQMD_REQUIRE(n_steps >= 0 && p->n >= 0 && p->id_base >= 0, "n_steps, n, id_base >= 0");
    PYQMD_REQUIRE(p->n < 2147483647LL / 2, "at most 2^30 nuclei per launch");
    PYQMD_REQUIRE((reinterpret_cast<uintptr_t>(p->zn) & 15) == 0, "zn must be 16-byte aligned");
    if (p->n == 0 || n_steps == 0) return PYQMD_OK;
    PYQMD_REQUIRE(p->zn && p->half_life && p->p_decay && p->table, "state arrays / table");
    PYQMD_REQUIRE(p->n_watch >= 0 && p->n_watch <= 8, "n_watch in [0, 8]");
    // quads of global ids covered by [id_base, id_base + n): grid-stride over them
    const int64_t n_quads = ((p->id_base + p->n + 3) >> 2) - (p->id_base >> 2);
    int dev = 0, sms = 0;
    PYQMD_CUDA_CHECK(cudaGetDevice(&dev));
    PYQMD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t blocks = (n_quads + kPopThreads - 1) / kPopThreads;
    if (blocks > (int64_t)sms * PYQMD_POP_MINBLOCKS) blocks = (int64_t)sms * PYQMD_POP_MINBLOCKS;   // persistent grid
    const bool fast = !p->uniforms && !p->decided && (p->id_base & 3) == 0 && (p->n & 3) == 0;
    if (fast)
        population_kernel<true><<<(unsigned)blocks, kPopThreads, 0, (cudaStream_t)stream>>>(*p, n_steps, n_quads);
    else
        population_kernel<false><<<(unsigned)blocks, kPopThreads, 0, (cudaStream_t)stream>>>(*p, n_steps, n_quads);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
