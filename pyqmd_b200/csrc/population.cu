// population.cu -- decay-only Monte Carlo over a population of particle-less nuclei
// (BASELINE config 5: 1e8 C-14 / U-238 nuclei, half-life statistics) for sm_100a.
//
// Per nucleus and sub-step this is the reference's light Nucleus.should_decay
// (OtsoBear/PyQMD decay_chains.py:400-421, same code as particles.py:126-147) followed, on a
// hit, by the (Z, N) / half-life part of handle_decay (nuclear_sim.py:213,288-289,353).
//
// One thread owns TWO neighbouring nuclei (global ids 2k, 2k+1) for all n_steps sub-steps of a
// launch: one Philox4x32-10 call yields both every-step draws (DrawSource, slot 0), which halves the
// integer work that bounded the round-1 kernel (ncu r01h: ALU 41 %, issue 66 %, HBM 32 %).  State is
// kept in registers; HBM traffic per nucleus-launch is the 4-byte (Z, N) word: half-life and per-step
// probability of a tabulated nuclide come from its (L2/L1-resident) table row, the per-nucleus side
// arrays are read only for nuclides with an ESTIMATED half-life (10**uniform(a, b), one value per
// nucleus) or when the caller supplied its own per-nucleus values (PYQMD_POP_PER_NUCLEUS_STATE), and
// written only for nuclei that decayed.  Draws: Philox keyed by (seed; global id, step, slot) or, for
// the bit-exact parity path, a caller-supplied array.  Per-step decay counts are reduced
// block-wide -> one atomicAdd per block and column.
#include "common.cuh"
#include "decay_device.cuh"

namespace pyqmd {

constexpr int kPopThreads = 256;

struct PopNucleus {
    int i;            // local index, valid when ok (populations are < 2^31 nuclei per GPU)
    bool ok, dirty;
    int32_t zn;
    double T, p;      // T is only defined after a decay in this launch (it is never read before)
};

// table row of an in-range (Z, N): callers validate on the host, daughters come from the table
__device__ __forceinline__ const pyqmd_nuclide_entry* row_of(const pyqmd_nuclide_entry* table, int32_t zn)
{
    return table + ((zn >> 16) * PYQMD_TABLE_NDIM + (zn & 0xffff));
}

__device__ __forceinline__ void pop_load(const pyqmd_population& P, PopNucleus& a)
{
    a.dirty = false;
    a.zn = 0; a.T = 0.0; a.p = -1.0;
    if (!a.ok) return;
    a.zn = P.zn[a.i];
    const pyqmd_nuclide_entry* row = row_of(P.table, a.zn);
    if ((P.flags & PYQMD_POP_PER_NUCLEUS_STATE) || row->kind == PYQMD_HL_BAND)
        a.p = P.p_decay[a.i];
    else
        a.p = row->p_decay;
}

// should_decay with the draw u0 (decay_chains.py:400-421) and, on a hit, the (Z, N) / half-life part
// of handle_decay (nuclear_sim.py:213,288-289,353).  Returns the decay mode that was counted, or NONE.
__device__ __forceinline__ int pop_step(const pyqmd_population& P, const DrawSource& draws, PopNucleus& a,
                                        double u0, uint32_t step_abs, int s, bool& fired, int& watch)
{
    fired = false;
    watch = -1;
    if (!a.ok || !(a.p >= 0.0)) return PYQMD_DECAY_NONE;    // stable: no draw, decay_chains.py:403
    fired = u0 < a.p;                                       // :421
    if (!fired) return PYQMD_DECAY_NONE;
    const uint64_t gid = (uint64_t)(P.id_base + a.i);
    const pyqmd_nuclide_entry* cur = row_of(P.table, a.zn);
    int k = 0;
    if (cur->n_opt > 1) k = pick_option(cur, draws.one(gid, a.i, step_abs, s, 1));   // :218-229
    const int mode = cur->opt_mode[k];
    if (mode == PYQMD_DECAY_NONE) return mode;              // :231-232
    for (int wch = 0; wch < P.n_watch; ++wch)
        if (P.watch_zn[wch] == a.zn) watch = wch;
    a.dirty = true;
    a.zn = cur->opt_zn[k];                                  // nuclear_sim.py:288-289
    const pyqmd_nuclide_entry* nxt = lookup(P.table, a.zn);
    const double u3 = (nxt->kind == PYQMD_HL_BAND) ? draws.one(gid, a.i, step_abs, s, 3) : 0.0;
    bool used3;
    daughter_half_life(nxt, u3, P.dt_decay, a.T, a.p, used3);                        // nuclear_sim.py:353
    return mode;
}

__global__ void __launch_bounds__(kPopThreads) population_kernel(const pyqmd_population P,
                                                                 const int n_steps)
{
    __shared__ unsigned int scount[PYQMD_COUNT_COLS];
    // global pair index: both nuclei of a pair share one Philox counter, whatever the sharding
    const int j = (int)(blockIdx.x * kPopThreads + threadIdx.x);
    const uint64_t pair = (uint64_t)(P.id_base >> 1) + (uint64_t)j;
    const int n = (int)P.n;
    PopNucleus a, b;
    a.i = 2 * j - (int)(P.id_base & 1);
    b.i = a.i + 1;
    a.ok = a.i >= 0 && a.i < n;
    b.ok = b.i < n;
    pop_load(P, a);
    pop_load(P, b);
    const DrawSource draws{P.uniforms, P.seed, P.uniforms_n};

    for (int s = 0; s < n_steps; ++s) {
        const uint32_t step_abs = P.step0 + (uint32_t)s;
        double ua = 0.0, ub = 0.0;
        if (P.uniforms) {
            if (a.ok) ua = draws.one(0, a.i, step_abs, s, 0);
            if (b.ok) ub = draws.one(0, b.i, step_abs, s, 0);
        } else {
            // unconditional (a stable nucleus simply ignores its draw): the integer work of Philox does
            // not wait for the zn -> table-row loads and hides their latency
            draws.slot0_pair(pair, step_abs, ua, ub);
        }
        bool fa, fb;
        int wa, wb;
        const int ma = pop_step(P, draws, a, ua, step_abs, s, fa, wa);
        const int mb = pop_step(P, draws, b, ub, step_abs, s, fb, wb);
        if (P.decided) {
            if (a.ok) P.decided[(int64_t)s * n + a.i] = fa ? 1 : 0;
            if (b.ok) P.decided[(int64_t)s * n + b.i] = fb ? 1 : 0;
        }
        // block-aggregated counters; one barrier per sub-step when nothing in the block decayed
        const bool counted = ma != PYQMD_DECAY_NONE || mb != PYQMD_DECAY_NONE;
        if (__syncthreads_or(counted)) {
            if (threadIdx.x < PYQMD_COUNT_COLS) scount[threadIdx.x] = 0;
            __syncthreads();
            if (ma != PYQMD_DECAY_NONE) {
                atomicAdd(&scount[ma], 1u);
                if (wa >= 0) atomicAdd(&scount[8 + wa], 1u);
            }
            if (mb != PYQMD_DECAY_NONE) {
                atomicAdd(&scount[mb], 1u);
                if (wb >= 0) atomicAdd(&scount[8 + wb], 1u);
            }
            __syncthreads();
            if (threadIdx.x < PYQMD_COUNT_COLS && scount[threadIdx.x] && P.step_counts)
                atomicAdd(P.step_counts + (int64_t)s * PYQMD_COUNT_COLS + threadIdx.x,
                          (unsigned long long)scount[threadIdx.x]);
        }
    }
    if (a.ok && a.dirty) { P.zn[a.i] = a.zn; P.half_life[a.i] = a.T; P.p_decay[a.i] = a.p; }
    if (b.ok && b.dirty) { P.zn[b.i] = b.zn; P.half_life[b.i] = b.T; P.p_decay[b.i] = b.p; }
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int pyqmd_population_step(const pyqmd_population* p, int32_t n_steps, void* stream)
{
    PYQMD_REQUIRE(p != nullptr, "population descriptor is NULL");
    PYQMD_REQUIRE(n_steps >= 0 && p->n >= 0 && p->id_base >= 0, "n_steps, n, id_base >= 0");
    PYQMD_REQUIRE(p->n < 2147483647LL / 2, "at most 2^30 nuclei per launch");
    if (p->n == 0 || n_steps == 0) return PYQMD_OK;
    PYQMD_REQUIRE(p->zn && p->half_life && p->p_decay && p->table, "state arrays / table");
    PYQMD_REQUIRE(p->n_watch >= 0 && p->n_watch <= 8, "n_watch in [0, 8]");
    // pairs of global ids covered by [id_base, id_base + n): one thread each
    const int64_t n_pairs = ((p->id_base + p->n + 1) >> 1) - (p->id_base >> 1);
    const int64_t blocks = (n_pairs + kPopThreads - 1) / kPopThreads;
    PYQMD_REQUIRE(blocks <= 2147483647LL, "population too large for one launch");
    population_kernel<<<(unsigned)blocks, kPopThreads, 0, (cudaStream_t)stream>>>(*p, n_steps);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
