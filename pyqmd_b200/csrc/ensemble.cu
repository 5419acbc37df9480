// ensemble.cu -- ensembles of independent nuclei: fused decay -> force -> integrate,
// K sub-steps per launch with the nucleus resident in shared memory (sm_100a).
//
// Replaces, for many nuclei at once, the sub-step loop body of the reference
// (OtsoBear/PyQMD nuclear_sim.py:165-173): Nucleus.should_decay (particles.py:126-147),
// the physics slice of handle_decay (nuclear_sim.py:213,288-294,349,353) and
// NuclearForces.update_particles_cpu (nuclear_forces.py:236-323).
//
// Layout: a block of T threads holds G = T / cap nuclei (cap = largest nucleon count in the
// launch's size bin); thread t owns nucleon (t % cap) of nucleus (t / cap).  Positions and
// types live in shared memory as float4 (x, y, isProton, 0) so the j loop is one broadcast
// LDS.128 per pair; velocities and force accumulators stay in registers.  The update is
// Jacobi (double-buffered through registers + a barrier), like the reference CPU path and
// unlike its racy OpenCL kernel.  HBM is touched once on entry and once on exit, whatever
// n_steps is.
#include "common.cuh"
#include "decay_device.cuh"
#include "pair_law.cuh"

namespace pyqmd {

// Serial transmutation by the leader thread of one nucleus (rare event).
// Follows handle_decay's physics slice, nuclear_sim.py:213,215,288-294,349,353.
__device__ void leader_decay(const pyqmd_ensemble& e, const DrawSource& draws, float4* sp,
                             float2* sv, int gbase, int& cnt, int nuc, uint32_t step_abs,
                             uint32_t step_rel, int32_t& zn, double& T, double& p)
{
    const uint64_t gid = (uint64_t)(e.id_base + nuc);
    const pyqmd_nuclide_entry* cur = lookup(e.table, zn);
    int k = 0;
    if (cur->n_opt > 1) {                                   // decay_chains.py:218-229
        const double u1 = draws.one(gid, nuc, step_abs, step_rel, 1);
        k = pick_option(cur, u1);
    }
    const int mode = cur->opt_mode[k];
    if (mode == PYQMD_DECAY_NONE) return;                   // decay_chains.py:231-232; :215
    zn = cur->opt_zn[k];                                    // nuclear_sim.py:288-289

    // Nucleus.adjust_particles, particles.py:149-203
    if (mode == PYQMD_DECAY_BETA_MINUS || mode == PYQMD_DECAY_BETA_PLUS) {
        const float from = (mode == PYQMD_DECAY_BETA_MINUS) ? 0.0f : 1.0f;   // :158-171
        for (int j = 0; j < cnt; ++j) {
            if (sp[gbase + j].z == from) {
                sp[gbase + j].z = 1.0f - from;
                break;
            }
        }
    } else if (mode == PYQMD_DECAY_ALPHA || mode == PYQMD_DECAY_NEUTRON ||
               mode == PYQMD_DECAY_PROTON) {
        int rp = (mode == PYQMD_DECAY_ALPHA) ? 2 : (mode == PYQMD_DECAY_PROTON ? 1 : 0);
        int rn = (mode == PYQMD_DECAY_ALPHA) ? 2 : (mode == PYQMD_DECAY_NEUTRON ? 1 : 0);
        int w = 0;
        for (int j = 0; j < cnt; ++j) {                     // :183-198, order preserving
            const float4 q = sp[gbase + j];
            if (rp > 0 && q.z == 1.0f) { --rp; continue; }
            if (rn > 0 && q.z == 0.0f) { --rn; continue; }
            float2 v = sv[gbase + j];
            v.x *= 0.8f;                                    // :201-203
            v.y *= 0.8f;
            sp[gbase + w] = q;
            sv[gbase + w] = v;
            ++w;
        }
        cnt = w;
    }

    // Nucleus.update_center_of_mass, particles.py:205-208 (float64 accumulate, list order)
    double cx = 0.0, cy = 0.0;
    if (cnt > 0) {
        for (int j = 0; j < cnt; ++j) {
            cx += (double)sp[gbase + j].x;
            cy += (double)sp[gbase + j].y;
        }
        cx /= (double)cnt;
        cy /= (double)cnt;
    }
    if (e.origin) {
        cx += e.origin[2 * (int64_t)nuc];
        cy += e.origin[2 * (int64_t)nuc + 1];
    }

    // products(x, y), nuclear_sim.py:294 -> decay_chains.py:331-371
    int ptype = -1;
    double speed = 0.0, vx = 0.0, vy = 0.0;
    double u2 = 0.0, u3 = 0.0;
    draws.pair(gid, nuc, step_abs, step_rel, 1, u2, u3);
    if (emission_of(mode, ptype, speed)) {
        const double ang = __dmul_rn(6.283185307179586, u2);   // uniform(0, 2*pi)
        vx = speed * cos(ang);
        vy = speed * sin(ang);
    }
    if (e.event_count) {
        const unsigned long long slot = atomicAdd(e.event_count, 1ULL);
        if (e.events && (int64_t)slot < e.event_capacity) {
            pyqmd_decay_event ev;
            ev.nucleus = (int64_t)gid;
            ev.step = (int32_t)step_abs;
            ev.mode = mode;
            ev.zn_new = zn;
            ev.ptype = ptype;
            ev.x = cx; ev.y = cy; ev.vx = vx; ev.vy = vy;
            e.events[slot] = ev;
        }
    }
    if (e.mode_counts) atomicAdd(e.mode_counts + mode, 1ULL);

    // nucleus.stability = get_half_life(Z', N'), nuclear_sim.py:353
    bool used3;
    daughter_half_life(lookup(e.table, zn), u3, e.dt_decay, T, p, used3);
}

// Per-warp sums of the nucleon positions of a block (G == 1: the whole block is one nucleus),
// written next to the positions so that the centre of mass (nuclear_forces.py:242-243) needs no
// barrier of its own; summed by every thread in warp order, i.e. deterministically.
__device__ __forceinline__ void publish_warp_sum(float2* wsum, float x, float y, bool active)
{
    float sx = active ? x : 0.f, sy = active ? y : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = make_float2(sx, sy);
}

// N3 = true: every unordered pair is evaluated once (F_ij = -F_ji holds exactly for this law:
// it depends on d and on symmetric type predicates only) on a ring schedule -- nucleon i visits
// partners i+1 .. i+(n-1)/2 (mod n), plus i+n/2 for the lower half when n is even -- and the
// reaction is accumulated in a per-warp shared-memory row (no atomics, fixed order, so results
// are reproducible).  Halves the MUFU and FMA work per ordered pair.  N3 = false is the plain
// ordered-pair loop, kept for blocks of more than 256 threads where the per-warp reaction rows
// would not fit in shared memory.
template <int MAXT, bool N3>
__global__ void __launch_bounds__(MAXT) ensemble_kernel(const pyqmd_ensemble e, const LawParams L,
                                                         const int n_steps, const int G)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    const int nW = T >> 5;
    float4* sp = reinterpret_cast<float4*>(smem_raw);
    float2* sv = reinterpret_cast<float2*>(sp + T);
    float2* react = sv + T;                                  // [nW][T], N3 only
    float2* wsum = react + (N3 ? nW * T : 0);                // [nW]
    int* scnt = reinterpret_cast<int*>(wsum + nW);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int cap = e.cap;
    const int g = tid / cap;
    const int li = tid - g * cap;
    const int gbase = g * cap;
    const int64_t q = (int64_t)blockIdx.x * G + g;
    const bool has_nuc = (g < G) && (q < e.n_list);
    const int nuc = has_nuc ? (e.list ? e.list[q] : (int)q) : -1;
    const bool leader = has_nuc && li == 0;

    int cnt = 0;
    int64_t off = 0;
    if (has_nuc) {
        cnt = e.count[nuc];
        off = e.offset[nuc];
    }
    float x = 0.f, y = 0.f, tp = 0.f;
    float2 vel = make_float2(0.f, 0.f);
    if (has_nuc && li < cnt) {
        const float2 p2 = reinterpret_cast<const float2*>(e.pos)[off + li];
        vel = reinterpret_cast<const float2*>(e.vel)[off + li];
        x = p2.x; y = p2.y;
        tp = e.is_proton[off + li] ? 1.0f : 0.0f;
    }
    sp[tid] = make_float4(x, y, tp, 0.f);
    if (li == 0 && g < G) scnt[g] = cnt;
    if (N3)
        for (int w = 0; w < nW; ++w) react[w * T + tid] = make_float2(0.f, 0.f);
    publish_warp_sum(wsum, x, y, has_nuc && li < cnt);
    // warps that can hold nucleons of this thread's nucleus
    const int w_lo = gbase >> 5;
    const int w_hi = min((gbase + cap - 1) >> 5, nW - 1);

    // leader-held nucleus state
    int32_t zn = 0;
    double T_half = 0.0, p_dec = -1.0;
    if (leader && e.decay_enabled) {
        zn = e.zn[nuc];
        T_half = e.half_life[nuc];
        p_dec = e.p_decay[nuc];
    }
    DrawSource draws{e.uniforms, e.seed, e.uniforms_n};
    float R = 2.4f * cbrtf((float)cnt);                     // nuclear_forces.py:304

    for (int s = 0; s < n_steps; ++s) {
        // ---- decay test: Nucleus.should_decay, particles.py:126-147 --------------------------
        if (e.decay_enabled) {
            bool fire = false;
            const uint32_t step_abs = e.step0 + (uint32_t)s;
            if (leader && p_dec >= 0.0) {                   // stable: no draw (:129-130)
                const double u0 = draws.one((uint64_t)(e.id_base + nuc), nuc, step_abs, s, 0);
                fire = u0 < p_dec;                          // :147
            }
            if (__syncthreads_or(fire)) {
                sv[tid] = vel;
                __syncthreads();
                if (fire) {
                    leader_decay(e, draws, sp, sv, gbase, cnt, nuc, step_abs, s, zn, T_half, p_dec);
                    scnt[g] = cnt;
                }
                __syncthreads();
                if (g < G) cnt = scnt[g];
                vel = sv[tid];
                const float4 me = sp[tid];
                x = me.x; y = me.y; tp = me.z;
                R = 2.4f * cbrtf((float)cnt);
                publish_warp_sum(wsum, x, y, has_nuc && li < cnt);
                __syncthreads();
            }
        } else {
            __syncthreads();
        }

        const bool active = has_nuc && li < cnt;
        float fx = 0.f, fy = 0.f;
        float cx = 0.f, cy = 0.f;
        if (active) {
            // ---- centre of mass, nuclear_forces.py:242-243 ----------------------------------------
            const float4* tile = sp + gbase;
            if (e.centre) {                                 // caller-supplied `center`, :64
                cx = e.centre[2 * (int64_t)nuc];
                cy = e.centre[2 * (int64_t)nuc + 1];
            } else {
                float sx = 0.f, sy = 0.f;
                if (G == 1) {
                    for (int w = 0; w < nW; ++w) { sx += wsum[w].x; sy += wsum[w].y; }
                } else {
                    for (int j = 0; j < cnt; ++j) { sx += tile[j].x; sy += tile[j].y; }
                }
                const float inv_n = 1.0f / (float)cnt;
                cx = sx * inv_n;
                cy = sy * inv_n;
            }
            // ---- all-pairs force, nuclear_forces.py:248-298 ---------------------------------------
            if (N3) {
                float2* row = react + warp * T + gbase;
                const int half = (cnt - 1) >> 1;
                int j = li;
#pragma unroll 2
                for (int k = 0; k < half; ++k) {
                    j = (j + 1 == cnt) ? 0 : j + 1;         // partner (li + k + 1) mod cnt
                    const float4 o = tile[j];
                    const float dx = o.x - x, dy = o.y - y;
                    const float sc = pair_general(dx, dy, tp, o.z, L);
                    const float px = dx * sc, py = dy * sc;
                    fx += px;
                    fy += py;
                    float2 r = row[j];                      // reaction on the partner
                    r.x -= px;
                    r.y -= py;
                    row[j] = r;
                }
                if (!(cnt & 1) && li < (cnt >> 1)) {        // antipodal partner, even n
                    j = li + (cnt >> 1);
                    const float4 o = tile[j];
                    const float dx = o.x - x, dy = o.y - y;
                    const float sc = pair_general(dx, dy, tp, o.z, L);
                    const float px = dx * sc, py = dy * sc;
                    fx += px;
                    fy += py;
                    float2 r = row[j];
                    r.x -= px;
                    r.y -= py;
                    row[j] = r;
                }
            } else {
#pragma unroll 4
                for (int j = 0; j < cnt; ++j) {
                    const float4 o = tile[j];
                    const float dx = o.x - x, dy = o.y - y;
                    const float sc = pair_general(dx, dy, tp, o.z, L);
                    fx = fmaf(dx, sc, fx);
                    fy = fmaf(dy, sc, fy);
                }
            }
        }
        __syncthreads();                 // Jacobi: all reads (and all reactions) before any write
        if (active) {
            if (N3) {
                for (int w = w_lo; w <= w_hi; ++w) {        // fixed order: reproducible
                    const float2 r = react[w * T + tid];
                    react[w * T + tid] = make_float2(0.f, 0.f);
                    fx += r.x;
                    fy += r.y;
                }
            }
            contain_and_integrate(x, y, vel.x, vel.y, fx, fy, cx, cy, R, e.dt_phys);   // :301-323
            sp[tid] = make_float4(x, y, tp, 0.f);
            if (e.force && s == n_steps - 1)
                reinterpret_cast<float2*>(e.force)[off + li] = make_float2(fx, fy);
        }
        if (G == 1) publish_warp_sum(wsum, x, y, active);
    }

    if (has_nuc && li < cnt) {
        reinterpret_cast<float2*>(e.pos)[off + li] = make_float2(x, y);
        reinterpret_cast<float2*>(e.vel)[off + li] = vel;
        e.is_proton[off + li] = (tp != 0.f) ? 1 : 0;
    }
    if (leader) {
        e.count[nuc] = cnt;
        if (e.decay_enabled) {
            e.zn[nuc] = zn;
            e.half_life[nuc] = T_half;
            e.p_decay[nuc] = p_dec;
        }
    }
}

static int pick_block_threads(int cap, int* G_out)
{
    if (cap > 128) {
        *G_out = 1;
        return (cap + 31) / 32 * 32;
    }
    int best_T = 256, best_G = 256 / cap;
    double best_u = (double)best_G * cap / 256.0;
    for (int T = 224; T >= 128; T -= 32) {
        if (T < cap) break;
        const int G = T / cap;
        const double u = (double)G * cap / T;
        if (u > best_u + 1e-9) { best_u = u; best_T = T; best_G = G; }
    }
    *G_out = best_G;
    return best_T;
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int pyqmd_ensemble_step(const pyqmd_ensemble* e, int32_t n_steps, void* stream)
{
    PYQMD_REQUIRE(e != nullptr, "ensemble descriptor is NULL");
    PYQMD_REQUIRE(n_steps >= 0, "n_steps must be >= 0");
    PYQMD_REQUIRE(e->pos && e->vel && e->is_proton && e->offset && e->count, "state arrays");
    PYQMD_REQUIRE(e->cap >= 1 && e->cap <= 1024, "cap must be in [1, 1024]");
    if (e->decay_enabled)
        PYQMD_REQUIRE(e->zn && e->half_life && e->p_decay && e->table, "decay arrays / table");
    const int64_t n_list = e->list ? e->n_list : e->n_nuclei;
    if (n_list == 0 || n_steps == 0) return PYQMD_OK;
    pyqmd_ensemble d = *e;
    d.n_list = n_list;
    int G = 1;
    const int T = pick_block_threads(e->cap, &G);
    const int64_t grid = (n_list + G - 1) / G;
    PYQMD_REQUIRE(grid <= 2147483647LL, "too many nuclei for one launch");
    const int nW = T / 32;
    const bool n3 = T <= 256;
    const size_t smem = (size_t)T * (sizeof(float4) + sizeof(float2)) +
                        (n3 ? (size_t)nW * T * sizeof(float2) : 0) + (size_t)nW * sizeof(float2) +
                        (size_t)G * sizeof(int);
    const LawParams L = make_law_params(e->strong, e->coulomb, e->pauli);
    cudaStream_t st = (cudaStream_t)stream;
    if (n3)
        ensemble_kernel<256, true><<<(unsigned)grid, T, smem, st>>>(d, L, n_steps, G);
    else
        ensemble_kernel<1024, false><<<(unsigned)grid, T, smem, st>>>(d, L, n_steps, G);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
