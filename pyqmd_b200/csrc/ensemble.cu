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

// ---------------------------------------------------------------------------------------------------
// ensemble_pair_kernel: the production kernel for nuclei of up to 512 nucleons.
//
// Each thread owns TWO nucleons (2t, 2t+1) of its nucleus, so every evaluation of the law is a
// packed f32x2 evaluation of (i_a, j) and (i_b, j): all FMA-pipe arithmetic issues as
// FADD2/FMUL2/FFMA2 (half the issue slots), compares/selects/MUFU stay per element.  Newton's
// third law is used on a ring of "super-particles" (= the nucleon pairs of the threads): thread t
// visits super-partners t+1 .. t+(m-1)/2 (mod m) -- plus the antipode for the lower half when m is
// even -- and its own pair (2t, 2t+1) once; the reaction on partner j (summed over i_a, i_b) goes
// to a per-warp shared-memory row (no atomics; fixed order => bit-reproducible).
//
// Shared memory, per block of G nuclei x capT threads (capS = 2 capT slots per nucleus):
//   A4   float4[2 * G * capS]  (x, x, y, y) per slot, canonical [0, M) then mirror [M, 2M) so the
//                              ring never wraps (M = cnt rounded up to even)
//   T2   float2[2 * G * capS]  (isProton, isProton), same indexing
//   spc  float4[G * capS], sv float2[G * capS]   canonical staging used only when a nucleus decays
//   react float2[nW][G * capS] per-warp reaction rows
//   wsum float4[nW]            per-warp position sums (first / second nucleus present in the warp)
// An odd nucleon count is padded with a ghost neutron parked at (1e5, 1e5): every term of the law
// is exactly 0 at that distance, so it needs no masking in the pair loop.
constexpr float kGhost = 1.0e5f;

struct PairSmem {
    float4* A4;
    float2* T2;
    float4* spc;
    float2* sv;
    float2* react;
    float4* wsum;
    int* scnt;
};

// Slots are stored split by parity (even slots of a nucleus first, then the odd ones) so that the
// lanes of a warp, which visit slots 2(t+k) resp. 2(t+k)+1, touch consecutive 16-byte words: no
// shared-memory bank conflicts (ncu r01c showed 2-way conflicts with the interleaved layout).
// Within each parity half, super-index u = s >> 1 runs over [0, m) and is mirrored at [m, 2m).
__device__ __forceinline__ void put_slot(const PairSmem& S, int gb2, int capS, int m, int s, float x,
                                         float y, float tp)
{
    const float4 a = make_float4(x, x, y, y);
    const float2 t = make_float2(tp, tp);
    const int idx = gb2 + (s & 1) * capS + (s >> 1);
    S.A4[idx] = a;
    S.A4[idx + m] = a;
    S.T2[idx] = t;
    S.T2[idx + m] = t;
}

// per-warp position sums, split by nucleus (a warp spans at most two nuclei when capT >= 32)
__device__ __forceinline__ void publish_pair_sums(float4* wsum, int capT, int g, float vx, float vy)
{
    const int gfirst = ((threadIdx.x & ~31)) / capT;
    const bool first = (g == gfirst);
    float a = first ? vx : 0.f, b = first ? vy : 0.f;
    float c = first ? 0.f : vx, d = first ? 0.f : vy;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = make_float4(a, b, c, d);
}

#ifndef PYQMD_ENS_MINBLOCKS
#define PYQMD_ENS_MINBLOCKS 4      // A/B on B200: 2 -> 1.01e12, 3 -> 1.07e12, 4 -> 1.12e12, 5 -> spills
#endif
template <int MAXT>
__global__ void __launch_bounds__(MAXT, PYQMD_ENS_MINBLOCKS) ensemble_pair_kernel(const pyqmd_ensemble e,
                                                              const LawParams L, const int n_steps,
                                                              const int G, const int capT)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    const int nW = T >> 5;
    const int capS = 2 * capT;
    const int nSlots = G * capS;
    PairSmem S;
    S.A4 = reinterpret_cast<float4*>(smem_raw);
    S.spc = S.A4 + 2 * nSlots;
    S.wsum = S.spc + nSlots;
    S.T2 = reinterpret_cast<float2*>(S.wsum + nW);
    S.sv = S.T2 + 2 * nSlots;
    S.react = S.sv + nSlots;
    S.scnt = reinterpret_cast<int*>(S.react + nW * nSlots);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int g = tid / capT;
    const int t = tid - g * capT;
    const int gb = g * capS;             // canonical slot base (spc, sv, react)
    const int gb2 = 2 * gb;              // base in the mirrored arrays
    const int s0 = 2 * t, s1 = 2 * t + 1;
    const int64_t q = (int64_t)blockIdx.x * G + g;
    const bool has_nuc = (g < G) && (q < e.n_list);
    const int nuc = has_nuc ? (e.list ? e.list[q] : (int)q) : -1;
    const bool leader = has_nuc && t == 0;

    int cnt = 0;
    int64_t off = 0;
    if (has_nuc) {
        cnt = e.count[nuc];
        off = e.offset[nuc];
    }
    int m = (cnt + 1) >> 1, M = 2 * m;
    bool active = has_nuc && t < m;
    bool real1 = active && s1 < cnt;
    float x0 = kGhost, y0 = kGhost, t0 = 0.f, x1 = kGhost, y1 = kGhost, t1 = 0.f;
    float2 v0 = make_float2(0.f, 0.f), v1 = make_float2(0.f, 0.f);
    if (active) {
        const float2 p = reinterpret_cast<const float2*>(e.pos)[off + s0];
        v0 = reinterpret_cast<const float2*>(e.vel)[off + s0];
        x0 = p.x; y0 = p.y;
        t0 = e.is_proton[off + s0] ? 1.0f : 0.0f;
    }
    if (real1) {
        const float2 p = reinterpret_cast<const float2*>(e.pos)[off + s1];
        v1 = reinterpret_cast<const float2*>(e.vel)[off + s1];
        x1 = p.x; y1 = p.y;
        t1 = e.is_proton[off + s1] ? 1.0f : 0.0f;
    }
    if (active) {
        put_slot(S, gb2, capS, m, s0, x0, y0, t0);
        put_slot(S, gb2, capS, m, s1, x1, y1, t1);
    }
    if (t == 0 && g < G) S.scnt[g] = cnt;
    for (int k = tid; k < nW * nSlots; k += T) S.react[k] = make_float2(0.f, 0.f);
    publish_pair_sums(S.wsum, capT, g, (active ? x0 : 0.f) + (real1 ? x1 : 0.f),
                      (active ? y0 : 0.f) + (real1 ? y1 : 0.f));
    const int w_lo = (g * capT) >> 5;
    const int w_hi = min((g * capT + capT - 1) >> 5, nW - 1);

    int32_t zn = 0;
    double T_half = 0.0, p_dec = -1.0;
    if (leader && e.decay_enabled) {
        zn = e.zn[nuc];
        T_half = e.half_life[nuc];
        p_dec = e.p_decay[nuc];
    }
    const DrawSource draws{e.uniforms, e.seed, e.uniforms_n};
    const GenConsts gc = make_gen_consts(L);
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
                if (active) {
                    S.spc[gb + s0] = make_float4(x0, y0, t0, 0.f);
                    S.sv[gb + s0] = v0;
                }
                if (real1) {
                    S.spc[gb + s1] = make_float4(x1, y1, t1, 0.f);
                    S.sv[gb + s1] = v1;
                }
                __syncthreads();
                if (fire) {
                    leader_decay(e, draws, S.spc, S.sv, gb, cnt, nuc, step_abs, s, zn, T_half, p_dec);
                    S.scnt[g] = cnt;
                }
                __syncthreads();
                if (g < G) cnt = S.scnt[g];
                m = (cnt + 1) >> 1;
                M = 2 * m;
                active = has_nuc && t < m;
                real1 = active && s1 < cnt;
                x0 = y0 = x1 = y1 = kGhost;
                t0 = t1 = 0.f;
                v0 = v1 = make_float2(0.f, 0.f);
                if (active) {
                    const float4 a = S.spc[gb + s0];
                    x0 = a.x; y0 = a.y; t0 = a.z;
                    v0 = S.sv[gb + s0];
                }
                if (real1) {
                    const float4 a = S.spc[gb + s1];
                    x1 = a.x; y1 = a.y; t1 = a.z;
                    v1 = S.sv[gb + s1];
                }
                __syncthreads();                            // staging consumed before slots move
                if (active) {
                    put_slot(S, gb2, capS, m, s0, x0, y0, t0);
                    put_slot(S, gb2, capS, m, s1, x1, y1, t1);
                }
                R = 2.4f * cbrtf((float)cnt);
                publish_pair_sums(S.wsum, capT, g, (active ? x0 : 0.f) + (real1 ? x1 : 0.f),
                                  (active ? y0 : 0.f) + (real1 ? y1 : 0.f));
                __syncthreads();
            }
        } else {
            __syncthreads();
        }

        float f0x = 0.f, f0y = 0.f, f1x = 0.f, f1y = 0.f;
        float cx = 0.f, cy = 0.f;
        if (active) {
            // ---- centre of mass, nuclear_forces.py:242-243 ----------------------------------------
            if (e.centre) {                                 // caller-supplied `center`, :64
                cx = e.centre[2 * (int64_t)nuc];
                cy = e.centre[2 * (int64_t)nuc + 1];
            } else {
                float sx = 0.f, sy = 0.f;
                if (capT >= 32) {
                    for (int w = w_lo; w <= w_hi; ++w) {
                        const float4 ws = S.wsum[w];
                        const int gf = (w << 5) / capT;
                        if (gf == g) { sx += ws.x; sy += ws.y; }
                        else if (gf + 1 == g) { sx += ws.z; sy += ws.w; }
                    }
                } else {
                    for (int j = 0; j < cnt; ++j) {
                        const float4 a = S.A4[gb2 + (j & 1) * capS + (j >> 1)];
                        sx += a.x;
                        sy += a.z;
                    }
                }
                const float inv_n = 1.0f / (float)cnt;
                cx = sx * inv_n;
                cy = sy * inv_n;
            }
            // ---- all-pairs force, nuclear_forces.py:248-298 ---------------------------------------
            const ulonglong2* Ag = reinterpret_cast<const ulonglong2*>(S.A4 + gb2);
            const float2* Tg = S.T2 + gb2;
            float2* row = S.react + warp * nSlots + gb;
            const f32x2 xi2 = pk(x0, x1), yi2 = pk(y0, y1);
            const f32x2 nq2 = pk(-L.C * t0, -L.C * t1);
            f32x2 fx2 = pk(0.f, 0.f), fy2 = pk(0.f, 0.f);
            const int capT_ = capT;
            auto visit = [&](int u, int par) {              // partner slot 2u + par
                const ulonglong2 o = Ag[par * capS + u];
                const float2 tt = Tg[par * capS + u];
                const f32x2 dx2 = sub2(o.x, xi2), dy2 = sub2(o.y, yi2);
                const f32x2 sc2 = pair_general2(dx2, dy2, t0, t1, tt.x, pk(tt.x, tt.y), nq2, gc, L);
                const f32x2 px2 = mul2(dx2, sc2), py2 = mul2(dy2, sc2);
                fx2 = add2(fx2, px2);
                fy2 = add2(fy2, py2);
                float pa, pb, qa, qb;
                upk(px2, pa, pb);
                upk(py2, qa, qb);
                // u < 2m: un-mirror with one unsigned min (u - m wraps to a huge value when u < m)
                const unsigned ur = min((unsigned)u, (unsigned)(u - m));
                float2* rp = row + par * capT_ + ur;         // reaction on the partner
                float2 r = *rp;
                r.x -= pa + pb;
                r.y -= qa + qb;
                *rp = r;
            };
            const int hs = (m - 1) >> 1;
            for (int k = 1; k <= hs; ++k) {
                visit(t + k, 0);
                visit(t + k, 1);
            }
            if (!(m & 1) && t < (m >> 1)) {                 // antipodal super-partner, even m
                visit(t + (m >> 1), 0);
                visit(t + (m >> 1), 1);
            }
            upk(fx2, f0x, f1x);
            upk(fy2, f0y, f1y);
            {                                               // the thread's own pair (2t, 2t+1)
                const float dx = x1 - x0, dy = y1 - y0;
                const float sc = pair_general(dx, dy, t0, t1, L);
                f0x = fmaf(dx, sc, f0x);  f0y = fmaf(dy, sc, f0y);
                f1x = fmaf(-dx, sc, f1x); f1y = fmaf(-dy, sc, f1y);
            }
        }
        __syncthreads();                 // Jacobi: all reads (and all reactions) before any write
        if (active) {
            for (int w = w_lo; w <= w_hi; ++w) {            // fixed order: reproducible
                float2* rw = S.react + w * nSlots + gb;
                const float2 ra = rw[t], rb = rw[capT + t];   // slots 2t (even half), 2t+1 (odd half)
                rw[t] = make_float2(0.f, 0.f);
                rw[capT + t] = make_float2(0.f, 0.f);
                f0x += ra.x; f0y += ra.y;
                f1x += rb.x; f1y += rb.y;
            }
            contain_and_integrate(x0, y0, v0.x, v0.y, f0x, f0y, cx, cy, R, e.dt_phys);   // :301-323
            if (real1) contain_and_integrate(x1, y1, v1.x, v1.y, f1x, f1y, cx, cy, R, e.dt_phys);
            put_slot(S, gb2, capS, m, s0, x0, y0, t0);
            if (real1) put_slot(S, gb2, capS, m, s1, x1, y1, t1);
            if (e.force && s == n_steps - 1) {
                reinterpret_cast<float2*>(e.force)[off + s0] = make_float2(f0x, f0y);
                if (real1) reinterpret_cast<float2*>(e.force)[off + s1] = make_float2(f1x, f1y);
            }
        }
        publish_pair_sums(S.wsum, capT, g, (active ? x0 : 0.f) + (real1 ? x1 : 0.f),
                          (active ? y0 : 0.f) + (real1 ? y1 : 0.f));
    }

    if (active) {
        reinterpret_cast<float2*>(e.pos)[off + s0] = make_float2(x0, y0);
        reinterpret_cast<float2*>(e.vel)[off + s0] = v0;
        e.is_proton[off + s0] = (t0 != 0.f) ? 1 : 0;
    }
    if (real1) {
        reinterpret_cast<float2*>(e.pos)[off + s1] = make_float2(x1, y1);
        reinterpret_cast<float2*>(e.vel)[off + s1] = v1;
        e.is_proton[off + s1] = (t1 != 0.f) ? 1 : 0;
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

static size_t pair_smem_bytes(int T, int G, int capT)
{
    const size_t nW = T / 32, nSlots = (size_t)G * 2 * capT;
    return sizeof(float4) * (2 * nSlots + nSlots + nW) +
           sizeof(float2) * (2 * nSlots + nSlots + nW * nSlots) + sizeof(int) * G + 16;
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
    // production path: two nucleons per thread, packed f32x2, Newton-3 ring
    if (e->cap <= 512) {
        const int capT = (e->cap + 1) / 2;
        int Gp = 1;
        const int Tp = pick_block_threads(capT, &Gp);
        const size_t sm = pair_smem_bytes(Tp, Gp, capT);
        if (sm <= 200 * 1024) {
            const int64_t gridp = (n_list + Gp - 1) / Gp;
            PYQMD_REQUIRE(gridp <= 2147483647LL, "too many nuclei for one launch");
            const LawParams Lp = make_law_params(e->strong, e->coulomb, e->pauli);
            static bool attr_set = false;
            if (!attr_set) {
                PYQMD_CUDA_CHECK(cudaFuncSetAttribute(ensemble_pair_kernel<224>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      200 * 1024));
                PYQMD_CUDA_CHECK(cudaFuncSetAttribute(ensemble_pair_kernel<224>,
                                                      cudaFuncAttributePreferredSharedMemoryCarveout,
                                                      cudaSharedmemCarveoutMaxShared));
                PYQMD_CUDA_CHECK(cudaFuncSetAttribute(ensemble_pair_kernel<256>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      200 * 1024));
                PYQMD_CUDA_CHECK(cudaFuncSetAttribute(ensemble_pair_kernel<256>,
                                                      cudaFuncAttributePreferredSharedMemoryCarveout,
                                                      cudaSharedmemCarveoutMaxShared));
                attr_set = true;
            }
            // blocks of <= 224 threads: 4 blocks / SM at 72 registers per thread
            if (Tp <= 224)
                ensemble_pair_kernel<224><<<(unsigned)gridp, Tp, sm, (cudaStream_t)stream>>>(
                    d, Lp, n_steps, Gp, capT);
            else
                ensemble_pair_kernel<256><<<(unsigned)gridp, Tp, sm, (cudaStream_t)stream>>>(
                    d, Lp, n_steps, Gp, capT);
            PYQMD_CUDA_CHECK(cudaGetLastError());
            return PYQMD_OK;
        }
    }
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
