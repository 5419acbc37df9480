// ensemble.cu -- ensembles of independent nuclei: fused decay -> force -> integrate,
// K sub-steps per launch with the nucleus resident in shared memory (sm_100a).
//
// Replaces, for many nuclei at once, the sub-step loop body of the reference
// (OtsoBear/PyQMD nuclear_sim.py:165-173): Nucleus.should_decay (particles.py:126-147),
// the physics slice of handle_decay (nuclear_sim.py:213,288-294,349,353) and
// NuclearForces.update_particles_cpu (nuclear_forces.py:236-323).
//
// Layout: see ensemble_ring_kernel below -- every unordered pair once on warp-local rings, positions
// and types in shared memory as SoA, velocities and force accumulators in registers.  The update is
// Jacobi (double-buffered through registers + a barrier), like the reference CPU path and unlike
// its racy OpenCL kernel.  HBM is touched once on entry and once on exit, whatever n_steps is.
#include <stdlib.h>
#include <string.h>

#include <cooperative_groups.h>

#include <mutex>

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
    double u2 = 0.0;
    if (emission_of(mode, ptype, speed)) {
        u2 = draws.one(gid, nuc, step_abs, step_rel, 2);
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
    const pyqmd_nuclide_entry* nxt = lookup(e.table, zn);
    const double u3 = (nxt->kind == PYQMD_HL_BAND) ? draws.one(gid, nuc, step_abs, step_rel, 3) : 0.0;
    bool used3;
    daughter_half_life(nxt, u3, e.dt_decay, T, p, used3);
}

// ---------------------------------------------------------------------------------------------------
// ensemble_ring_kernel: the production kernel for nuclei of up to 1024 nucleons.
//
// Every unordered pair of a nucleus is evaluated ONCE (F_ij = -F_ji holds exactly for this law: it
// depends on d and on symmetric type predicates only) by warp-local rings:
//
//   * a lane owns kQ = 4 consecutive nucleons (a "subgroup"), a warp a group of Pe <= 32 subgroups,
//     a nucleus nG = 1..8 groups (one warp each); nuclei of <= 64 nucleons share a warp (K per warp).
//   * diagonal block (group x itself): at ring step m lane l meets subgroup (l + m) mod Pe of its own
//     group, m = 1 .. (Pe-1)/2 (+ the antipode for the lower half when Pe is even), 16 pairs per lane
//     and step, all packed f32x2 (two j per instruction); the pairs inside a subgroup are done once
//     more in ordered form (m = 0, no reaction).
//   * off-diagonal blocks: group w meets the groups w+1 .. w+(nG-1)/2 (mod nG) in full (Pe steps) and,
//     for even nG, half of the steps against group w + nG/2 (the partner does the other half).
//   * the REACTION accumulators of a j subgroup travel with it from lane to lane (SHFL), exactly as in
//     cloud_sym_kernel: no shared-memory read-modify-write, no atomics, fixed summation order =>
//     bit-reproducible.  After a diagonal ring they are shuffled home; after an off-diagonal block
//     they are written once to the per-distance reaction row of the target group.
//
// Shared memory per nucleus (slots = nG * P * 4): X, Y, T float[slots] (SoA: one LDS.128 fetches the
// 4 x / y / types of a subgroup, conflict-free because lanes read consecutive 16-byte words),
// react float2[nG/2][slots], and a canonical staging area (float4 + float2 per slot) that is touched
// only when a nucleus decays.  Padding slots hold ghost neutrons parked at (1e5, 1e5): every term of
// the law is exactly 0 at that distance (and two ghosts coincide: skipped), so no masking is needed.
constexpr float kGhost = 1.0e5f;
constexpr int kQ = 4;

struct RingGeom {
    int nG;      // warps (groups) per nucleus
    int P;       // lanes per group at full capacity (cap nucleons)
    int K;       // nuclei per warp (nG == 1 only)
    int slots;   // shared-memory slots per nucleus = nG * P * kQ
    int warps;   // warps per block
    int G;       // nuclei per block
};

static RingGeom ring_geom(int cap)
{
    RingGeom g;
    int S = (cap + kQ - 1) / kQ;
    if (S < 1) S = 1;
    if (S <= 32) {
        g.nG = 1; g.P = S; g.K = 32 / S; g.warps = 4;
    } else {
        g.nG = (S + 31) / 32; g.P = (S + g.nG - 1) / g.nG; g.K = 1; g.warps = g.nG;
    }
    g.slots = g.nG * g.P * kQ;
    g.G = (g.nG == 1) ? g.warps * g.K : 1;
    return g;
}

static size_t ring_smem_bytes(const RingGeom& g)
{
    const size_t nSl = (size_t)g.G * g.slots;
    return nSl * (3 * sizeof(float) + (size_t)(g.nG / 2) * sizeof(float2) + sizeof(float4) +
                  sizeof(float2)) + (size_t)g.warps * sizeof(float2) + (size_t)g.G * sizeof(int) + 16;
}

struct Reacts {
    f32x2 x01, y01, x23, y23;     // sum of dx * s (resp. dy * s) over the i side, per j of the subgroup
};

__device__ __forceinline__ void rotate(Reacts& r, int src)
{
    r.x01 = shfl64(r.x01, src); r.y01 = shfl64(r.y01, src);
    r.x23 = shfl64(r.x23, src); r.y23 = shfl64(r.y23, src);
}

// The 4 i-nucleons of a lane against the 4 j-nucleons of one subgroup: 16 pairs, 8 packed evaluations
// of the general law (nuclear_forces.py:253-298).
template <bool REACT>
__device__ __forceinline__ void ring_visit(const ulonglong2 X, const ulonglong2 Y, const float4 T,
                                           const f32x2 (&xi2)[kQ], const f32x2 (&yi2)[kQ],
                                           const float (&ti)[kQ], f32x2 (&ax)[kQ], f32x2 (&ay)[kQ],
                                           Reacts& r, const GenConsts& gc, const LawParams& L,
                                           const f32x2 negC)
{
    const f32x2 nq01 = mul2(negC, pk(T.x, T.y)), nq23 = mul2(negC, pk(T.z, T.w));
#pragma unroll
    for (int k = 0; k < kQ; ++k) {
        const f32x2 ti2 = pk1(ti[k]);
        {
            const f32x2 dx = sub2(X.x, xi2[k]), dy = sub2(Y.x, yi2[k]);
            const f32x2 s = pair_general2(dx, dy, T.x, T.y, ti[k], ti2, nq01, gc, L);
            ax[k] = fma2(dx, s, ax[k]);
            ay[k] = fma2(dy, s, ay[k]);
            if (REACT) { r.x01 = fma2(dx, s, r.x01); r.y01 = fma2(dy, s, r.y01); }
        }
        {
            const f32x2 dx = sub2(X.y, xi2[k]), dy = sub2(Y.y, yi2[k]);
            const f32x2 s = pair_general2(dx, dy, T.z, T.w, ti[k], ti2, nq23, gc, L);
            ax[k] = fma2(dx, s, ax[k]);
            ay[k] = fma2(dy, s, ay[k]);
            if (REACT) { r.x23 = fma2(dx, s, r.x23); r.y23 = fma2(dy, s, r.y23); }
        }
    }
}

#ifndef PYQMD_RING_MINBLOCKS_SINGLE
#define PYQMD_RING_MINBLOCKS_SINGLE 4
#endif
#ifndef PYQMD_RING_MINBLOCKS_PAIR
#define PYQMD_RING_MINBLOCKS_PAIR 8
#endif

// MULTI = true: one nucleus per block of nG >= 2 warps (MAXT = 64: the two-warp case, 129..256
// nucleons, tuned on Pb-208 / U-238; MAXT = 256: up to 8 warps); MULTI = false: every warp is on its
// own (K >= 1 nuclei per warp, MAXT = 128), so only warp-level synchronisation is used.
template <bool MULTI, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
ensemble_ring_kernel(const pyqmd_ensemble e, const LawParams L, const int n_steps, const RingGeom geo)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nSl = geo.G * geo.slots;
    const int nR = geo.nG >> 1;
    float* sX = reinterpret_cast<float*>(smem_raw);
    float* sY = sX + nSl;
    float* sT = sY + nSl;
    float2* sReact = reinterpret_cast<float2*>(sT + nSl);          // [nR][nSl]
    float4* spc = reinterpret_cast<float4*>(sReact + (size_t)nR * nSl);
    float2* sv = reinterpret_cast<float2*>(spc + nSl);
    float2* wsum = sv + nSl;                                        // [warps]
    int* scnt = reinterpret_cast<int*>(wsum + geo.warps);           // [G]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // MULTI or K == 1: the whole warp serves one nucleus (lanes >= Pe idle along: every lane must see
    // the same count, the ring trip counts have to be warp-uniform); K > 1: `sub` picks the team.
    int sub = 0, l = lane, grp = warp, g = 0;
    if (!MULTI) {
        grp = 0;
        if (geo.K > 1) {
            sub = lane / geo.P;
            l = lane - sub * geo.P;
        }
        g = warp * geo.K + sub;
    }
    const bool in_team = sub < geo.K;
    if (!in_team) g = warp * geo.K;                                 // keep shared-memory indices in range
    const int tbase = MULTI ? 0 : sub * geo.P;                      // first lane of the team
    const int64_t q = (int64_t)blockIdx.x * geo.G + g;
    const bool has_nuc = in_team && q < e.n_list;
    if (!MULTI && !__any_sync(0xffffffffu, has_nuc)) return;        // this warp has nothing to do
    const int nuc = has_nuc ? (e.list ? e.list[q] : (int)q) : -1;
    const bool leader = has_nuc && l == 0 && grp == 0;
    const int gb = g * geo.slots;                                   // first slot of the nucleus
    const bool per_count_ring = MULTI || geo.K == 1;                // ring length follows the live count

    int cnt = 0;
    int64_t off = 0;
    if (has_nuc) {
        cnt = e.count[nuc];
        off = e.offset[nuc];
    }
    auto ring_len = [&](int c) {
        if (!per_count_ring) return geo.P;                          // trip counts must be warp-uniform
        int S = (c + kQ - 1) / kQ;
        if (S < 1) S = 1;
        return MULTI ? (S + geo.nG - 1) / geo.nG : S;
    };
    int Pe = ring_len(cnt);
    bool active = in_team && l < Pe;
    int s0 = (grp * Pe + (active ? l : 0)) * kQ;                    // first slot of this lane

    float xi[kQ], yi[kQ], ti[kQ];
    float2 vi[kQ];
    auto load_global = [&]() {
#pragma unroll
        for (int k = 0; k < kQ; ++k) {
            xi[k] = kGhost; yi[k] = kGhost; ti[k] = 0.f;
            vi[k] = make_float2(0.f, 0.f);
            if (active && s0 + k < cnt) {
                const float2 p = reinterpret_cast<const float2*>(e.pos)[off + s0 + k];
                vi[k] = reinterpret_cast<const float2*>(e.vel)[off + s0 + k];
                xi[k] = p.x; yi[k] = p.y;
                ti[k] = e.is_proton[off + s0 + k] ? 1.0f : 0.0f;
            }
        }
    };
    // own subgroup -> shared memory, and the nucleus' coordinate sums (nuclear_forces.py:242-243)
    float sumx = 0.f, sumy = 0.f;
    auto publish = [&]() {
        if (active) {
            reinterpret_cast<float4*>(sX + gb)[s0 >> 2] = make_float4(xi[0], xi[1], xi[2], xi[3]);
            reinterpret_cast<float4*>(sY + gb)[s0 >> 2] = make_float4(yi[0], yi[1], yi[2], yi[3]);
            reinterpret_cast<float4*>(sT + gb)[s0 >> 2] = make_float4(ti[0], ti[1], ti[2], ti[3]);
        }
        float px = 0.f, py = 0.f;
#pragma unroll
        for (int k = 0; k < kQ; ++k)
            if (active && s0 + k < cnt) { px += xi[k]; py += yi[k]; }
        if (MULTI || geo.K == 1) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                px += __shfl_xor_sync(0xffffffffu, px, o);
                py += __shfl_xor_sync(0xffffffffu, py, o);
            }
            if (MULTI && lane == 0) wsum[warp] = make_float2(px, py);
            sumx = px; sumy = py;
        } else {
            float ax_ = 0.f, ay_ = 0.f;
            for (int j = 0; j < geo.P; ++j) {                       // fixed order, team by team
                ax_ += __shfl_sync(0xffffffffu, px, min(tbase + j, 31));
                ay_ += __shfl_sync(0xffffffffu, py, min(tbase + j, 31));
            }
            sumx = ax_; sumy = ay_;
        }
    };
    auto team_sync = [&]() {
        if (MULTI) __syncthreads(); else __syncwarp();
    };

    load_global();
    if (leader) scnt[g] = cnt;
    publish();

    int32_t zn = 0;
    double T_half = 0.0, p_dec = -1.0;
    if (leader && e.decay_enabled) {
        zn = e.zn[nuc];
        T_half = e.half_life[nuc];
        p_dec = e.p_decay[nuc];
    }
    const DrawSource draws{e.uniforms, e.seed, e.uniforms_n};
    const GenConsts gc = make_gen_consts(L);
    const f32x2 negC = pk1(-L.C);
    float R = 2.4f * cbrtf((float)cnt);                     // nuclear_forces.py:304
    team_sync();

    for (int s = 0; s < n_steps; ++s) {
        // ---- decay test: Nucleus.should_decay, particles.py:126-147 --------------------------
        if (e.decay_enabled) {
            bool fire = false;
            const uint32_t step_abs = e.step0 + (uint32_t)s;
            if (leader && p_dec >= 0.0) {                   // stable: no draw (:129-130)
                const double u0 = draws.one((uint64_t)(e.id_base + nuc), nuc, step_abs, s, 0);
                fire = u0 < p_dec;                          // :147
            }
            const bool any_fire = MULTI ? (__syncthreads_or(fire) != 0)
                                        : (__any_sync(0xffffffffu, fire) != 0);
            if (any_fire) {
                // canonical (list-order) staging: slot index == position in Nucleus.particles
#pragma unroll
                for (int k = 0; k < kQ; ++k)
                    if (active && s0 + k < cnt) {
                        spc[gb + s0 + k] = make_float4(xi[k], yi[k], ti[k], 0.f);
                        sv[gb + s0 + k] = vi[k];
                    }
                team_sync();
                if (fire) {
                    leader_decay(e, draws, spc, sv, gb, cnt, nuc, step_abs, s, zn, T_half, p_dec);
                    scnt[g] = cnt;
                }
                team_sync();
                if (has_nuc) cnt = scnt[g];
                Pe = ring_len(cnt);
                active = in_team && l < Pe;
                s0 = (grp * Pe + (active ? l : 0)) * kQ;
#pragma unroll
                for (int k = 0; k < kQ; ++k) {
                    xi[k] = kGhost; yi[k] = kGhost; ti[k] = 0.f;
                    vi[k] = make_float2(0.f, 0.f);
                    if (active && s0 + k < cnt) {
                        const float4 a = spc[gb + s0 + k];
                        xi[k] = a.x; yi[k] = a.y; ti[k] = a.z;
                        vi[k] = sv[gb + s0 + k];
                    }
                }
                R = 2.4f * cbrtf((float)cnt);
                team_sync();                                // staging consumed before X / Y / T move
                publish();
                team_sync();
            }
        }

        // ---- centre of mass, nuclear_forces.py:242-243 --------------------------------------------
        float cx = 0.f, cy = 0.f;
        if (has_nuc) {
            if (e.centre) {                                 // caller-supplied `center`, :64
                cx = e.centre[2 * (int64_t)nuc];
                cy = e.centre[2 * (int64_t)nuc + 1];
            } else {
                float sx = sumx, sy = sumy;
                if (MULTI) {
                    sx = 0.f; sy = 0.f;
                    for (int w = 0; w < geo.nG; ++w) { sx += wsum[w].x; sy += wsum[w].y; }
                }
                const float inv_n = 1.0f / (float)max(cnt, 1);
                cx = sx * inv_n;
                cy = sy * inv_n;
            }
        }

        // ---- all-pairs force, nuclear_forces.py:248-298 ---------------------------------------------
        f32x2 ax[kQ], ay[kQ], xi2[kQ], yi2[kQ];
#pragma unroll
        for (int k = 0; k < kQ; ++k) {
            ax[k] = 0ull; ay[k] = 0ull;
            xi2[k] = pk1(xi[k]);
            yi2[k] = pk1(yi[k]);
        }
        const int lq = active ? l : 0;
        const int src = active ? ((l + 1 == Pe) ? tbase : lane + 1) : lane;   // the ring neighbour
        const ulonglong2* X4 = reinterpret_cast<const ulonglong2*>(sX + gb);
        const ulonglong2* Y4 = reinterpret_cast<const ulonglong2*>(sY + gb);
        const float4* T4 = reinterpret_cast<const float4*>(sT + gb);
        Reacts rd = {0ull, 0ull, 0ull, 0ull};
        {
            // pairs inside the subgroup: ordered, no reaction (the self pairs are skipped by d2 < 0.01)
            Reacts none = {0ull, 0ull, 0ull, 0ull};
            ring_visit<false>(make_ulonglong2(pk(xi[0], xi[1]), pk(xi[2], xi[3])),
                              make_ulonglong2(pk(yi[0], yi[1]), pk(yi[2], yi[3])),
                              make_float4(ti[0], ti[1], ti[2], ti[3]), xi2, yi2, ti, ax, ay, none, gc, L,
                              negC);
            // diagonal ring of the lane's own group
            const int own = grp * Pe;
            const int hs = (Pe - 1) >> 1;
            int qd = lq;
#pragma unroll 1
            for (int m = 1; m <= hs; ++m) {
                qd = (qd + 1 == Pe) ? 0 : qd + 1;
                ring_visit<true>(X4[own + qd], Y4[own + qd], T4[own + qd], xi2, yi2, ti, ax, ay, rd, gc,
                                 L, negC);
                rotate(rd, src);
            }
            if (!(Pe & 1)) {                                // antipodal subgroup, lower half only
                if (active && l < (Pe >> 1)) {
                    const int qa = own + l + (Pe >> 1);
                    ring_visit<true>(X4[qa], Y4[qa], T4[qa], xi2, yi2, ti, ax, ay, rd, gc, L, negC);
                }
                rotate(rd, src);
            }
            // lane l now holds the reaction of subgroup (l + Pe/2 + 1) mod Pe: bring it home
            int from = lq - (Pe >> 1) - 1;
            from += (from < 0) ? Pe : 0;
            from += (from < 0) ? Pe : 0;
            rotate(rd, active ? tbase + from : lane);
        }
        if (MULTI) {
            // off-diagonal blocks: groups grp+1 .. grp+nG/2 (the last one shared with the partner)
            for (int dlt = 1; dlt <= nR; ++dlt) {
                int tg = grp + dlt;
                tg -= (tg >= geo.nG) ? geo.nG : 0;
                int m0 = 0, m1 = Pe;
                if (2 * dlt == geo.nG) {
                    const int h = (Pe + 1) >> 1;
                    if (grp < nR) { m0 = 0; m1 = h; } else { m0 = 1; m1 = Pe - h + 1; }
                }
                Reacts ro = {0ull, 0ull, 0ull, 0ull};
                const int base = tg * Pe;
                int qo = lq + m0;
                qo -= (qo >= Pe) ? Pe : 0;
#pragma unroll 1
                for (int m = m0; m < m1; ++m) {
                    ring_visit<true>(X4[base + qo], Y4[base + qo], T4[base + qo], xi2, yi2, ti, ax, ay, ro,
                                     gc, L, negC);
                    rotate(ro, src);
                    qo = (qo + 1 == Pe) ? 0 : qo + 1;
                }
                if (active) {                               // lane l holds subgroup (l + m1) mod Pe of tg
                    float a, b, c2, d;
                    float4* row = reinterpret_cast<float4*>(sReact + (size_t)(dlt - 1) * nSl + gb) +
                                  2 * (base + qo);
                    upk(ro.x01, a, b);
                    upk(ro.y01, c2, d);
                    row[0] = make_float4(a, c2, b, d);
                    upk(ro.x23, a, b);
                    upk(ro.y23, c2, d);
                    row[1] = make_float4(a, c2, b, d);
                }
            }
        }
        team_sync();                 // Jacobi: all reads (and all reactions) before any write
        {
            float rx[kQ], ry[kQ];
            upk(rd.x01, rx[0], rx[1]); upk(rd.x23, rx[2], rx[3]);
            upk(rd.y01, ry[0], ry[1]); upk(rd.y23, ry[2], ry[3]);
#pragma unroll
            for (int k = 0; k < kQ; ++k) {
                float a, b;
                upk(ax[k], a, b);
                float fx = (a + b) - rx[k];
                upk(ay[k], a, b);
                float fy = (a + b) - ry[k];
                if (MULTI && active) {
                    for (int dlt = 1; dlt <= nR; ++dlt) {   // fixed order: reproducible
                        const float2 rr = sReact[(size_t)(dlt - 1) * nSl + gb + s0 + k];
                        fx -= rr.x;
                        fy -= rr.y;
                    }
                }
                if (active && s0 + k < cnt) {
                    contain_and_integrate(xi[k], yi[k], vi[k].x, vi[k].y, fx, fy, cx, cy, R,
                                          e.dt_phys);                                // :301-323
                    if (e.force && s == n_steps - 1)
                        reinterpret_cast<float2*>(e.force)[off + s0 + k] = make_float2(fx, fy);
                }
            }
        }
        publish();
        team_sync();
    }

#pragma unroll
    for (int k = 0; k < kQ; ++k)
        if (active && s0 + k < cnt) {
            reinterpret_cast<float2*>(e.pos)[off + s0 + k] = make_float2(xi[k], yi[k]);
            reinterpret_cast<float2*>(e.vel)[off + s0 + k] = vi[k];
            if (e.decay_enabled) e.is_proton[off + s0 + k] = (ti[k] != 0.f) ? 1 : 0;
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
// ensemble_pair_kernel: the block-wide ring, used where the warp-local rings of ensemble_ring_kernel
// would leave too many lanes idle (Pb-208: 52 subgroups on 2 x 32 lanes = 81 %, against 208 / 224
// threads = 93 % here; measured on B200, r02: 1.15e12 vs 1.05e12 pairs/s).  See pyqmd_ensemble_step for
// the dispatch rule.
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
                // lane t - 1 updates the row slot lane t has just written in the next ring step:
                // order the two accesses (warp-level memory fence; the lanes may not be converged
                // when a warp spans two nuclei with different counts)
                __syncwarp(__activemask());
            }
            if (!(m & 1) && t < (m >> 1)) {                 // antipodal super-partner, even m
                __syncwarp(__activemask());
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


// ---------------------------------------------------------------------------------------------------
// ensemble_quad_kernel: the block-wide ring with FOUR nucleons per thread.
//
// The warp-local rings above execute ~20 % fewer instructions per pair than ensemble_pair_kernel (16 pairs
// per visit, no per-pair index arithmetic), but a nucleus whose subgroups do not fill its warps leaves
// lanes idle (Pb-208: 52 subgroups on 64 lanes).  Here the subgroups of G nuclei are packed into one block
// -- thread tid owns subgroup tid % capT of nucleus tid / capT, so a warp may span two nuclei and no lane
// idles but the block's last few -- and the ring runs over the whole nucleus: at step k thread t meets
// subgroup (t + k) mod P (P = live subgroups), 16 pairs per visit through ring_visit.  The reaction on the
// partner subgroup cannot travel by SHFL across warps, so it is accumulated into a per-warp row in shared
// memory (two 16-byte read-modify-writes per 16 pairs; lane t - 1 touches the same words one step later,
// hence the warp fence after every step) and the rows are summed in fixed order after the barrier.
// Decay handling, staging area and centre sums are those of ensemble_pair_kernel.  Needs capT >= 32 (a
// warp spans at most two nuclei): 125 .. 1024 nucleons.
struct QuadSmem {
    float4 *X4, *Y4, *T4;     // [2 * G * capT] subgroups, mirrored at +P so that t + k needs no modulo
    float4* spc;              // [G * 4 capT] staging (x, y, type, -), touched when a nucleus decays
    float4* wsum;             // [warps] coordinate sums by nucleus
    float2* sv;               // [G * 4 capT] staging velocities
    float2* react;            // [warps][G * 4 capT] reaction rows
    int* scnt;                // [G]
};

static size_t quad_smem_bytes(int T, int G, int capT)
{
    const size_t nW = T / 32, nSub = (size_t)G * capT, nSlots = 4 * nSub;
    return sizeof(float4) * (3 * 2 * nSub + nSlots + nW) + sizeof(float2) * (nSlots + nW * nSlots) +
           sizeof(int) * G + 16;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 512 / MAXT) ensemble_quad_kernel(const pyqmd_ensemble e,
                                                                         const LawParams L,
                                                                         const int n_steps, const int G,
                                                                         const int capT)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x;
    const int nW = T >> 5;
    const int capS = kQ * capT;
    const int nSub = G * capT;
    const int nSlots = G * capS;
    QuadSmem S;
    S.X4 = reinterpret_cast<float4*>(smem_raw);
    S.Y4 = S.X4 + 2 * nSub;
    S.T4 = S.Y4 + 2 * nSub;
    S.spc = S.T4 + 2 * nSub;
    S.wsum = S.spc + nSlots;
    S.sv = reinterpret_cast<float2*>(S.wsum + nW);
    S.react = S.sv + nSlots;
    S.scnt = reinterpret_cast<int*>(S.react + nW * nSlots);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int g = tid / capT;
    const int t = tid - g * capT;
    const int gb = g * capS;             // slot base (spc, sv, react)
    const int gb2 = 2 * g * capT;        // subgroup base in the mirrored arrays
    const int s0 = kQ * t;
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
    int P = (cnt + kQ - 1) / kQ;
    bool active = has_nuc && t < P;
    float xi[kQ], yi[kQ], ti[kQ];
    float2 vi[kQ];
#pragma unroll
    for (int k = 0; k < kQ; ++k) {
        xi[k] = kGhost; yi[k] = kGhost; ti[k] = 0.f;
        vi[k] = make_float2(0.f, 0.f);
        if (active && s0 + k < cnt) {
            const float2 p = reinterpret_cast<const float2*>(e.pos)[off + s0 + k];
            vi[k] = reinterpret_cast<const float2*>(e.vel)[off + s0 + k];
            xi[k] = p.x; yi[k] = p.y;
            ti[k] = e.is_proton[off + s0 + k] ? 1.0f : 0.0f;
        }
    }
    // own subgroup -> shared memory (both copies), and the nucleus' coordinate sums (:242-243)
    auto publish = [&]() {
        if (active) {
            const float4 X = make_float4(xi[0], xi[1], xi[2], xi[3]);
            const float4 Y = make_float4(yi[0], yi[1], yi[2], yi[3]);
            const float4 Tt = make_float4(ti[0], ti[1], ti[2], ti[3]);
            S.X4[gb2 + t] = X; S.X4[gb2 + t + P] = X;
            S.Y4[gb2 + t] = Y; S.Y4[gb2 + t + P] = Y;
            S.T4[gb2 + t] = Tt; S.T4[gb2 + t + P] = Tt;
        }
        float px = 0.f, py = 0.f;
#pragma unroll
        for (int k = 0; k < kQ; ++k)
            if (active && s0 + k < cnt) { px += xi[k]; py += yi[k]; }
        publish_pair_sums(S.wsum, capT, g, px, py);
    };
    if (t == 0 && g < G) S.scnt[g] = cnt;
    for (int k = tid; k < nW * nSlots; k += T) S.react[k] = make_float2(0.f, 0.f);
    publish();
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
    const f32x2 negC = pk1(-L.C);
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
#pragma unroll
                for (int k = 0; k < kQ; ++k)
                    if (active && s0 + k < cnt) {
                        S.spc[gb + s0 + k] = make_float4(xi[k], yi[k], ti[k], 0.f);
                        S.sv[gb + s0 + k] = vi[k];
                    }
                __syncthreads();
                if (fire) {
                    leader_decay(e, draws, S.spc, S.sv, gb, cnt, nuc, step_abs, s, zn, T_half, p_dec);
                    S.scnt[g] = cnt;
                }
                __syncthreads();
                if (g < G) cnt = S.scnt[g];
                P = (cnt + kQ - 1) / kQ;
                active = has_nuc && t < P;
#pragma unroll
                for (int k = 0; k < kQ; ++k) {
                    xi[k] = kGhost; yi[k] = kGhost; ti[k] = 0.f;
                    vi[k] = make_float2(0.f, 0.f);
                    if (active && s0 + k < cnt) {
                        const float4 a = S.spc[gb + s0 + k];
                        xi[k] = a.x; yi[k] = a.y; ti[k] = a.z;
                        vi[k] = S.sv[gb + s0 + k];
                    }
                }
                __syncthreads();                            // staging consumed before the subgroups move
                R = 2.4f * cbrtf((float)cnt);
                publish();
                __syncthreads();
            }
        } else {
            __syncthreads();             // subgroups and sums of this sub-step are in shared memory
        }

        float fx[kQ], fy[kQ];
        float cx = 0.f, cy = 0.f;
        f32x2 ax[kQ], ay[kQ], xi2[kQ], yi2[kQ];
#pragma unroll
        for (int k = 0; k < kQ; ++k) {
            ax[k] = 0ull; ay[k] = 0ull;
            xi2[k] = pk1(xi[k]);
            yi2[k] = pk1(yi[k]);
        }
        if (active) {
            // ---- centre of mass, nuclear_forces.py:242-243 ----------------------------------------
            if (e.centre) {                                 // caller-supplied `center`, :64
                cx = e.centre[2 * (int64_t)nuc];
                cy = e.centre[2 * (int64_t)nuc + 1];
            } else {
                float sx = 0.f, sy = 0.f;
                for (int w = w_lo; w <= w_hi; ++w) {
                    const float4 ws = S.wsum[w];
                    const int gf = (w << 5) / capT;
                    if (gf == g) { sx += ws.x; sy += ws.y; }
                    else if (gf + 1 == g) { sx += ws.z; sy += ws.w; }
                }
                const float inv_n = 1.0f / (float)cnt;
                cx = sx * inv_n;
                cy = sy * inv_n;
            }
            // pairs inside the subgroup: ordered, no reaction (the self pairs are skipped by d2 < 0.01)
            Reacts none = {0ull, 0ull, 0ull, 0ull};
            ring_visit<false>(make_ulonglong2(pk(xi[0], xi[1]), pk(xi[2], xi[3])),
                              make_ulonglong2(pk(yi[0], yi[1]), pk(yi[2], yi[3])),
                              make_float4(ti[0], ti[1], ti[2], ti[3]), xi2, yi2, ti, ax, ay, none, gc, L,
                              negC);
        }
        // ---- all-pairs force, nuclear_forces.py:248-298: ring over the nucleus' P subgroups -----------
        {
            const ulonglong2* Xg = reinterpret_cast<const ulonglong2*>(S.X4 + gb2);
            const ulonglong2* Yg = reinterpret_cast<const ulonglong2*>(S.Y4 + gb2);
            const float4* Tg = S.T4 + gb2;
            float4* row = reinterpret_cast<float4*>(S.react + (size_t)warp * nSlots + gb);
            auto visit = [&](int u) {                       // partner subgroup u in [0, 2P): mirrored
                const unsigned ur = min((unsigned)u, (unsigned)(u - P));   // u mod P
                Reacts r = {0ull, 0ull, 0ull, 0ull};
                ring_visit<true>(Xg[u], Yg[u], Tg[u], xi2, yi2, ti, ax, ay, r, gc, L, negC);
                float a0, a1, b0, b1;
                // (loading the row words before the 16 pairs are evaluated was measured: 1.5 % slower)
                float4 v = row[2 * ur], w = row[2 * ur + 1];
                upk(r.x01, a0, a1);                          // slots 4 ur, 4 ur + 1: (x, y, x, y)
                upk(r.y01, b0, b1);
                v.x -= a0; v.y -= b0; v.z -= a1; v.w -= b1;
                row[2 * ur] = v;
                upk(r.x23, a0, a1);                          // slots 4 ur + 2, 4 ur + 3
                upk(r.y23, b0, b1);
                w.x -= a0; w.y -= b0; w.z -= a1; w.w -= b1;
                row[2 * ur + 1] = w;
            };
            // trip counts differ between the (at most two) nuclei of a warp: every lane runs the warp's
            // maximum and masks its own visits, so that the fence after each step is warp-wide
            const int hs = active ? ((P - 1) >> 1) : 0;
            const int hs_w = __reduce_max_sync(0xffffffffu, hs);
            for (int k = 1; k <= hs_w; ++k) {
                if (k <= hs) visit(t + k);
                __syncwarp();            // lane t - 1 updates at step k + 1 the words lane t wrote at step k
            }
            if (active && !(P & 1) && t < (P >> 1)) visit(t + (P >> 1));   // antipodal subgroup, even P
        }
        __syncthreads();                 // Jacobi: all reads (and all reactions) before any write
        if (active) {
#pragma unroll
            for (int k = 0; k < kQ; ++k) {
                float a, b;
                upk(ax[k], a, b); fx[k] = a + b;
                upk(ay[k], a, b); fy[k] = a + b;
            }
            for (int w = w_lo; w <= w_hi; ++w) {            // fixed order: reproducible
                float4* rw = reinterpret_cast<float4*>(S.react + (size_t)w * nSlots + gb) + 2 * t;
                const float4 ra = rw[0], rb = rw[1];
                rw[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                rw[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                fx[0] += ra.x; fy[0] += ra.y; fx[1] += ra.z; fy[1] += ra.w;
                fx[2] += rb.x; fy[2] += rb.y; fx[3] += rb.z; fy[3] += rb.w;
            }
#pragma unroll
            for (int k = 0; k < kQ; ++k)
                if (s0 + k < cnt) {
                    contain_and_integrate(xi[k], yi[k], vi[k].x, vi[k].y, fx[k], fy[k], cx, cy, R,
                                          e.dt_phys);              // :301-323
                    if (e.force && s == n_steps - 1)
                        reinterpret_cast<float2*>(e.force)[off + s0 + k] = make_float2(fx[k], fy[k]);
                }
        }
        publish();
    }

#pragma unroll
    for (int k = 0; k < kQ; ++k)
        if (active && s0 + k < cnt) {
            reinterpret_cast<float2*>(e.pos)[off + s0 + k] = make_float2(xi[k], yi[k]);
            reinterpret_cast<float2*>(e.vel)[off + s0 + k] = vi[k];
            e.is_proton[off + s0 + k] = (ti[k] != 0.f) ? 1 : 0;
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

// block shape of the quad kernel: G nuclei x capT threads in T in {128, 160, 224, 256} threads, the choice
// that keeps most lanes busy on most resident warps (128 registers per thread: 512 threads per SM)
static int pick_quad_threads(int capT, int* G_out)
{
    static const int shapes[4] = {128, 160, 224, 256};
    int best_T = 0, best_G = 0;
    double best = -1.0;
    const char* pin = getenv("PYQMD_QUAD_THREADS");         // tuning runs
    for (int T : shapes) {
        if (pin && atoi(pin) != T) continue;
        const int G = T / capT;
        if (G < 1) continue;
        const size_t sm = quad_smem_bytes(T, G, capT);
        int blocks = 512 / T;
        const int by_smem = (int)((size_t)220 * 1024 / (sm + 1024));
        if (by_smem < blocks) blocks = by_smem;
        if (blocks < 1) continue;
        const double score = (double)G * capT / T * (blocks * (T / 32));
        if (score > best + 1e-9) { best = score; best_T = T; best_G = G; }
    }
    *G_out = best_G;
    return best_T;
}


// ---------------------------------------------------------------------------------------------------
// ensemble_cluster_kernel: ONE nucleus on a thread-block CLUSTER of 8 CTAs (8 SMs), for launches of a
// handful of nuclei (BASELINE config 1: a single U-238; the interactive app).  Such a launch cannot
// fill the GPU, its sub-steps are a latency chain -- 21 us each in one CTA -- so the chain is cut
// eight ways instead:
//   * every CTA keeps a full replica of the positions and types in its shared memory (double
//     buffered: Jacobi) and OWNS cnt/8 nucleons; a thread evaluates its nucleon against one slice of
//     all partners (ordered pairs: with 8 SMs on the job the reaction exchange of Newton's third law
//     would cost more than it saves), the slices' partial forces meet in shared memory;
//   * the owner integrates and stores the new position straight into all eight replicas through
//     distributed shared memory (cluster.map_shared_rank), together with its share of the coordinate
//     sums and -- CTA 0 -- next sub-step's decay decision; ONE cluster barrier per sub-step;
//   * a decay (rare) gathers the velocities in CTA 0, whose leader thread runs the serial
//     transmutation (leader_decay), and redistributes the state.
#ifndef PYQMD_CLUSTER_THREADS
#define PYQMD_CLUSTER_THREADS 256
#endif
constexpr int kClusterSize = 8;
constexpr int kClusterThreads = PYQMD_CLUSTER_THREADS;

struct ClusterSmem {
    float* X[2];
    float* Y[2];
    float* Tt;          // [capP] types
    float2* part;       // [slices][ipc] partial forces
    float2* csum[2];    // [8] per-CTA coordinate sums, per buffer
    float4* spc;        // [capP] canonical staging (CTA 0 holds the authoritative copy during a decay)
    float2* sv;         // [capP]
    int* flags;         // [0..1] fire flag per buffer parity, [2] cnt
};

__device__ __forceinline__ ClusterSmem carve_cluster(unsigned char* raw, int capP, int slices, int ipc)
{
    ClusterSmem S;
    float* f = reinterpret_cast<float*>(raw);
    S.X[0] = f; S.X[1] = f + capP; S.Y[0] = f + 2 * capP; S.Y[1] = f + 3 * capP; S.Tt = f + 4 * capP;
    S.spc = reinterpret_cast<float4*>(f + 5 * capP + (capP & 3 ? 4 - (capP & 3) : 0));
    S.sv = reinterpret_cast<float2*>(S.spc + capP);
    S.part = S.sv + capP;
    S.csum[0] = S.part + slices * ipc;
    S.csum[1] = S.csum[0] + kClusterSize;
    S.flags = reinterpret_cast<int*>(S.csum[1] + kClusterSize);
    return S;
}

static size_t cluster_smem_bytes(int capP, int slices, int ipc)
{
    return sizeof(float) * (5 * (size_t)capP + 4) + sizeof(float4) * capP + sizeof(float2) * capP +
           sizeof(float2) * ((size_t)slices * ipc + 2 * kClusterSize) + sizeof(int) * 4 + 32;
}

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kClusterThreads, 1)
ensemble_cluster_kernel(const pyqmd_ensemble e, const LawParams L, const int n_steps, const int capP,
                        const int ipc)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int slices = kClusterThreads / ipc;
    const ClusterSmem S = carve_cluster(smem_raw, capP, slices, ipc);
    // the same address in CTA r of the cluster (distributed shared memory)
    auto remote = [&](auto* ptr, int r) { return cluster.map_shared_rank(ptr, r); };

    const int c = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x / kClusterSize;            // one cluster per listed nucleus
    const int nuc = e.list ? e.list[q] : (int)q;
    const int64_t off = e.offset[nuc];
    int cnt = e.count[nuc];
    const int il = tid % ipc, sl = tid / ipc;
    const bool leader = (c == 0 && tid == 0);

    // layout derived from the live count
    int per, i, jper, j0, j1;
    bool owner;
    auto layout = [&]() {
        per = (cnt + kClusterSize - 1) / kClusterSize;      // nucleons per CTA (<= ipc)
        i = c * per + il;
        owner = (sl == 0) && il < per && i < cnt;
        jper = (((cnt + slices - 1) / slices) + 1) & ~1;    // partners per slice, even (packed pairs)
        j0 = sl * jper;
        j1 = min(j0 + jper, (cnt + 1) & ~1);
    };
    layout();

    // replicas: every CTA loads the whole nucleus; velocities stay with the owners
    for (int k = tid; k < capP; k += kClusterThreads) {
        float x = kGhost, y = kGhost, t = 0.f;
        if (k < cnt) {
            const float2 p2 = reinterpret_cast<const float2*>(e.pos)[off + k];
            x = p2.x; y = p2.y;
            t = e.is_proton[off + k] ? 1.0f : 0.0f;
        }
        S.X[0][k] = x; S.Y[0][k] = y; S.Tt[k] = t;
        S.X[1][k] = kGhost; S.Y[1][k] = kGhost;
    }
    float2 vel = make_float2(0.f, 0.f);
    if (owner) vel = reinterpret_cast<const float2*>(e.vel)[off + i];
    __syncthreads();
    // this CTA's share of the coordinate sums of buffer 0, published to every CTA
    auto publish_sum = [&](int buf) {
        if (tid < 32) {
            float sx = 0.f, sy = 0.f;
            for (int k = c * per + tid; k < min((c + 1) * per, cnt); k += 32) { sx += S.X[buf][k]; sy += S.Y[buf][k]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, o);
                sy += __shfl_xor_sync(0xffffffffu, sy, o);
            }
            if (tid < kClusterSize) remote(S.csum[buf], tid)[c] = make_float2(sx, sy);
        }
    };
    publish_sum(0);

    int32_t zn = 0;
    double T_half = 0.0, p_dec = -1.0;
    if (leader && e.decay_enabled) {
        zn = e.zn[nuc];
        T_half = e.half_life[nuc];
        p_dec = e.p_decay[nuc];
    }
    const DrawSource draws{e.uniforms, e.seed, e.uniforms_n};
    const GenConsts gc = make_gen_consts(L);
    // decay decision of sub-step `s` (Nucleus.should_decay, particles.py:126-147): taken by the leader and
    // stored in every CTA one cluster barrier ahead of its use
    auto decide = [&](int s, int parity) {
        bool fire = false;
        if (e.decay_enabled && p_dec >= 0.0 && s < n_steps) {      // stable: no draw (:129-130)
            const double u0 = draws.one((uint64_t)(e.id_base + nuc), nuc, e.step0 + (uint32_t)s, s, 0);
            fire = u0 < p_dec;                                      // :147
        }
#pragma unroll
        for (int r = 0; r < kClusterSize; ++r) remote(S.flags, r)[parity] = fire ? 1 : 0;
    };
    if (leader) decide(0, 0);
    cluster.sync();

    int buf = 0;
    for (int s = 0; s < n_steps; ++s) {
        if (e.decay_enabled && S.flags[s & 1]) {
            // ---- decay: physics slice of handle_decay, nuclear_sim.py:213,288-294,349,353 -----------
            if (owner) remote(S.sv, 0)[i] = vel;            // velocities to CTA 0 (canonical order)
            cluster.sync();
            if (c == 0) {
                for (int k = tid; k < cnt; k += kClusterThreads)
                    S.spc[k] = make_float4(S.X[buf][k], S.Y[buf][k], S.Tt[k], 0.f);
                __syncthreads();
                if (tid == 0) {
                    leader_decay(e, draws, S.spc, S.sv, 0, cnt, nuc, e.step0 + (uint32_t)s, s, zn, T_half, p_dec);
#pragma unroll
                    for (int r = 0; r < kClusterSize; ++r) remote(S.flags, r)[2] = cnt;
                }
            }
            cluster.sync();
            cnt = S.flags[2];
            layout();
            const float4* spc0 = remote(S.spc, 0);
            for (int k = tid; k < capP; k += kClusterThreads) {      // every CTA pulls the new state
                float4 a = make_float4(kGhost, kGhost, 0.f, 0.f);
                if (k < cnt) a = spc0[k];
                S.X[buf][k] = a.x; S.Y[buf][k] = a.y; S.Tt[k] = a.z;
                S.X[buf ^ 1][k] = kGhost; S.Y[buf ^ 1][k] = kGhost;    // no stale nucleon beyond the new count
            }
            vel = make_float2(0.f, 0.f);
            if (owner) vel = remote(S.sv, 0)[i];
            __syncthreads();
            publish_sum(buf);
            cluster.sync();                                  // CTA 0's staging consumed; sums in place
        }

        // ---- centre of mass, nuclear_forces.py:242-243 ------------------------------------------------
        float cx, cy;
        if (e.centre) {
            cx = e.centre[2 * (int64_t)nuc];
            cy = e.centre[2 * (int64_t)nuc + 1];
        } else {
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int r = 0; r < kClusterSize; ++r) { sx += S.csum[buf][r].x; sy += S.csum[buf][r].y; }
            const float inv_n = 1.0f / (float)max(cnt, 1);
            cx = sx * inv_n; cy = sy * inv_n;
        }
        // ---- pair forces of nucleon i against the partners of slice sl, :248-298 ----------------------
        const bool has_i = il < per && i < cnt;
        const float xi = has_i ? S.X[buf][i] : kGhost, yi = has_i ? S.Y[buf][i] : kGhost;
        const float ti = has_i ? S.Tt[i] : 0.f;
        const f32x2 xi2 = pk1(xi), yi2 = pk1(yi), ti2 = pk1(ti), negC = pk1(-L.C);
        f32x2 fx2 = 0ull, fy2 = 0ull;
        const float2* X2 = reinterpret_cast<const float2*>(S.X[buf]);
        const float2* Y2 = reinterpret_cast<const float2*>(S.Y[buf]);
        const float2* T2 = reinterpret_cast<const float2*>(S.Tt);
#pragma unroll 2
        for (int j = j0; j < j1; j += 2) {                  // the self pair is skipped by d2 < 0.01 (:257)
            const float2 xj = X2[j >> 1], yj = Y2[j >> 1], tj = T2[j >> 1];
            const f32x2 dx = sub2(pk(xj.x, xj.y), xi2), dy = sub2(pk(yj.x, yj.y), yi2);
            const f32x2 sc = pair_general2(dx, dy, tj.x, tj.y, ti, ti2, mul2(negC, pk(tj.x, tj.y)), gc, L);
            fx2 = fma2(dx, sc, fx2);
            fy2 = fma2(dy, sc, fy2);
        }
        {
            float a, b, c2, d;
            upk(fx2, a, b);
            upk(fy2, c2, d);
            S.part[sl * ipc + il] = make_float2(a + b, c2 + d);
        }
        __syncthreads();
        const int nb = buf ^ 1;
        if (owner) {
            float fx = 0.f, fy = 0.f;
            for (int k = 0; k < slices; ++k) {              // fixed order: reproducible
                const float2 pf = S.part[k * ipc + il];
                fx += pf.x; fy += pf.y;
            }
            float x = xi, y = yi;
            const float Rn = 2.4f * cbrtf((float)cnt);      // nuclear_forces.py:304
            contain_and_integrate(x, y, vel.x, vel.y, fx, fy, cx, cy, Rn, e.dt_phys);   // :301-323
            if (e.force && s == n_steps - 1) reinterpret_cast<float2*>(e.force)[off + i] = make_float2(fx, fy);
#pragma unroll
            for (int r = 0; r < kClusterSize; ++r) {        // the new position into every replica
                remote(S.X[nb], r)[i] = x;
                remote(S.Y[nb], r)[i] = y;
            }
            S.X[nb][i] = x; S.Y[nb][i] = y;                 // (own replica: visible below after __syncthreads)
        }
        __syncthreads();                                    // this CTA's new positions are in its own replica
        publish_sum(nb);
        if (leader) decide(s + 1, (s + 1) & 1);
        cluster.sync();                                     // Jacobi: every replica complete before the next read
        buf = nb;
    }

    if (owner) {
        reinterpret_cast<float2*>(e.pos)[off + i] = make_float2(S.X[buf][i], S.Y[buf][i]);
        reinterpret_cast<float2*>(e.vel)[off + i] = vel;
        if (e.decay_enabled) e.is_proton[off + i] = (S.Tt[i] != 0.f) ? 1 : 0;
    }
    if (leader) {
        e.count[nuc] = cnt;
        if (e.decay_enabled) {
            e.zn[nuc] = zn;
            e.half_life[nuc] = T_half;
            e.p_decay[nuc] = p_dec;
        }
    }
    cluster.sync();                                         // nobody leaves while peers may still address its memory
}

}  // namespace pyqmd

using namespace pyqmd;

static const auto ring_single = ensemble_ring_kernel<false, 128, PYQMD_RING_MINBLOCKS_SINGLE>;
static const auto ring_pair = ensemble_ring_kernel<true, 64, PYQMD_RING_MINBLOCKS_PAIR>;
static const auto ring_multi = ensemble_ring_kernel<true, 256, 2>;

// cudaFuncSetAttribute is per device: remember which devices have been configured
static int configure_ensemble_kernels(void)
{
    static unsigned char done[64] = {0};
    static std::mutex mu;
    int dev = 0;
    PYQMD_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return PYQMD_OK;
    const void* kernels[3] = {(const void*)ring_single, (const void*)ring_pair, (const void*)ring_multi};
    for (const void* k : kernels) {
        PYQMD_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        PYQMD_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              cudaSharedmemCarveoutMaxShared));
    }
    const void* blockwide[2] = {(const void*)ensemble_pair_kernel<224>, (const void*)ensemble_pair_kernel<256>};
    for (const void* k : blockwide) {
        PYQMD_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        PYQMD_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              cudaSharedmemCarveoutMaxShared));
    }
    const void* quads[4] = {(const void*)ensemble_quad_kernel<128>, (const void*)ensemble_quad_kernel<160>,
                            (const void*)ensemble_quad_kernel<224>, (const void*)ensemble_quad_kernel<256>};
    for (const void* k : quads) {
        PYQMD_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        PYQMD_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              cudaSharedmemCarveoutMaxShared));
    }
    if (dev >= 0 && dev < 64) done[dev] = 1;
    return PYQMD_OK;
}

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
    const LawParams L = make_law_params(e->strong, e->coulomb, e->pauli);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = configure_ensemble_kernels();
    if (rc != PYQMD_OK) return rc;
    const RingGeom geo = ring_geom(e->cap);
    // Which ring?  Lane utilisation decides: warp-local rings (ensemble_ring_kernel) execute ~20 % fewer
    // instructions per pair, but a nucleus whose 4-nucleon subgroups do not fill its warps (Pb-208:
    // 26 of 32 lanes) loses more than that; the block-wide ring packs G nuclei into one block instead.
    // Measured on B200 (r02, pairs/s, ring vs block): U-238 1.21e12 / 1.10e12, Pb-208 1.05 / 1.15,
    // Au-197 0.98 / 1.08, Ag-107 1.02 / 1.05, Fe-56 0.85 / 0.75, C-14 0.36 / 0.31.
    // a handful of heavy nuclei: one nucleus per 8-CTA cluster (latency path)
    {
        const char* force = getenv("PYQMD_ENSEMBLE_KERNEL");
        const bool eligible = e->cap >= 64 && e->cap <= 512;
        // 8 SMs per nucleus, twice the pair evaluations (ordered pairs): 3.4 us per U-238 sub-step against
        // 21 us in one CTA, so it pays while the clusters run in at most ~4 waves (16-18 clusters fit at once)
        const bool want = force ? !strcmp(force, "cluster") : (n_list <= 64);
        if (eligible && want) {
            const int capP = (e->cap + 3) & ~3;
            const int ipc = e->cap <= 256 ? 32 : 64;
            const size_t sm = cluster_smem_bytes(capP, kClusterThreads / ipc, ipc);
            PYQMD_REQUIRE(n_list <= 2147483647LL / kClusterSize, "too many nuclei for one launch");
            ensemble_cluster_kernel<<<(unsigned)(n_list * kClusterSize), kClusterThreads, sm, st>>>(
                d, L, n_steps, capP, ipc);
            PYQMD_CUDA_CHECK(cudaGetLastError());
            return PYQMD_OK;
        }
    }
    // block-wide ring with four nucleons per thread (ensemble_quad_kernel): ring_visit's instruction count
    // at the block ring's lane utilisation.  PYQMD_ENSEMBLE_KERNEL=quad selects it.
    {
        const int capQ = (e->cap + kQ - 1) / kQ;
        int Gq = 0;
        const int Tq = capQ >= 32 ? pick_quad_threads(capQ, &Gq) : 0;
        const char* force = getenv("PYQMD_ENSEMBLE_KERNEL");
        bool use_quad = false;
        if (Tq > 0) {
            // Not dispatched automatically.  Measured on B200 (r02m, pairs/s, quad / block ring / warp rings):
            // one sub-step per launch Pb-208 1.166e12 / 1.144e12 / 1.045e12, Au-197 1.089e12 / 1.073e12 / 0.98e12,
            // U-238 1.170e12 / - / 1.200e12; FOUR fused sub-steps per launch (the app's frame) Pb-208 1.100e12
            // against the block ring's 1.171e12: its best shape, 3 nuclei on 160 threads, puts 5 warps per block
            // and 15 per SM on 4 schedulers, and the two block barriers of every sub-step wait for the scheduler
            // that holds two of them (ncu: barrier stalls 0.27 -> 1.06 per issued instruction from 1 to 4
            // sub-steps).  Kept as a pinned alternative; every parity case runs through it.
            if (force) use_quad = !strcmp(force, "quad");
        }
        if (use_quad) {
            const size_t sm = quad_smem_bytes(Tq, Gq, capQ);
            const int64_t gridq = (n_list + Gq - 1) / Gq;
            PYQMD_REQUIRE(gridq <= 2147483647LL, "too many nuclei for one launch");
            if (Tq == 128) ensemble_quad_kernel<128><<<(unsigned)gridq, Tq, sm, st>>>(d, L, n_steps, Gq, capQ);
            else if (Tq == 160) ensemble_quad_kernel<160><<<(unsigned)gridq, Tq, sm, st>>>(d, L, n_steps, Gq, capQ);
            else if (Tq == 224) ensemble_quad_kernel<224><<<(unsigned)gridq, Tq, sm, st>>>(d, L, n_steps, Gq, capQ);
            else ensemble_quad_kernel<256><<<(unsigned)gridq, Tq, sm, st>>>(d, L, n_steps, Gq, capQ);
            PYQMD_CUDA_CHECK(cudaGetLastError());
            return PYQMD_OK;
        }
    }
    bool use_ring = true;
    if (e->cap > 64 && e->cap <= 512) {
        const int S = (e->cap + kQ - 1) / kQ;
        const double u_ring = (double)S / (geo.nG * 32.0);
        const int capT = (e->cap + 1) / 2;
        int Gp = 1;
        const int Tp = pick_block_threads(capT, &Gp);
        const double u_block = (double)Gp * capT / Tp;
        use_ring = u_ring >= 0.92 * u_block || pair_smem_bytes(Tp, Gp, capT) > 200 * 1024;
        // a few hundred nuclei cannot fill the GPU: the step is a latency chain, and the block ring puts
        // twice as many threads on a nucleus (one U-238 alone: 21 us per sub-step against 27 us); below
        // 64 nuclei the cluster kernel above has already taken the launch
        int dev = 0, sms = 0;
        PYQMD_CUDA_CHECK(cudaGetDevice(&dev));
        PYQMD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (n_list < 2 * (int64_t)sms && pair_smem_bytes(Tp, Gp, capT) <= 200 * 1024) use_ring = false;
        // tests pin the choice so that both kernels see every parity case: PYQMD_ENSEMBLE_KERNEL=ring|block
        if (const char* force = getenv("PYQMD_ENSEMBLE_KERNEL")) {
            if (!strcmp(force, "ring")) use_ring = true;
            else if (!strcmp(force, "block") && pair_smem_bytes(Tp, Gp, capT) <= 200 * 1024) use_ring = false;
        }
    }
    if (!use_ring) {
        const int capT = (e->cap + 1) / 2;
        int Gp = 1;
        const int Tp = pick_block_threads(capT, &Gp);
        const size_t sm = pair_smem_bytes(Tp, Gp, capT);
        const int64_t gridp = (n_list + Gp - 1) / Gp;
        PYQMD_REQUIRE(gridp <= 2147483647LL, "too many nuclei for one launch");
        // blocks of <= 224 threads: 4 blocks / SM at 72 registers per thread
        if (Tp <= 224)
            ensemble_pair_kernel<224><<<(unsigned)gridp, Tp, sm, st>>>(d, L, n_steps, Gp, capT);
        else
            ensemble_pair_kernel<256><<<(unsigned)gridp, Tp, sm, st>>>(d, L, n_steps, Gp, capT);
        PYQMD_CUDA_CHECK(cudaGetLastError());
        return PYQMD_OK;
    }
    const int64_t grid = (n_list + geo.G - 1) / geo.G;
    PYQMD_REQUIRE(grid <= 2147483647LL, "too many nuclei for one launch");
    const size_t smem = ring_smem_bytes(geo);
    PYQMD_REQUIRE(smem <= 100 * 1024, "shared memory budget");
    if (geo.nG == 2)
        ring_pair<<<(unsigned)grid, 64, smem, st>>>(d, L, n_steps, geo);
    else if (geo.nG > 2)
        ring_multi<<<(unsigned)grid, geo.warps * 32, smem, st>>>(d, L, n_steps, geo);
    else
        ring_single<<<(unsigned)grid, geo.warps * 32, smem, st>>>(d, L, n_steps, geo);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
