/*
 * pyqmd_oracle.c -- CPU float64 restatement of PyQMD's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA library in
 * pyqmd_b200/csrc; it is never linked into, imported by, or called from the product path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it.
 *
 * Parity status: the reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the *reference itself*, imported unmodified in
 * the build container by tests/golden/gen_golden.py (fixtures in tests/golden/).  On the
 * build container it reproduces the reference's float64 results bit-for-bit (same glibc
 * libm, same operation order, CPython-3.12 compensated sum for the centre of mass).
 *
 * Every function cites the reference lines it follows; paths are relative to the
 * reference root (OtsoBear/PyQMD).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- CPython >= 3.12 builtin sum() over floats -----------------------------------------
 * nuclear_forces.py:242-243 and particles.py:207-208 compute the centre with
 * sum(p.x for p in particles) / len(particles).  Since CPython 3.12 sum() of floats uses
 * Neumaier compensated summation (Python/bltinmodule.c, cs_add / cs_to_double); restated
 * here so the oracle is bit-identical to the reference run under the image's Python 3.12.
 */
static double py312_sum(const double *v, int64_t n)
{
    double hi = 0.0, lo = 0.0;
    for (int64_t k = 0; k < n; ++k) {
        double x = v[k];
        double t = hi + x;
        if (fabs(hi) >= fabs(x))
            lo += (hi - t) + x;
        else
            lo += (x - t) + hi;
        hi = t;
    }
    if (lo != 0.0 && isfinite(lo))
        return hi + lo;
    return hi;
}

double orc_py312_mean(const double *v, int64_t n)
{
    return py312_sum(v, n) / (double)n;
}

/* Branch statistics for one step, used for the FLOP accounting of SURVEY.md section 8(d). */
typedef struct {
    int64_t evaluated;  /* ordered pairs that passed the d2 >= 0.01 test */
    int64_t skipped;    /* ordered pairs with d2 < 0.01 (nuclear_forces.py:257) */
    int64_t hard;       /* d < 4.25 */
    int64_t core;       /* d < 2.8 */
    int64_t attr;       /* 2.8 <= d < 9 */
    int64_t tail;       /* d >= 9 */
    int64_t pp;         /* both protons */
    int64_t pauli;      /* same type and d < 8 */
    int64_t clamped;    /* |net| hit the +-12 cap */
    int64_t contained;  /* nucleons that received the containment force */
} orc_branch_stats;

/* relative closeness of d (or d2) to a branch threshold */
static inline int near_thr(double v, double thr, double tol)
{
    return fabs(v - thr) <= tol * thr;
}

/*
 * One Jacobi force + damped-Euler step; follows NuclearForces.update_particles_cpu,
 * nuclear_forces.py:236-323, statement by statement.
 *
 *   x,y,vx,vy  in/out, length n  (Particle.x/.y/.vx/.vy, particles.py:24-29)
 *   is_proton  1 = ParticleType.PROTON, 0 = NEUTRON (particles.py:5-7)
 *   S,C,P      strong/coulomb/pauli strengths (nuclear_forces.py:13-15)
 *   fx,fy      optional out: the force on each nucleon *before* integration
 *   amb        optional out: 1 where nucleon i has a pair (or its containment test) within
 *              relative amb_tol of a discontinuous branch threshold (d2 = 0.01, d = 2.8, 4.25,
 *              8, 9; containment radius), i.e. where an FP32 evaluation may legitimately take
 *              the other branch ("hard parts"); the +-12 clamp is
 *              continuous and is not flagged
 *   st         optional out: branch statistics
 *   integrate  0 = compute forces only, leave state untouched
 */
void orc_force_step(int64_t n, double *x, double *y, double *vx, double *vy,
                    const uint8_t *is_proton, double S, double C, double P, double dt,
                    double *fx, double *fy, uint8_t *amb, double amb_tol,
                    orc_branch_stats *st, int integrate)
{
    if (n <= 0) return;                                  /* :238-239 */
    orc_branch_stats z;
    memset(&z, 0, sizeof z);

    /* :242-243 centre of mass of the step-start positions */
    double center_x = py312_sum(x, n) / (double)n;
    double center_y = py312_sum(y, n) / (double)n;

    double *Fx = (double *)malloc(sizeof(double) * (size_t)n);   /* :246 */
    double *Fy = (double *)malloc(sizeof(double) * (size_t)n);
    /* :304 -- pow(len(particles), 1.0/3.0): int ** float goes through libm pow */
    double nuclear_radius = 1.2 * pow((double)n, 1.0 / 3.0) * 2.0;

    for (int64_t i = 0; i < n; ++i) {                    /* :248 */
        double f0 = 0.0, f1 = 0.0;
        int a = 0;
        for (int64_t j = 0; j < n; ++j) {                /* :249 */
            if (i == j) continue;                        /* :250-251 */
            double dx = x[j] - x[i];                     /* :253 */
            double dy = y[j] - y[i];                     /* :254 */
            double dist2 = dx * dx + dy * dy;            /* :255 */
            if (amb && near_thr(dist2, 0.01, amb_tol)) a = 1;
            if (dist2 < 0.01) { z.skipped++; continue; } /* :257-258 */
            double dist = sqrt(dist2);                   /* :260 */
            double net = 0.0;                            /* :261 */
            z.evaluated++;
            if (amb && (near_thr(dist, 4.25, amb_tol) || near_thr(dist, 2.8, amb_tol) ||
                        near_thr(dist, 9.0, amb_tol)))
                a = 1;

            const double min_allowed = 4.25;             /* :264 */
            if (dist < min_allowed) {                    /* :265 */
                double overlap = min_allowed - dist;     /* :266 */
                net -= 60.0 * pow(overlap / min_allowed, 1.5);   /* :267 */
                z.hard++;
            }
            double r_ratio = dist / 7.0;                 /* :270-271 */
            if (dist < 2.8) {                            /* :273 */
                net -= 0.7 * S / (dist2 + 0.15);         /* :275 */
                z.core++;
            } else if (dist < 9.0) {                     /* :276 */
                net += 1.25 * S * exp(-r_ratio) / (dist + 0.15);         /* :278 */
                z.attr++;
            } else {
                net += 0.15 * S * exp(-r_ratio * 1.8) / (dist + 0.15);   /* :281 */
                z.tail++;
            }
            if (is_proton[i] && is_proton[j]) {          /* :284 */
                net -= C / (dist2 + 0.15);               /* :285 */
                z.pp++;
            }
            if (is_proton[i] == is_proton[j]) {          /* :288 */
                if (amb && near_thr(dist, 8.0, amb_tol)) a = 1;
                if (dist < 8.0) {                        /* :289-290 */
                    net -= P * exp(-dist / 8.0 * 2.0);   /* :291 */
                    z.pauli++;
                }
            }
            /* :294  max(-12.0, min(12.0, net)) with Python's first-wins tie rule */
            double m = (net < 12.0) ? net : 12.0;
            double c = (m > -12.0) ? m : -12.0;
            if (c != net) z.clamped++;
            net = c;
            if (dist > 0) {                              /* :296 */
                f0 += dx * net / dist;                   /* :297 */
                f1 += dy * net / dist;                   /* :298 */
            }
        }
        /* :301-309 centre-of-mass containment */
        double cdx = center_x - x[i];
        double cdy = center_y - y[i];
        double cdist = sqrt(pow(cdx, 2.0) + pow(cdy, 2.0));       /* :303 (x**2) */
        if (amb && (near_thr(cdist, nuclear_radius * 1.5, amb_tol))) a = 1;
        if (cdist > nuclear_radius * 1.5 && cdist > 0.01) {      /* :306 */
            double cf = 0.03 * (cdist - nuclear_radius);         /* :307 */
            f0 += cf * cdx / cdist;                              /* :308 */
            f1 += cf * cdy / cdist;                              /* :309 */
            z.contained++;
        }
        Fx[i] = f0;
        Fy[i] = f1;
        if (amb) amb[i] = (uint8_t)a;
    }

    if (integrate) {
        for (int64_t i = 0; i < n; ++i) {                /* :312-323 */
            vx[i] += Fx[i] * dt;
            vy[i] += Fy[i] * dt;
            vx[i] *= 0.85;
            vy[i] *= 0.85;
            x[i] += vx[i] * dt;
            y[i] += vy[i] * dt;
        }
    }
    if (fx) memcpy(fx, Fx, sizeof(double) * (size_t)n);
    if (fy) memcpy(fy, Fy, sizeof(double) * (size_t)n);
    if (st) *st = z;
    free(Fx);
    free(Fy);
}

/*
 * n_steps consecutive steps of many independent nuclei stored CSR-style
 * (offsets[k]..offsets[k]+count[k]); OpenMP over nuclei.  This is the loop
 * nuclear_sim.py:169-173 would run for each nucleus, batched for the CPU baseline.
 * Returns the number of ordered pairs visited (sum of n*(n-1) per nucleus-step).
 */
int64_t orc_ensemble_force_steps(int64_t n_nuclei, const int64_t *offsets, const int32_t *count,
                                 double *x, double *y, double *vx, double *vy,
                                 const uint8_t *is_proton, double S, double C, double P,
                                 double dt, int n_steps, int n_threads)
{
    int64_t pairs = 0;
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());   /* 0 = all host cores */
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : pairs)
    for (int64_t k = 0; k < n_nuclei; ++k) {
        int64_t o = offsets[k];
        int64_t n = count[k];
        for (int s = 0; s < n_steps; ++s)
            orc_force_step(n, x + o, y + o, vx + o, vy + o, is_proton + o, S, C, P, dt, NULL,
                           NULL, NULL, 0.0, NULL, 1);
        pairs += n * (n - 1) * (int64_t)n_steps;
    }
    return pairs;
}

/*
 * Forces only, i-range [i0, i1) of a large cloud, OpenMP over i; for the cloud CPU
 * baseline and for parity checks of cloud sub-blocks.  Same law as orc_force_step
 * (nuclear_forces.py:248-309) with the centre supplied by the caller.
 */
void orc_cloud_forces(int64_t n, const double *x, const double *y, const uint8_t *is_proton,
                      double S, double C, double P, double center_x, double center_y,
                      int64_t i0, int64_t i1, double *fx, double *fy, int n_threads)
{
    double nuclear_radius = 1.2 * pow((double)n, 1.0 / 3.0) * 2.0;
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());   /* 0 = all host cores */
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i) {
        double f0 = 0.0, f1 = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            if (i == j) continue;
            double dx = x[j] - x[i];
            double dy = y[j] - y[i];
            double dist2 = dx * dx + dy * dy;
            if (dist2 < 0.01) continue;
            double dist = sqrt(dist2);
            double net = 0.0;
            if (dist < 4.25) net -= 60.0 * pow((4.25 - dist) / 4.25, 1.5);
            double r_ratio = dist / 7.0;
            if (dist < 2.8)
                net -= 0.7 * S / (dist2 + 0.15);
            else if (dist < 9.0)
                net += 1.25 * S * exp(-r_ratio) / (dist + 0.15);
            else
                net += 0.15 * S * exp(-r_ratio * 1.8) / (dist + 0.15);
            if (is_proton[i] && is_proton[j]) net -= C / (dist2 + 0.15);
            if (is_proton[i] == is_proton[j] && dist < 8.0) net -= P * exp(-dist / 8.0 * 2.0);
            double m = (net < 12.0) ? net : 12.0;
            net = (m > -12.0) ? m : -12.0;
            f0 += dx * net / dist;
            f1 += dy * net / dist;
        }
        double cdx = center_x - x[i];
        double cdy = center_y - y[i];
        double cdist = sqrt(pow(cdx, 2.0) + pow(cdy, 2.0));
        if (cdist > nuclear_radius * 1.5 && cdist > 0.01) {
            double cf = 0.03 * (cdist - nuclear_radius);
            f0 += cf * cdx / cdist;
            f1 += cf * cdy / cdist;
        }
        fx[i - i0] = f0;
        fy[i - i0] = f1;
    }
}

/*
 * Per-frame overlap projection; follows NuclearSimulation.resolve_overlaps,
 * nuclear_sim.py:355-379: sequential Gauss-Seidel over i < j with immediate updates, minimum
 * distance 5.0 (= 2 * radius).  The degenerate case dist < 0.001 (:367-370) draws
 * random.uniform(0, 2*pi); the draws are taken in order from `uniforms` (as u in [0,1)).
 * Returns the number of draws consumed, or -1 if `uniforms` ran out.
 */
int64_t orc_resolve_overlaps(int64_t n, double *x, double *y, const double *uniforms,
                             int64_t n_uniforms, int64_t *n_pushes)
{
    const double min_dist = 5.0;                         /* :357 */
    int64_t used = 0, pushes = 0;
    for (int64_t i = 0; i < n; ++i) {                    /* :359 */
        for (int64_t j = i + 1; j < n; ++j) {            /* :360 */
            double dx = x[j] - x[i];                     /* :361 */
            double dy = y[j] - y[i];                     /* :362 */
            double dist2 = dx * dx + dy * dy;            /* :363 */
            if (dist2 < min_dist * min_dist) {           /* :365 */
                double dist = sqrt(dist2);               /* :366 */
                if (dist < 0.001) {                      /* :367 */
                    if (used >= n_uniforms) return -1;
                    double angle = 0.0 + (2.0 * 3.141592653589793 - 0.0) * uniforms[used++];  /* :368 */
                    dx = cos(angle);                     /* :369 */
                    dy = sin(angle);
                    dist = 0.001;                        /* :370 */
                } else {
                    dx /= dist;                          /* :372 */
                    dy /= dist;                          /* :373 */
                }
                double push = (min_dist - dist) * 0.5;   /* :375 */
                x[i] -= dx * push;                       /* :376 */
                y[i] -= dy * push;                       /* :377 */
                x[j] += dx * push;                       /* :378 */
                y[j] += dy * push;                       /* :379 */
                ++pushes;
            }
        }
    }
    if (n_pushes) *n_pushes = pushes;
    return used;
}

/*
 * Decay probability for one sub-step; follows Nucleus.should_decay,
 * particles.py:126-147 (identical copy at decay_chains.py:400-421).
 * Returns -1.0 for a stable nucleus (T = inf): the reference returns False *without*
 * drawing (particles.py:129-130).  Otherwise the caller decides with  u < p  (:147).
 */
double orc_decay_probability(double T, double dt)
{
    if (isinf(T) && T > 0) return -1.0;                  /* :129-130 */
    double p;
    if (dt > T * 0.01)                                   /* :134 */
        p = 1.0 - pow(0.5, dt / T);                      /* :136 */
    else {
        double decay_constant = 0.693 / T;               /* :140 */
        p = decay_constant * dt;                         /* :141 */
    }
    /* :144  max(0.0, min(1.0, p)) */
    double m = (p < 1.0) ? p : 1.0;
    p = (m > 0.0) ? m : 0.0;
    return p;
}

/*
 * Decay decisions for a population of particle-less nuclei (decay_chains.py:390-421),
 * one draw per unstable nucleus: out[k] = u[k] < p(T[k], dt).  Stable nuclei give 0 and
 * are reported in consumed[k] = 0 (no draw taken).  OpenMP over nuclei.
 */
int64_t orc_decay_decisions(int64_t n, const double *T, double dt, const double *u,
                            uint8_t *out, uint8_t *consumed, int n_threads)
{
    int64_t fired = 0;
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());   /* 0 = all host cores */
#endif
#pragma omp parallel for schedule(static) reduction(+ : fired)
    for (int64_t k = 0; k < n; ++k) {
        double p = orc_decay_probability(T[k], dt);
        int d = 0, c = 0;
        if (p >= 0.0) {
            c = 1;
            d = u[k] < p;                                /* particles.py:147 */
        }
        out[k] = (uint8_t)d;
        if (consumed) consumed[k] = (uint8_t)c;
        fired += d;
    }
    return fired;
}

/*
 * CPython's random.random() word-to-double map (Modules/_randommodule.c,
 * _random_Random_random_impl): a = w0 >> 5, b = w1 >> 6, (a*2^26 + b) / 2^53.
 * The CUDA decay kernels build their uniforms from two Philox words the same way.
 */
double orc_u53(uint32_t w0, uint32_t w1)
{
    uint32_t a = w0 >> 5, b = w1 >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

/* Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11;
 * Random123 v1.x philox.h).  The reference has no counter-based RNG; this restates the
 * published algorithm the CUDA kernels use, pinned by the Random123 known-answer vectors
 * in tests/test_oracle.py. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Uniform for (seed, nucleus id, step, slot) exactly as the CUDA kernels define it
 * (csrc/decay_device.cuh, DrawSource), key = (seed_lo, seed_hi):
 *   slot 0     counter (id >> 1, step, 0), words (0,1) for even ids, (2,3) for odd ids
 *   slots 1,2  counter (id, step, 1), words (0,1) / (2,3)
 *   slot 3     counter (id, step, 2), words (0,1) */
double orc_philox_uniform(uint64_t seed, uint64_t id, uint32_t step, uint32_t slot)
{
    uint64_t c = (slot == 0) ? (id >> 1) : id;
    uint32_t ctr[4] = {(uint32_t)c, (uint32_t)(c >> 32), step, slot == 0 ? 0u : (slot == 3 ? 2u : 1u)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t w[4];
    int hi = (slot == 0) ? (int)(id & 1) : (slot == 2);
    orc_philox4x32_10(ctr, key, w);
    return hi ? orc_u53(w[2], w[3]) : orc_u53(w[0], w[1]);
}

void orc_philox_uniforms(uint64_t seed, uint64_t id0, int64_t n, uint32_t step, uint32_t slot,
                         double *out)
{
    for (int64_t k = 0; k < n; ++k) out[k] = orc_philox_uniform(seed, id0 + (uint64_t)k, step, slot);
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();      /* not omp_get_max_threads(): that follows the last omp_set_num_threads */
#else
    return 1;
#endif
}
