/*
 * pyqmd_b200.h -- C ABI of libpyqmd_b200.so: PyQMD's hot path (all-pairs nucleon force ->
 * damped Euler integrate -> per-nucleus stochastic decay) on NVIDIA B200 (sm_100a).
 *
 * Plain pointers and sizes only; no torch / C++ types.  All device pointers are caller
 * owned (e.g. torch tensors).  Entry points that take DEVICE pointers are stream-ordered and
 * non-blocking.  Of the entry points that take HOST arrays, section (A) --
 * pyqmd_update_forces_and_positions, pyqmd_update_particles_f64, pyqmd_cloud_step_host -- BLOCKS until
 * the result is back in the caller's arrays (like the reference's event.wait() + .get(),
 * nuclear_forces.py:221,227), while pyqmd_ensemble_step_host only ENQUEUES its chunked copies and
 * kernels and joins them into `stream`: its host arrays are valid after that stream has been
 * synchronised.  Return value: 0 on success, negative PYQMD_ERR_* otherwise (no
 * exceptions cross the ABI); pyqmd_last_error() gives the message for the calling thread.
 *
 * Each entry point names the reference interface it replaces (file:line relative to the
 * root of OtsoBear/PyQMD).  The reference binds its device code through PyOpenCL
 * (nuclear_forces.py:174-183, 212-220); INTEGRATION.md shows the ctypes stub that replaces
 * that binding.
 */
#ifndef PYQMD_B200_H
#define PYQMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYQMD_ABI_VERSION 1

#define PYQMD_OK 0
#define PYQMD_ERR_INVALID (-1)   /* bad argument */
#define PYQMD_ERR_CUDA (-2)      /* CUDA runtime error (see pyqmd_last_error) */
#define PYQMD_ERR_NO_DEVICE (-3) /* no usable sm_100 device */
#define PYQMD_ERR_CAPACITY (-4)  /* a caller-supplied buffer is too small */

/* ParticleType, particles.py:5-11 */
enum {
    PYQMD_PT_PROTON = 0, PYQMD_PT_NEUTRON = 1, PYQMD_PT_ALPHA = 2, PYQMD_PT_ELECTRON = 3,
    PYQMD_PT_GAMMA = 4, PYQMD_PT_POSITRON = 5
};
/* DecayType, particles.py:13-21 */
enum {
    PYQMD_DECAY_NONE = 0, PYQMD_DECAY_ALPHA = 1, PYQMD_DECAY_BETA_MINUS = 2,
    PYQMD_DECAY_BETA_PLUS = 3, PYQMD_DECAY_GAMMA = 4, PYQMD_DECAY_NEUTRON = 5,
    PYQMD_DECAY_PROTON = 6, PYQMD_DECAY_FISSION = 7
};

/* ---------------------------------------------------------------------------------------- */
/* library                                                                                   */
int pyqmd_abi_version(void);
const char *pyqmd_last_error(void);
/* out[0]=SM count, [1]=cc major, [2]=cc minor, [3]=max SM clock kHz, [4]=L2 bytes,
 * [5]=max dynamic smem per block, [6]=total global memory MiB, [7]=reserved */
int pyqmd_device_props(int device, int64_t out[8]);
/* sizeof of the four ABI structs below, for binding self-checks:
 * nuclide_entry, decay_event, ensemble, population */
int pyqmd_struct_sizes(int64_t out[4]);
/* FP32 FMA peak microbenchmark on the current device (dependent FFMA chains, 8 per thread):
 * writes achieved TFLOP/s for scalar FFMA and for packed fma.rn.f32x2. */
int pyqmd_fp32_peak(int iters, double *tflops_ffma, double *tflops_ffma2, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* (A) reference-shaped, host buffers, blocking                                              */

/*
 * Replaces the OpenCL kernel launch inside NuclearForces.update_particles_gpu
 * (nuclear_forces.py:202-227: H2D x2, kernel `update_forces_and_positions` :60-68, wait, D2H)
 * with the same argument list: `particles` = float32[n][4] (x, y, vx, vy) as h_particles
 * (:190-198), `types` = int32[n], 0 proton / 1 neutron (:199), centre (:206-208), strengths
 * (:216-218), dt (:219).  Updated in place.  Unlike the reference kernel (in-place, racy,
 * :168-171) the step is Jacobi, i.e. the semantics of update_particles_cpu (:236-323).
 */
int pyqmd_update_forces_and_positions(float *particles, const int32_t *types, int32_t num_particles,
                                      float center_x, float center_y, float strong_strength,
                                      float coulomb_strength, float pauli_strength, float dt);

/*
 * Replaces NuclearForces.update_particles_cpu / _gpu as a whole (nuclear_forces.py:185-323)
 * for callers that keep float64 state (Particle.x/.y/.vx/.vy are Python floats,
 * particles.py:24-29): centre of mass (:242-243) is taken in float64, positions are made
 * nucleus-relative before the FP32 device step and restored afterwards, `n_steps` sub-steps
 * run back to back on the device (nuclear_sim.py:161-173 without decay).  n <= 1024 runs as one
 * thread block (ensemble kernel); larger systems take the sorted symmetric scheme of
 * pyqmd_cloud_step_host.  Blocking.
 */
int pyqmd_update_particles_f64(double *x, double *y, double *vx, double *vy,
                               const uint8_t *is_proton, int64_t n, double strong_strength,
                               double coulomb_strength, double pauli_strength, double dt,
                               int32_t n_steps);

/*
 * The reference's per-step call (nuclear_forces.py:185-234: pack -> upload -> kernel -> wait ->
 * download -> write back) for ONE system of any size whose state lives in host arrays:
 *   h_pos, h_vel  float2[n] (in/out; pinned memory copies fastest, pageable works)
 *   h_is_proton   uint8[n]
 *   h_force       optional float2[n] (may be NULL): force of the last step, containment included
 * Uploads, sorts once on the device (type bit + Morton code, stable radix sort), runs n_steps
 * Jacobi steps of the symmetric scheme (every unordered pair once, as pyqmd_cloud_pair_forces +
 * pyqmd_cloud_integrate below), un-sorts and downloads.  The caller's nucleon order is kept.
 * Positions should be relative to a nearby origin (FP32).  Blocking.
 */
int pyqmd_cloud_step_host(float *h_pos, float *h_vel, const uint8_t *h_is_proton, float *h_force,
                          int64_t n, float strong, float coulomb, float pauli, float dt,
                          int32_t n_steps);

/* ---------------------------------------------------------------------------------------- */
/* (B) one large nucleon cloud resident on the device (BASELINE config 4)                    */

/* bytes of scratch the cloud entry points need for n nucleons */
int64_t pyqmd_cloud_workspace_bytes(int64_t n);

/*
 * One Jacobi step of nuclear_forces.py:236-323 for nucleons [i0, i1) of an n-nucleon cloud.
 *   pos_in   float2[n]  positions at step start (a full replica on every GPU)
 *   pos_out  float2[n]  new positions, written for [i0, i1) only (all-gather afterwards)
 *   vel      float2[n]  velocities, updated in place for [i0, i1)
 *   force    float2[n]  optional (may be NULL): pre-integration force for [i0, i1)
 *   is_proton uint8[n]  1 proton / 0 neutron
 * Correct for any particle order; fastest when the caller keeps the cloud partitioned by
 * type and spatially sorted (pyqmd_cloud_sort_keys), which lets whole 256-nucleon tiles take
 * the far-field fast path.  The centre of mass (:242-243) is reduced on the device in a
 * fixed order, identically on every rank.
 */
int pyqmd_cloud_step(const float *pos_in, float *pos_out, float *vel, float *force,
                     const uint8_t *is_proton, int64_t n, int64_t i0, int64_t i1, float strong,
                     float coulomb, float pauli, float dt, void *workspace, void *stream);

/*
 * The same step split in two, with every unordered pair evaluated once (the law is symmetric, so
 * the force on j is the negated force on i, nuclear_forces.py:253-298) -- about twice as fast as
 * the ordered i-block scheme of pyqmd_cloud_step on any number of GPUs:
 *
 *   pyqmd_cloud_pair_forces  adds this part's share (`part` of `n_parts`: i-block rows of 1024
 *       nucleons dealt boustrophedon-wise, each with the j tiles at and after its diagonal) of the
 *       pair forces on ALL n nucleons into force_acc, int64[n][2] fixed point with scale
 *       2^pyqmd_cloud_force_scale_log2(n).  Integer accumulation is associative: the sum over
 *       parts (an integer reduce-scatter between GPUs) is bit-identical for any n_parts.
 *       force_acc must be zero before the first part is added.
 *   pyqmd_cloud_integrate    consumes and CLEARS the accumulators of [i0, i1) (force_acc_i0 points
 *       at the entry of nucleon i0): containment (:301-309) + damped Euler (:312-323).  Uses the
 *       centre of mass left in `workspace` by pyqmd_cloud_pair_forces.
 */
int32_t pyqmd_cloud_force_scale_log2(int64_t n);
int pyqmd_cloud_pair_forces(const float *pos, const uint8_t *is_proton, int64_t n, int32_t part,
                            int32_t n_parts, float strong, float coulomb, float pauli,
                            long long *force_acc, void *workspace, void *stream);
/*
 * pyqmd_cloud_pair_forces with option flags.
 * PYQMD_CLOUD_SKIP_EXACT_ZEROS (off by default): beyond d = 353 the tail term
 * 0.15 S exp(-1.8 d / 7) / (d + eps) is exactly +0 in the kernel's FP32 arithmetic (its 2^x argument is
 * below -126 and ex2.approx.ftz flushes to zero), so a pair that far apart contributes its Coulomb term
 * only, i.e. nothing unless both nucleons are protons.  With the flag, 256-nucleon tiles whose bounding
 * boxes are further apart than that skip the exponential, and skip the tile altogether when one side has
 * no protons.  The accumulators come out BIT-IDENTICAL to the default (tests/test_gpu_cloud.py); only the
 * time changes.  It is opt-in because the benchmark metric counts N (N - 1) pair evaluations per step
 *  and a kernel that proves most of them zero is no longer "evaluating all pairs".
 * Ignored (default behaviour) when the strengths do not allow the proof (S > 180).
 */
#define PYQMD_CLOUD_SKIP_EXACT_ZEROS 1u
int pyqmd_cloud_pair_forces_ex(const float *pos, const uint8_t *is_proton, int64_t n, int32_t part,
                               int32_t n_parts, float strong, float coulomb, float pauli,
                               long long *force_acc, void *workspace, uint32_t flags, void *stream);
int pyqmd_cloud_integrate(const float *pos_in, float *pos_out, float *vel, float *force, int64_t n,
                          int64_t i0, int64_t i1, float dt, long long *force_acc_i0,
                          void *workspace, void *stream);

/*
 * Multi-GPU epilogue of the symmetric scheme over PEER MEMORY (NVLink / NVSwitch, no collective
 * library on the data path): reduce-scatter of the force accumulators + integrate + all-gather of
 * the new positions in one kernel.  acc_peers / pos_out_peers are DEVICE arrays of n_peers
 * pointers: rank p's accumulator array (int64[n][2]) and rank p's replica of the next positions
 * (float2[n]), all mapped into this process (CUDA IPC / symmetric memory).  The owner of [i0, i1)
 * pulls and clears acc_peers[p][i], integrates (:301-323) and stores the new position into every
 * pos_out_peers[p][i].  The caller must barrier all ranks before (every pyqmd_cloud_pair_forces
 * finished) and after (every push landed) the call.
 */
int pyqmd_cloud_exchange_integrate(const float *pos_in, float *vel, float *force, int64_t n,
                                   int64_t i0, int64_t i1, float dt, long long *const *acc_peers,
                                   float *const *pos_out_peers, int32_t n_peers, void *workspace,
                                   void *stream);

/* 64-bit sort keys: bit 62 = neutron (protons sort first, signed or unsigned), low 48 bits = 2-D
 * Morton code of the position inside [xmin, xmin+extent) x [ymin, ymin+extent).  A STABLE sort by
 * key gives the layout above (stability matters when several ranks sort their replicas
 * independently: equal keys must end up in the same order everywhere). */
int pyqmd_cloud_sort_keys(const float *pos, const uint8_t *is_proton, int64_t n, float xmin,
                          float ymin, float extent, uint64_t *keys, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* (C)+(D) decay tables                                                                      */

#define PYQMD_TABLE_ZDIM 128
#define PYQMD_TABLE_NDIM 192
enum { PYQMD_HL_INF = 0, PYQMD_HL_TABLE = 1, PYQMD_HL_BAND = 2 };

/* One (Z, N) row (88 bytes) of the dense nuclide table, index Z * PYQMD_TABLE_NDIM + N.
 * Built on the host from HALF_LIVES / DECAY_CHAINS and the two heuristics
 * (decay_chains.py:13-167, 169-201, 247-328) by pyqmd_b200/nuclides.py. */
typedef struct {
    double half_life;   /* PYQMD_HL_TABLE: database value, seconds (may be +inf)  :257-262 */
    double p_decay;     /* probability per sub-step for this run's dt_decay, computed on the
                           host exactly as particles.py:134-144; < 0 = stable (no draw); NaN for
                           PYQMD_HL_BAND rows (the value is per nucleus, not per nuclide) */
    double band_a, band_b, band_unit; /* PYQMD_HL_BAND: 10**uniform(a,b)*unit     :311-328 */
    double opt_cum[2];  /* running sums of the branch probabilities               :222-227 */
    int32_t opt_zn[2];  /* daughter (Z << 16) | N                                             */
    int32_t opt_mode[2];/* PYQMD_DECAY_*                                                      */
    int32_t n_opt;      /* 1 or 2                                                             */
    int32_t kind;       /* PYQMD_HL_*                                                         */
    uint64_t p_thr;     /* ceil(p_decay * 2^53): `random() < p` (particles.py:147) with random() =
                           m / 2^53, m a 53-bit integer, is EXACTLY `m < p_thr` -- the decay-only
                           kernel compares integers; 0 = stable, PYQMD_THR_PER_NUCLEUS for
                           PYQMD_HL_BAND rows */
} pyqmd_nuclide_entry;
#define PYQMD_THR_PER_NUCLEUS 0xFFFFFFFFFFFFFFFFull

/* One decay event (what handle_decay appends to self.particles, nuclear_sim.py:294,349). */
typedef struct {
    int64_t nucleus;    /* global nucleus id */
    int32_t step;       /* sub-step index */
    int32_t mode;       /* PYQMD_DECAY_* */
    int32_t zn_new;     /* daughter (Z << 16) | N */
    int32_t ptype;      /* emitted PYQMD_PT_*, -1 if none */
    double x, y;        /* emission point = new centre of mass (:291-294) + origin */
    double vx, vy;      /* speed * (cos, sin)(2 pi u)  decay_chains.py:331-371 */
} pyqmd_decay_event;

/* ---------------------------------------------------------------------------------------- */
/* (C) ensembles of independent nuclei, one (or several small) nuclei per thread block       */

typedef struct {
    /* nucleon state, CSR by nucleus, nucleus-relative FP32 (Particle, particles.py:23-39) */
    float *pos;               /* float2[total slots] */
    float *vel;               /* float2[total slots] */
    uint8_t *is_proton;       /* [total slots] */
    float *force;             /* optional float2[total slots]: force of the last sub-step,
                                 containment included, before integration; NULL = not kept */
    const int64_t *offset;    /* [n_nuclei] first slot of each nucleus */
    int32_t *count;           /* [n_nuclei] live nucleons (shrinks on alpha / n / p emission) */
    /* nucleus state (Nucleus, particles.py:52-60) */
    int32_t *zn;              /* (Z << 16) | N */
    double *half_life;        /* Nucleus.stability */
    double *p_decay;          /* probability per sub-step; < 0 = stable */
    const double *origin;     /* double[n_nuclei][2] added to emission points, or NULL */
    const float *centre;      /* optional float2[n_nuclei]: containment centre supplied by the
                                 caller (the `center` kernel argument, nuclear_forces.py:64)
                                 instead of the mean position; NULL = compute (:242-243) */
    int64_t n_nuclei;
    int64_t id_base;          /* global id of nucleus 0 (RNG counters, event records) */
    /* the nuclei this launch handles */
    const int32_t *list;      /* [n_list] local nucleus indices, or NULL = 0..n_nuclei-1 */
    int64_t n_list;
    int32_t cap;              /* max nucleons of any listed nucleus (<= 1024) */
    int32_t decay_enabled;
    /* physics */
    float strong, coulomb, pauli, dt_phys;   /* nuclear_forces.py:13-15; nuclear_sim.py:145 */
    double dt_decay;                         /* nuclear_sim.py:165 */
    /* decay inputs */
    const pyqmd_nuclide_entry *table;
    const double *uniforms;   /* optional [n_steps][uniforms_n][4] draws, else Philox(seed) */
    int64_t uniforms_n;
    uint64_t seed;
    uint32_t step0;           /* index of the first sub-step of this call */
    uint32_t reserved;
    /* outputs */
    pyqmd_decay_event *events;
    int64_t event_capacity;
    unsigned long long *event_count;   /* device counter; events beyond capacity are counted
                                          but not stored */
    unsigned long long *mode_counts;   /* [8] decays by PYQMD_DECAY_* */
} pyqmd_ensemble;

/*
 * n_steps passes of the sub-step loop body nuclear_sim.py:165-173 for every listed nucleus:
 * should_decay (particles.py:126-147) -> physics slice of handle_decay (nuclear_sim.py:213,
 * 288-294, 349, 353) -> force + integrate (nuclear_forces.py:236-323).  State stays in shared
 * memory between the n_steps sub-steps.
 */
int pyqmd_ensemble_step(const pyqmd_ensemble *e, int32_t n_steps, void *stream);

/*
 * The same sub-steps for an ensemble whose state lives in HOST memory -- the shape of the reference's
 * per-step call (nuclear_forces.py:190-234: pack, upload, kernel, download, write back) for many
 * nuclei: every call uploads pos / vel / is_proton, runs n_steps sub-steps and downloads pos / vel
 * (and is_proton / count / zn when decay is enabled).  The work is cut into `n_chunks` chunks that
 * flow through three in-order lanes (upload, compute, download) linked by events, so both copy
 * engines run back to back and the kernels hide under them.  Host arrays should be pinned.
 *   e          descriptor whose pointers are DEVICE buffers of the same layout (staging area)
 *   h_*        host arrays with the layout of e->pos, e->vel, e->is_proton, e->count, e->zn;
 *              h_is_proton may be NULL when decay is disabled (the types already on the device are
 *              kept: without decay they cannot change, so one upload is enough)
 *   chunks     n_chunks descriptors: nuclei [nuc0, nuc1), slots [slot0, slot1) and the kernel
 *              launches of the chunk (one per size bin: cap, DEVICE index list, its length)
 * Stream-ordered on `stream`, non-blocking.
 */
#define PYQMD_MAX_CHUNK_LAUNCHES 12
typedef struct {
    int64_t nuc0, nuc1, slot0, slot1;
    int32_t n_launch;
    int32_t cap[PYQMD_MAX_CHUNK_LAUNCHES];
    const int32_t *list[PYQMD_MAX_CHUNK_LAUNCHES];
    int64_t n_list[PYQMD_MAX_CHUNK_LAUNCHES];
} pyqmd_host_chunk;

int pyqmd_ensemble_step_host(const pyqmd_ensemble *e, float *h_pos, float *h_vel, uint8_t *h_is_proton,
                             int32_t *h_count, int32_t *h_zn, const pyqmd_host_chunk *chunks,
                             int32_t n_chunks, int32_t n_steps, void *stream);

/*
 * Per-frame overlap projection of every listed nucleus: NuclearSimulation.resolve_overlaps,
 * nuclear_sim.py:355-379 (sequential i < j sweep, minimum distance 5.0, immediate updates), run
 * once per frame after the sub-steps (:175-176).  Only pos / offset / count / list / cap /
 * id_base / seed / step0 of the descriptor are read.  `uniforms` (optional,
 * double[n_nuclei][uniforms_per_nucleus], values in [0,1)) feeds the random direction of the
 * degenerate case dist < 0.001 (:367-370) in consumption order; otherwise Philox(seed).
 * `n_pushes` (optional device counter) accumulates the number of pushes applied.
 */
int pyqmd_resolve_overlaps(const pyqmd_ensemble *e, const double *uniforms,
                           int32_t uniforms_per_nucleus, unsigned long long *n_pushes,
                           void *stream);

/*
 * Initial nucleon layout of every listed nucleus: Nucleus.initialize_particles, particles.py:62-124
 * (shell-by-shell proton/neutron pairs, 20 candidate angles per nucleon, keep the one farthest from
 * its nearest same-type neighbour).  Reads zn / offset / list / cap / id_base; writes pos (nucleus
 * relative, FP32), vel = 0, is_proton and count = Z + N.  All listed nuclei must have the same
 * nucleon count A <= cap; `shell_radii` = double[7], initial_radius * (i + 1) / 7 with
 * initial_radius = 1.2 * A**(1/3) * 0.7 computed by the caller (:64-68).  `uniforms` (optional,
 * device double[n_list][cap][21]): the draws of each placement in consumption order (radius factor,
 * then 20 angles); otherwise Philox(seed).
 */
int pyqmd_ensemble_init_layout(const pyqmd_ensemble *e, const double *shell_radii,
                               const double *uniforms, unsigned long long seed, void *stream);

/*
 * Branch census of the pair law over every ordered pair of the listed nuclei
 * (nuclear_forces.py:257-291), for the algorithmic-FLOP accounting of the roofline; accumulates
 * into counts[8] (device): evaluated, skipped, hard core, core, attractive, tail, p-p, Pauli.
 */
int pyqmd_ensemble_census(const pyqmd_ensemble *e, unsigned long long *counts, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* (C') emitted-particle life cycle (nuclear_sim.py:162,178-210,294-349)                     */

/* One free particle (what the app keeps in self.particles, nuclear_sim.py:349). */
typedef struct {
    double x, y, vx, vy;   /* absolute position (origin included), renormalised velocity */
    double age, lifetime;  /* Particle.age / Particle.lifetime, particles.py:29-38 */
    int64_t nucleus;       /* global id of the nucleus that emitted it */
    int32_t type;          /* PYQMD_PT_* */
    int32_t pad;
} pyqmd_free_particle;

/* Per-frame constants, computed by the caller with the reference's own expressions (they depend on
 * the frame only, not on the particle): see pyqmd_b200/sim.py:frame_constants. */
typedef struct {
    int32_t num_steps;      /* sub-steps of the frame = update_particle calls per particle, :153,162 */
    uint32_t step0;         /* ensemble step index of the frame's first sub-step */
    int32_t fast_forward;   /* time_scale > 1.0, :320 */
    int32_t reserved;
    double speed_scale;     /* 0.3 * (10 / max(1, substeps_used)), :189-190 */
    double aging_scale;     /* min(1, 1 / (sqrt(max(1, ts/100)) * sqrt(max(1, sub/10)))), :199-200 */
    double age_dt;          /* desired_dt / num_steps, :162 */
    double nucleon_dt;      /* effective_physics_dt * time_scale ** 0.5, :207 */
    double lifetime_fast;   /* lifetime of every product when time_scale > 1, :320-338 */
    double lifetime_floor;  /* 5 * max(1, substeps_used / 5), :341 */
} pyqmd_free_frame;

/*
 * One frame of the free-particle list: (1) every particle of pool_in[0 .. *n_in) is advanced by
 * frame->num_steps update_particle calls (nuclear_sim.py:178-210) and, unless it expired, appended to
 * pool_out; (2) every decay event of events[0 .. *event_count) with an emitted particle gets the speed /
 * lifetime rewrite of handle_decay (:295-342), the sub-steps that were left in its frame, and is
 * appended as well; (3) *event_count is reset when reset_event_count != 0.  *n_out is the new length
 * (particles beyond `capacity` are counted in *dropped).  All counts are DEVICE counters: the host
 * never synchronises.  The order of pool_out is unspecified.
 */
int pyqmd_free_particles_frame(const pyqmd_free_particle *pool_in, const unsigned long long *n_in,
                               pyqmd_free_particle *pool_out, unsigned long long *n_out,
                               int64_t capacity, const pyqmd_decay_event *events,
                               unsigned long long *event_count, int64_t event_capacity,
                               const pyqmd_free_frame *frame, unsigned long long *dropped,
                               int32_t reset_event_count, void *stream);

/* ---------------------------------------------------------------------------------------- */
/* (D) decay-only population of particle-less nuclei (decay_chains.py:390-421; config 5)     */

#define PYQMD_COUNT_COLS 16   /* per step: decays by mode [0..7], decays of watch_zn[k] [8..15] */
/* half_life / p_decay hold caller-chosen per-nucleus values and must be read for every nucleus.  When
 * clear (the default), they are read only for nuclides whose half-life is an ESTIMATE drawn per nucleus
 * (PYQMD_HL_BAND); for tabulated nuclides the table row is used instead (4 B instead of 20 B of HBM
 * traffic per nucleus and launch).  Both arrays must always be allocated and initialised: nuclei that
 * decay write their new values back. */
#define PYQMD_POP_PER_NUCLEUS_STATE 1

typedef struct {
    int32_t *zn;
    double *half_life;
    double *p_decay;
    int64_t n;
    int64_t id_base;
    const pyqmd_nuclide_entry *table;
    double dt_decay;
    const double *uniforms;   /* optional [n_steps][uniforms_n][4] */
    int64_t uniforms_n;
    uint64_t seed;
    uint32_t step0;
    int32_t n_watch;
    int32_t watch_zn[8];
    unsigned long long *step_counts;   /* [n_steps][PYQMD_COUNT_COLS], accumulated into */
    uint8_t *decided;         /* optional [n_steps][n]: 1 where should_decay fired */
    int32_t flags;            /* PYQMD_POP_* */
    int32_t reserved;
} pyqmd_population;

int pyqmd_population_step(const pyqmd_population *p, int32_t n_steps, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PYQMD_B200_H */
