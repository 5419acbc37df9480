// cloud_sym.cu -- one large nucleon cloud, every unordered pair evaluated ONCE (sm_100a).
//
// Same law and same Jacobi step as cloud.cu (NuclearForces.update_particles_cpu,
// OtsoBear/PyQMD nuclear_forces.py:236-323); the reference loops over ordered pairs (:248-251),
// but F_ij = -F_ji holds exactly for this law (it depends on |r_i - r_j| and on symmetric type
// predicates only), so the force on j is the negated force on i and half the special-function
// and FMA work disappears.  BASELINE config 4 (N = 1M).
//
// Decomposition
//   * i-blocks ("rows") of kSymIBlock = 1024 nucleons, j-tiles of 256.  Row b owns the tiles t >= 4b: its own
//     4 diagonal tiles (ordered evaluation, no reaction) and every tile after them (each pair once,
//     reaction on j).  A work unit is (row, run of tiles_per_unit tiles); rows are dealt to the
//     `n_parts` GPUs boustrophedon-wise so the triangular work is balanced.
//   * a warp keeps 128 i-nucleons in registers (4 per lane) and walks a 128-nucleon half tile in 32
//     steps: at step m lane L meets the 4 j-nucleons of group (L + m) mod 32 (read from shared
//     memory, conflict-free), i.e. 16 pairs per lane per step, all packed f32x2.  The reaction
//     accumulators of a j group travel with it from lane to lane (8 SHFL per step) and are home
//     after 32 steps: no atomics, no cross-lane reduction trees, fixed summation order.
//   * per tile the 8 warps' reaction rows are summed in a fixed order and added to global
//     FIXED-POINT accumulators (int64, scale 2^k chosen from N so that nothing can overflow) with
//     integer atomics: integer addition is associative, so the result does not depend on the order
//     in which units finish -- bit-reproducible, also across GPU counts after the integer
//     reduce-scatter.  The i side accumulates FP32 per tile -> float64 per unit -> the same
//     accumulators.
//   * far / near classification per (warp, tile) by bounding boxes as in cloud.cu.
#include "cloud.cuh"

namespace pyqmd {

#ifndef PYQMD_SYM_UNROLL
#define PYQMD_SYM_UNROLL 2
#endif
#ifndef PYQMD_SYM_THREADS
#define PYQMD_SYM_THREADS 256
#endif
constexpr int kSymUnroll = PYQMD_SYM_UNROLL;
#ifndef PYQMD_SYM_MINBLOCKS
#define PYQMD_SYM_MINBLOCKS 2
#endif
constexpr float kGhostI = -3.0e18f;   // padding i-nucleons: every term of the law is exactly 0
constexpr float kGhostJ = 3.0e18f;    // padding j-nucleons (distinct from the i ghosts: d2 finite, > 0)
// 256 threads x 2 blocks per SM = 16 warps at 128 registers.  Measured on B200 at N = 1M (round 2): 320
// threads x 2 = 20 warps at 96 registers is 10 % slower although its spills sit outside the sweep loops,
// one block of 256 at 252 registers (8 warps) 14 % slower: the sweep wants both the registers and the warps.
constexpr int kSymThreads = PYQMD_SYM_THREADS;
constexpr int kWarps = kSymThreads / 32;
constexpr int kSymIBlock = kSymThreads * kIPT;
static_assert(kSymThreads >= kTile && kSymIBlock % kTile == 0, "rows are whole tiles; one j per fetching thread");
constexpr int kHalf = 128;            // j-nucleons per 32-step sweep (32 lanes x 4)

struct SymParams {
    int64_t n;
    int nb, nt;                       // rows (i-blocks), j-tiles
    int part, n_parts;
    int tiles_per_unit;
    int far_enabled;
    int skip_exact_zeros;             // PYQMD_CLOUD_SKIP_EXACT_ZEROS
    float scale;                      // 2^k, fixed-point scale of the accumulators
};

// Beyond this bounding-box gap the tail term 0.15 S exp(-1.8 d / 7) / (d + eps) is EXACTLY zero in the
// arithmetic of far_s2: its exponent argument log2(0.15 S) - 0.371 d (+ a correction < 1e-3) is
// below -126, and ex2.approx.ftz flushes 2^x to +0 there.  What is left of such a pair is the Coulomb
// term (p-p only).  d >= 353 covers every S <= 180 (log2(0.15 S) <= 4.76); the host checks the bound and
// ignores the flag otherwise.
constexpr float kUltraGap = 353.0f;

// Two far pairs (one i, two j), action on i and (REACT) reaction on the two j; far_s2 (cloud.cuh) is
// the law.  NOEXP (MODE 1 / 2 only): the tail term of every pair of the tile is exactly zero (kUltraGap).
template <int MODE, bool REACT, bool NOEXP = false>
__device__ __forceinline__ void far_pair2_sym(f32x2 xj, f32x2 yj, f32x2 xi, f32x2 yi, f32x2 cq,
                                              const FarConsts& c, f32x2& fx, f32x2& fy, f32x2& rx,
                                              f32x2& ry)
{
    const f32x2 dx = sub2(xj, xi), dy = sub2(yj, yi);
    const f32x2 s = far_s2<MODE, NOEXP>(fma2(dy, dy, mul2(dx, dx)), cq, c);
    fx = fma2(dx, s, fx);
    fy = fma2(dy, s, fy);
    if (REACT) {
        rx = fma2(dx, s, rx);
        ry = fma2(dy, s, ry);
    }
}

// One 32-step sweep of a warp's 128 i-nucleons over a 128-nucleon half tile.
//   PATH 0/1/2: far field, MODE = PATH;  PATH 3: general law.  NOEXP: see far_pair2_sym.
template <int PATH, bool REACT, bool NOEXP = false>
__device__ __forceinline__ void sweep_half(const float* __restrict__ sx, const float* __restrict__ sy,
                                           const float* __restrict__ st, int lane,
                                           const float (&xi)[kIPT], const float (&yi)[kIPT],
                                           const float (&ti)[kIPT], f32x2 (&ax)[kIPT],
                                           f32x2 (&ay)[kIPT], const LawParams& L, float2* row)
{
    const ulonglong2* x4 = reinterpret_cast<const ulonglong2*>(sx);
    const ulonglong2* y4 = reinterpret_cast<const ulonglong2*>(sy);
    const float4* t4 = reinterpret_cast<const float4*>(st);
    f32x2 xi2[kIPT], yi2[kIPT];
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        xi2[k] = pk1(xi[k]);
        yi2[k] = pk1(yi[k]);
    }
    f32x2 rx01 = 0ull, ry01 = 0ull, rx23 = 0ull, ry23 = 0ull;
    const int nxt = (lane + 1) & 31;
    if (PATH < 3) {
        const FarConsts c = make_far_consts(L);
        f32x2 nq2[kIPT];
#pragma unroll
        for (int k = 0; k < kIPT; ++k) nq2[k] = pk1(-L.C * ti[k]);
#pragma unroll kSymUnroll
        for (int m = 0; m < 32; ++m) {
            const int q = (lane + m) & 31;
            const ulonglong2 X = x4[q], Y = y4[q];
            f32x2 T01 = 0ull, T23 = 0ull;
            if (PATH == 2) {
                const float4 T = t4[q];
                T01 = pk(T.x, T.y);
                T23 = pk(T.z, T.w);
            }
#pragma unroll
            for (int k = 0; k < kIPT; ++k) {
                far_pair2_sym<PATH, REACT, NOEXP>(X.x, Y.x, xi2[k], yi2[k],
                                                  PATH == 2 ? mul2(nq2[k], T01) : 0ull, c, ax[k], ay[k],
                                                  rx01, ry01);
                far_pair2_sym<PATH, REACT, NOEXP>(X.y, Y.y, xi2[k], yi2[k],
                                                  PATH == 2 ? mul2(nq2[k], T23) : 0ull, c, ax[k], ay[k],
                                                  rx23, ry23);
            }
            if (REACT) {                   // the accumulators follow their j group to lane - 1
                rx01 = shfl64(rx01, nxt); ry01 = shfl64(ry01, nxt);
                rx23 = shfl64(rx23, nxt); ry23 = shfl64(ry23, nxt);
            }
        }
    } else {
        const GenConsts gc = make_gen_consts(L);
        const f32x2 negC = pk1(-L.C);
#pragma unroll 1
        for (int m = 0; m < 32; ++m) {
            const int q = (lane + m) & 31;
            const ulonglong2 X = x4[q], Y = y4[q];
            const float4 T = t4[q];
            const f32x2 nq01 = mul2(negC, pk(T.x, T.y)), nq23 = mul2(negC, pk(T.z, T.w));
#pragma unroll
            for (int k = 0; k < kIPT; ++k) {
                const f32x2 ti2 = pk1(ti[k]);
                {
                    const f32x2 dx = sub2(X.x, xi2[k]), dy = sub2(Y.x, yi2[k]);
                    const f32x2 s = pair_general2(dx, dy, T.x, T.y, ti[k], ti2, nq01, gc, L);
                    ax[k] = fma2(dx, s, ax[k]);
                    ay[k] = fma2(dy, s, ay[k]);
                    if (REACT) { rx01 = fma2(dx, s, rx01); ry01 = fma2(dy, s, ry01); }
                }
                {
                    const f32x2 dx = sub2(X.y, xi2[k]), dy = sub2(Y.y, yi2[k]);
                    const f32x2 s = pair_general2(dx, dy, T.z, T.w, ti[k], ti2, nq23, gc, L);
                    ax[k] = fma2(dx, s, ax[k]);
                    ay[k] = fma2(dy, s, ay[k]);
                    if (REACT) { rx23 = fma2(dx, s, rx23); ry23 = fma2(dy, s, ry23); }
                }
            }
            if (REACT) {
                rx01 = shfl64(rx01, nxt); ry01 = shfl64(ry01, nxt);
                rx23 = shfl64(rx23, nxt); ry23 = shfl64(ry23, nxt);
            }
        }
    }
    if (REACT) {                           // after 32 passes the accumulators of group `lane` are home
        float a, b, c2, d;
        upk(rx01, a, b);
        upk(ry01, c2, d);
        reinterpret_cast<float4*>(row)[2 * lane] = make_float4(a, c2, b, d);          // j = 4 lane, 4 lane + 1
        upk(rx23, a, b);
        upk(ry23, c2, d);
        reinterpret_cast<float4*>(row)[2 * lane + 1] = make_float4(a, c2, b, d);      // j = 4 lane + 2, + 3
    }
}

// path 0..3 as sweep_half; 4: every pair of the tile is exactly zero (tail underflow, no p-p pair);
// 5 / 6: MODE 1 / 2 without the exponential.
template <bool REACT>
__device__ __forceinline__ void sweep_tile(int path, const float* sx, const float* sy, const float* st,
                                           int lane, const float (&xi)[kIPT], const float (&yi)[kIPT],
                                           const float (&ti)[kIPT], float (&fx)[kIPT], float (&fy)[kIPT],
                                           const LawParams& L, float2* row)
{
    if (path == 4) {
#pragma unroll
        for (int k = 0; k < kIPT; ++k) { fx[k] = 0.f; fy[k] = 0.f; }
        if (REACT)                                           // the flush sums all eight warp rows
            for (int k = lane; k < kTile / 2; k += 32)
                reinterpret_cast<float4*>(row)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    f32x2 ax[kIPT], ay[kIPT];
#pragma unroll
    for (int k = 0; k < kIPT; ++k) { ax[k] = 0ull; ay[k] = 0ull; }
#pragma unroll 1
    for (int h = 0; h < kTile / kHalf; ++h) {
        const float* hx = sx + h * kHalf;
        const float* hy = sy + h * kHalf;
        const float* ht = st + h * kHalf;
        float2* hr = row + h * kHalf;
        if (path == 0) sweep_half<0, REACT>(hx, hy, ht, lane, xi, yi, ti, ax, ay, L, hr);
        else if (path == 1) sweep_half<1, REACT>(hx, hy, ht, lane, xi, yi, ti, ax, ay, L, hr);
        else if (path == 2) sweep_half<2, REACT>(hx, hy, ht, lane, xi, yi, ti, ax, ay, L, hr);
        else if (path == 5) sweep_half<1, REACT, true>(hx, hy, ht, lane, xi, yi, ti, ax, ay, L, hr);
        else if (path == 6) sweep_half<2, REACT, true>(hx, hy, ht, lane, xi, yi, ti, ax, ay, L, hr);
        else sweep_half<3, REACT>(hx, hy, ht, lane, xi, yi, ti, ax, ay, L, hr);
    }
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        float a, b;
        upk(ax[k], a, b); fx[k] = a + b;
        upk(ay[k], a, b); fy[k] = a + b;
    }
}

__device__ __forceinline__ void acc_add(long long* acc, int64_t i, double fx, double fy, float scale)
{
    atomicAdd(reinterpret_cast<unsigned long long*>(acc + 2 * i),
              (unsigned long long)__double2ll_rn(fx * (double)scale));
    atomicAdd(reinterpret_cast<unsigned long long*>(acc + 2 * i + 1),
              (unsigned long long)__double2ll_rn(fy * (double)scale));
}

__global__ void __launch_bounds__(kSymThreads, PYQMD_SYM_MINBLOCKS)
cloud_sym_kernel(const float2* __restrict__ pos, const uint8_t* __restrict__ isp, SymParams sp,
                 CloudWorkspace w, LawParams L, long long* __restrict__ acc)
{
    __shared__ __align__(16) float sxb[2][kTile];
    __shared__ __align__(16) float syb[2][kTile];
    __shared__ __align__(16) float stb[2][kTile];
    __shared__ __align__(16) float2 react[2][kWarps][kTile];

    // this block's row (rows dealt boustrophedon-wise to the parts) and run of tiles
    const int g = blockIdx.y;
    const int b = g * sp.n_parts + ((g & 1) ? sp.n_parts - 1 - sp.part : sp.part);
    if (b >= sp.nb) return;
    const int first = b * (kSymIBlock / kTile);
    const int t_begin = first + (int)blockIdx.x * sp.tiles_per_unit;
    if (t_begin >= sp.nt) return;
    const int t_end = min(t_begin + sp.tiles_per_unit, sp.nt);
    const int64_t n = sp.n;

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t wbase = (int64_t)b * kSymIBlock + (int64_t)wid * (32 * kIPT);
    float xi[kIPT], yi[kIPT], ti[kIPT], fx[kIPT], fy[kIPT];
    double Fx[kIPT], Fy[kIPT];
    float bxmin = INFINITY, bymin = INFINITY, bxmax = -INFINITY, bymax = -INFINITY;
    bool allp = true, alln = true;
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        xi[k] = kGhostI; yi[k] = kGhostI; ti[k] = 0.f;
        Fx[k] = 0.0; Fy[k] = 0.0;
        if (i < n) {
            const float2 p = pos[i];
            const bool pr = isp[i] != 0;
            xi[k] = p.x; yi[k] = p.y;
            ti[k] = pr ? 1.0f : 0.0f;
            bxmin = fminf(bxmin, p.x); bxmax = fmaxf(bxmax, p.x);
            bymin = fminf(bymin, p.y); bymax = fmaxf(bymax, p.y);
            allp = allp && pr;
            alln = alln && !pr;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bxmin = fminf(bxmin, __shfl_xor_sync(0xffffffffu, bxmin, o));
        bymin = fminf(bymin, __shfl_xor_sync(0xffffffffu, bymin, o));
        bxmax = fmaxf(bxmax, __shfl_xor_sync(0xffffffffu, bxmax, o));
        bymax = fmaxf(bymax, __shfl_xor_sync(0xffffffffu, bymax, o));
    }
    allp = __all_sync(0xffffffffu, allp);
    alln = __all_sync(0xffffffffu, alln);

    // software pipeline of the j tiles (register prefetch -> the other shared-memory buffer)
    const bool loader = threadIdx.x < kTile;              // one j of the tile per loading thread
    float2 nxt = make_float2(kGhostJ, kGhostJ);
    float nxt_t = 0.f;
    auto fetch = [&](int tile) {
        nxt = make_float2(kGhostJ, kGhostJ);
        nxt_t = 0.f;
        const int64_t j = (int64_t)tile * kTile + threadIdx.x;
        if (loader && tile < t_end && j < n) { nxt = pos[j]; nxt_t = isp[j] ? 1.0f : 0.0f; }
    };
    fetch(t_begin);
    if (loader) {
        sxb[0][threadIdx.x] = nxt.x;
        syb[0][threadIdx.x] = nxt.y;
        stb[0][threadIdx.x] = nxt_t;
    }
    fetch(t_begin + 1);

    auto flush = [&](int rb, int tile) {                  // reaction of `tile`: the warps' rows -> accumulators
        if (!loader) return;
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int k = 0; k < kWarps; ++k) {                // fixed order
            const float2 v = react[rb][k][threadIdx.x];
            sx += v.x;
            sy += v.y;
        }
        const int64_t j = (int64_t)tile * kTile + threadIdx.x;
        if (j < n) acc_add(acc, j, -(double)sx, -(double)sy, sp.scale);
    };

    int buf = 0;
    bool prev_react = false;
    for (int tile = t_begin; tile < t_end; ++tile, buf ^= 1) {
        __syncthreads();          // tile data complete in `buf`; everybody is done with tile - 1
        if (prev_react) flush(buf ^ 1, tile - 1);
        if (tile + 1 < t_end) {
            if (loader) {
                sxb[buf ^ 1][threadIdx.x] = nxt.x;
                syb[buf ^ 1][threadIdx.x] = nxt.y;
                stb[buf ^ 1][threadIdx.x] = nxt_t;
            }
            fetch(tile + 2);
        }
        const bool diag = tile < first + kSymIBlock / kTile;
        const float4 bb = w.bbox[tile];
        const int tf = w.flags[tile];
        const float gx = fmaxf(0.f, fmaxf(bb.x - bxmax, bxmin - bb.z));
        const float gy = fmaxf(0.f, fmaxf(bb.y - bymax, bymin - bb.w));
        const float gap2 = fmaf(gx, gx, gy * gy);
        const bool far = sp.far_enabled && (gap2 > 81.01f);
        int path = !far ? 3
                   : ((alln || (tf & kTileAllNeutron)) ? 0
                      : ((allp && (tf & kTileAllProton)) ? 1 : 2));
        if (far && sp.skip_exact_zeros && gap2 > kUltraGap * kUltraGap) path = (path == 0) ? 4 : path + 4;
        // the whole block agrees that every pair of this tile is exactly zero: no sweep, no flush
        if (sp.skip_exact_zeros && __syncthreads_and(path == 4)) {
            prev_react = false;
            continue;
        }
        if (diag)
            sweep_tile<false>(path, sxb[buf], syb[buf], stb[buf], lane, xi, yi, ti, fx, fy, L, nullptr);
        else
            sweep_tile<true>(path, sxb[buf], syb[buf], stb[buf], lane, xi, yi, ti, fx, fy, L,
                             react[buf][wid]);
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            Fx[k] += (double)fx[k];
            Fy[k] += (double)fy[k];
        }
        prev_react = !diag;
    }
    __syncthreads();
    if (prev_react) flush(buf ^ 1, t_end - 1);
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        if (i < n) acc_add(acc, i, Fx[k], Fy[k], sp.scale);
    }
}

// Consumes (and clears) the accumulators of [i0, i1): containment + damped Euler, :301-323.
__global__ void __launch_bounds__(256)
cloud_sym_integrate_kernel(const float2* __restrict__ pos_in, float2* __restrict__ pos_out,
                           float2* __restrict__ vel, float2* __restrict__ force, int64_t n, int64_t i0,
                           int64_t i1, CloudWorkspace w, long long* __restrict__ acc_i0, float inv_scale,
                           float dt)
{
    const int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    long long* a = acc_i0 + 2 * (i - i0);
    const longlong2 v = *reinterpret_cast<const longlong2*>(a);
    *reinterpret_cast<longlong2*>(a) = make_longlong2(0, 0);
    float Fx = (float)((double)v.x * (double)inv_scale);
    float Fy = (float)((double)v.y * (double)inv_scale);
    const float cx = (float)w.centre[0], cy = (float)w.centre[1];
    const float R = 2.4f * cbrtf((float)n);            // :304
    const float2 p = pos_in[i];
    float2 vv = vel[i];
    float x = p.x, y = p.y;
    contain_and_integrate(x, y, vv.x, vv.y, Fx, Fy, cx, cy, R, dt);
    pos_out[i] = make_float2(x, y);
    vel[i] = vv;
    if (force) force[i] = make_float2(Fx, Fy);
}

// Fused exchange + integrate for several GPUs (peer memory over NVLink, no NCCL on the data path):
// the owner of [i0, i1) PULLS the int64 force accumulators of its nucleons from every rank (16-byte
// peer loads), sums them (integer: exact, order-independent), clears the consumed entries on the
// peers, integrates, and PUSHES the new positions into every rank's replica (8-byte peer stores) --
// i.e. reduce-scatter(forces) + integrate + all-gather(positions) in one kernel.  The caller
// brackets it with two device-side barriers (all pair-force kernels done / all pushes landed).
__global__ void __launch_bounds__(256)
cloud_sym_exchange_integrate_kernel(const float2* __restrict__ pos_in, float2* __restrict__ vel,
                                    float2* __restrict__ force, int64_t n, int64_t i0, int64_t i1,
                                    CloudWorkspace w, long long* const* __restrict__ acc_peers,
                                    float2* const* __restrict__ pos_out_peers, int n_peers,
                                    float inv_scale, float dt)
{
    const int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    long long sx = 0, sy = 0;
    for (int p = 0; p < n_peers; ++p) {
        longlong2* a = reinterpret_cast<longlong2*>(acc_peers[p]) + i;
        const longlong2 v = *a;
        *a = make_longlong2(0, 0);
        sx += v.x;
        sy += v.y;
    }
    float Fx = (float)((double)sx * (double)inv_scale);
    float Fy = (float)((double)sy * (double)inv_scale);
    const float cx = (float)w.centre[0], cy = (float)w.centre[1];
    const float R = 2.4f * cbrtf((float)n);            // :304
    const float2 p0 = pos_in[i];
    float2 vv = vel[i];
    float x = p0.x, y = p0.y;
    contain_and_integrate(x, y, vv.x, vv.y, Fx, Fy, cx, cy, R, dt);
    vel[i] = vv;
    if (force) force[i] = make_float2(Fx, Fy);
    const float2 out = make_float2(x, y);
    for (int p = 0; p < n_peers; ++p) pos_out_peers[p][i] = out;
}

int cloud_prepass(const float* pos, const uint8_t* is_proton, int64_t n, const CloudWorkspace& w,
                  cudaStream_t st);

static int scale_log2_for(int64_t n)
{
    // |F| <= 12 (n - 1) [nuclear_forces.py:294]; keep 2 bits of head-room below 2^63
    int bits = 4;                                        // 12 < 2^4
    while (((int64_t)1 << (bits - 4)) < n) ++bits;       // 12 n < 2^bits
    int k = 61 - bits;
    if (k > 44) k = 44;
    return k;
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int32_t pyqmd_cloud_force_scale_log2(int64_t n) { return scale_log2_for(n < 1 ? 1 : n); }

extern "C" int pyqmd_cloud_pair_forces(const float* pos, const uint8_t* is_proton, int64_t n,
                                       int32_t part, int32_t n_parts, float strong, float coulomb,
                                       float pauli, long long* force_acc, void* workspace, void* stream)
{
    return pyqmd_cloud_pair_forces_ex(pos, is_proton, n, part, n_parts, strong, coulomb, pauli, force_acc,
                                      workspace, 0u, stream);
}

extern "C" int pyqmd_cloud_pair_forces_ex(const float* pos, const uint8_t* is_proton, int64_t n,
                                          int32_t part, int32_t n_parts, float strong, float coulomb,
                                          float pauli, long long* force_acc, void* workspace,
                                          uint32_t flags, void* stream)
{
    PYQMD_REQUIRE(n >= 0 && n_parts >= 1 && part >= 0 && part < n_parts, "0 <= part < n_parts");
    if (n == 0) return PYQMD_OK;
    PYQMD_REQUIRE(pos && is_proton && force_acc && workspace, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const CloudWorkspace w = carve(workspace, n);
    const int64_t nt = n_tiles_of(n);
    PYQMD_REQUIRE(nt <= 2147483647LL / 4, "cloud too large");
    const int rc = cloud_prepass(pos, is_proton, n, w, st);
    if (rc != PYQMD_OK) return rc;
    const LawParams L = make_law_params(strong, coulomb, pauli);
    SymParams sp;
    sp.n = n;
    sp.nb = (int)((n + kSymIBlock - 1) / kSymIBlock);
    sp.nt = (int)nt;
    sp.part = part;
    sp.n_parts = n_parts;
    sp.far_enabled = (strong > 0.f && !L.far_needs_clamp) ? 1 : 0;
    // exact-zero skipping needs the tail exponent below the flush-to-zero point at the ultra-far gap
    const float arg_at_gap = L.log2TailK + kUltraGap * (-1.8f * kLog2e / 7.0f);
    sp.skip_exact_zeros = ((flags & PYQMD_CLOUD_SKIP_EXACT_ZEROS) && sp.far_enabled && arg_at_gap < -126.2f) ? 1 : 0;
    sp.scale = ldexpf(1.0f, scale_log2_for(n));
    // ~96 units per resident-block slot of this part (148 SMs x 2), 4..64 tiles each: the hardware
    // hands blocks to SMs as slots free up, so the idle tail of a launch is about half a unit -- 0.5 %
    // of the step at ~100 units per slot whatever the number of parts (round 1 used 24 per slot: 2 % at
    // 8 GPUs).  A sweep of 8..128 tiles per unit on B200 at N = 1M changed the step time by < 0.8 %.
    const double tiles_total = 0.5 * (double)sp.nb * (double)nt / n_parts;
    int tpu = (int)(tiles_total / (296.0 * 96.0));
    if (tpu > 64) tpu = 64;
    if (tpu < 4) tpu = 4;
    sp.tiles_per_unit = tpu;
    const int rows_mine = (sp.nb + n_parts - 1) / n_parts;
    const int units_max = (int)((nt + tpu - 1) / tpu);
    PYQMD_REQUIRE(rows_mine <= 65535, "too many rows for one launch");
    const dim3 grid((unsigned)units_max, (unsigned)rows_mine);
    cloud_sym_kernel<<<grid, kSymThreads, 0, st>>>(reinterpret_cast<const float2*>(pos), is_proton, sp, w,
                                                L, force_acc);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}

extern "C" int pyqmd_cloud_integrate(const float* pos_in, float* pos_out, float* vel, float* force,
                                     int64_t n, int64_t i0, int64_t i1, float dt,
                                     long long* force_acc_i0, void* workspace, void* stream)
{
    PYQMD_REQUIRE(n >= 0 && i0 >= 0 && i0 <= i1 && i1 <= n, "0 <= i0 <= i1 <= n");
    if (n == 0 || i0 == i1) return PYQMD_OK;
    PYQMD_REQUIRE(pos_in && pos_out && vel && force_acc_i0 && workspace, "NULL pointer");
    PYQMD_REQUIRE(pos_in != pos_out, "pos_in and pos_out must differ (Jacobi update)");
    const CloudWorkspace w = carve(workspace, n);
    const float inv_scale = ldexpf(1.0f, -scale_log2_for(n));
    cloud_sym_integrate_kernel<<<(unsigned)((i1 - i0 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(pos_in), reinterpret_cast<float2*>(pos_out),
        reinterpret_cast<float2*>(vel), reinterpret_cast<float2*>(force), n, i0, i1, w, force_acc_i0,
        inv_scale, dt);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}

extern "C" int pyqmd_cloud_exchange_integrate(const float* pos_in, float* vel, float* force, int64_t n,
                                              int64_t i0, int64_t i1, float dt,
                                              long long* const* acc_peers, float* const* pos_out_peers,
                                              int32_t n_peers, void* workspace, void* stream)
{
    PYQMD_REQUIRE(n >= 0 && i0 >= 0 && i0 <= i1 && i1 <= n, "0 <= i0 <= i1 <= n");
    PYQMD_REQUIRE(n_peers >= 1, "n_peers >= 1");
    if (n == 0 || i0 == i1) return PYQMD_OK;
    PYQMD_REQUIRE(pos_in && vel && acc_peers && pos_out_peers && workspace, "NULL pointer");
    const CloudWorkspace w = carve(workspace, n);
    const float inv_scale = ldexpf(1.0f, -scale_log2_for(n));
    cloud_sym_exchange_integrate_kernel<<<(unsigned)((i1 - i0 + 255) / 256), 256, 0,
                                          (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(pos_in), reinterpret_cast<float2*>(vel),
        reinterpret_cast<float2*>(force), n, i0, i1, w, acc_peers,
        reinterpret_cast<float2* const*>(pos_out_peers), n_peers, inv_scale, dt);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
