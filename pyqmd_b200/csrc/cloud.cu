// cloud.cu -- one large nucleon cloud: j-tiled all-pairs force + containment + damped Euler
// for an i-block of the cloud (sm_100a).  BASELINE config 4 (N = 1M, i-block sharded).
//
// Replaces NuclearForces.update_particles_cpu / the OpenCL kernel of the reference
// (OtsoBear/PyQMD nuclear_forces.py:236-323 / :57-173) for a single big system.
//
// Design (B200-first, not a translation of the untiled OpenCL loop):
//  * positions float2 SoA; 256-nucleon j tiles staged in shared memory, read back as
//    broadcast LDS.128 (two j per load); each thread keeps kIPT i-nucleons in registers, so
//    one smem word feeds kIPT pair evaluations;
//  * a tile-statistics pre-pass gives every tile a bounding box and a type flag, and the
//    fixed-order float64 centre of mass; a warp whose i bounding box is >= 9 away from a
//    tile's box takes the branch-free far-field path (2 MUFU + 14 FMA-pipe ops per pair),
//    everything else the general path -- the classification is warp-uniform, so there is
//    no divergence.  With the cloud kept type-partitioned and Morton-sorted by the caller
//    (pyqmd_cloud_sort_keys) > 99.9 % of tile visits are far;
//  * Jacobi update: reads pos_in, writes pos_out (the reference's OpenCL kernel updates in
//    place and races; the CPU path, which is the parity target, is Jacobi).
#include "cloud.cuh"

namespace pyqmd {

// ---- pass 1: per-tile bounding box, type flag, coordinate sums ---------------------------------
__global__ void __launch_bounds__(kTile) cloud_tile_stats(const float2* __restrict__ pos,
                                                          const uint8_t* __restrict__ isp,
                                                          int64_t n, CloudWorkspace w)
{
    const int64_t j = (int64_t)blockIdx.x * kTile + threadIdx.x;
    const bool ok = j < n;
    float2 p = ok ? pos[j] : make_float2(0.f, 0.f);
    float xmin = ok ? p.x : INFINITY, xmax = ok ? p.x : -INFINITY;
    float ymin = ok ? p.y : INFINITY, ymax = ok ? p.y : -INFINITY;
    double sx = ok ? (double)p.x : 0.0, sy = ok ? (double)p.y : 0.0;
    int np = (ok && isp[j]) ? 1 : 0, nn = (ok && !isp[j]) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        np += __shfl_xor_sync(0xffffffffu, np, o);
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    __shared__ float4 sb[kTile / 32];
    __shared__ double2 ss[kTile / 32];
    __shared__ int2 sc[kTile / 32];
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        sb[wid] = make_float4(xmin, ymin, xmax, ymax);
        ss[wid] = make_double2(sx, sy);
        sc[wid] = make_int2(np, nn);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float4 b = sb[0];
        double2 s = ss[0];
        int2 c = sc[0];
        for (int k = 1; k < kTile / 32; ++k) {       // fixed order: deterministic
            b.x = fminf(b.x, sb[k].x); b.y = fminf(b.y, sb[k].y);
            b.z = fmaxf(b.z, sb[k].z); b.w = fmaxf(b.w, sb[k].w);
            s.x += ss[k].x; s.y += ss[k].y;
            c.x += sc[k].x; c.y += sc[k].y;
        }
        w.bbox[blockIdx.x] = b;
        w.sums[blockIdx.x] = s;
        w.flags[blockIdx.x] = (c.y == 0 ? kTileAllProton : 0) | (c.x == 0 ? kTileAllNeutron : 0);
    }
}

// ---- pass 2: centre of mass (nuclear_forces.py:242-243), fixed summation order -----------------
__global__ void __launch_bounds__(256) cloud_centre(CloudWorkspace w, int64_t n_tiles, int64_t n)
{
    __shared__ double2 s[256];
    double sx = 0.0, sy = 0.0;
    for (int64_t k = threadIdx.x; k < n_tiles; k += 256) {
        sx += w.sums[k].x;
        sy += w.sums[k].y;
    }
    s[threadIdx.x] = make_double2(sx, sy);
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s[threadIdx.x].x += s[threadIdx.x + o].x;
            s[threadIdx.x].y += s[threadIdx.x + o].y;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        w.centre[0] = s[0].x / (double)n;
        w.centre[1] = s[0].y / (double)n;
    }
}

// ---- pass 3: partial forces of one work unit = (i-block, j-segment) --------------------------------
//
// Packed FP32 (Blackwell add/mul/fma.f32x2 -> FADD2/FMUL2/FFMA2): one instruction works on two
// pairs, so the FMA-pipe operations of a far pair cost half the issue slots and the kernel moves
// from issue-bound (ncu r01a: 82 % issue, 18 instr/pair) to pipe-bound.  ncu r01e (14 FMA-pipe ops
// + 2 special-function ops per pair): XU pipe 84 %, FMA pipe ~80 % busy -- both near saturation,
// so the FMA work was trimmed: now 11 ops per ordered pair without a Coulomb term (dx, dy, d2: 4;
// far_s2: 5; accumulation: 2).  (Moving a share of the 2^x evaluations from the XU pipe to an FMA-pipe
// polynomial was measured and rejected: every share was slower.)
template <int MODE>
__device__ __forceinline__ void far_tile_packed(const float* __restrict__ sx,
                                                const float* __restrict__ sy,
                                                const float* __restrict__ st, int jmax,
                                                const float (&xi)[kIPT], const float (&yi)[kIPT],
                                                const float (&qi)[kIPT], float (&fx)[kIPT],
                                                float (&fy)[kIPT], const LawParams& L)
{
    const FarConsts c = make_far_consts(L);
    f32x2 xi2[kIPT], yi2[kIPT], nq2[kIPT], ax[kIPT], ay[kIPT];
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        xi2[k] = pk(xi[k], xi[k]);
        yi2[k] = pk(yi[k], yi[k]);
        nq2[k] = pk(-qi[k], -qi[k]);
        ax[k] = pk(0.f, 0.f);
        ay[k] = pk(0.f, 0.f);
    }
    const int quads = jmax >> 2;
    const ulonglong2* x4 = reinterpret_cast<const ulonglong2*>(sx);
    const ulonglong2* y4 = reinterpret_cast<const ulonglong2*>(sy);
    const ulonglong2* t4 = reinterpret_cast<const ulonglong2*>(st);
#pragma unroll 1
    for (int jq = 0; jq < quads; ++jq) {
        const ulonglong2 X = x4[jq], Y = y4[jq];            // 4 j per LDS.128
        ulonglong2 T = make_ulonglong2(0ull, 0ull);
        if (MODE == 2) T = t4[jq];
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            far_pair2<MODE>(X.x, Y.x, xi2[k], yi2[k], MODE == 2 ? mul2(nq2[k], T.x) : 0ull, c, ax[k],
                            ay[k]);
            far_pair2<MODE>(X.y, Y.y, xi2[k], yi2[k], MODE == 2 ? mul2(nq2[k], T.y) : 0ull, c, ax[k],
                            ay[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        float a, b;
        upk(ax[k], a, b); fx[k] += a + b;
        upk(ay[k], a, b); fy[k] += a + b;
    }
    for (int j = quads << 2; j < jmax; ++j) {               // ragged tail of the last tile
        const float ox = sx[j], oy = sy[j], tj = st[j];
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            const float dx = ox - xi[k], dy = oy - yi[k];
            const float sc = pair_far_impl<MODE, false>(dx, dy, qi[k] * tj, L);
            fx[k] = fmaf(dx, sc, fx[k]);
            fy[k] = fmaf(dy, sc, fy[k]);
        }
    }
}

// Scalar far path with the +-12 clamp, only used when the strengths make |net| >= 12 possible
// beyond d = 9 (never with the reference's defaults).
template <int MODE>
__device__ __forceinline__ void far_tile_clamped(const float* __restrict__ sx,
                                                 const float* __restrict__ sy,
                                                 const float* __restrict__ st, int jmax,
                                                 const float (&xi)[kIPT], const float (&yi)[kIPT],
                                                 const float (&qi)[kIPT], float (&fx)[kIPT],
                                                 float (&fy)[kIPT], const LawParams& L)
{
    for (int j = 0; j < jmax; ++j) {
        const float ox = sx[j], oy = sy[j], tj = st[j];
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            const float dx = ox - xi[k], dy = oy - yi[k];
            const float sc = pair_far_impl<MODE, true>(dx, dy, qi[k] * tj, L);
            fx[k] = fmaf(dx, sc, fx[k]);
            fy[k] = fmaf(dy, sc, fy[k]);
        }
    }
}

__device__ __forceinline__ void near_tile(const float* __restrict__ sx, const float* __restrict__ sy,
                                          const float* __restrict__ st, int jmax,
                                          const float (&xi)[kIPT], const float (&yi)[kIPT],
                                          const float (&ti)[kIPT], float (&fx)[kIPT],
                                          float (&fy)[kIPT], const LawParams& L)
{
    for (int j = 0; j < jmax; ++j) {
        const float ox = sx[j], oy = sy[j], tj = st[j];
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            const float dx = ox - xi[k], dy = oy - yi[k];
            const float s = pair_general(dx, dy, ti[k], tj, L);
            fx[k] = fmaf(dx, s, fx[k]);
            fy[k] = fmaf(dy, s, fy[k]);
        }
    }
}

// grid = (n_iblocks, n_seg): block (b, s) accumulates the force of j tiles
// [s * tiles_per_seg, (s+1) * tiles_per_seg) on the 1024 nucleons of i-block b and writes
// float64 partial sums to partial[s][i - i0].  Splitting the j range keeps >= ~10 work units
// per resident-block slot whatever N and the number of ranks are (wave quantisation was
// costing 18 % at N = 1M on one GPU and > 50 % on eight).
template <bool CLAMP>
__global__ void __launch_bounds__(kThreads, 2)
cloud_force_kernel(const float2* __restrict__ pos_in, const uint8_t* __restrict__ isp, int64_t n,
                   int64_t i0, int64_t i1, CloudWorkspace w, LawParams L, int far_enabled,
                   int tiles_per_seg)
{
    // double-buffered j tiles: one barrier per tile (the store of tile t+1 goes to the other
    // buffer while tile t is being consumed)
    __shared__ __align__(16) float sxb[2][kTile];
    __shared__ __align__(16) float syb[2][kTile];
    __shared__ __align__(16) float stb[2][kTile];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t wbase = i0 + (int64_t)blockIdx.x * kIBlock + (int64_t)wid * (32 * kIPT);

    float xi[kIPT], yi[kIPT], ti[kIPT], qi[kIPT], fx[kIPT], fy[kIPT];
    // Tile-local FP32 partial sums are folded into float64 totals once per tile: with 10^5..10^6
    // partners a single FP32 running sum loses ~sqrt(N) ulp and would break the 1e-5 budget.
    double Fx[kIPT], Fy[kIPT];
    float bxmin = INFINITY, bymin = INFINITY, bxmax = -INFINITY, bymax = -INFINITY;
    bool allp = true, alln = true;
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        int64_t i = wbase + k * 32 + lane;
        if (i > i1 - 1) i = i1 - 1;                    // clamp: duplicates do not move the bbox
        if (i < i0) i = i0;
        const float2 p = pos_in[i];
        const bool pr = isp[i] != 0;
        xi[k] = p.x; yi[k] = p.y;
        ti[k] = pr ? 1.0f : 0.0f;
        qi[k] = pr ? L.C : 0.0f;
        fx[k] = 0.f; fy[k] = 0.f;
        Fx[k] = 0.0; Fy[k] = 0.0;
        bxmin = fminf(bxmin, p.x); bxmax = fmaxf(bxmax, p.x);
        bymin = fminf(bymin, p.y); bymax = fmaxf(bymax, p.y);
        allp = allp && pr;
        alln = alln && !pr;
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

    const int64_t n_tiles = n_tiles_of(n);
    const int64_t t_begin = (int64_t)blockIdx.y * tiles_per_seg;
    int64_t t_end = t_begin + tiles_per_seg;
    if (t_end > n_tiles) t_end = n_tiles;
    // software pipeline: the next tile's element is fetched into registers while the current
    // tile is being consumed
    float2 nxt = make_float2(0.f, 0.f);
    float nxt_t = 0.f;
    {
        const int64_t j = t_begin * kTile + threadIdx.x;
        if (j < n) { nxt = pos_in[j]; nxt_t = isp[j] ? 1.0f : 0.0f; }
    }
    if (t_begin < t_end) {
        sxb[0][threadIdx.x] = nxt.x;
        syb[0][threadIdx.x] = nxt.y;
        stb[0][threadIdx.x] = nxt_t;
        const int64_t j = (t_begin + 1) * kTile + threadIdx.x;
        if (j < n && t_begin + 1 < t_end) { nxt = pos_in[j]; nxt_t = isp[j] ? 1.0f : 0.0f; }
    }
    int buf = 0;
    for (int64_t tile = t_begin; tile < t_end; ++tile, buf ^= 1) {
        __syncthreads();              // tile `tile` is complete in buffer `buf`; buffer buf^1 is free
        const float* sx = sxb[buf];
        const float* sy = syb[buf];
        const float* st = stb[buf];
        if (tile + 1 < t_end) {
            sxb[buf ^ 1][threadIdx.x] = nxt.x;
            syb[buf ^ 1][threadIdx.x] = nxt.y;
            stb[buf ^ 1][threadIdx.x] = nxt_t;
            const int64_t j = (tile + 2) * kTile + threadIdx.x;
            if (j < n && tile + 2 < t_end) { nxt = pos_in[j]; nxt_t = isp[j] ? 1.0f : 0.0f; }
        }
        const int64_t rem = n - tile * kTile;
        const int jmax = rem < kTile ? (int)rem : kTile;

        // warp-uniform classification against the tile's bounding box
        const float4 bb = w.bbox[tile];
        const int tf = w.flags[tile];
        const float gx = fmaxf(0.f, fmaxf(bb.x - bxmax, bxmin - bb.z));
        const float gy = fmaxf(0.f, fmaxf(bb.y - bymax, bymin - bb.w));
        const bool far = far_enabled && (fmaf(gx, gx, gy * gy) > 81.01f);
        if (far) {
            const int mode = (alln || (tf & kTileAllNeutron)) ? 0
                             : ((allp && (tf & kTileAllProton)) ? 1 : 2);
            if (CLAMP) {
                if (mode == 0) far_tile_clamped<0>(sx, sy, st, jmax, xi, yi, qi, fx, fy, L);
                else if (mode == 1) far_tile_clamped<1>(sx, sy, st, jmax, xi, yi, qi, fx, fy, L);
                else far_tile_clamped<2>(sx, sy, st, jmax, xi, yi, qi, fx, fy, L);
            } else {
                if (mode == 0) far_tile_packed<0>(sx, sy, st, jmax, xi, yi, qi, fx, fy, L);
                else if (mode == 1) far_tile_packed<1>(sx, sy, st, jmax, xi, yi, qi, fx, fy, L);
                else far_tile_packed<2>(sx, sy, st, jmax, xi, yi, qi, fx, fy, L);
            }
        } else {
            near_tile(sx, sy, st, jmax, xi, yi, ti, fx, fy, L);
        }
#pragma unroll
        for (int k = 0; k < kIPT; ++k) {
            Fx[k] += (double)fx[k];
            Fy[k] += (double)fy[k];
            fx[k] = 0.f;
            fy[k] = 0.f;
        }
    }

    double2* part = w.partial + (int64_t)blockIdx.y * (i1 - i0);
#pragma unroll
    for (int k = 0; k < kIPT; ++k) {
        const int64_t i = wbase + k * 32 + lane;
        if (i >= i0 && i < i1) part[i - i0] = make_double2(Fx[k], Fy[k]);
    }
}

// ---- pass 4: reduce the segment partials in a fixed order, containment + integrate ------------------
// nuclear_forces.py:301-323
__global__ void __launch_bounds__(256)
cloud_integrate_kernel(const float2* __restrict__ pos_in, float2* __restrict__ pos_out,
                       float2* __restrict__ vel, float2* __restrict__ force, int64_t n, int64_t i0,
                       int64_t i1, CloudWorkspace w, int n_seg, float dt)
{
    const int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    double sfx = 0.0, sfy = 0.0;
    for (int s = 0; s < n_seg; ++s) {
        const double2 p = w.partial[(int64_t)s * (i1 - i0) + (i - i0)];
        sfx += p.x;
        sfy += p.y;
    }
    const float cx = (float)w.centre[0], cy = (float)w.centre[1];
    const float R = 2.4f * cbrtf((float)n);            // :304
    const float2 p = pos_in[i];
    float2 v = vel[i];
    float x = p.x, y = p.y;
    float Fx = (float)sfx, Fy = (float)sfy;
    contain_and_integrate(x, y, v.x, v.y, Fx, Fy, cx, cy, R, dt);
    pos_out[i] = make_float2(x, y);
    vel[i] = v;
    if (force) force[i] = make_float2(Fx, Fy);         // containment is part of the reported force
}

// ---- sort keys (cloud_sort_key: cloud.cuh) ---------------------------------------------------------
__global__ void cloud_sort_keys_kernel(const float2* __restrict__ pos, const uint8_t* __restrict__ isp,
                                       int64_t n, float xmin, float ymin, float inv_extent,
                                       uint64_t* __restrict__ keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = cloud_sort_key(pos[i], isp[i] != 0, xmin, ymin, inv_extent);
}

// tile statistics + centre of mass, shared with the symmetric scheme (cloud_sym.cu)
int cloud_prepass(const float* pos, const uint8_t* is_proton, int64_t n, const CloudWorkspace& w,
                  cudaStream_t st)
{
    const int64_t nt = n_tiles_of(n);
    cloud_tile_stats<<<(unsigned)nt, kTile, 0, st>>>(reinterpret_cast<const float2*>(pos), is_proton,
                                                     n, w);
    cloud_centre<<<1, 256, 0, st>>>(w, nt, n);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int64_t pyqmd_cloud_workspace_bytes(int64_t n)
{
    if (n < 0) return PYQMD_ERR_INVALID;
    const int64_t nt = n_tiles_of(n);
    return 32 + nt * (int64_t)(sizeof(float4) + sizeof(double2) + sizeof(int)) + 512 +
           partial_entries(n) * (int64_t)sizeof(double2);
}

extern "C" int pyqmd_cloud_step(const float* pos_in, float* pos_out, float* vel, float* force,
                                const uint8_t* is_proton, int64_t n, int64_t i0, int64_t i1,
                                float strong, float coulomb, float pauli, float dt, void* workspace,
                                void* stream)
{
    PYQMD_REQUIRE(n >= 0 && i0 >= 0 && i0 <= i1 && i1 <= n, "0 <= i0 <= i1 <= n");
    if (n == 0 || i0 == i1) return PYQMD_OK;            // nuclear_forces.py:238-239
    PYQMD_REQUIRE(pos_in && pos_out && vel && is_proton && workspace, "NULL pointer");
    PYQMD_REQUIRE(pos_in != pos_out, "pos_in and pos_out must differ (Jacobi update)");
    cudaStream_t st = (cudaStream_t)stream;
    const CloudWorkspace w = carve(workspace, n);
    const int64_t nt = n_tiles_of(n);
    PYQMD_REQUIRE(nt <= 2147483647LL, "cloud too large");
    const LawParams L = make_law_params(strong, coulomb, pauli);
    {
        const int rc = cloud_prepass(pos_in, is_proton, n, w, st);
        if (rc != PYQMD_OK) return rc;
    }
    const int64_t blocks = (i1 - i0 + kIBlock - 1) / kIBlock;
    const int far_enabled = strong > 0.f ? 1 : 0;
    const int n_seg = segments_for(n, i1 - i0);
    const int tiles_per_seg = (int)((nt + n_seg - 1) / n_seg);
    PYQMD_REQUIRE((int64_t)n_seg * (i1 - i0) <= partial_entries(n), "workspace too small");
    const dim3 grid((unsigned)blocks, (unsigned)n_seg);
    const float2* pin = reinterpret_cast<const float2*>(pos_in);
    if (L.far_needs_clamp)
        cloud_force_kernel<true><<<grid, kThreads, 0, st>>>(pin, is_proton, n, i0, i1, w, L,
                                                            far_enabled, tiles_per_seg);
    else
        cloud_force_kernel<false><<<grid, kThreads, 0, st>>>(pin, is_proton, n, i0, i1, w, L,
                                                             far_enabled, tiles_per_seg);
    cloud_integrate_kernel<<<(unsigned)((i1 - i0 + 255) / 256), 256, 0, st>>>(
        pin, reinterpret_cast<float2*>(pos_out), reinterpret_cast<float2*>(vel),
        reinterpret_cast<float2*>(force), n, i0, i1, w, n_seg, dt);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}

extern "C" int pyqmd_cloud_sort_keys(const float* pos, const uint8_t* is_proton, int64_t n,
                                     float xmin, float ymin, float extent, uint64_t* keys,
                                     void* stream)
{
    PYQMD_REQUIRE(n >= 0 && extent > 0.f, "n >= 0 and extent > 0");
    if (n == 0) return PYQMD_OK;
    PYQMD_REQUIRE(pos && is_proton && keys, "NULL pointer");
    const int64_t blocks = (n + 255) / 256;
    cloud_sort_keys_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(pos), is_proton, n, xmin, ymin, 1.0f / extent, keys);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
