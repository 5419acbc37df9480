// cloud.cuh -- pieces shared by the two evaluation schemes of the single-cloud path
// (cloud.cu: ordered pairs per i-block; cloud_sym.cu: every unordered pair once, Newton's third law).
#pragma once
#include "common.cuh"
#include "pair_law.cuh"

namespace pyqmd {

constexpr int kTile = 256;      // j-tile size == tile-statistics granularity
constexpr int kThreads = 256;   // threads per block of the force kernel
constexpr int kIPT = 4;         // i-nucleons per thread
constexpr int kIBlock = kThreads * kIPT;

enum : int { kTileAllProton = 1, kTileAllNeutron = 2 };

struct CloudWorkspace {
    double* centre;   // [2] float64 centre of mass
    float4* bbox;     // [n_tiles] xmin, ymin, xmax, ymax
    double2* sums;    // [n_tiles] partial coordinate sums
    int* flags;       // [n_tiles] kTileAll*
    double2* partial; // [n_seg][i1 - i0] per-segment partial forces (ordered scheme)
};

constexpr int64_t kTargetUnits = 8192;   // ~28 work units per resident-block slot (148 SMs x 2)
constexpr int kMaxSeg = 64;

// number of j segments for an i-range of n_i nucleons of an n-nucleon cloud
__host__ inline int segments_for(int64_t n, int64_t n_i)
{
    const int64_t iblocks = (n_i + kIBlock - 1) / kIBlock;
    const int64_t tiles = (n + kTile - 1) / kTile;
    int64_t s = (kTargetUnits + iblocks - 1) / (iblocks > 0 ? iblocks : 1);
    if (s > kMaxSeg) s = kMaxSeg;
    if (s > tiles) s = tiles;
    if (s < 1) s = 1;
    return (int)s;
}

__host__ __device__ inline int64_t n_tiles_of(int64_t n) { return (n + kTile - 1) / kTile; }

inline int64_t partial_entries(int64_t n)
{
    // worst case over all i-ranges [i0, i1) of an n-nucleon cloud of n_seg * (i1 - i0)
    return n + (kTargetUnits + kMaxSeg) * (int64_t)kIBlock;
}

inline CloudWorkspace carve(void* ws, int64_t n)
{
    const int64_t nt = n_tiles_of(n);
    unsigned char* p = reinterpret_cast<unsigned char*>(ws);
    CloudWorkspace w;
    w.centre = reinterpret_cast<double*>(p);            p += 32;
    w.bbox = reinterpret_cast<float4*>(p);              p += sizeof(float4) * nt;
    w.sums = reinterpret_cast<double2*>(p);             p += sizeof(double2) * nt;
    w.flags = reinterpret_cast<int*>(p);              p += (sizeof(int) * nt + 255) / 256 * 256;
    w.partial = reinterpret_cast<double2*>(p);
    return w;
}

// ---- sort keys -----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread_bits(uint32_t v)
{
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

// 64-bit sort key: bit 62 = neutron (protons first for signed and unsigned sorts alike), low 48 bits
// = 2-D Morton code of the position inside [xmin, xmin + extent) x [ymin, ymin + extent)
__device__ __forceinline__ uint64_t cloud_sort_key(float2 p, bool proton, float xmin, float ymin, float inv_extent)
{
    const float ux = fminf(fmaxf((p.x - xmin) * inv_extent, 0.f), 0.99999994f);
    const float uy = fminf(fmaxf((p.y - ymin) * inv_extent, 0.f), 0.99999994f);
    const uint32_t qx = (uint32_t)(ux * 16777216.f), qy = (uint32_t)(uy * 16777216.f);
    uint64_t key = spread_bits(qx) | (spread_bits(qy) << 1);
    if (!proton) key |= (1ull << 62);
    return key;
}

struct FarConsts {
    f32x2 kexp, logA, l1, l2, g1, g2, one, negC, cg1, cg2;
};

// Far-field approximations (all inside the 1e-5 budget, errors relative to the exact term):
//   d/(d+eps) = 2^(-log2(1 + eps r)), r = 1/d <= 1/9: the correction is folded into the exponent of
//       the tail, -log2(1 + eps r) ~= r (l1 + l2 r)           (minimax, 5.8e-8 relative in the term)
//   1/(d2+eps) = r2 g(eps r2), g(w) ~= 1 + w (g1 + g2 w)      (minimax on w <= eps/81, 2e-10)
constexpr float kL1 = -0.2163957936310277f;      // in r (eps = 0.15 folded in)
constexpr float kL2 = 0.01598281877193875f;
constexpr float kG1 = -0.99999808f * kEps;       // in r2
constexpr float kG2 = 0.99722842f * kEps * kEps;

__device__ __forceinline__ FarConsts make_far_consts(const LawParams& L)
{
    FarConsts c;
    c.kexp = pk1(-1.8f * kLog2e / 7.0f);
    c.logA = pk1(L.log2TailK);
    c.l1 = pk1(kL1);
    c.l2 = pk1(kL2);
    c.g1 = pk1(kG1);
    c.g2 = pk1(kG2);
    c.one = pk1(1.0f);
    c.negC = pk1(-L.C);
    c.cg1 = pk1(-L.C * kG1);
    c.cg2 = pk1(-L.C * kG2);
    return c;
}

// s = net / d of two far pairs (packed), given d2 = dx^2 + dy^2 >= 81:
//   net = 0.15 S exp(-1.8 d/7)/(d + eps) - [pp] C/(d2 + eps)                     [281,285]
// With r = rsqrt(d2) the exponent of the tail is  logA + k d + r (l1 + l2 r) = logA + r (k d2 + l1 + l2 r)
// -- k d2 + l1 does not wait for the rsqrt -- and the Coulomb term of an all-proton tile pair is
// r2 r G(r2), G = -C g(r2) with the charge folded into the polynomial's coefficients.
// FMA-pipe operations per pair besides dx, dy, d2 and the force accumulation: 5 (MODE 0), 8 (MODE 1),
// 10 (MODE 2; cq = -C t_i t_j comes from the caller); special-function operations: 2 (rsqrt, ex2).
// NOEXP (MODE 1 / 2 only): the caller has proved that the tail term is exactly zero (cloud_sym.cu,
// kUltraGap), so e = +0 enters the same expression: identical bits without the exponential.
// A degree-5 polynomial 2^x on the FMA pipe for a share of the pairs was measured (r01e+) and
// rejected: every share was slower, the FMA pipe is as busy as the XU pipe.
template <int MODE, bool NOEXP = false>
__device__ __forceinline__ f32x2 far_s2(f32x2 d2, f32x2 cq, const FarConsts& c)
{
    float a0, a1;
    upk(d2, a0, a1);
    const f32x2 r = pk(mufu_rsqrt(a0), mufu_rsqrt(a1));
    f32x2 e = 0ull;
    if (!NOEXP) {
        const f32x2 u = fma2(d2, c.kexp, c.l1);
        const f32x2 arg = fma2(r, fma2(r, c.l2, u), c.logA);
        upk(arg, a0, a1);
        e = pk(mufu_ex2(a0), mufu_ex2(a1));
    }
    const f32x2 r2 = mul2(r, r);
    if (MODE == 0) return mul2(e, r2);
    if (MODE == 1) {
        const f32x2 G = fma2(r2, fma2(r2, c.cg2, c.cg1), c.negC);     // -C d2/(d2+eps)
        return mul2(r2, fma2(G, r, e));
    }
    const f32x2 g = fma2(r2, fma2(r2, c.g2, c.g1), c.one);            // d2/(d2+eps)
    return mul2(r2, fma2(mul2(cq, r), g, e));                         // cq carries the minus sign
}

// Two far pairs (one i, two j) in packed form: action on i.
template <int MODE>
__device__ __forceinline__ void far_pair2(f32x2 xj, f32x2 yj, f32x2 xi, f32x2 yi, f32x2 cq,
                                          const FarConsts& c, f32x2& fx, f32x2& fy)
{
    const f32x2 dx = sub2(xj, xi), dy = sub2(yj, yi);
    const f32x2 s = far_s2<MODE>(fma2(dy, dy, mul2(dx, dx)), cq, c);
    fx = fma2(dx, s, fx);
    fy = fma2(dy, s, fy);
}

}  // namespace pyqmd
