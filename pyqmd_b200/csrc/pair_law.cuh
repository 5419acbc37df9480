// pair_law.cuh -- the PyQMD nucleon pair-force law in FP32 for sm_100a.
//
// Behavioural spec: NuclearForces.update_particles_cpu, nuclear_forces.py:248-298 (reference
// root = OtsoBear/PyQMD), restated as branch-free FP32 with MUFU approximations whose error
// (rsqrt/ex2/rcp/sqrt.approx: <= 2 ulp each) stays inside the 1e-5 per-step budget.
// This is new code written for the B200; it is not a translation of the reference's
// OpenCL kernel (nuclear_forces.py:57-173), which is untiled, in-place and racy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pyqmd {

// Law constants (nuclear_forces.py line numbers in brackets)
constexpr float kEps        = 0.15f;    // [275,278,281,285]
constexpr float kSkipD2     = 0.01f;    // [257]
constexpr float kHardD      = 4.25f;    // [264]
constexpr float kCoreD      = 2.8f;     // [273]
constexpr float kAttrD      = 9.0f;     // [276]
constexpr float kPauliD     = 8.0f;     // [289]
constexpr float kMaxForce   = 12.0f;    // [294]
constexpr float kLog2e      = 1.4426950408889634f;
constexpr float kDamp       = 0.85f;    // [318-319]

struct LawParams {
    float S, C, P;          // strong / coulomb / pauli strength [13-15]
    // derived on the host once per call
    float coreK;            // 0.7 * S            [275]
    float attrK;            // 1.25 * S           [278]
    float tailK;            // 0.15 * S           [281]
    float log2TailK;        // log2(0.15 * S)  (tail coefficient folded into the exponent)
    float log2AttrK;        // log2(|1.25 * S|)
    float sgnS;             // sign of S: carried by 1/(d+eps) so that the folded coefficients stay positive
    int   far_needs_clamp;  // 1 if |net| could reach 12 for d >= 9 with these strengths
};

__host__ inline LawParams make_law_params(float S, float C, float P)
{
    LawParams p;
    p.S = S; p.C = C; p.P = P;
    p.coreK = 0.7f * S;
    p.attrK = 1.25f * S;
    p.tailK = 0.15f * S;
    p.log2TailK = (p.tailK != 0.f) ? log2f(fabsf(p.tailK)) : -150.f;
    p.log2AttrK = (p.attrK != 0.f) ? log2f(fabsf(p.attrK)) : -150.f;
    p.sgnS = (S < 0.f) ? -1.0f : 1.0f;
    // bound of |net| on d >= 9: tail <= tailK*exp(-1.8*9/7)/9.15, coulomb <= C/81.15
    double bound = fabs(0.15 * (double)S) * 0.09885 / 9.15 + fabs((double)C) / 81.15;
    p.far_needs_clamp = !(bound < 11.9) || !(S > 0.f);
    return p;
}

__device__ __forceinline__ float mufu_rsqrt(float x)
{
    float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
__device__ __forceinline__ float mufu_rcp(float x)
{
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
__device__ __forceinline__ float mufu_ex2(float x)
{
    float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
__device__ __forceinline__ float mufu_sqrt(float x)
{
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}

// ---- packed FP32 (Blackwell add/sub/mul/fma.f32x2 -> FADD2/FMUL2/FFMA2) -----------------------------
// One instruction works on two pairs: FMA-pipe work costs half the issue slots, which is what
// both hot kernels were limited by (ncu r01a: 82 % issue utilisation).
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk(float lo, float hi)
{
    f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ f32x2 pk1(float v) { return pk(v, v); }

// warp shuffle of a packed pair
__device__ __forceinline__ f32x2 shfl64(f32x2 v, int src)
{
    const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)(v & 0xffffffffull), src);
    const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
    return ((f32x2)hi << 32) | lo;
}

// General pair: every branch of the law, evaluated with selects (no divergence).
//   dx,dy  = r_j - r_i                                  [253-254]
//   ti,tj  = 1.0f for a proton, 0.0f for a neutron
// returns s such that  F_i += (dx,dy) * s   with  s = clamp(net)/d, or 0 for a skipped pair.
// The self pair (dx = dy = 0) is skipped by the d2 < 0.01 test [257], so callers need no i != j.
__device__ __forceinline__ float pair_general(float dx, float dy, float ti, float tj,
                                              const LawParams& L)
{
    const float d2 = fmaf(dy, dy, dx * dx);                       // [255]
    // d2 == 0 (self pair / coincident nucleons) gives inf/NaN below; the final select on
    // d2 < 0.01 discards them (FSEL does not propagate the unselected NaN)
    const float rinv = mufu_rsqrt(d2);
    const float d = d2 * rinv;                                    // [260]

    // hard core: -60 * ((4.25-d)/4.25)^1.5 for d < 4.25           [264-267]
    const float ov = __saturatef(fmaf(d, -1.0f / kHardD, 1.0f));
    float net = -60.0f * ov * mufu_sqrt(ov);

    // strong force: one reciprocal serves 1/(d+eps) and 1/(d2+eps) [275,278,281,285]
    const float a = fmaf(d, L.sgnS, L.sgnS * kEps);   // sgn(S) (d + eps)
    const float b = d2 + kEps;
    const float rab = mufu_rcp(a * b);
    const float inv_a = rab * b;          // sgn(S)/(d+eps)
    const float inv_b = rab * a;          // 1/(d2+eps): the signs cancel
    const bool is_core = d2 < kCoreD * kCoreD;                    // [273]
    const bool is_attr = d2 < kAttrD * kAttrD;                    // [276]
    // coefficient folded into the exponent: coef * exp(-k d) = 2^(log2 coef - k' d)
    const float kexp = is_attr ? (-kLog2e / 7.0f) : (-1.8f * kLog2e / 7.0f);
    const float lexp = is_attr ? L.log2AttrK : L.log2TailK;
    const float e = mufu_ex2(fmaf(d, kexp, lexp));                // coef * exp(-d/7) or coef * exp(-1.8 d/7)
    const float strong = is_core ? (-L.coreK * inv_b) : (e * inv_a);
    net += strong;

    // Coulomb between protons                                     [284-285]
    net = fmaf(-(L.C * ti * tj), inv_b, net);

    // Pauli for equal types and d < 8                              [288-291]
    const float pe = mufu_ex2(d * (-2.0f * kLog2e / kPauliD));
    const bool pauli = (ti == tj) && (d2 < kPauliD * kPauliD);
    net = pauli ? fmaf(-L.P, pe, net) : net;

    net = fminf(fmaxf(net, -kMaxForce), kMaxForce);               // [294]
    const float s = net * rinv;                                   // [297-298]: (dx*net)/d
    return (d2 < kSkipD2) ? 0.0f : s;                             // [257]
}

// Two general pairs at once: nucleons (i_a, i_b) of one thread against the same partner j, all
// FMA-pipe arithmetic packed (f32x2), compares / selects / min-max / MUFU per element.
//   dx, dy = (r_j - r_ia, r_j - r_ib);  tj2 = (t_j, t_j);  nq = (-C t_ia, -C t_ib).
// Same arithmetic as pair_general.
struct GenConsts {
    f32x2 nInvHard, one, eps, negCoreK, n60, negP, kPauli;
    f32x2 kAttr, lAttr, kTail, lTail;      // 2^(k d + l) forms of the attractive / tail strong terms
    f32x2 sgn, sgnEps;                     // sign of S and sgn * eps (see LawParams::sgnS)
};

__device__ __forceinline__ GenConsts make_gen_consts(const LawParams& L)
{
    GenConsts c;
    c.nInvHard = pk1(-1.0f / kHardD);
    c.one = pk1(1.0f);
    c.eps = pk1(kEps);
    c.negCoreK = pk1(-L.coreK);
    c.n60 = pk1(-60.0f);
    c.negP = pk1(-L.P);
    c.kPauli = pk1(-2.0f * kLog2e / kPauliD);
    c.kAttr = pk1(-kLog2e / 7.0f);
    c.kTail = pk1(-1.8f * kLog2e / 7.0f);
    c.lAttr = pk1(L.log2AttrK);
    c.lTail = pk1(L.log2TailK);
    c.sgn = pk1(L.sgnS);
    c.sgnEps = pk1(L.sgnS * kEps);
    return c;
}

__device__ __forceinline__ f32x2 pair_general2(f32x2 dx, f32x2 dy, float ta, float tb, float tj,
                                               f32x2 tj2, f32x2 nq, const GenConsts& c,
                                               const LawParams& L)
{
    const f32x2 d2 = fma2(dy, dy, mul2(dx, dx));                    // [255]
    float d2a, d2b;
    upk(d2, d2a, d2b);
    const f32x2 rinv = pk(mufu_rsqrt(d2a), mufu_rsqrt(d2b));
    const f32x2 d = mul2(d2, rinv);                                 // [260]
    // hard core                                                     [264-267]
    // 1 - d/4.25 clamped at 0 is a saturating FMA (d >= 0, so the upper bound 1 never binds; a NaN d
    // of a self pair saturates to 0): two scalar FFMA.SAT instead of FFMA2 + two FMNMX
    float da, db;
    upk(d, da, db);
    const float ova = __saturatef(fmaf(da, -1.0f / kHardD, 1.0f));
    const float ovb = __saturatef(fmaf(db, -1.0f / kHardD, 1.0f));
    const f32x2 hc = mul2(pk(ova, ovb), pk(mufu_sqrt(ova), mufu_sqrt(ovb)));
    // strong                                                        [273-281]
    const f32x2 a = fma2(d, c.sgn, c.sgnEps), b = add2(d2, c.eps);   // a = sgn(S) (d + eps)
#ifdef PYQMD_GEN_TWO_RCP
    float aa, ab, ba, bb;
    upk(a, aa, ab);
    upk(b, ba, bb);
    const f32x2 inv_a = pk(mufu_rcp(aa), mufu_rcp(ab)), inv_b = pk(mufu_rcp(ba), mufu_rcp(bb));
#else
    float aba, abb;
    upk(mul2(a, b), aba, abb);
    const f32x2 rab = pk(mufu_rcp(aba), mufu_rcp(abb));
    const f32x2 inv_a = mul2(rab, b), inv_b = mul2(rab, a);
#endif
    const bool attr_a = d2a < kAttrD * kAttrD, attr_b = d2b < kAttrD * kAttrD;
    // coefficient folded into the exponent: |coef| * 2^(k d) = 2^(k d + log2 |coef|) (its sign rides on
    // inv_a); both candidate arguments are computed packed, one select per element picks the branch
    // (A/B on B200, C2: +5 % over select-k / select-coef / multiply)
    float xa, xb, ya, yb;
    upk(fma2(d, c.kAttr, c.lAttr), xa, xb);
    upk(fma2(d, c.kTail, c.lTail), ya, yb);
    const f32x2 e = pk(mufu_ex2(attr_a ? xa : ya), mufu_ex2(attr_b ? xb : yb));
    float sfa, sfb, sca, scb;
    upk(mul2(e, inv_a), sfa, sfb);
    upk(mul2(c.negCoreK, inv_b), sca, scb);
    const f32x2 strong = pk(d2a < kCoreD * kCoreD ? sca : sfa, d2b < kCoreD * kCoreD ? scb : sfb);
    f32x2 net = fma2(hc, c.n60, strong);
    // Coulomb                                                       [284-285]
    net = fma2(mul2(nq, tj2), inv_b, net);
    // Pauli                                                         [288-291]
#ifdef PYQMD_PAULI_POLY
    // exp(-d/4) on the FMA pipe: (p6(d))^2, p6 ~ exp(-d/8) minimax on [0, 8] (5.6e-7 relative after
    // squaring; the value is discarded by the select below for d >= 8).  The general law is bound by
    // the MUFU pipe (5 per pair, 16/clk/SM), which this takes to 4.
    f32x2 q = fma2(d, pk1(3.188497021966441e-09f), pk1(-2.323615291288661e-07f));
    q = fma2(q, d, pk1(1.0052935977000743e-05f));
    q = fma2(q, d, pk1(-0.0003251763409934938f));
    q = fma2(q, d, pk1(0.0078120157122612f));
    q = fma2(q, d, pk1(-0.12499973922967911f));
    q = fma2(q, d, c.one);
    const f32x2 netp = fma2(c.negP, mul2(q, q), net);
#else
    float pa, pb;
    upk(mul2(d, c.kPauli), pa, pb);
    const f32x2 netp = fma2(c.negP, pk(mufu_ex2(pa), mufu_ex2(pb)), net);
#endif
    float na, nb, npa, npb;
    upk(net, na, nb);
    upk(netp, npa, npb);
    na = (ta == tj && d2a < kPauliD * kPauliD) ? npa : na;
    nb = (tb == tj && d2b < kPauliD * kPauliD) ? npb : nb;
    na = fminf(fmaxf(na, -kMaxForce), kMaxForce);                   // [294]
    nb = fminf(fmaxf(nb, -kMaxForce), kMaxForce);
    float sa, sb;
    upk(mul2(pk(na, nb), rinv), sa, sb);                            // [297-298]
    return pk(d2a < kSkipD2 ? 0.f : sa, d2b < kSkipD2 ? 0.f : sb);  // [257]
}

// Far pair, valid only when the caller has proved d >= 9 for the pair (tile bounding boxes):
//   net = 0.15 S exp(-1.8 d / 7)/(d + eps) - [pp] C/(d2 + eps)     [281,285]
// Neither the hard core, the Pauli term (d < 8) nor the d2 < 0.01 skip can apply.
// 1/(d+eps) is expanded around 1/d (eps/d <= 1/60): (1 - z + z^2 - z^3), z = eps/d, relative
// truncation error z^4 <= 7.7e-8; this moves one MUFU op onto the FMA pipe.
// The tail coefficient is folded into the exponent: 0.15 S exp(-k d) = 2^(log2(0.15 S) - k' d).
// MODE 0: no Coulomb term; 1: every pair is p-p (cq ignored, C used); 2: per-pair charge
// product cq = C * t_i * t_j supplied by the caller.
template <int MODE, bool CLAMP>
__device__ __forceinline__ float pair_far_impl(float dx, float dy, float cq, const LawParams& L)
{
    const float d2 = fmaf(dy, dy, dx * dx);
    const float r = mufu_rsqrt(d2);
    const float d = d2 * r;
    const float e = mufu_ex2(fmaf(d, -1.8f * kLog2e / 7.0f, L.log2TailK));
    float h = fmaf(r, -kEps * kEps * kEps, kEps * kEps);
    h = fmaf(r, h, -kEps);
    h = fmaf(r, h, 1.0f);                   // d/(d+eps)
    const float r2 = r * r;
    const float c = (MODE == 1) ? L.C : cq;
    if (!CLAMP) {
        // s = net/d = e * h / d2  (- c / ((d2+eps) d) with Coulomb)
        float s = e * (h * r2);
        if (MODE != 0) {
            // 1/(d2+eps) = r2 * (1 - eps r2 + eps^2 r2^2), eps r2 <= 1.9e-3: error < 7e-9
            float g = fmaf(r2, kEps * kEps, -kEps);
            g = fmaf(r2, g, 1.0f);
            s = fmaf(-c * r, g * r2, s);
        }
        return s;
    } else {
        float net = e * (h * r);
        if (MODE != 0) {
            float g = fmaf(r2, kEps * kEps, -kEps);
            g = fmaf(r2, g, 1.0f);
            net = fmaf(-c, g * r2, net);
        }
        net = fminf(fmaxf(net, -kMaxForce), kMaxForce);
        return net * r;
    }
}

template <bool PP, bool CLAMP>
__device__ __forceinline__ float pair_far(float dx, float dy, const LawParams& L)
{
    return pair_far_impl<PP ? 1 : 0, CLAMP>(dx, dy, 0.f, L);
}

template <bool CLAMP>
__device__ __forceinline__ float pair_far_q(float dx, float dy, float cq, const LawParams& L)
{
    return pair_far_impl<2, CLAMP>(dx, dy, cq, L);
}

// Centre-of-mass containment + damped Euler, per nucleon.
//   [301-309]  if |c - x| > 1.5 R and > 0.01:  F += 0.03 (|c - x| - R) (c - x)/|c - x|
//   [312-323]  v += F dt; v *= 0.85; x += v dt
// R = 1.2 * n^(1/3) * 2.0 is computed by the caller [304].
__device__ __forceinline__ void contain_and_integrate(float& x, float& y, float& vx, float& vy,
                                                      float& fx, float& fy, float cx, float cy,
                                                      float R, float dt)
{
    const float cdx = cx - x, cdy = cy - y;
    const float cd2 = fmaf(cdy, cdy, cdx * cdx);
    const float cd = sqrtf(cd2);
    if (cd > R * 1.5f && cd > 0.01f) {
        const float cf = 0.03f * (cd - R) / cd;
        fx = fmaf(cf, cdx, fx);
        fy = fmaf(cf, cdy, fy);
    }
    vx = fmaf(fx, dt, vx) * kDamp;
    vy = fmaf(fy, dt, vy) * kDamp;
    x = fmaf(vx, dt, x);
    y = fmaf(vy, dt, y);
}

}  // namespace pyqmd
