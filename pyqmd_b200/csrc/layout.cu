// layout.cu -- device-side initial nucleon layout: Nucleus.initialize_particles
// (OtsoBear/PyQMD particles.py:62-124) for every nucleus of an ensemble (sm_100a).
//
// The reference places nucleons one by one: pairs (proton, neutron) shell by shell (capacities
// 2, 8, 20, 28, 50, 82, 126, :67,:106-119), then the surplus protons, then the surplus neutrons
// (:121-124).  Each placement (:71-104) draws a radius factor, then 20 candidate angles, and keeps
// the candidate whose nearest SAME-TYPE neighbour is farthest (first strict maximum; with no
// same-type nucleon placed yet every candidate "wins", so the last one is kept).  0.25 s per U-238
// in Python, i.e. days for 10^6 nuclei -- here: one warp per nucleus, placements sequential (each
// depends on all earlier ones), the 20 candidates on 20 lanes, nearest-neighbour scan over the
// nucleons already placed (shared memory), float64 throughout like the reference.
//
// Draw order per placement k: slot 0 = radius factor (:74 random.random()), slots 1..20 = angles
// (:78 uniform(0, 2*pi) = 2*pi*random()).  `uniforms` (optional, double[n_list][cap][21]) injects
// them for parity tests; otherwise Philox4x32-10 keyed by (seed; global nucleus id, k, slot pair).
#include "common.cuh"
#include "decay_device.cuh"

namespace pyqmd {

constexpr int kTries = 20;              // particles.py:77
constexpr int kLayoutDraws = 21;
constexpr int kLayoutWarps = 4;

struct ShellRadii { double r[7]; };     // initial_radius * (i + 1) / 7, computed on the host (:64-68)

__global__ void __launch_bounds__(32 * kLayoutWarps)
init_layout_kernel(const pyqmd_ensemble e, const ShellRadii radii, const double* __restrict__ uniforms,
                   const uint64_t seed)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cap = e.cap;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double2* sp = reinterpret_cast<double2*>(smem_raw) + (size_t)wid * cap;
    uint8_t* st = smem_raw + sizeof(double2) * (size_t)kLayoutWarps * cap + (size_t)wid * cap;

    const int64_t q = (int64_t)blockIdx.x * kLayoutWarps + wid;
    if (q >= e.n_list) return;
    const int nuc = e.list ? e.list[q] : (int)q;
    const int32_t zn = e.zn[nuc];
    const int Z = zn >> 16, N = zn & 0xffff;
    const int A = min(Z + N, cap);
    const int64_t off = e.offset[nuc];
    const uint64_t gid = (uint64_t)(e.id_base + nuc);
    const int caps[7] = {2, 8, 20, 28, 50, 82, 126};

    auto draw = [&](int k, int slot) -> double {
        if (uniforms) return uniforms[((int64_t)q * cap + k) * kLayoutDraws + slot];
        uint32_t w[4];
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), 0x80000000u | (uint32_t)k,
                      (uint32_t)(slot >> 1), (uint32_t)seed, (uint32_t)(seed >> 32), w);
        return (slot & 1) ? u53(w[2], w[3]) : u53(w[0], w[1]);
    };

    // placement schedule (:106-124), warp-uniform scalar state
    int placed_p = 0, placed_n = 0, shell = 0, pairs_left = 0;
    bool pair_phase = true, next_is_proton = true;
    if (Z > 0 && N > 0) pairs_left = min(caps[0] / 2, min(Z, N));
    else pair_phase = false;
    int n_same_p = 0, n_same_n = 0;

    for (int k = 0; k < A; ++k) {
        bool want_p;
        int sh;
        if (pair_phase) {
            want_p = next_is_proton;
            sh = shell;
        } else {
            want_p = placed_p < Z;
            sh = shell;
        }
        const double shell_radius = radii.r[min(sh, 6)];
        const double radius = shell_radius * (0.8 + 0.2 * draw(k, 0));          // :74
        // candidate of this lane
        double px = 0.0, py = 0.0, gap = -1.0;
        if (lane < kTries) {
            const double angle = 6.283185307179586 * draw(k, 1 + lane);          // :78
            px = radius * cos(angle);                                            // :79-80 (centre 0,0)
            py = radius * sin(angle);
            double g2 = INFINITY;
            for (int j = 0; j < k; ++j) {                                        // :83-88
                if ((st[j] != 0) != want_p) continue;
                const double2 o = sp[j];
                const double ddx = o.x - px, ddy = o.y - py;
                g2 = fmin(g2, ddx * ddx + ddy * ddy);
            }
            gap = sqrt(g2);            // sqrt is monotonic: min of roots = root of min
        }
        const int n_same = want_p ? n_same_p : n_same_n;
        // :90-92 -- first strict maximum of the nearest-neighbour distance; all-inf: last candidate
        double best = gap;
        int who = lane;
        if (n_same == 0) {
            who = kTries - 1;
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int ow = __shfl_xor_sync(0xffffffffu, who, o);
                if (ob > best || (ob == best && ow < who)) { best = ob; who = ow; }
            }
            if (!(best > 0.0)) who = -1;          // every gap 0: best_angle stays 0 (:76)
        }
        double bx, by;
        if (who >= 0) {
            bx = __shfl_sync(0xffffffffu, px, who);
            by = __shfl_sync(0xffffffffu, py, who);
        } else {
            bx = radius;                          // cos(0), sin(0)
            by = 0.0;
        }
        if (lane == 0) {
            sp[k] = make_double2(bx, by);
            st[k] = want_p ? 1 : 0;
            reinterpret_cast<float2*>(e.pos)[off + k] = make_float2((float)bx, (float)by);
            reinterpret_cast<float2*>(e.vel)[off + k] = make_float2(0.f, 0.f);
            e.is_proton[off + k] = want_p ? 1 : 0;
        }
        __syncwarp();
        // advance the schedule
        if (want_p) { ++placed_p; ++n_same_p; } else { ++placed_n; ++n_same_n; }
        if (pair_phase) {
            if (next_is_proton) {
                next_is_proton = false;
            } else {
                next_is_proton = true;
                if (--pairs_left == 0) {
                    shell = min(shell + 1, 6);                                   // :117-119
                    if (placed_p < Z && placed_n < N)
                        pairs_left = min(caps[min(shell, 6)] / 2, min(Z - placed_p, N - placed_n));
                    else
                        pair_phase = false;
                }
            }
        }
    }
    if (lane == 0) e.count[nuc] = A;
}

}  // namespace pyqmd

using namespace pyqmd;

extern "C" int pyqmd_ensemble_init_layout(const pyqmd_ensemble* e, const double* shell_radii,
                                          const double* uniforms, unsigned long long seed, void* stream)
{
    PYQMD_REQUIRE(e != nullptr && shell_radii != nullptr, "NULL argument");
    PYQMD_REQUIRE(e->pos && e->vel && e->is_proton && e->offset && e->count && e->zn, "state arrays");
    PYQMD_REQUIRE(e->cap >= 1 && e->cap <= 1024, "cap must be in [1, 1024]");
    const int64_t n_list = e->list ? e->n_list : e->n_nuclei;
    if (n_list == 0) return PYQMD_OK;
    pyqmd_ensemble d = *e;
    d.n_list = n_list;
    ShellRadii r;
    for (int i = 0; i < 7; ++i) r.r[i] = shell_radii[i];
    const int64_t blocks = (n_list + kLayoutWarps - 1) / kLayoutWarps;
    PYQMD_REQUIRE(blocks <= 2147483647LL, "too many nuclei for one launch");
    const size_t smem = (sizeof(double2) + 1) * (size_t)kLayoutWarps * e->cap + 16;
    init_layout_kernel<<<(unsigned)blocks, 32 * kLayoutWarps, smem, (cudaStream_t)stream>>>(
        d, r, uniforms, (uint64_t)seed);
    PYQMD_CUDA_CHECK(cudaGetLastError());
    return PYQMD_OK;
}
