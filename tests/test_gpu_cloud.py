"""GPU parity of the single-cloud path (BASELINE config 4) against the oracle, through
pyqmd_cloud_step (ordered i-block scheme), pyqmd_cloud_pair_forces + pyqmd_cloud_integrate
(symmetric scheme: every unordered pair once) and pyqmd_cloud_sort_keys."""
import numpy as np
import pytest
import torch

from gpu_util import FORCE_TOL, POS_TOL, extent_of
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def make_cloud(n, seed=1234, frac_p=0.4, density=1 / 25):
    """i.i.d. uniform in a disc of number density 1/25, exactly
    round(frac_p * n) protons at shuffled indices, PCG64(seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R = np.sqrt(n / density / np.pi)
    r = R * np.sqrt(rng.random(n))
    th = 2 * np.pi * rng.random(n)
    pos = np.stack([r * np.cos(th), r * np.sin(th)], 1).astype(np.float32)
    isp = np.zeros(n, np.uint8)
    isp[rng.permutation(n)[: int(round(frac_p * n))]] = 1
    return pos, isp


def oracle_forces(pos, isp, i0, i1, strengths=(150.0, 30.0, 35.0), amb_check=True):
    x = pos[:, 0].astype(np.float64); y = pos[:, 1].astype(np.float64)
    cx, cy = float(x.mean()), float(y.mean())
    return orc.cloud_forces(x, y, isp, i0, i1, *strengths, center=(cx, cy))


def ambiguous_mask(pos, i0, i1, tol=1e-6):
    """Nucleons of [i0,i1) with a partner within tol (relative) of a branch threshold."""
    p = pos.astype(np.float64)
    out = np.zeros(i1 - i0, bool)
    for k, i in enumerate(range(i0, i1)):
        d = np.hypot(p[:, 0] - p[i, 0], p[:, 1] - p[i, 1])
        d[i] = 1e9
        for thr in (0.1, 2.8, 4.25, 8.0, 9.0):
            if (np.abs(d - thr) <= tol * thr * 2).any():
                out[k] = True
    return out


def rel_l2(a, b, mask=None):
    if mask is not None:
        a, b = a[~mask], b[~mask]
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


SCHEMES = ["symmetric", "ordered"]


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("sort", [True, False])
def test_subcloud_4096_forces_and_step(sort, scheme):
    from pyqmd_b200.state import NucleonCloud
    n = 4096
    pos, isp = make_cloud(n)
    cloud = NucleonCloud(pos, isp, sort=sort, keep_force=True, scheme=scheme)
    cloud.step(1)
    F = cloud.forces().cpu().numpy().astype(np.float64)
    fx, fy = oracle_forces(pos, isp, 0, n)
    amb = ambiguous_mask(pos, 0, n)
    assert amb.sum() < 20
    Fo = np.stack([fx, fy], 1)
    assert rel_l2(F, Fo, amb) <= FORCE_TOL
    # one step from rest: x' = x + 0.85 F dt^2
    dt = cloud.dt
    want = pos.astype(np.float64) + 0.85 * Fo * dt * dt
    got = cloud.positions().cpu().numpy().astype(np.float64)
    err = np.hypot(*(got - want).T)[~amb].max() / extent_of(pos)
    assert err <= POS_TOL
    v = cloud.velocities().cpu().numpy().astype(np.float64)
    assert rel_l2(v, 0.85 * Fo * dt, amb) <= 2e-5


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("n", [1, 2, 31, 255, 256, 257, 1023, 1024, 1025, 3000])
def test_ragged_sizes(n, scheme):
    """Tile (256) and i-block (1024) boundaries, partial last tile, single nucleon."""
    from pyqmd_b200.state import NucleonCloud
    pos, isp = make_cloud(n, seed=n)
    cloud = NucleonCloud(pos, isp, keep_force=True, scheme=scheme)
    cloud.step(1)
    F = cloud.forces().cpu().numpy().astype(np.float64)
    fx, fy = oracle_forces(pos, isp, 0, n)
    amb = ambiguous_mask(pos, 0, n)
    Fo = np.stack([fx, fy], 1)
    if np.abs(Fo[~amb]).max() > 0:
        assert rel_l2(F, Fo, amb) <= FORCE_TOL
    else:
        assert np.abs(F[~amb]).max() == 0.0


@pytest.mark.parametrize("scheme", SCHEMES)
def test_dense_cloud_all_branches(scheme):
    """A compressed cloud (mean spacing ~2) exercises hard core, core, attractive, Pauli and
    the clamp in the tiled kernel's near path."""
    from pyqmd_b200.state import NucleonCloud
    n = 2000
    pos, isp = make_cloud(n, seed=7, density=1 / 4)
    cloud = NucleonCloud(pos, isp, keep_force=True, scheme=scheme)
    cloud.step(1)
    F = cloud.forces().cpu().numpy().astype(np.float64)
    fx, fy = oracle_forces(pos, isp, 0, n)
    amb = ambiguous_mask(pos, 0, n)
    assert rel_l2(F, np.stack([fx, fy], 1), amb) <= FORCE_TOL


@pytest.mark.parametrize("scheme", SCHEMES)
def test_clamped_far_path_with_large_strengths(scheme):
    from pyqmd_b200.state import NucleonCloud
    n = 3000
    pos, isp = make_cloud(n, seed=9)
    st = (20000.0, 3000.0, 35.0)          # tail + Coulomb can hit the +-12 cap beyond d = 9
    cloud = NucleonCloud(pos, isp, keep_force=True, strengths=st, scheme=scheme)
    cloud.step(1)
    F = cloud.forces().cpu().numpy().astype(np.float64)
    fx, fy = oracle_forces(pos, isp, 0, n, strengths=st)
    amb = ambiguous_mask(pos, 0, n)
    assert rel_l2(F, np.stack([fx, fy], 1), amb) <= FORCE_TOL


def test_i_block_sharding_is_bit_identical():
    """Two 'ranks' computing [0, n/2) and [n/2, n) of the same replica reproduce the single-GPU
    step exactly (what the all-gather path relies on)."""
    from pyqmd_b200 import _lib
    from pyqmd_b200.state import NucleonCloud
    n = 5000
    pos, isp = make_cloud(n, seed=3)
    full = NucleonCloud(pos, isp, scheme="ordered")
    full.step(1)
    a = NucleonCloud(pos, isp, rank=0, world=1, scheme="ordered")
    b = NucleonCloud(pos, isp, rank=0, world=1, scheme="ordered")
    # emulate world = 2 without a process group: restrict the i-range by hand
    h = (n + 1) // 2
    a.i0, a.i1 = 0, h
    b.i0, b.i1 = h, n
    a.step(1); b.step(1)
    merged = torch.cat([a.pos[:h], b.pos[h:n]])
    assert torch.equal(merged, full.pos[:n])


def test_symmetric_parts_sum_bit_identically():
    """The fixed-point accumulators do not depend on how the rows are dealt to parts (what makes the
    multi-GPU reduce-scatter exact): 1 part == 3 parts == 8 parts, bit for bit; and the symmetric
    and ordered schemes agree to FP32 rounding."""
    import ctypes as C
    from pyqmd_b200 import _lib
    from pyqmd_b200.state import NucleonCloud
    n = 20_000
    pos, isp = make_cloud(n, seed=5)
    cloud = NucleonCloud(pos, isp, keep_force=True)
    lib = _lib.lib()
    accs = []
    for parts in (1, 3, 8):
        acc = torch.zeros(n, 2, dtype=torch.int64, device="cuda")
        for part in range(parts):
            _lib.check(lib.pyqmd_cloud_pair_forces(
                cloud.pos.data_ptr(), cloud.is_proton.data_ptr(), n, part, parts, 150.0, 30.0, 35.0,
                acc.data_ptr(), cloud.workspace.data_ptr(), _lib.current_stream()), "pair_forces")
        torch.cuda.synchronize()
        accs.append(acc)
    assert torch.equal(accs[0], accs[1]) and torch.equal(accs[0], accs[2])
    assert int(accs[0].abs().max()) > 0
    # Newton's third law: action and reaction are rounded separately (FP32 partial sums), so the pair
    # forces cancel over the cloud to FP32 rounding of the summed magnitudes
    assert accs[0].sum(0).abs().max().item() <= 1e-6 * accs[0].abs().sum().item()
    scale = 2.0 ** cloud.force_scale_log2
    F_sym = (accs[0].double() / scale).cpu().numpy()
    ordered = NucleonCloud(pos, isp, keep_force=True, scheme="ordered")
    ordered.step(1)
    cloud.step(1)
    Fo, Fs = ordered.forces().cpu().numpy().astype(np.float64), cloud.forces().cpu().numpy().astype(np.float64)
    assert rel_l2(Fs, Fo) <= 2e-6
    assert torch.count_nonzero(cloud.acc).item() == 0          # integrate leaves the accumulators cleared
    # pair part of the integrated force = accumulators (containment is added by the integrate pass)
    far_from_edge = np.hypot(*cloud.pos[:n].cpu().numpy().T) < 90     # containment starts at 1.5 R = 97.7
    assert far_from_edge.any()
    Fs_sorted = cloud.force.cpu().numpy().astype(np.float64)
    assert np.abs(Fs_sorted[far_from_edge] - F_sym[far_from_edge]).max() <= 1e-5


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("st", [(0.0, 30.0, 35.0), (-40.0, 30.0, 35.0)])
def test_zero_and_negative_strong_strength(scheme, st):
    """S <= 0 disables the far-field fast path (its folded log2 coefficient needs S > 0); the general
    path carries the sign of S on 1/(d+eps)."""
    from pyqmd_b200.state import NucleonCloud
    n = 2500
    pos, isp = make_cloud(n, seed=13)
    cloud = NucleonCloud(pos, isp, keep_force=True, strengths=st, scheme=scheme)
    cloud.step(1)
    F = cloud.forces().cpu().numpy().astype(np.float64)
    fx, fy = oracle_forces(pos, isp, 0, n, strengths=st)
    amb = ambiguous_mask(pos, 0, n)
    assert rel_l2(F, np.stack([fx, fy], 1), amb) <= FORCE_TOL


@pytest.mark.parametrize("scheme", SCHEMES)
def test_large_cloud_sampled_against_oracle(scheme):
    """N = 200,000: sorted, tile-classified fast path; 384 sampled nucleons against the oracle."""
    from pyqmd_b200.state import NucleonCloud
    n = 200_000
    pos, isp = make_cloud(n, seed=11)
    cloud = NucleonCloud(pos, isp, keep_force=True, scheme=scheme)
    cloud.step(1)
    F = cloud.forces().cpu().numpy().astype(np.float64)
    x = pos[:, 0].astype(np.float64); y = pos[:, 1].astype(np.float64)
    c = (float(x.mean()), float(y.mean()))
    for i0 in (0, 77_777, n - 128):
        fx, fy = orc.cloud_forces(x, y, isp, i0, i0 + 128, center=c)
        amb = ambiguous_mask(pos, i0, i0 + 128)
        assert rel_l2(F[i0:i0 + 128], np.stack([fx, fy], 1), amb) <= FORCE_TOL
    # Newton 3 on the whole cloud: pair forces cancel, only containment remains
    assert np.isfinite(F).all()


@pytest.mark.parametrize("scheme", SCHEMES)
def test_multi_step_cloud_teacher_forced(scheme):
    from pyqmd_b200.state import NucleonCloud
    n = 1500
    pos, isp = make_cloud(n, seed=21, density=1 / 9)
    cloud = NucleonCloud(pos, isp, keep_force=True, sort=False, scheme=scheme)
    for s in range(4):
        p0 = cloud.pos[:n].cpu().numpy().copy()
        cloud.step(1)
        fx, fy = oracle_forces(p0, isp, 0, n)
        amb = ambiguous_mask(p0, 0, n)
        assert rel_l2(cloud.force.cpu().numpy().astype(np.float64), np.stack([fx, fy], 1), amb) <= FORCE_TOL


def test_full_size_cloud_1M_properties():
    """Config C4 at full size (N = 10^6, 40 % protons): sampled nucleons against the oracle, Newton's
    third law over the whole cloud, and the symmetric scheme against the ordered kernel on a block."""
    from pyqmd_b200 import _lib
    from pyqmd_b200.state import NucleonCloud
    n = 1_000_000
    pos, isp = make_cloud(n)
    cloud = NucleonCloud(pos, isp, keep_force=True)
    p_sorted = cloud.pos[:n].cpu().numpy().copy()
    t_sorted = cloud.is_proton.cpu().numpy().copy()
    lib = _lib.lib()
    acc = torch.zeros(n, 2, dtype=torch.int64, device="cuda")
    _lib.check(lib.pyqmd_cloud_pair_forces(cloud.pos.data_ptr(), cloud.is_proton.data_ptr(), n, 0, 1,
                                           150.0, 30.0, 35.0, acc.data_ptr(), cloud.workspace.data_ptr(),
                                           _lib.current_stream()), "pair_forces")
    torch.cuda.synchronize()
    assert acc.sum(0).abs().max().item() <= 1e-6 * acc.abs().sum().item()      # pair forces cancel
    cloud.step(1)
    F = cloud.force.cpu().numpy().astype(np.float64)                           # sorted order
    x, y = p_sorted[:, 0].astype(np.float64), p_sorted[:, 1].astype(np.float64)
    c = (float(x.mean()), float(y.mean()))
    for i0 in (0, 399_900, 654_321, n - 64):                                   # p block, p/n border, n block
        fx, fy = orc.cloud_forces(x, y, t_sorted, i0, i0 + 64, center=c)
        amb = ambiguous_mask(p_sorted, i0, i0 + 64)
        assert rel_l2(F[i0:i0 + 64], np.stack([fx, fy], 1), amb) <= FORCE_TOL, i0
    # ordered kernel on one 4096-nucleon block of the same (sorted) cloud
    o = NucleonCloud(p_sorted, t_sorted, keep_force=True, scheme="ordered", sort=False)
    o.i0, o.i1 = 500_000, 504_096
    o.step(1)
    Fo = o.force[o.i0:o.i1].cpu().numpy().astype(np.float64)
    assert rel_l2(F[o.i0:o.i1], Fo) <= 2e-6
    # one step from rest: x' = x + 0.85 F dt^2
    got = cloud.pos[:n].cpu().numpy().astype(np.float64)
    want = p_sorted.astype(np.float64) + 0.85 * F * cloud.dt ** 2
    assert np.abs(got - want).max() / extent_of(p_sorted) <= POS_TOL


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_random_clouds_against_oracle(seed):
    """Randomised differential test: random size, density, proton fraction, strengths, scheme and
    ordering; forces of one step against the float64 oracle."""
    from pyqmd_b200.state import NucleonCloud
    rng = np.random.default_rng(seed)
    for _ in range(6):
        n = int(rng.integers(2, 7000))
        density = float(rng.choice([1 / 4, 1 / 25, 1 / 100]))
        frac_p = float(rng.choice([0.0, 0.1, 0.4, 0.9, 1.0]))
        st = (float(rng.uniform(1, 300)), float(rng.uniform(0, 60)), float(rng.uniform(0, 60)))
        scheme = str(rng.choice(SCHEMES))
        sort = bool(rng.random() < 0.7)
        pos, isp = make_cloud(n, seed=int(rng.integers(1 << 30)), frac_p=frac_p, density=density)
        cloud = NucleonCloud(pos, isp, keep_force=True, strengths=st, scheme=scheme, sort=sort)
        cloud.step(1)
        F = cloud.forces().cpu().numpy().astype(np.float64)
        fx, fy = oracle_forces(pos, isp, 0, n, strengths=st)
        amb = ambiguous_mask(pos, 0, n)
        Fo = np.stack([fx, fy], 1)
        if np.abs(Fo[~amb]).max() > 0:
            assert rel_l2(F, Fo, amb) <= FORCE_TOL, (n, density, frac_p, st, scheme, sort)


# ---- host-buffer entry point: pyqmd_cloud_step_host / the large-n branch of pyqmd_update_particles_f64 --
@pytest.mark.parametrize("n", [1025, 4096, 20000])
def test_host_buffer_cloud_step_matches_oracle(n):
    """The reference-shaped call for one large system (nuclear_forces.py:185-234): host arrays in,
    host arrays out, caller's order kept; force-L2 and position gates against the oracle."""
    from pyqmd_b200.forces import NuclearForces
    pos, isp = make_cloud(n, seed=n + 1)
    rng = np.random.default_rng(n)
    vel = (rng.standard_normal((n, 2)) * 0.05).astype(np.float32)
    p, v = pos.copy(), vel.copy()
    F = np.zeros((n, 2), np.float32)
    nf = NuclearForces()
    nf.step_cloud(p, v, isp, 1 / 240, 1, force=F)
    fx, fy = oracle_forces(pos, isp, 0, n)
    amb = ambiguous_mask(pos, 0, n) if n <= 4096 else np.zeros(n, bool)
    Fo = np.stack([fx, fy], 1)
    assert rel_l2(F.astype(np.float64), Fo, amb) <= FORCE_TOL
    dt = 1 / 240
    vw = (vel.astype(np.float64) + Fo * dt) * 0.85
    want = pos.astype(np.float64) + vw * dt
    err = np.hypot(*(p.astype(np.float64) - want).T)[~amb].max() / extent_of(pos)
    assert err <= POS_TOL
    assert rel_l2(v.astype(np.float64), vw, amb) <= 2e-5


def test_host_buffer_cloud_equals_device_resident_steps():
    """3 steps through the host entry point == 3 steps of the device-resident NucleonCloud (same
    scheme, same sort keys up to the FP32 rounding of the bounding box), to FP32 rounding."""
    from pyqmd_b200.forces import NuclearForces
    from pyqmd_b200.state import NucleonCloud
    n = 30000
    pos, isp = make_cloud(n, seed=5)
    p, v = pos.copy(), np.zeros_like(pos)
    NuclearForces().step_cloud(p, v, isp, 1 / 240, 3)
    cloud = NucleonCloud(pos, isp)
    cloud.step(3)
    got = cloud.positions().cpu().numpy()
    assert np.abs(got.astype(np.float64) - p).max() / extent_of(pos) <= 1e-6
    assert rel_l2(v.astype(np.float64), cloud.velocities().cpu().numpy().astype(np.float64)) <= 1e-5


def test_large_system_through_step_arrays_uses_the_sorted_symmetric_scheme():
    """NuclearForces.step_arrays (float64 state, nuclear_forces.py:236 at N = 200k): force-L2 gate on a
    sample of nucleons (the oracle evaluates rows [i0, i1) against all partners)."""
    from pyqmd_b200.forces import NuclearForces
    n = 200_000
    pos, isp = make_cloud(n, seed=11)
    x, y = pos[:, 0].astype(np.float64) + 400.0, pos[:, 1].astype(np.float64) + 400.0
    vx, vy = np.zeros(n), np.zeros(n)
    x0, y0 = x.copy(), y.copy()
    NuclearForces().step_arrays(x, y, vx, vy, isp, 1 / 240, 1)
    i0, i1 = 70_000, 70_512
    fx, fy = oracle_forces(pos, isp, i0, i1)
    dt = 1 / 240
    Fg = np.stack([vx[i0:i1], vy[i0:i1]], 1) / (0.85 * dt)        # v' = 0.85 F dt from rest
    assert rel_l2(Fg, np.stack([fx, fy], 1)) <= FORCE_TOL
    want = np.stack([x0[i0:i1], y0[i0:i1]], 1) + 0.85 * np.stack([fx, fy], 1) * dt * dt
    err = np.hypot(x[i0:i1] - want[:, 0], y[i0:i1] - want[:, 1]).max() / extent_of(pos)
    assert err <= POS_TOL


# ---- the multi-GPU exchange on ONE device: virtual ranks ------------------------------------------------
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_symmetric_scheme_with_virtual_ranks_on_one_gpu(world):
    """The strong-scaling path of bench.py -- rows dealt to `world` parts, one int64 accumulator array
    per part, pyqmd_cloud_exchange_integrate pulling / clearing every part's accumulators and pushing the
    new positions into every replica -- run with all parts on ONE device (the peers' pointer tables simply
    point at buffers of the same GPU).  Must be bit-identical to the single-part step, for any `world`:
    this keeps the fused exchange kernel under test on single-GPU boxes (tests/test_gpu_multi.py needs 2+)."""
    import ctypes as C

    from pyqmd_b200 import _lib
    from pyqmd_b200.sharding import cloud_chunk, shard_range
    from pyqmd_b200.state import NucleonCloud
    n = 20_011
    pos, isp = make_cloud(n, seed=21)
    single = NucleonCloud(pos, isp)
    single.step(2)
    # sorted start state shared by all virtual ranks
    ref = NucleonCloud(pos, isp)
    dev = ref.device
    lib = _lib.lib()
    chunk = cloud_chunk(n, world)
    padded = chunk * world
    pos_a = [torch.zeros(padded, 2, device=dev) for _ in range(world)]
    pos_b = [torch.zeros(padded, 2, device=dev) for _ in range(world)]
    acc = [torch.zeros(padded, 2, dtype=torch.int64, device=dev) for _ in range(world)]
    vel = [ref.vel.clone() for _ in range(world)]
    ws = [torch.zeros_like(ref.workspace) for _ in range(world)]
    for r in range(world):
        pos_a[r][:n] = ref.pos[:n]
    ptrs = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)
    acc_ptrs = ptrs(acc)
    S, Cc, P = ref.strengths
    st = _lib.current_stream()
    for step in range(2):
        cur, nxt = (pos_a, pos_b) if step % 2 == 0 else (pos_b, pos_a)
        nxt_ptrs = ptrs(nxt)
        for r in range(world):
            _lib.check(lib.pyqmd_cloud_pair_forces(cur[r].data_ptr(), ref.is_proton.data_ptr(), n, r, world,
                                                   S, Cc, P, acc[r].data_ptr(), ws[r].data_ptr(), st), "pair")
        for r in range(world):
            i0, i1 = shard_range(n, r, world)
            _lib.check(lib.pyqmd_cloud_exchange_integrate(cur[r].data_ptr(), vel[r].data_ptr(), None, n, i0, i1,
                                                          ref.dt, acc_ptrs.data_ptr(), nxt_ptrs.data_ptr(),
                                                          world, ws[r].data_ptr(), st), "exchange")
    final = pos_a                                         # two steps: the replicas are back in pos_a
    for r in range(world):
        assert torch.equal(final[r][:n], single.pos[:n]), r      # every replica, bit for bit
        assert int(acc[r].abs().max()) == 0                      # consumed entries were cleared
    i0, i1 = shard_range(n, 1, world)
    assert torch.equal(vel[1][i0:i1], single.vel[i0:i1])


# ---- opt-in exact-zero skipping (PYQMD_CLOUD_SKIP_EXACT_ZEROS) -----------------------------------------
@pytest.mark.parametrize("n", [60_001, 300_000])
def test_skipping_exact_zeros_is_bit_identical(n):
    """Beyond d = 353 the tail term underflows to exactly +0 in the kernel's FP32 arithmetic; skipping the
    exponential (and whole tiles without a p-p pair) there must not change a single bit of the step."""
    from pyqmd_b200.state import NucleonCloud
    pos, isp = make_cloud(n, seed=31)                 # radius 690 / 1545: plenty of tile pairs beyond 353
    a = NucleonCloud(pos, isp, keep_force=True)
    b = NucleonCloud(pos, isp, keep_force=True, skip_exact_zeros=True)
    a.step(1); b.step(1)
    assert torch.equal(a.force, b.force)              # (the accumulators themselves are consumed by the step)
    a.step(2); b.step(2)
    assert torch.equal(a.pos, b.pos) and torch.equal(a.vel, b.vel)


def test_skipping_exact_zeros_is_ignored_when_the_strengths_do_not_allow_it():
    from pyqmd_b200.state import NucleonCloud
    n = 60_001
    pos, isp = make_cloud(n, seed=32)
    st = (2000.0, 30.0, 35.0)                         # log2(0.15 S) = 8.2: the tail is not yet zero at 353
    a = NucleonCloud(pos, isp, strengths=st)
    b = NucleonCloud(pos, isp, strengths=st, skip_exact_zeros=True)
    a.step(2); b.step(2)
    assert torch.equal(a.pos, b.pos)
