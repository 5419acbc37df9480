"""The CPU oracle (oracle/) against the golden vectors produced by the unmodified reference
(tests/golden/gen_golden.py).  Bit-exact: the oracle restates the reference's float64
arithmetic operation by operation on the same libm."""
import math

import numpy as np
import pytest

from conftest import unhex
from oracle import decay_oracle as dor
from oracle import oracle as orc
from oracle import ref_loader


def _run_case(case):
    inp, out = case["input"], case["output"]
    x, y, vx, vy = (unhex(inp[k]).copy() for k in ("x", "y", "vx", "vy"))
    t = np.array(inp["is_proton"], np.uint8)
    for _ in range(case["steps"]):
        orc.force_step(x, y, vx, vy, t, float.fromhex(case["dt"]), float.fromhex(case["S"]),
                       float.fromhex(case["C"]), float.fromhex(case["P"]))
    return (x, y, vx, vy), tuple(unhex(out[k]) for k in ("x", "y", "vx", "vy"))


def test_force_kats_bit_exact(force_kats):
    assert len(force_kats["cases"]) >= 50
    for case in force_kats["cases"]:
        got, want = _run_case(case)
        for g, w in zip(got, want):
            assert np.array_equal(g, w), case["name"]


def test_survey_kat_values(force_kats):
    """The literal known answers listed in SURVEY.md section 4."""
    by = {c["name"]: c for c in force_kats["cases"]}
    a = by["A_pp_d3"]["output"]
    assert float.fromhex(a["x"][0]) == 0.00013862902538497969
    assert float.fromhex(a["vx"][0]) == 0.033270966092395125
    assert float.fromhex(a["x"][1]) == 2.999861370974615
    b = by["B_pn_d2"]["output"]
    assert float.fromhex(b["y"][0]) == -0.00017708333333333335
    assert float.fromhex(b["vy"][0]) == -0.0425
    assert float.fromhex(by["C_nn_d10"]["output"]["vx"][0]) == 0.00048599506919328627
    assert float.fromhex(by["C_pp_d10"]["output"]["vx"][0]) == -0.00015055011303337376
    d = by["D_pp_skip"]["output"]
    assert float.fromhex(d["vx"][0]) == 0.0 and float.fromhex(d["x"][1]) == 0.05
    assert float.fromhex(by["E_nn_containment"]["output"]["vx"][0]) == 0.004991220132282217


def test_u238_trajectory_teacher_forced(u238_traj):
    """From the reference's state at step s the oracle must land exactly on its state at s+1."""
    steps = list(u238_traj["steps"])
    states = u238_traj["states"]
    t = u238_traj["is_proton"]
    dt = float(u238_traj["dt"])
    for s in u238_traj["pair_steps"]:
        a = states[steps.index(s)].copy()
        b = states[steps.index(s + 1)]
        x, y, vx, vy = (np.ascontiguousarray(a[:, k]) for k in range(4))
        orc.force_step(x, y, vx, vy, t, dt)
        assert np.array_equal(np.stack([x, y, vx, vy], 1), b), f"step {s}"


def test_u238_trajectory_free_running(u238_traj):
    steps = list(u238_traj["steps"])
    st = u238_traj["states"]
    t = u238_traj["is_proton"]
    x, y, vx, vy = (np.ascontiguousarray(st[0][:, k]) for k in range(4))
    for s in range(1, 101):
        orc.force_step(x, y, vx, vy, t, float(u238_traj["dt"]))
        if s in steps:
            assert np.array_equal(np.stack([x, y, vx, vy], 1), st[steps.index(s)]), f"step {s}"


def test_force_step_empty_and_single():
    e = np.zeros(0)
    orc.force_step(e, e.copy(), e.copy(), e.copy(), np.zeros(0, np.uint8), 1 / 240)   # no-op
    x, y, vx, vy = np.array([1.0]), np.array([2.0]), np.array([0.5]), np.array([0.0])
    orc.force_step(x, y, vx, vy, np.array([1], np.uint8), 0.1)
    assert vx[0] == 0.5 * 0.85 and x[0] == 1.0 + 0.5 * 0.85 * 0.1


def test_branch_stats_and_flops():
    x = np.array([0.0, 3.0, 20.0]); y = np.zeros(3)
    r = orc.force_step(x, y, np.zeros(3), np.zeros(3), np.array([1, 1, 0], np.uint8), 1 / 240,
                       integrate=False, want_stats=True, want_forces=True)
    st = r["stats"].as_dict()
    assert st["evaluated"] == 6 and st["attr"] == 2 and st["tail"] == 4 and st["pp"] == 2
    assert st["hard"] == 2 and st["pauli"] == 2
    assert r["stats"].flops() == 15 * 6 + 7 * 2 + 8 * 4 + 5 * 2 + 3 * 2 + 6 * 2
    assert x[1] == 3.0     # integrate=False leaves the state alone


def test_ambiguity_flags():
    x = np.array([0.0, 9.0 * (1 + 1e-8), 30.0]); y = np.zeros(3)
    r = orc.force_step(x, y, np.zeros(3), np.zeros(3), np.zeros(3, np.uint8), 1 / 240,
                       integrate=False, amb_tol=1e-6)
    assert list(r["amb"]) == [True, True, False]


def test_decay_probabilities(decay_events):
    n = 0
    for row in decay_events["should_decay"]:
        T, dt = float.fromhex(row["T"]), float.fromhex(row["dt"])
        p = orc.decay_probability(T, dt)
        if not row["consumed"]:
            assert p == -1.0
            continue
        assert p == float.fromhex(row["p"]), (T, dt)
        for uh, dec, used in row["decisions"]:
            assert (float.fromhex(uh) < p) == bool(dec) and used == 1
        n += 1
    assert n >= 60
    # SURVEY.md section 4 values
    T = 180825048000.0
    assert orc.decay_probability(T, T * 1e-3).hex() == "0x1.6b54e2b063e07p-11"
    assert orc.decay_probability(T, 0.1 * T).hex() == "0x1.124bff742a770p-4"
    assert orc.decay_probability(T, T) == 0.5


def test_seeded_decision_strings(decay_events):
    for row in decay_events["seeded"]:
        T, dt = float.fromhex(row["T"]), float.fromhex(row["dt"])
        u = unhex(row["uniforms"])
        dec, consumed = orc.decay_decisions(np.full(len(u), T), dt, u)
        assert "".join("1" if d else "0" for d in dec) == row["bits"]
        assert consumed.all()
    first = decay_events["seeded"][0]
    assert first["seed"] == 12345 and first["bits"][:32] == "01000000000000000000000100001000"


def test_cpython_uniform_map():
    """random.random() == ((a >> 5) * 2**26 + (b >> 6)) / 2**53 on the MT19937 word stream."""
    import random
    r = random.Random(2024)
    st = random.Random(2024)
    for _ in range(100):
        a, b = st.getrandbits(32), st.getrandbits(32)
        assert orc.u53(a, b) == r.random()


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xFFFFFFFF
    assert orc.philox4x32_10([f, f, f, f], [f, f]) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                             [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_uniform_range_and_slots():
    u = orc.philox_uniforms(7, 0, 20000, 3, 0)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    assert orc.philox_uniform(7, 5, 3, 0) == u[5]
    assert orc.philox_uniform(7, 5, 3, 1) != u[5]


def test_decay_tables(decay_tables):
    rvals = [float.fromhex(h) for h in decay_tables["r_values"]]
    for z, n, hl, used, prods, pused in decay_tables["rows"]:
        if used:
            for u, want in zip((0.0, 0.5, 1.0), hl):
                got, c = dor.half_life(z, n, u)
                assert c == 1 and got == float.fromhex(want), (z, n, u)
        else:
            got, c = dor.half_life(z, n, 0.5)
            assert c == 0 and got == float.fromhex(hl[0]), (z, n)
        for k, r in enumerate(rvals):
            want = prods[k] if len(prods) > 1 else prods[0]
            nz, nn, mode, c = dor.decay_product(z, n, r)
            assert [nz, nn, -1 if mode is None else mode] == want and c == pused, (z, n, r)


def test_adjust_particles(decay_events):
    for row in decay_events["adjust"]:
        types = [dor.PROTON if t else dor.NEUTRON for t in row["types"]]
        new_types, removed, damp = dor.adjust_types(types, row["mode"])
        assert [int(t == dor.PROTON) for t in new_types] == row["out_types"]
        kept = [i for i in range(len(types)) if i not in removed]
        assert kept == row["out_index"]
        want_vx = unhex(row["out_vx"])
        got_vx = np.array([(1.0 + i) * (0.8 if damp else 1.0) for i in kept])
        assert np.array_equal(got_vx, want_vx)


def _nucleus_from(rec):
    return dor.OracleNucleus(rec["z"], rec["n"], unhex(rec["x"]), unhex(rec["y"]), rec["is_proton"],
                             unhex(rec["vx"]), unhex(rec["vy"]),
                             origin=(float.fromhex(rec["cx"]), float.fromhex(rec["cy"])),
                             T=float.fromhex(rec["T"]))


def _assert_state(nuc, rec, where):
    assert (nuc.z, nuc.n) == (rec["z"], rec["n"]), where
    assert nuc.T == float.fromhex(rec["T"]), where
    assert [int(t == dor.PROTON) for t in nuc.types] == rec["is_proton"], where
    for got, key in ((nuc.x, "x"), (nuc.y, "y"), (nuc.vx, "vx"), (nuc.vy, "vy")):
        assert np.array_equal(got, unhex(rec[key])), (where, key)
    assert (nuc.cx, nuc.cy) == (float.fromhex(rec["cx"]), float.fromhex(rec["cy"])), where


def test_chain_walks(decay_events):
    n_events = 0
    for walk in decay_events["walks"]:
        nuc = _nucleus_from(walk["start"])
        for k, ev in enumerate(walk["events"]):
            draws = [float.fromhex(h) for h in ev["draws"]]
            # the reference consumes its stream in order: branch?, angle?, half-life?
            has_branch = ev["n_opts"] > 1
            seq = list(draws)
            u1 = seq.pop(0) if has_branch else None
            nuc.emitted.clear()
            before = (nuc.z, nuc.n)
            opts = dor.decay_options(*before)
            mode_pre = opts[dor.pick_option(opts, u1)][2]
            u2 = seq.pop(0) if mode_pre in dor.EMISSION else None
            u3 = seq.pop(0) if seq else None
            mode, used = nuc.decay_event(u1, u2, u3)
            assert sum(used) == ev["used"], (walk["z"], walk["n"], k)
            assert (-1 if mode is None else mode) == ev["mode"]
            _assert_state(nuc, ev["after"], (walk["z"], walk["n"], k))
            assert len(nuc.emitted) == len(ev["emitted"])
            for got, want in zip(nuc.emitted, ev["emitted"]):
                assert got[0] == want[0]
                assert [got[1], got[2], got[3], got[4]] == [float.fromhex(h) for h in want[1:]]
            n_events += 1
    assert n_events >= 60


def test_substep_loops(decay_events):
    for loop in decay_events["loops"]:
        nuc = _nucleus_from(loop["start"])
        dt_decay, dt_phys = float.fromhex(loop["dt_decay"]), float.fromhex(loop["dt_phys"])
        for k, st in enumerate(loop["steps"]):
            draws = [float.fromhex(h) for h in st["draws"]]
            # gen_golden feeds slot 0 separately and slots 1.. as a stream (branch?, angle?, T?)
            p = orc.decay_probability(nuc.T, dt_decay)
            seq = draws[1:]
            slots = [draws[0], None, None, None]
            if p >= 0 and draws[0] < p:
                opts = dor.decay_options(nuc.z, nuc.n)
                if len(opts) > 1:
                    slots[1] = seq.pop(0)
                mode_pre = opts[dor.pick_option(opts, slots[1])][2]
                if mode_pre in dor.EMISSION:
                    slots[2] = seq.pop(0)
                slots[3] = seq.pop(0) if seq else None
            nuc.emitted.clear()
            decayed, mode, consumed = nuc.substep(dt_phys, dt_decay, slots)
            assert int(decayed) == st["decayed"] and consumed[0] == st["used0"], (loop["z"], k)
            assert sum(consumed[1:]) == st["used"]
            assert (-1 if mode is None else mode) == st["mode"]
            _assert_state(nuc, st["after"], (loop["z"], loop["n"], k))


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")
def test_live_against_reference():
    """When /root/reference is mounted (build container) re-pin the oracle live."""
    import random
    R = ref_loader.Ref()
    nf = R.forces(120.0, 25.0, 40.0)
    random.seed(11)
    nuc = R.particles.Nucleus(26, 30, 400, 400)
    ps = nuc.particles
    x = np.array([p.x for p in ps]); y = np.array([p.y for p in ps])
    vx = np.zeros(len(ps)); vy = np.zeros(len(ps))
    t = np.array([p.type == R.particles.ParticleType.PROTON for p in ps], np.uint8)
    for _ in range(5):
        nf.update_particles_cpu(ps, 1 / 120)
        orc.force_step(x, y, vx, vy, t, 1 / 120, 120.0, 25.0, 40.0)
    assert np.array_equal(x, np.array([p.x for p in ps]))
    assert np.array_equal(vy, np.array([p.vy for p in ps]))
    assert orc.py312_mean(x) == sum(p.x for p in ps) / len(ps)


def test_resolve_overlaps_bit_exact():
    """Per-frame projection nuclear_sim.py:355-379, incl. the degenerate pair and the frame loop
    (4 sub-steps + projection, :161-176)."""
    from conftest import load_json
    g = load_json("resolve_overlaps.json.gz")
    for case in g["cases"]:
        x, y = unhex(case["input"]["x"]).copy(), unhex(case["input"]["y"]).copy()
        used, pushes = orc.resolve_overlaps(x, y, [float.fromhex(h) for h in case["draws"]])
        assert used == case["used"], case["name"]
        assert np.array_equal(x, unhex(case["output"]["x"])), case["name"]
        assert np.array_equal(y, unhex(case["output"]["y"])), case["name"]
    f0 = g["frames"][0]
    x, y = unhex(f0["x"]).copy(), unhex(f0["y"]).copy()
    vx, vy = unhex(f0["vx"]).copy(), unhex(f0["vy"]).copy()
    t = np.array(f0["is_proton"], np.uint8)
    for k, fr in enumerate(g["frames"][1:]):
        for _ in range(g["substeps_per_frame"]):
            orc.force_step(x, y, vx, vy, t, float.fromhex(g["dt"]))
        orc.resolve_overlaps(x, y)
        assert np.array_equal(x, unhex(fr["x"])) and np.array_equal(vy, unhex(fr["vy"])), k
