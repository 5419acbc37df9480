"""Times the UNMODIFIED reference (baseline/_ref, see install_reference.py) on the host cores.

Measurement infrastructure for bench.py (`--impl reference` and the `cpu_baseline` leg), never on the
product path.  The reference is single-threaded Python (nuclear_forces.py:236-323 is a double `for`
loop), so "all host cores" means one independent process per core, each stepping its own bounded
sample of the workload with the reference's own functions:

  cloud     update_particles_cpu on an n_sub-nucleon cloud drawn like the benchmark's (PCG64, disc of
            number density 1/25, 40 % protons)
  ensemble  for each of m nuclei: Nucleus.should_decay(dt_decay) (particles.py:126-147) then
            update_particles_cpu(nucleus.particles, dt) -- the sub-step body nuclear_sim.py:165-173
  decay     decay_chains.Nucleus.should_decay (decay_chains.py:400-421) over m particle-less nuclei

Workers are plain subprocesses of this file (`python baseline/ref_arm.py --worker ...`, one per core,
driven over stdin/stdout: no fork of a process that may hold a CUDA context, no multiprocessing
start-method pitfalls); they import nothing but the reference, numpy and oracle.ref_loader (which stubs
pyopencl / pygame).
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_W = {}


def available():
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from oracle import ref_loader
    return ref_loader.available()


def _init(kind, size, isotopes, dt_decay, wid):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import random

    import numpy as np
    from oracle import ref_loader
    R = ref_loader.Ref()
    _W.update(kind=kind, R=R, nf=R.forces(), dt_decay=dt_decay)
    random.seed(1000 + wid)
    if kind == "cloud":
        rng = np.random.Generator(np.random.PCG64(1234 + wid))
        rad = np.sqrt(size * 25.0 / np.pi)
        r = rad * np.sqrt(rng.random(size))
        th = 2 * np.pi * rng.random(size)
        isp = np.zeros(size, np.uint8)
        isp[rng.permutation(size)[: int(round(0.4 * size))]] = 1
        x = (r * np.cos(th)).astype(np.float32)
        y = (r * np.sin(th)).astype(np.float32)
        _W["particles"] = R.make_particles(x, y, np.zeros(size), np.zeros(size), isp)
        _W["pairs"] = size * (size - 1)
    elif kind == "ensemble":
        nuclei = []
        for k in range(size):
            z, n = isotopes[(wid * size + k) % len(isotopes)]
            nuc = R.particles.Nucleus(z, n, 400, 400)            # reference layout, particles.py:62-124
            nuc.stability = R.decay_chains.get_half_life(z, n)   # nuclear_sim.py:116
            nuclei.append(nuc)
        _W["nuclei"] = nuclei
        _W["pairs"] = sum(len(a.particles) * (len(a.particles) - 1) for a in nuclei)
    else:
        T = 180825048000.0                                       # C-14, decay_chains.py HALF_LIVES
        nuclei = []
        for k in range(size):
            a = R.decay_chains.Nucleus(6, 8, 0, 0) if k % 2 == 0 else R.decay_chains.Nucleus(92, 146, 0, 0)
            nuclei.append(a)
        _W["nuclei"], _W["pairs"] = nuclei, 0
    return None


def _step(_):
    t0 = time.perf_counter()
    kind, nf = _W["kind"], _W["nf"]
    if kind == "cloud":
        nf.update_particles_cpu(_W["particles"], 1 / 240)
        units = 1
    elif kind == "ensemble":
        for nuc in _W["nuclei"]:
            nuc.should_decay(_W["dt_decay"])          # decision only: the sample keeps its size
            nf.update_particles_cpu(nuc.particles, 1 / 240)
        units = len(_W["nuclei"])
    else:
        dt = _W["dt_decay"]
        for nuc in _W["nuclei"]:
            nuc.should_decay(dt)
        units = len(_W["nuclei"])
    return _W["pairs"], units, time.perf_counter() - t0


def _worker_main(argv):
    kind, size, wid, dt_decay = argv[0], int(argv[1]), int(argv[2]), float(argv[3])
    isotopes = json.loads(argv[4])
    _init(kind, size, [tuple(v) for v in isotopes], dt_decay, wid)
    sys.stdout.write("ready\n")
    sys.stdout.flush()
    for line in sys.stdin:
        if line.strip() != "s":
            break
        pairs, units, sec = _step(None)
        sys.stdout.write(f"{pairs} {units} {sec!r}\n")
        sys.stdout.flush()


def run(kind, size, steps, warmup, procs=None, isotopes=((82, 126),), dt_decay=180825048000.0 * 1e-3,
        ready_timeout=300.0):
    """Returns dict(pairs_per_s, units_per_s, ms_per_step, cores, pairs_per_step, units_per_step)."""
    procs = procs or os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    workers = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", kind, str(size), str(w),
                                 repr(float(dt_decay)), json.dumps([list(v) for v in isotopes])],
                                stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, env=env)
               for w in range(procs)]
    try:
        for p in workers:
            line = p.stdout.readline()
            if line.strip() != "ready":
                raise RuntimeError("reference worker failed to start")

        def one_step():
            for p in workers:
                p.stdin.write("s\n")
                p.stdin.flush()
            out = [p.stdout.readline().split() for p in workers]
            return sum(int(o[0]) for o in out), sum(int(o[1]) for o in out)

        for _ in range(warmup):
            one_step()
        t0 = time.perf_counter()
        pairs = units = 0
        for _ in range(steps):
            a, b = one_step()
            pairs += a
            units += b
        sec = time.perf_counter() - t0
    finally:
        for p in workers:
            try:
                p.stdin.write("q\n")
                p.stdin.flush()
            except Exception:
                pass
        for p in workers:
            try:
                p.wait(timeout=10)
            except Exception:
                p.kill()
    return dict(pairs_per_s=pairs / sec, units_per_s=units / sec, ms_per_step=sec / steps * 1e3,
                cores=procs, pairs_per_step=pairs // max(steps, 1), units_per_step=units // max(steps, 1),
                seconds=sec)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        _worker_main(sys.argv[2:])
    else:
        print(json.dumps(run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]),
                             procs=int(sys.argv[5]) if len(sys.argv) > 5 else None)))
