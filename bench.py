#!/usr/bin/env python
"""bench.py -- PyQMD hot path on B200: pair interactions/s (+ nucleus-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload ensemble|mixed|cloud|decay] [--no-extras]

Contract (one JSON line from rank 0):
  metric   "pair interactions/s" -- ordered pairs N(N-1) per nucleus-step, no Newton-3 halving
           (nuclear_forces.py:248-251), whole job over all ranks
  step     one sub-step (decay test -> force -> integrate, nuclear_sim.py:165-173) of the whole
           per-GPU batch; the default workload is BASELINE.json configs[1]: 65,536 independent
           Pb-208 nuclei per GPU, one nucleus per thread block ("weak" scaling, sharded by
           nucleus, no data-path collective)
  value    device-resident throughput (state already in HBM), CUDA events, max over ranks
  e2e      same metric through the host-buffer API: pinned host state -> H2D -> kernel -> D2H
           every step
  roofline dominant kernel vs the FP32 FMA peak measured in this run (FFMA-chain
           microbenchmark; MEASURED_PEAKS.json has no FP32 figure), algorithmic FLOPs per
           pair by SURVEY.md section 8(d)'s convention from a device-side branch census of the state
  cpu_baseline  the oracle's C port (OpenMP, all host cores) on a bounded sample, rank 0, N=1
  also     (unless --no-extras) the single-cloud workload of configs[3] (N = 1M nucleons,
           i-block sharded + position all-gather, "strong" scaling) measured in the same run

--impl reference times the CPU port of the reference's path (the reference itself is pure
Python and cannot travel to the GPU box; the C oracle is bit-identical to it and ~35x faster,
i.e. a conservative baseline) with all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PB208 = (82, 126)
N_ENSEMBLE = 65536
N_MIXED = 1_000_000
N_CLOUD = 1_000_000
N_DECAY = 100_000_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ensemble", choices=["ensemble", "mixed", "cloud", "decay"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--nuclei", type=int, default=0, help="override nuclei per GPU (ensemble/mixed)")
    ap.add_argument("--cloud-n", type=int, default=N_CLOUD)
    ap.add_argument("--cloud-exchange", default="peer", choices=["peer", "nccl"],
                    help="symmetric scheme on several GPUs: fused peer-memory kernel, or NCCL collectives")
    ap.add_argument("--cloud-scheme", default="symmetric", choices=["symmetric", "ordered"],
                    help="symmetric: every unordered pair once + integer force reduce-scatter; "
                         "ordered: i-block rows x all j, position all-gather only")
    ap.add_argument("--substeps", type=int, default=1, help="sub-steps fused per step call")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def sample_once(self):
        if self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            except Exception:
                pass

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[kernel]["bytes"]
    except Exception:
        return None


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------------
def cpu_ensemble_sample(isotopes, n_nuclei, n_steps, threads):
    """(pairs/s, nucleus-steps/s, seconds) of the oracle port on n_nuclei template nuclei."""
    from oracle import oracle as orc
    from pyqmd_b200.state import layout_templates
    tm = layout_templates()
    xs, ys, ts, cnt = [], [], [], []
    for k in range(n_nuclei):
        z, n = isotopes[k % len(isotopes)]
        xy = tm[f"z{z}_n{n}_xy"][(k // len(isotopes)) % 64]
        xs.append(xy[:, 0].astype(np.float64)); ys.append(xy[:, 1].astype(np.float64))
        ts.append(tm[f"z{z}_n{n}_isp"][(k // len(isotopes)) % 64]); cnt.append(z + n)
    x, y, t = np.concatenate(xs), np.concatenate(ys), np.concatenate(ts)
    cnt = np.array(cnt, np.int32)
    off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
    vx, vy = np.zeros_like(x), np.zeros_like(x)
    t0 = time.perf_counter()
    pairs = orc.ensemble_force_steps(off, cnt, x, y, vx, vy, t, 1 / 240, n_steps, n_threads=threads)
    dt = time.perf_counter() - t0
    return pairs / dt, n_nuclei * n_steps / dt, dt


def cpu_cloud_sample(n, n_i, threads, seed=1234):
    from oracle import oracle as orc
    n_i = min(n_i, n)
    pos, isp = make_cloud(n, seed)
    x, y = pos[:, 0].astype(np.float64), pos[:, 1].astype(np.float64)
    t0 = time.perf_counter()
    orc.cloud_forces(x, y, isp, 0, n_i, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_i * (n - 1) / dt, dt


def make_cloud(n, seed=1234, frac_p=0.4, density=1 / 25):
    """uniform disc of number density 1/25, 40 % protons, PCG64(1234)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R = np.sqrt(n / density / np.pi)
    r = R * np.sqrt(rng.random(n))
    th = 2 * np.pi * rng.random(n)
    pos = np.stack([r * np.cos(th), r * np.sin(th)], 1).astype(np.float32)
    isp = np.zeros(n, np.uint8)
    isp[rng.permutation(n)[: int(round(frac_p * n))]] = 1
    return pos, isp


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the CPU port of the reference's path on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from pyqmd_b200.state import README_ISOTOPES
    threads = orc.max_threads()
    if args.workload == "cloud":
        n_i = 64 * threads
        vals = []
        for s in range(args.warmup + args.steps):
            v, dt = cpu_cloud_sample(65536, n_i, threads)
            if s >= args.warmup:
                vals.append((v, dt))
        value = float(np.mean([v for v, _ in vals]))
        ms = float(np.mean([d for _, d in vals])) * 1e3
        sample = f"{n_i} i-nucleons x 65,535 partners of a 65,536-nucleon PCG64(1234) cloud per step"
        cfg = {"workload": f"C4 single 2-D nucleon cloud N={args.cloud_n} (40% protons), all-pairs, "
                           f"i-block sharded x{args.gpus}"}
        extra = {}
    else:
        isotopes = README_ISOTOPES if args.workload == "mixed" else (PB208,)
        n_nuc = 64 * threads
        vals = []
        for s in range(args.warmup + args.steps):
            v, ns, dt = cpu_ensemble_sample(isotopes, n_nuc, 1, threads)
            if s >= args.warmup:
                vals.append((v, ns, dt))
        value = float(np.mean([v for v, _, _ in vals]))
        ms = float(np.mean([d for _, _, d in vals])) * 1e3
        sample = f"{n_nuc} reference-layout nuclei x 1 sub-step per step"
        cfg = {"workload": ("C3 mixed ensemble of 1M nuclei over the nine preset isotopes, decay on, "
                            "sharded by nucleus" if args.workload == "mixed" else
                            "C2 ensemble of 65,536 independent Pb-208 nuclei per GPU, one nucleus per "
                            "thread block")}
        extra = {"nucleus_steps_per_s": float(np.mean([n for _, n, _ in vals]))}
    line = {
        "impl": "reference", "metric": "pair interactions/s", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (reference-generated initial layouts)", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def timed_steps(fn, steps, warmup, dist, torch, sampler=None):
    """W warm-up calls, then K timed calls bracketed by barrier + synchronize; CUDA events on the
    current stream; returns seconds (max over ranks)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if sampler is not None:
        sampler.sample_once()         # the queue is still draining: a sample under load even for short runs
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    if dist is not None:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from pyqmd_b200 import _lib
    from pyqmd_b200.state import (README_ISOTOPES, DecayPopulation, HostEnsembleRunner,
                                  NucleonCloud, NucleusEnsemble)
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()                   # no CPU fallback: fail loudly before anything else
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.require_cuda()
    lib = _lib.lib()
    dev = f"cuda:{local_rank}"
    K, W = args.steps, max(args.warmup, 0)

    # FP32 FMA peak of this GPU, measured now (roofline denominator)
    f1, f2 = C.c_double(), C.c_double()
    _lib.check(lib.pyqmd_fp32_peak(4096, C.byref(f1), C.byref(f2), _lib.current_stream()), "fp32_peak")
    props = (C.c_int64 * 8)()
    lib.pyqmd_device_props(local_rank, props)
    fp32_peak = max(f1.value, f2.value)
    nominal = props[0] * 128 * 2 * props[3] * 1e3 / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    sampler = ClockSampler(physical_gpu_index(local_rank))
    line = {"metric": "pair interactions/s", "unit": "pairs/s", "n_gpus": world, "steps": K,
            "warmup": W, "higher_is_better": True, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    launches = 0

    ens = tot_nuc = tot_pairs = None
    if args.workload in ("ensemble", "mixed"):
        res, ens, tot_nuc, tot_pairs = bench_ensemble(args, args.workload, K, W, rank, world, dev, dist,
                                                      torch, sampler, fp32_peak,
                                                      (f1.value, f2.value, nominal), hbm_peak, hbm_src,
                                                      with_e2e=True)
        line.update(res)
        clocks = sampler.result()
        launches = K * len(ens.bins)
    elif args.workload == "cloud":
        line.update(bench_cloud(args, K, W, rank, world, dev, dist, torch, sampler, fp32_peak,
                                (f1.value, f2.value, nominal)))
        clocks = sampler.result()
        launches = K * 4
    else:
        line.update(bench_decay(args, K, W, rank, world, dev, dist, torch, sampler, hbm_peak, hbm_src))
        clocks = sampler.result()
        launches = K

    line["clocks"] = clocks
    line["gpu_launches"] = launches

    if rank == 0 and world == 1:
        from oracle import oracle as orc
        threads = orc.max_threads()
        if args.workload == "cloud":
            v, dt = cpu_cloud_sample(65536, 4096 * threads, threads)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                                    "sample": f"{min(4096 * threads, 65536)} i-nucleons of a 65,536-nucleon cloud, "
                                              f"{dt:.1f} s"}
        elif args.workload in ("ensemble", "mixed"):
            isotopes = (PB208,) if args.workload == "ensemble" else README_ISOTOPES
            n_s = 256 * threads
            v, ns, dt = cpu_ensemble_sample(isotopes, n_s, 30, threads)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                                    "nucleus_steps_per_s": ns,
                                    "sample": f"{n_s} nuclei x 30 sub-steps of the same workload, "
                                              f"{dt:.1f} s, OpenMP over nuclei"}

    if not args.no_extras and args.workload == "ensemble":
        # every other config of BASELINE.json in the same line; a failure here must not cost the headline
        try:
            # app-faithful frame (nuclear_sim.py:161-176): 4 sub-steps, then the overlap projection, from a
            # settled state (fresh layouts + 12 frames; the step-only state above has collapsed)
            from pyqmd_b200.state import NucleusEnsemble
            n_f = ens.n_nuclei
            del ens
            torch.cuda.empty_cache()
            ens = NucleusEnsemble.from_templates((PB208,), n_f, device=dev, id_base=rank * n_f, decay=False)
            for _ in range(12):
                ens.frame(4)
            p0 = int(ens.push_count.item())
            sec_f = timed_steps(lambda: ens.frame(4), 10, 1, dist, torch)
            pushes = (int(ens.push_count.item()) - p0) / 11 / n_f
            sec_s = timed_steps(lambda: ens.step(4), 10, 1, dist, torch)
            census_f, flops_f = ens.census(range(0, n_f, max(1, n_f // 256)))
            frame = {"ms_per_frame": sec_f / 10 * 1e3, "substeps_per_frame": 4,
                     "ms_4_substeps_alone": sec_s / 10 * 1e3,
                     "nucleus_frames_per_s": float(tot_nuc.item()) * 10 / sec_f,
                     "pairs_per_s": float(tot_pairs.item()) / args.substeps * 4 * 10 / sec_f,
                     "pushes_per_nucleus_frame": pushes, "flops_per_pair": flops_f,
                     "branch_census": census_f,
                     "note": "4 sub-steps + resolve_overlaps per frame, device resident, settled nuclei"}
            del ens
            torch.cuda.empty_cache()
            keep = ("metric", "unit", "value", "ms_per_step", "config", "roofline", "scaling",
                    "nucleus_steps_per_s", "decays")
            also = {"frame": frame}
            extra = bench_cloud(args, 2, 1, rank, world, dev, dist, torch, None, fp32_peak,
                                (f1.value, f2.value, nominal))
            also["cloud"] = {k: extra[k] for k in keep if k in extra}
            torch.cuda.empty_cache()
            extra, ens3, _, _ = bench_ensemble(args, "mixed", 5, 3, rank, world, dev, dist, torch, None,
                                               fp32_peak, (f1.value, f2.value, nominal), hbm_peak, hbm_src,
                                               with_e2e=False)
            also["mixed"] = {k: extra[k] for k in keep if k in extra}
            del ens3
            torch.cuda.empty_cache()
            extra = bench_decay(args, 3, 3, rank, world, dev, dist, torch, None, hbm_peak, hbm_src)
            also["decay"] = {k: extra[k] for k in keep if k in extra}
            torch.cuda.empty_cache()
            if rank == 0:
                also["c1"] = bench_c1(dev, torch)
            if rank == 0 and world == 1:
                # the reference's CPU path (C port, all host threads) beside each config
                from oracle import oracle as orc
                threads = orc.max_threads()
                v, dt_s = cpu_cloud_sample(65536, 1024 * threads, threads)
                also["cloud"]["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                                                 "sample": f"{1024 * threads} i-nucleons x 65,535 partners, {dt_s:.1f} s"}
                v, ns, dt_s = cpu_ensemble_sample(README_ISOTOPES, 9 * 32 * threads, 10, threads)
                also["mixed"]["cpu_baseline"] = {"value": v, "unit": "pairs/s", "nucleus_steps_per_s": ns,
                                                 "cores": threads, "kind": "port",
                                                 "sample": f"{9 * 32 * threads} nuclei x 10 sub-steps, {dt_s:.1f} s"}
                n_d = 4_000_000 * threads
                Tn = np.full(n_d, 180825048000.0)
                un = np.random.default_rng(1).random(n_d)
                t0 = time.perf_counter()
                orc.decay_decisions(Tn, 180825048000.0 * 1e-3, un, n_threads=threads)
                dt_s = time.perf_counter() - t0
                also["decay"]["cpu_baseline"] = {"value": n_d / dt_s, "unit": "nucleus-steps/s", "cores": threads,
                                                 "kind": "port",
                                                 "sample": f"{n_d} should_decay decisions with supplied uniforms, {dt_s:.2f} s"}
            line["also"] = also
        except Exception as exc:      # noqa: BLE001
            import traceback
            traceback.print_exc()
            line.setdefault("also", {})["error"] = repr(exc)[:300]

    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def bench_ensemble(args, workload, K, W, rank, world, dev, dist, torch, sampler, fp32_peak, peak_info,
                   hbm_peak, hbm_src, with_e2e):
    """C2 (``ensemble``: 65,536 Pb-208 per GPU, weak) or C3 (``mixed``: 1M nuclei over the nine
    preset isotopes in total, decay on, strong)."""
    from pyqmd_b200.state import README_ISOTOPES, HostEnsembleRunner, NucleusEnsemble
    isotopes = (PB208,) if workload == "ensemble" else README_ISOTOPES
    total = args.nuclei * world if args.nuclei else (N_ENSEMBLE * world if workload == "ensemble"
                                                     else N_MIXED)
    per = (total + world - 1) // world
    lo = rank * per
    n_mine = max(0, min(per, total - lo))
    decay = workload == "mixed"
    ens = NucleusEnsemble.from_templates(isotopes, n_mine, device=dev, id_base=lo, decay=decay,
                                         dt_decay=180825048000.0 * 1e-3, seed=2024)
    pairs_step = ens.pairs_per_step() * args.substeps
    nucleons = int(ens.count.sum().item())
    import math
    stride = max(1, ens.n_nuclei // 256)
    while math.gcd(stride, len(isotopes)) != 1:      # isotopes cycle with the nucleus id: hit them all
        stride += 1
    sample = range(0, ens.n_nuclei, stride)
    flops0 = ens.census(sample)[1]
    sec = timed_steps(lambda: ens.step(args.substeps), K, W, dist, torch, sampler)
    census1, flops1 = ens.census(sample)
    tot_pairs = torch.tensor([float(pairs_step)], device=dev, dtype=torch.float64)
    tot_nuc = torch.tensor([float(ens.n_nuclei)], device=dev, dtype=torch.float64)
    decays = ens.mode_counts.clone()
    if dist is not None:
        dist.all_reduce(tot_pairs); dist.all_reduce(tot_nuc); dist.all_reduce(decays)
    value = float(tot_pairs.item()) * K / sec
    flops_pair = 0.5 * (flops0 + flops1)
    my_rate = pairs_step * K / sec          # this rank's kernel
    res = {
        "value": value, "ms_per_step": sec / K * 1e3,
        "scaling": "weak" if workload == "ensemble" and not args.nuclei else "strong",
        "config": {"workload": ("C2 ensemble of 65,536 independent Pb-208 nuclei per GPU, one "
                                "nucleus per thread block" if workload == "ensemble" else
                                "C3 mixed ensemble of 1M nuclei over the nine preset isotopes, "
                                "decay on, sharded by nucleus"),
                   "nuclei_total": int(tot_nuc.item()), "nucleons_per_gpu": nucleons,
                   "substeps_per_step": args.substeps, "dt_phys": 1 / 240,
                   "l2_policy": "inputs larger than L2 (state %.0f MB per GPU)" % (nucleons * 17 / 1e6),
                   "parallelism": f"by-nucleus x{world}, no collective"},
        "nucleus_steps_per_s": float(tot_nuc.item()) * args.substeps * K / sec,
        "roofline": {"bound": "fp32", "achieved": my_rate * flops_pair / 1e12, "peak": fp32_peak,
                     "unit": "TFLOP/s", "frac": my_rate * flops_pair / 1e12 / fp32_peak,
                     "traffic": ncu_traffic("ensemble_pair_kernel") if workload == "ensemble"
                     and not args.nuclei and args.substeps == 1 else None,
                     "kernel": "ensemble_pair_kernel",
                     "flops_per_pair": flops_pair, "branch_census_end": census1,
                     "peak_source": "FFMA-chain microbenchmark in this run (scalar %.1f, "
                                    "f32x2 %.1f TFLOP/s); nominal %.1f" % peak_info,
                     "hbm_gbs": nucleons * 36 * args.substeps * K / sec / 1e9 / max(args.substeps, 1),
                     "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src},
    }
    if decay:
        res["decays"] = {"total_events": int(decays.sum().item()),
                         "by_mode": [int(v) for v in decays.tolist()]}
    if with_e2e:
        # e2e: host-buffer API, H2D + kernel + D2H every step
        runner = HostEnsembleRunner(ens, chunks=8)
        k2 = max(3, K // 4)
        sec_e2e = timed_steps(lambda: runner.step(args.substeps), k2, 2, dist, torch)
        res["e2e"] = {"value": float(tot_pairs.item()) * k2 / sec_e2e, "unit": "pairs/s",
                      "h2d_bytes_per_step": runner.h2d_bytes, "d2h_bytes_per_step": runner.d2h_bytes,
                      "chunks": runner.n_chunks}
        # for information: one host round trip per app frame (4 sub-steps, nuclear_sim.py:153) instead of
        # one per sub-step -- what NuclearForces.step(particles, dt, n) offers over the reference's call
        sec_e2e4 = timed_steps(lambda: runner.step(4 * args.substeps), max(3, k2 // 4), 1, dist, torch)
        res["e2e"]["value_4_substeps_per_call"] = (float(tot_pairs.item()) * 4 * max(3, k2 // 4) / sec_e2e4)
        del runner
    return res, ens, tot_nuc, tot_pairs


def bench_c1(dev, torch):
    """C1: one U-238 nucleus (configs[0]), 1000 force+integrate(+decay test) steps.
    (a) through the reference-shaped drop-in NuclearForces.update_particles_cpu(list[Particle], dt),
        one host round trip per sub-step, exactly how nuclear_sim.py:171-173 calls it;
    (b) the same 1000 sub-steps fused into one call (NuclearForces.step);
    (c) device resident (NucleusEnsemble of one nucleus, 1000 sub-steps in one launch, decay on)."""
    import random
    from pyqmd_b200.forces import NuclearForces
    from pyqmd_b200.state import NucleusEnsemble, layout_templates
    from pyqmd_b200.types import Particle, ParticleType
    tm = layout_templates()
    xy, isp = tm["z92_n146_xy"][0], tm["z92_n146_isp"][0]
    mk = lambda: [Particle(400.0 + float(x), 400.0 + float(y),
                           ParticleType.PROTON if t else ParticleType.NEUTRON) for (x, y), t in zip(xy, isp)]
    nf = NuclearForces()
    pairs = 238 * 237
    ps = mk()
    for _ in range(5):
        nf.update_particles_cpu(ps, 1 / 240)
    t0 = time.perf_counter()
    calls = 200
    for _ in range(calls):
        nf.update_particles_cpu(ps, 1 / 240)
    per_call = (time.perf_counter() - t0) / calls
    ps = mk()
    nf.step(ps, 1 / 240, 10)
    t0 = time.perf_counter()
    nf.step(ps, 1 / 240, 1000)
    fused = time.perf_counter() - t0
    ens = NucleusEnsemble.from_templates(((92, 146),), 1, device=dev, decay=True,
                                         dt_decay=1.409993568e17 * 1e-3, seed=1)
    ens.step(10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ens.step(1000)
    e1.record()
    torch.cuda.synchronize()
    resident = e0.elapsed_time(e1) * 1e-3
    from oracle import oracle as orc
    x, y = xy[:, 0].astype(np.float64), xy[:, 1].astype(np.float64)
    t0 = time.perf_counter()
    orc.ensemble_force_steps(np.array([0], np.int64), np.array([238], np.int32), x, y, np.zeros(238),
                             np.zeros(238), isp, 1 / 240, 200, n_threads=1)
    cpu = (time.perf_counter() - t0) / 200
    return {"config": {"workload": "C1 single U-238 nucleus (238 nucleons), 1000 sub-steps"},
            "cpu_port_1core": {"ms_per_substep": cpu * 1e3, "pairs_per_s": pairs / cpu,
                               "note": "C oracle port, one core (one nucleus does not parallelise)"},
            "dropin_per_substep_call": {"ms_per_call": per_call * 1e3, "pairs_per_s": pairs / per_call,
                                        "api": "NuclearForces.update_particles_cpu(list[Particle], dt)"},
            "dropin_fused_1000": {"ms_total": fused * 1e3, "pairs_per_s": pairs * 1000 / fused,
                                  "api": "NuclearForces.step(list[Particle], dt, 1000)"},
            "device_resident_1000": {"ms_total": resident * 1e3, "pairs_per_s": pairs * 1000 / resident,
                                     "nucleus_steps_per_s": 1000 / resident,
                                     "api": "NucleusEnsemble.step(1000), decay test on"},
            "reference_python_cpu": "75-91 ms per sub-step (SURVEY.md section 6 [probe]), 6.2-7.5e5 pairs/s"}


def bench_cloud(args, K, W, rank, world, dev, dist, torch, sampler, fp32_peak, peak_info):
    from pyqmd_b200.state import NucleonCloud
    n = args.cloud_n
    pos, isp = make_cloud(n)
    cloud = NucleonCloud(pos, isp, device=dev, rank=rank, world=world, scheme=args.cloud_scheme,
                         exchange=args.cloud_exchange)
    sec = timed_steps(lambda: cloud.step(1), K, W, dist, torch, sampler)
    pairs = float(n) * (n - 1)
    f_pp = (float(isp.sum()) / n) ** 2
    flops_pair = 23.0 + 3.0 * f_pp
    mine = cloud.pairs_per_step() * K / sec
    return {
        "value": pairs * K / sec, "ms_per_step": sec / K * 1e3, "scaling": "strong",
        "config": {"workload": f"C4 single 2-D nucleon cloud N={n} (40% protons), all-pairs, "
                               f"i-block sharded x{world}",
                   "l2_policy": "per-step working set (positions 8N B) is L2 resident by design; "
                                "compute bound", "dt_phys": 1 / 240,
                   "scheme": args.cloud_scheme, "exchange": cloud.exchange,
                   "parallelism": (f"i-block rows dealt x{world}; integer force reduce-scatter + integrate + "
                                   f"position all-gather " + ("fused in one peer-memory kernel (NVLink)"
                                                              if cloud.exchange == "peer" else "via NCCL")
                                   if args.cloud_scheme == "symmetric" and world > 1 else
                                   f"i-block x{world}" + (" + NCCL position all-gather" if world > 1 else ""))},
        "roofline": {"bound": "fp32", "achieved": mine * flops_pair / 1e12, "peak": fp32_peak,
                     "unit": "TFLOP/s", "frac": mine * flops_pair / 1e12 / fp32_peak,
                     "traffic": None,   # ncu capture is at N = 262,144 (profiles/traffic.json): 6.9 MB per launch
                     "kernel": "cloud_sym_kernel" if args.cloud_scheme == "symmetric" else "cloud_force_kernel",
                     "flops_per_pair": flops_pair,
                     "executed_pair_evaluations_per_step": pairs / 2 if args.cloud_scheme == "symmetric" else pairs,
                     "note": ("algorithmic FLOPs of all N(N-1) ordered pairs (SURVEY 8d) over time; the "
                              "symmetric scheme evaluates each unordered pair once, so frac may exceed 1"
                              if args.cloud_scheme == "symmetric" else "ordered pairs, each evaluated"),
                     "peak_source": "FFMA-chain microbenchmark in this run (scalar %.1f, f32x2 %.1f "
                                    "TFLOP/s); nominal %.1f" % peak_info},
    }


def bench_decay(args, K, W, rank, world, dev, dist, torch, sampler, hbm_peak, hbm_src):
    from pyqmd_b200.state import DecayPopulation
    total = N_DECAY
    per = (total + world - 1) // world
    lo = rank * per
    n_mine = max(0, min(per, total - lo))
    zn = torch.full((n_mine,), (6 << 16) | 8, dtype=torch.int32)
    zn[n_mine // 2:] = (92 << 16) | 146
    pop = DecayPopulation(zn, device=dev, dt_decay=180825048000.0 * 1e-3, seed=7, id_base=lo,
                          watch=((6, 8), (92, 146)))
    sub = max(args.substeps, 1)
    sec = timed_steps(lambda: pop.step(sub), K, W, dist, torch, sampler)
    return {
        "metric": "nucleus-steps/s", "unit": "nucleus-steps/s", "value": float(total) * sub * K / sec,
        "ms_per_step": sec / K * 1e3, "scaling": "strong", "dtype": "f64",
        "config": {"workload": "C5 decay-only Monte Carlo, 1e8 C-14 / U-238 nuclei, Philox draws",
                   "substeps_per_step": sub, "parallelism": f"by-nucleus x{world}"},
        "roofline": {"bound": "hbm", "achieved": n_mine * 20.0 * K / sec / 1e9, "peak": hbm_peak,
                     "unit": "GB/s", "frac": n_mine * 20.0 * K / sec / 1e9 / hbm_peak, "traffic": None,
                     "kernel": "population_kernel", "peak_source": hbm_src,
                     "bytes_per_nucleus_launch": 20,
                     "note": "read zn + half-life + p (20 B); written back only for decayed nuclei"},
    }


if __name__ == "__main__":
    main()
