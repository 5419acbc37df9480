#!/usr/bin/env python
"""bench.py -- PyQMD hot path on B200: pair interactions/s (+ nucleus-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cloud|ensemble|mixed|decay] [--no-extras] [--no-cpu]

Contract (one JSON line from rank 0):
  metric   "pair interactions/s" -- ordered pairs N(N-1) per step, no Newton-3 halving
           (nuclear_forces.py:248-251), whole job over all ranks
  step     one Jacobi step (force -> containment -> damped Euler, nuclear_forces.py:236-323) of the
           default workload, BASELINE.json configs[3]: ONE cloud of N = 1,000,000 nucleons (40 %
           protons), the configuration both numeric targets of north_star are quoted on and the
           largest single-GPU config.  "strong" scaling: the cloud is split over the ranks (i-block
           rows of the symmetric scheme dealt to the ranks; per step an integer reduce-scatter of
           the force accumulators + integrate + all-gather of positions, fused into ONE
           peer-memory kernel over NVLink -- `details.exchange_used` says which exchange ran; if
           symmetric memory cannot be set up the bench falls back to NCCL collectives and says so)
  value    device-resident throughput (state already in HBM), CUDA events, max over ranks
  e2e      the same metric through the host-buffer API, state in pinned HOST memory, copies inside
           the timed region every step.  1 GPU: NuclearForces.step_cloud -> pyqmd_cloud_step_host
           (upload, sort, step, un-sort, download -- the reference's per-step call
           nuclear_forces.py:185-234 at N = 1M).  N GPUs: every rank uploads its block, positions
           are all-gathered, the step runs, every rank downloads its block.
  roofline dominant kernel (cloud_sym_kernel) vs the FP32 FMA peak measured in this run (FFMA-chain
           microbenchmark; MEASURED_PEAKS.json has no FP32 figure), algorithmic FLOPs per ordered
           pair by SURVEY.md section 8(d)'s convention; `frac_executed` counts only the pair
           evaluations the kernel really executes (half: Newton's third law)
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, kind "reference") on the host cores, one
           process per core, bounded sample; the C/OpenMP port of the oracle beside it
  also     (unless --no-extras) every other BASELINE config in the same run -- C2 ensemble
           (free-running AND settled state), C3 mixed, C5 decay, C1, the app frame -- each with its
           own warm-up, >= 10 timed steps and its own clock record
  summary  LAST key: one number per config, so that a truncated tail still carries them

--impl reference times the reference's own CPU implementation (baseline/_ref: the unmodified
update_particles_cpu, nuclear_forces.py:236-323) on all host cores, on a bounded sample of the same
workload (one independent sub-cloud per core).  Rank 0 only.
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
README_ISO = ((1, 0), (2, 2), (6, 6), (6, 8), (26, 30), (47, 60), (79, 118), (82, 126), (92, 146))
N_ENSEMBLE = 65536
N_MIXED = 1_000_000
N_CLOUD = 1_000_000
N_DECAY = 100_000_000
T_C14 = 180825048000.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cloud", choices=["cloud", "ensemble", "mixed", "decay"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs (tuning runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (tuning runs)")
    ap.add_argument("--nuclei", type=int, default=0, help="override nuclei per GPU (ensemble/mixed)")
    ap.add_argument("--cloud-n", type=int, default=N_CLOUD)
    ap.add_argument("--cloud-exchange", default="peer", choices=["peer", "nccl"],
                    help="symmetric scheme on several GPUs: fused peer-memory kernel, or NCCL collectives")
    ap.add_argument("--cloud-scheme", default="symmetric", choices=["symmetric", "ordered"],
                    help="symmetric: every unordered pair once + integer force reduce-scatter; "
                         "ordered: i-block rows x all j, position all-gather only")
    ap.add_argument("--substeps", type=int, default=1, help="sub-steps fused per step call")
    ap.add_argument("--isotope", default="", help="Z,N of the ensemble workload (tuning runs; default Pb-208)")
    ap.add_argument("--settled", action="store_true",
                    help="ensemble workload from the settled (app-faithful) state instead of free-running")
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
            time.sleep(0.02)

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


def workload_config(args):
    """The `config` object, identical in both arms (ours / --impl reference)."""
    g = args.gpus
    if args.workload == "cloud":
        return {"workload": f"C4 single 2-D nucleon cloud N={args.cloud_n} (40% protons), all-pairs "
                            f"strong+Coulomb+Pauli, one Jacobi step per step",
                "n_nucleons": args.cloud_n, "dt_phys": 1 / 240, "gpus": g,
                "parallelism": f"one cloud split over {g} GPU(s) (strong scaling)",
                "l2_policy": "L2 flushed before every timed step (inside the timed region)",
                # the split this arm was asked for (the reference arm has no GPU side: same words, so that
                # both arms describe one configuration)
                "scheme": args.cloud_scheme,
                "exchange": "none (1 GPU)" if g == 1 else (args.cloud_exchange if args.cloud_scheme == "symmetric"
                                                           else "nccl position all-gather")}
    if args.workload == "ensemble":
        return {"workload": "C2 ensemble of 65,536 independent Pb-208 nuclei per GPU, one sub-step "
                            "(force, integrate) per step",
                "nuclei_per_gpu": args.nuclei or N_ENSEMBLE, "dt_phys": 1 / 240, "gpus": g,
                "parallelism": f"by nucleus over {g} GPU(s), no collective (weak scaling)",
                "l2_policy": "inputs larger than L2 (232 MB of state per GPU)"}
    if args.workload == "mixed":
        return {"workload": "C3 mixed ensemble of 1M nuclei over the nine preset isotopes, decay on, "
                            "one sub-step per step", "nuclei_total": N_MIXED, "dt_phys": 1 / 240, "gpus": g,
                "parallelism": f"by nucleus over {g} GPU(s), no collective (strong scaling)",
                "l2_policy": "inputs larger than L2 (1.6 GB of state in total)"}
    return {"workload": "C5 decay-only Monte Carlo, 1e8 C-14 / U-238 nuclei, one should_decay per nucleus "
                        "per step", "nuclei_total": N_DECAY, "gpus": g,
            "parallelism": f"by nucleus over {g} GPU(s), no collective (strong scaling)",
            "l2_policy": "inputs larger than L2 up to 2 GPUs (400 MB in total); beyond, L2 flushed between the "
                         "individually timed steps"}


# ---- CPU legs: the oracle's C port (OpenMP) and the unmodified reference (baseline/_ref) ------------
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


REF_SIZES = {"cloud": 640, "ensemble": 2, "mixed": 9, "decay": 400_000}


def reference_leg(workload, steps, warmup):
    """The unmodified reference on all host cores (baseline/ref_arm.py); None if baseline/_ref is absent."""
    from baseline import ref_arm
    if not ref_arm.available():
        return None
    kind = {"cloud": "cloud", "ensemble": "ensemble", "mixed": "ensemble", "decay": "decay"}[workload]
    iso = (PB208,) if workload == "ensemble" else README_ISO
    r = ref_arm.run(kind, REF_SIZES[workload], steps, warmup, isotopes=iso, dt_decay=T_C14 * 1e-3)
    what = {"cloud": f"{r['cores']} independent {REF_SIZES['cloud']}-nucleon sub-clouds (same generator and "
                     f"density as the 1M cloud), one update_particles_cpu each per step",
            "ensemble": f"{r['cores']} x {REF_SIZES['ensemble']} reference-built Pb-208 nuclei, should_decay + "
                        f"update_particles_cpu each per step",
            "mixed": f"{r['cores']} x {REF_SIZES['mixed']} reference-built nuclei over the nine preset isotopes, "
                     f"should_decay + update_particles_cpu each per step",
            "decay": f"{r['cores']} x {REF_SIZES['decay']} decay_chains.Nucleus.should_decay calls per step"}[workload]
    r["sample"] = what + f"; {steps} steps, {r['seconds']:.1f} s"
    return r


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    decay = args.workload == "decay"
    unit = "nucleus-steps/s" if decay else "pairs/s"
    metric = "nucleus-steps/s" if decay else "pair interactions/s"
    r = reference_leg(args.workload, args.steps, args.warmup)
    kind = "reference"
    extra = {}
    if r is None:       # baseline/_ref absent: the oracle's C port stands in (kind "port")
        from oracle import oracle as orc
        threads, kind = orc.max_threads(), "port"
        vals = []
        for s in range(args.warmup + args.steps):
            if args.workload == "cloud":
                v, dt = cpu_cloud_sample(65536, 64 * threads, threads)
            elif decay:
                n_d = 1_000_000 * threads
                un = np.random.default_rng(s).random(n_d)
                t0 = time.perf_counter()
                orc.decay_decisions(np.full(n_d, T_C14), T_C14 * 1e-3, un, n_threads=threads)
                dt = time.perf_counter() - t0
                v = n_d / dt
            else:
                v, _, dt = cpu_ensemble_sample((PB208,) if args.workload == "ensemble" else README_ISO,
                                               64 * threads, 1, threads)
            if s >= args.warmup:
                vals.append((v, dt))
        value, ms = float(np.mean([v for v, _ in vals])), float(np.mean([d for _, d in vals])) * 1e3
        cores, sample = threads, "oracle C port (OpenMP), bounded sample per step; baseline/_ref not installed"
    else:
        value = r["units_per_s"] if decay else r["pairs_per_s"]
        ms, cores, sample = r["ms_per_step"], r["cores"], r["sample"]
        if args.workload in ("ensemble", "mixed"):
            extra["nucleus_steps_per_s"] = r["units_per_s"]
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak" if args.workload == "ensemble" else "strong",
        "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
class Ctx:
    """Everything the workload functions share."""

    def __init__(self, args, torch, dist, rank, world, local_rank, lib):
        self.args, self.torch, self.dist = args, torch, dist
        self.rank, self.world, self.local_rank, self.lib = rank, world, local_rank, lib
        self.dev = f"cuda:{local_rank}"
        self.gpu_index = physical_gpu_index(local_rank)

    L2_FLUSH_BYTES = 256 << 20          # > 126 MB of L2

    def flush_l2(self):
        if getattr(self, "_flush_buf", None) is None:
            self._flush_buf = self.torch.empty(self.L2_FLUSH_BYTES, dtype=self.torch.uint8, device=self.dev)
        self._flush_buf.fill_(1)

    def timed(self, fn, steps, warmup, clocks=True, flush=None):
        """W warm-up calls, then K timed calls bracketed by barrier + synchronize; CUDA events on the
        current stream; returns (seconds as the max over ranks, clock record).
        flush="inside": an L2 flush (256 MB write) before every timed call, INSIDE the timed region (for
        steps that are long against its ~0.05 ms); flush="between": every call timed by its own event
        pair with the flush between the pairs, the K durations summed (for short steps)."""
        torch, dist = self.torch, self.dist
        if flush == "between":
            return self._timed_between(fn, steps, warmup, clocks)
        if flush == "inside":
            inner = fn

            def fn():
                self.flush_l2()
                inner()
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(self.gpu_index) if clocks else None
        if sampler is not None:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if sampler is not None:
            sampler.sample_once()     # the queue is still draining: a sample under load even for short runs
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        if dist is not None:
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, (sampler.result() if sampler is not None else None)

    def _timed_between(self, fn, steps, warmup, clocks):
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            self.flush_l2()
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(self.gpu_index) if clocks else None
        if sampler is not None:
            sampler.start()
        pairs = []
        for _ in range(steps):
            self.flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            pairs.append((e0, e1))
        if sampler is not None:
            sampler.sample_once()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sec = sum(a.elapsed_time(b) for a, b in pairs) * 1e-3
        if dist is not None:
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, (sampler.result() if sampler is not None else None)

    def timed_wall(self, fn, steps, warmup):
        """Same bracket for BLOCKING host-buffer calls (they synchronise internally): wall clock between
        two device-wide synchronisations, max over ranks."""
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C

    import torch
    from pyqmd_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()                   # no CPU fallback: fail loudly before anything else
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.lib()
    ctx = Ctx(args, torch, dist, rank, world, local_rank, lib)
    K, W = args.steps, max(args.warmup, 0)

    # FP32 FMA peak of this GPU, measured now (roofline denominator)
    f1, f2 = C.c_double(), C.c_double()
    _lib.check(lib.pyqmd_fp32_peak(4096, C.byref(f1), C.byref(f2), _lib.current_stream()), "fp32_peak")
    props = (C.c_int64 * 8)()
    lib.pyqmd_device_props(local_rank, props)
    ctx.fp32_peak = max(f1.value, f2.value)
    ctx.sms, ctx.sm_hz = int(props[0]), float(props[3]) * 1e3
    ctx.peak_info = (f1.value, f2.value, props[0] * 128 * 2 * props[3] * 1e3 / 1e12)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ctx.hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ctx.hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    line = {"metric": "pair interactions/s", "unit": "pairs/s", "n_gpus": world, "steps": K,
            "warmup": W, "higher_is_better": True, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    fn = {"cloud": bench_cloud, "ensemble": bench_ensemble, "mixed": bench_ensemble, "decay": bench_decay}[
        args.workload]
    res = fn(ctx, args.workload, K, W, with_e2e=not args.no_e2e)
    details = dict(res.pop("config", {}), **res.pop("details", {}))
    details.pop("workload", None)
    line.update(res)
    line["config"] = workload_config(args)          # identical in both arms
    line["details"] = details                       # what this arm measured about the workload

    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.workload)

    summary = {args.workload: line["value"], "e2e": (line.get("e2e") or {}).get("value"),
               "roofline_frac": line["roofline"]["frac"]}
    if not args.no_extras and args.workload == "cloud":
        also = {}
        for name, f, wl, k, w in (("ensemble_c2", bench_ensemble, "ensemble", 20, 5),
                                  ("ensemble_c2_settled", bench_ensemble_settled, "ensemble", 10, 3),
                                  ("mixed_c3", bench_ensemble, "mixed", 10, 3),
                                  ("decay_c5", bench_decay, "decay", 10, 3),
                                  ("cloud_c4_skip_exact_zeros_optin", bench_cloud_skip, "cloud", 10, 3)):
            try:        # a failure here must not cost the headline
                r = f(ctx, wl, k, w, with_e2e=(name == "ensemble_c2"))
                r["steps"], r["warmup"] = k, w
                also[name] = r
                summary[name] = r["value"]
                if r["roofline"].get("frac") is not None:
                    summary[name + "_frac"] = r["roofline"]["frac"]
                if "e2e" in r:
                    summary[name + "_e2e"] = r["e2e"]["value"]
                torch.cuda.empty_cache()
            except Exception as exc:      # noqa: BLE001
                import traceback
                traceback.print_exc()
                also[name] = {"error": repr(exc)[:300]}
        try:
            if rank == 0:
                also["c1"] = bench_c1(ctx)
                summary["c1_ms_per_1000_substeps_resident"] = also["c1"]["device_resident_1000"]["ms_total"]
        except Exception as exc:      # noqa: BLE001
            also["c1"] = {"error": repr(exc)[:300]}
        if rank == 0 and world == 1 and not args.no_cpu:
            for name, wl in (("ensemble_c2", "ensemble"), ("mixed_c3", "mixed"), ("decay_c5", "decay")):
                if name in also and "error" not in also[name]:
                    also[name]["cpu_baseline"] = cpu_baseline(wl, short=True)
        line["also"] = also
    line["summary"] = summary

    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(workload, short=False):
    """The reference's CPU path beside the GPU number: the unmodified reference (kind "reference",
    baseline/_ref) and the oracle's C/OpenMP port, both on all host cores, bounded samples."""
    from oracle import oracle as orc
    threads = orc.max_threads()
    decay = workload == "decay"
    unit = "nucleus-steps/s" if decay else "pairs/s"
    if workload == "cloud":
        n_i = (1024 if short else 4096) * threads
        v, dt = cpu_cloud_sample(65536, n_i, threads)
        port = {"value": v, "sample": f"{min(n_i, 65536)} i-nucleons x 65,535 partners, {dt:.1f} s"}
    elif decay:
        n_d = 4_000_000 * threads
        un = np.random.default_rng(1).random(n_d)
        t0 = time.perf_counter()
        orc.decay_decisions(np.full(n_d, T_C14), T_C14 * 1e-3, un, n_threads=threads)
        dt = time.perf_counter() - t0
        port = {"value": n_d / dt, "sample": f"{n_d} should_decay decisions with supplied uniforms, {dt:.2f} s"}
    else:
        iso = (PB208,) if workload == "ensemble" else README_ISO
        n_s, n_st = (64 if short else 256) * threads, (10 if short else 30)
        v, ns, dt = cpu_ensemble_sample(iso, n_s, n_st, threads)
        port = {"value": v, "nucleus_steps_per_s": ns,
                "sample": f"{n_s} nuclei x {n_st} sub-steps, {dt:.1f} s, OpenMP over nuclei"}
    port.update(unit=unit, cores=threads, kind="port")
    ref = None
    try:
        ref = reference_leg(workload, 3 if short else 8, 1)
    except Exception as exc:      # noqa: BLE001
        port["reference_error"] = repr(exc)[:200]
    if ref is None:
        return port
    out = {"value": ref["units_per_s"] if decay else ref["pairs_per_s"], "unit": unit, "cores": ref["cores"],
           "kind": "reference", "sample": ref["sample"], "port": port}
    if workload in ("ensemble", "mixed"):
        out["nucleus_steps_per_s"] = ref["units_per_s"]
    return out


# ---------------------------------------------------------------------------------------------
def mufu_roofline(ctx, evaluations_per_s, mufu_per_evaluation, what):
    """The pipe that really bounds the force kernels: special-function (MUFU) results, 16 per clock and
    SM; `frac` = MUFU operations the executed pair evaluations need / that peak."""
    peak = 16.0 * ctx.sms * ctx.sm_hz
    need = evaluations_per_s * mufu_per_evaluation
    return {"mufu_per_executed_evaluation": mufu_per_evaluation, "what": what, "achieved_per_s": need,
            "peak_per_s": peak, "frac": need / peak,
            "peak_source": "16 MUFU results per clock and SM (measured 15.9, scripts/ubench/pipes.cu) x %d SMs x "
                           "%.0f MHz" % (ctx.sms, ctx.sm_hz / 1e6)}


def bench_cloud(ctx, workload, K, W, with_e2e=True):
    """C4: one cloud of N nucleons, strong scaling."""
    from pyqmd_b200.forces import NuclearForces
    from pyqmd_b200.state import NucleonCloud
    args, torch = ctx.args, ctx.torch
    n = args.cloud_n
    pos, isp = make_cloud(n)
    cloud = NucleonCloud(pos, isp, device=ctx.dev, rank=ctx.rank, world=ctx.world, scheme=args.cloud_scheme,
                         exchange=args.cloud_exchange, allow_nccl_fallback=True)
    sec, clocks = ctx.timed(lambda: cloud.step(1), K, W, flush="inside")
    pairs = float(n) * (n - 1)
    f_pp = (float(isp.sum()) / n) ** 2
    flops_pair = 23.0 + 3.0 * f_pp
    mine = cloud.pairs_per_step() * K / sec
    sym = args.cloud_scheme == "symmetric"
    kernel = "cloud_sym_kernel" if sym else "cloud_force_kernel"
    achieved = mine * flops_pair / 1e12
    res = {
        "value": pairs * K / sec, "ms_per_step": sec / K * 1e3, "scaling": "strong",
        "details": {"l2_policy": "L2 flushed (256 MB write) before every timed step, inside the timed region "
                                 "(0.05 ms against a step of tens of ms); the step's own working set (positions "
                                 "8N B + accumulators 16N B) is smaller than L2",
                    "scheme": args.cloud_scheme, "exchange_used": cloud.exchange or "none (1 GPU)",
                    "exchange_detail": ("integer force reduce-scatter + integrate + position all-gather "
                                        + ("fused in one peer-memory kernel (NVLink/NVSwitch, symmetric memory)"
                                           if cloud.exchange == "peer" else "via NCCL")
                                        if sym and ctx.world > 1 else
                                        ("none (1 GPU)" if ctx.world == 1 else "NCCL position all-gather"))},
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": ctx.fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / ctx.fp32_peak,
                     "frac_executed": achieved / ctx.fp32_peak * (0.5 if sym else 1.0),
                     "traffic": ncu_traffic(kernel) if n == N_CLOUD and ctx.world == 1 else None,
                     "kernel": kernel, "flops_per_pair": flops_pair,
                     "algorithmic_bytes_per_launch": 36 * n,
                     "executed_pair_evaluations_per_step": pairs / 2 if sym else pairs,
                     "mufu": mufu_roofline(ctx, cloud.pairs_per_step() * (0.5 if sym else 1.0) * K / sec, 2.0,
                                           "far pair: rsqrt + ex2"),
                     "note": ("algorithmic FLOPs of all N(N-1) ordered pairs (SURVEY 8d) over the kernel time; "
                              "the symmetric scheme evaluates each unordered pair once, so `frac` may exceed 1 "
                              "and `frac_executed` (half) is the pipe utilisation; the kernel's own bound is "
                              "the MUFU pipe, 2 per unordered pair at 16/clk/SM"
                              if sym else "ordered pairs, each evaluated"),
                     "peak_source": "FFMA-chain microbenchmark in this run (scalar %.1f, f32x2 %.1f "
                                    "TFLOP/s); nominal %.1f" % ctx.peak_info},
        "clocks": clocks, "gpu_launches": K * 4,
    }
    if ctx.world > 1 and sym:
        # where the rest of the scaling goes: this rank's share of the pair forces, timed alone
        cloud.profile = []
        ctx.timed(lambda: cloud.step(1), 3, 0, clocks=False)
        mine_ms = torch.tensor([cloud.pair_kernel_ms()], device=ctx.dev, dtype=torch.float64)
        cloud.profile = None
        allms = [torch.zeros_like(mine_ms) for _ in range(ctx.world)]
        ctx.dist.all_gather(allms, mine_ms)
        per_rank = [float(t.item()) for t in allms]
        res["details"]["pair_kernel_ms_per_rank"] = per_rank
        res["details"]["pair_kernel_imbalance"] = max(per_rank) / (sum(per_rank) / len(per_rank)) - 1.0
        res["details"]["step_ms_outside_pair_kernel"] = sec / K * 1e3 - max(per_rank)
    if with_e2e:
        k2, w2 = max(3, K // 2), 2
        if ctx.world == 1:
            # the reference-shaped call: host arrays in, host arrays out, caller's order, one C-ABI call
            pin = lambda a: torch.from_numpy(a).pin_memory()
            h_pos, h_vel, h_isp = pin(pos.copy()), pin(np.zeros_like(pos)), pin(isp.copy())
            nf = NuclearForces()
            sec_e = ctx.timed_wall(lambda: nf.step_cloud(h_pos, h_vel, h_isp, 1 / 240, 1), k2, w2)
            h2d, d2h = 8 * n + 8 * n + n, 16 * n
            api = "NuclearForces.step_cloud -> pyqmd_cloud_step_host (upload, sort, step, un-sort, download)"
        else:
            blk = cloud.i1 - cloud.i0
            h_pos = torch.empty(blk, 2, dtype=torch.float32).pin_memory()
            h_vel = torch.empty(blk, 2, dtype=torch.float32).pin_memory()
            cloud.download_block(h_pos, h_vel)
            sec_e = ctx.timed_wall(lambda: cloud.step_host(h_pos, h_vel), k2, w2)
            h2d, d2h = 16 * blk, 16 * blk
            api = ("NucleonCloud.step_host: every rank uploads its block (pos, vel), positions are "
                   "all-gathered, the step runs, every rank downloads its block")
        res["e2e"] = {"value": pairs * k2 / sec_e, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "ms_per_step": sec_e / k2 * 1e3, "steps": k2, "api": api,
                      "timing": "wall clock between device-wide synchronisations (the call blocks), max over ranks"}
    del cloud
    torch.cuda.empty_cache()
    return res


def bench_cloud_skip(ctx, workload, K, W, with_e2e=False):
    """C4 with PYQMD_CLOUD_SKIP_EXACT_ZEROS (opt-in, NOT the headline): tiles further apart than 353 skip
    the tail exponential (exactly +0 in FP32 there) and, without a p-p pair, the whole tile.  Same bits,
    less work -- reported apart because the headline metric is defined over evaluated pairs."""
    from pyqmd_b200.state import NucleonCloud
    args, torch = ctx.args, ctx.torch
    n = args.cloud_n
    pos, isp = make_cloud(n)
    cloud = NucleonCloud(pos, isp, device=ctx.dev, rank=ctx.rank, world=ctx.world, scheme="symmetric",
                         exchange=args.cloud_exchange, allow_nccl_fallback=True, skip_exact_zeros=True)
    check = NucleonCloud(pos, isp, device=ctx.dev, rank=ctx.rank, world=ctx.world, scheme="symmetric",
                         exchange=args.cloud_exchange, allow_nccl_fallback=True)
    cloud.step(1); check.step(1)
    same = bool(torch.equal(cloud.pos, check.pos) and torch.equal(cloud.vel, check.vel))
    del check
    sec, clocks = ctx.timed(lambda: cloud.step(1), K, W, flush="inside")
    pairs = float(n) * (n - 1)
    res = {"metric": "pair interactions/s", "unit": "pairs/s", "value": pairs * K / sec,
           "ms_per_step": sec / K * 1e3, "scaling": "strong",
           "config": {"workload": f"C4 cloud N={n}, symmetric scheme with PYQMD_CLOUD_SKIP_EXACT_ZEROS (opt-in)"},
           "bit_identical_to_default_after_1_step": same,
           "roofline": {"bound": "fp32", "frac": None, "kernel": "cloud_sym_kernel (skip paths)",
                        "note": "no roofline fraction: most of the N(N-1) pairs are PROVED zero (tail underflow "
                                "beyond d = 353, no p-p pair in the tile), not evaluated; the number says how "
                                "fast the bit-identical step can be, not how busy the pipes are"},
           "clocks": clocks, "gpu_launches": K * 4}
    del cloud
    torch.cuda.empty_cache()
    return res


def _ensemble_result(ctx, workload, ens, K, sec, clocks, flops_pair, census, substeps, isotopes):
    torch, dist = ctx.torch, ctx.dist
    pairs_step = ens.pairs_per_step() * substeps
    nucleons = int(ens.count.sum().item())
    tot_pairs = torch.tensor([float(pairs_step)], device=ctx.dev, dtype=torch.float64)
    tot_nuc = torch.tensor([float(ens.n_nuclei)], device=ctx.dev, dtype=torch.float64)
    decays = ens.mode_counts.clone()
    if dist is not None:
        dist.all_reduce(tot_pairs); dist.all_reduce(tot_nuc); dist.all_reduce(decays)
    my_rate = pairs_step * K / sec
    achieved = my_rate * flops_pair / 1e12
    res = {
        "metric": "pair interactions/s", "unit": "pairs/s",
        "value": float(tot_pairs.item()) * K / sec, "ms_per_step": sec / K * 1e3,
        "scaling": "weak" if workload == "ensemble" and not ctx.args.nuclei else "strong",
        "config": {"workload": ("C2 ensemble of 65,536 independent Pb-208 nuclei per GPU"
                                if workload == "ensemble" and isotopes == (PB208,) else
                                "C3 mixed ensemble of 1M nuclei over the nine preset isotopes, decay on, "
                                "sharded by nucleus" if workload == "mixed" else f"ensemble of {isotopes}"),
                   "nuclei_total": int(tot_nuc.item()), "nucleons_per_gpu": nucleons,
                   "substeps_per_step": substeps, "dt_phys": 1 / 240,
                   "l2_policy": "inputs larger than L2 (state %.0f MB per GPU)" % (nucleons * 17 / 1e6),
                   "parallelism": f"by-nucleus x{ctx.world}, no collective"},
        "nucleus_steps_per_s": float(tot_nuc.item()) * substeps * K / sec,
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": ctx.fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / ctx.fp32_peak,
                     "traffic": ncu_traffic("ensemble_kernel") if workload == "ensemble"
                     and not ctx.args.nuclei and substeps == 1 else None,
                     "kernel": "ensemble kernel (see DESIGN.md section 4)", "flops_per_pair": flops_pair,
                     "algorithmic_bytes_per_launch": 36 * nucleons, "branch_census_end": census,
                     "mufu": mufu_roofline(ctx, my_rate / 2, 5.0,
                                           "general law: rsqrt, sqrt, rcp, 2 x ex2 per unordered pair "
                                           "(idle lanes of partly filled warps not counted)"),
                     "peak_source": "FFMA-chain microbenchmark in this run (scalar %.1f, "
                                    "f32x2 %.1f TFLOP/s); nominal %.1f" % ctx.peak_info,
                     "hbm_gbs": nucleons * 36 * K / sec / 1e9, "hbm_peak_gbs": ctx.hbm_peak,
                     "hbm_peak_source": ctx.hbm_src},
        "clocks": clocks, "gpu_launches": K * len(ens.bins),
    }
    if ens.decay:
        res["decays"] = {"total_events": int(decays.sum().item()), "by_mode": [int(v) for v in decays.tolist()]}
    return res, tot_pairs


def bench_ensemble(ctx, workload, K, W, with_e2e=True):
    """C2 (``ensemble``: 65,536 Pb-208 per GPU, weak) or C3 (``mixed``: 1M nuclei over the nine preset
    isotopes in total, decay on, strong).  Free-running sub-steps (no overlap projection): the state the
    reference's own un-projected dynamics collapse into."""
    import math

    from pyqmd_b200.state import README_ISOTOPES, HostEnsembleRunner, NucleusEnsemble
    args = ctx.args
    isotopes = (PB208,) if workload == "ensemble" else README_ISOTOPES
    if workload == "ensemble" and args.isotope:
        isotopes = (tuple(int(v) for v in args.isotope.split(",")),)
    total = args.nuclei * ctx.world if args.nuclei else (N_ENSEMBLE * ctx.world if workload == "ensemble"
                                                         else N_MIXED)
    per = (total + ctx.world - 1) // ctx.world
    lo = ctx.rank * per
    n_mine = max(0, min(per, total - lo))
    decay = workload == "mixed"
    ens = NucleusEnsemble.from_templates(isotopes, n_mine, device=ctx.dev, id_base=lo, decay=decay,
                                         dt_decay=T_C14 * 1e-3, seed=2024)
    stride = max(1, ens.n_nuclei // 256)
    while math.gcd(stride, len(isotopes)) != 1:      # isotopes cycle with the nucleus id: hit them all
        stride += 1
    sample = range(0, ens.n_nuclei, stride)
    if args.settled and workload == "ensemble":
        for _ in range(12):
            ens.frame(4)
    flops0 = ens.census(sample)[1]
    sec, clocks = ctx.timed(lambda: ens.step(args.substeps), K, W)
    census1, flops1 = ens.census(sample)
    res, tot_pairs = _ensemble_result(ctx, workload, ens, K, sec, clocks, 0.5 * (flops0 + flops1), census1,
                                      args.substeps, isotopes)
    res["state"] = ("settled (12 app frames), then free-running" if args.settled else
                    "free-running sub-steps from the reference layout (collapses, SURVEY section 7)")
    if with_e2e:
        # e2e: host-buffer API, H2D + kernel + D2H every step
        runner = HostEnsembleRunner(ens, chunks=16)
        k2 = max(3, K // 4)
        sec_e2e, _ = ctx.timed(lambda: runner.step(args.substeps), k2, 2, clocks=False)
        res["e2e"] = {"value": float(tot_pairs.item()) * k2 / sec_e2e, "unit": "pairs/s",
                      "h2d_bytes_per_step": runner.h2d_bytes, "d2h_bytes_per_step": runner.d2h_bytes,
                      "chunks": runner.n_chunks, "api": "HostEnsembleRunner.step -> pyqmd_ensemble_step_host"}
        sec_e2e4, _ = ctx.timed(lambda: runner.step(4 * args.substeps), max(3, k2 // 4), 1, clocks=False)
        res["e2e"]["value_4_substeps_per_call"] = (float(tot_pairs.item()) * 4 * max(3, k2 // 4) / sec_e2e4)
        del runner
    del ens
    ctx.torch.cuda.empty_cache()
    return res


def bench_ensemble_settled(ctx, workload, K, W, with_e2e=False):
    """C2 from the SETTLED, app-faithful state: fresh layouts + 12 app frames (4 sub-steps + the overlap
    projection, nuclear_sim.py:161-176); then (a) K app frames, (b) K calls of 4 sub-steps alone -- the
    roofline of the same kernel on a physical state (94 % tail pairs instead of 98 % hard core)."""
    from pyqmd_b200.state import NucleusEnsemble
    n_f = (ctx.args.nuclei or N_ENSEMBLE)
    ens = NucleusEnsemble.from_templates((PB208,), n_f, device=ctx.dev, id_base=ctx.rank * n_f, decay=False)
    for _ in range(12):
        ens.frame(4)
    sample = range(0, n_f, max(1, n_f // 256))
    p0 = int(ens.push_count.item())
    sec_f, clocks_f = ctx.timed(lambda: ens.frame(4), K, W)
    pushes = (int(ens.push_count.item()) - p0) / (K + W) / n_f
    census0, flops0 = ens.census(sample)
    sec_s, clocks = ctx.timed(lambda: ens.step(4), K, 1)
    census1, flops1 = ens.census(sample)
    res, tot_pairs = _ensemble_result(ctx, "ensemble", ens, K, sec_s, clocks, 0.5 * (flops0 + flops1), census1, 4,
                                      (PB208,))
    res["state"] = "settled: 12 app frames (4 sub-steps + resolve_overlaps) from the reference layout"
    res["frame"] = {"ms_per_frame": sec_f / K * 1e3, "substeps_per_frame": 4,
                    "pairs_per_s": float(tot_pairs.item()) * K / sec_f,
                    "pushes_per_nucleus_frame": pushes, "clocks": clocks_f,
                    "note": "4 sub-steps + resolve_overlaps per frame (nuclear_sim.py:161-176), device resident"}
    del ens
    ctx.torch.cuda.empty_cache()
    return res


def bench_c1(ctx):
    """C1: one U-238 nucleus (configs[0]), 1000 force+integrate(+decay test) steps.
    (a) through the reference-shaped drop-in NuclearForces.update_particles_cpu(list[Particle], dt),
        one host round trip per sub-step, exactly how nuclear_sim.py:171-173 calls it;
    (b) the same 1000 sub-steps fused into one call (NuclearForces.step);
    (c) device resident (NucleusEnsemble of one nucleus, 1000 sub-steps in one launch, decay on)."""
    from pyqmd_b200.forces import NuclearForces
    from pyqmd_b200.state import NucleusEnsemble, layout_templates
    from pyqmd_b200.types import Particle, ParticleType
    torch = ctx.torch
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
    ens = NucleusEnsemble.from_templates(((92, 146),), 1, device=ctx.dev, decay=True,
                                         dt_decay=1.409993568e17 * 1e-3, seed=1)
    ens.step(10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ens.step(1000)
    e1.record()
    torch.cuda.synchronize()
    resident = e0.elapsed_time(e1) * 1e-3
    out = {"config": {"workload": "C1 single U-238 nucleus (238 nucleons), 1000 sub-steps"},
           "dropin_per_substep_call": {"ms_per_call": per_call * 1e3, "pairs_per_s": pairs / per_call,
                                       "api": "NuclearForces.update_particles_cpu(list[Particle], dt)"},
           "dropin_fused_1000": {"ms_total": fused * 1e3, "pairs_per_s": pairs * 1000 / fused,
                                 "api": "NuclearForces.step(list[Particle], dt, 1000)"},
           "device_resident_1000": {"ms_total": resident * 1e3, "pairs_per_s": pairs * 1000 / resident,
                                    "nucleus_steps_per_s": 1000 / resident,
                                    "api": "NucleusEnsemble.step(1000), decay test on"}}
    if not ctx.args.no_cpu and ctx.world == 1:
        from oracle import oracle as orc
        x, y = xy[:, 0].astype(np.float64), xy[:, 1].astype(np.float64)
        t0 = time.perf_counter()
        orc.ensemble_force_steps(np.array([0], np.int64), np.array([238], np.int32), x, y, np.zeros(238),
                                 np.zeros(238), isp, 1 / 240, 200, n_threads=1)
        cpu = (time.perf_counter() - t0) / 200
        out["cpu_port_1core"] = {"ms_per_substep": cpu * 1e3, "pairs_per_s": pairs / cpu,
                                 "note": "C oracle port, one core (one nucleus does not parallelise)"}
        try:    # the unmodified reference on this box: the same U-238, 3 sub-steps, one core
            from oracle import ref_loader
            if ref_loader.available():
                R = ref_loader.Ref()
                rp = R.make_particles(400.0 + x, 400.0 + y, np.zeros(238), np.zeros(238), isp)
                rf = R.forces()
                rf.update_particles_cpu(rp, 1 / 240)
                t0 = time.perf_counter()
                for _ in range(3):
                    rf.update_particles_cpu(rp, 1 / 240)
                ref = (time.perf_counter() - t0) / 3
                out["reference_python_1core"] = {"ms_per_substep": ref * 1e3, "pairs_per_s": pairs / ref,
                                                 "kind": "reference",
                                                 "note": "unmodified update_particles_cpu (baseline/_ref), this box"}
        except Exception as exc:      # noqa: BLE001
            out["reference_python_1core"] = {"error": repr(exc)[:200]}
    return out


def bench_decay(ctx, workload, K, W, with_e2e=False):
    from pyqmd_b200.state import DecayPopulation
    torch = ctx.torch
    total = N_DECAY
    per = (total + ctx.world - 1) // ctx.world
    lo = ctx.rank * per
    n_mine = max(0, min(per, total - lo))
    zn = torch.full((n_mine,), (6 << 16) | 8, dtype=torch.int32)
    zn[n_mine // 2:] = (92 << 16) | 146
    pop = DecayPopulation(zn, device=ctx.dev, dt_decay=T_C14 * 1e-3, seed=7, id_base=lo,
                          watch=((6, 8), (92, 146)))
    sub = max(ctx.args.substeps, 1)
    bytes_nuc, note = pop.bytes_per_nucleus_launch()
    small = n_mine * bytes_nuc <= (160 << 20)          # the per-GPU input would fit the 126 MB L2
    sec, clocks = ctx.timed(lambda: pop.step(sub), K, W, flush="between" if small else None)
    res = {
        "metric": "nucleus-steps/s", "unit": "nucleus-steps/s", "value": float(total) * sub * K / sec,
        "ms_per_step": sec / K * 1e3, "scaling": "strong", "dtype": "f64",
        "config": {"workload": "C5 decay-only Monte Carlo, 1e8 C-14 / U-238 nuclei, Philox draws",
                   "substeps_per_step": sub, "parallelism": f"by-nucleus x{ctx.world}",
                   "l2_policy": ("L2 flushed between the timed steps (each step timed by its own event pair); "
                                 "%.0f MB per GPU" if small else "inputs larger than L2 (%.0f MB per GPU)")
                   % (n_mine * bytes_nuc / 1e6)},
        "roofline": {"bound": "hbm", "achieved": n_mine * bytes_nuc * K / sec / 1e9, "peak": ctx.hbm_peak,
                     "unit": "GB/s", "frac": n_mine * bytes_nuc * K / sec / 1e9 / ctx.hbm_peak,
                     "traffic": ncu_traffic("population_kernel") if ctx.world == 1 and sub == 1 else None,
                     "kernel": "population_kernel", "peak_source": ctx.hbm_src,
                     "bytes_per_nucleus_launch": bytes_nuc,
                     "algorithmic_bytes_per_launch": n_mine * bytes_nuc, "note": note},
        "clocks": clocks, "gpu_launches": K,
    }
    del pop
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    main()
