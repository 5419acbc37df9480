"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the keys
the driver reads, and our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                          text=True, timeout=600, env=e)


def _check_reference_line(r, kind):
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pair interactions/s" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 1e5 and d["ms_per_step"] > 0 and d["scaling"] == "strong"
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert "N=1000000" in d["config"]["workload"]          # the headline config: the 1M-nucleon cloud
    return d


def test_reference_arm_times_the_unmodified_reference():
    """baseline/_ref (baseline/install_reference.py) present: the arm runs the reference's own
    update_particles_cpu, one process per host core."""
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("baseline/_ref not installed")
    d = _check_reference_line(run("--impl", "reference", "--steps", "2", "--warmup", "1"), "reference")
    assert d["value"] < 1e8             # CPython: ~7e5 pairs/s per core


def test_reference_arm_falls_back_to_the_port_without_the_reference():
    r = run("--impl", "reference", "--steps", "2", "--warmup", "1", env={"PYQMD_NO_REF": "1"})
    _check_reference_line(r, "port")


def test_both_arms_describe_the_same_config():
    """`config` comes from one function of the command line only, so the reference arm and ours print the
    same object for the same flags (measured details of our arm go to `details`)."""
    sys.path.insert(0, ROOT)
    import argparse

    import bench
    ns = argparse.Namespace(workload="cloud", gpus=4, cloud_n=1_000_000, nuclei=0, cloud_scheme="symmetric",
                            cloud_exchange="peer")
    cfg = bench.workload_config(ns)
    assert cfg == bench.workload_config(ns) and cfg["gpus"] == 4 and cfg["exchange"] == "peer"
    assert "l2_policy" in cfg and "model" not in cfg
    ns.gpus = 1
    assert bench.workload_config(ns)["exchange"].startswith("none")
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "4", env={"PYQMD_NO_REF": "1"})
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    ns.gpus = 4
    assert d["config"] == bench.workload_config(ns)


def test_reference_arm_other_ranks_exit_without_work():
    r = run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_cuda():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    r = run("--steps", "1", "--warmup", "0", "--no-extras")
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
