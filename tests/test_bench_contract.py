"""The bench.py output contract, checked on the CPU through the reference arm (no GPU needed) and, for
the B200 arm, on the GPU with a small particle count."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "2e5")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("Gparticles/s deposited (2D TSC, fp64)") and d["unit"] == "Gparticles/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("histogram2d x-px phase space") and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run("--steps", "3", "--warmup", "3", "--n", "4e6", "--cpu-sample", "2e5", "--e2e-steps", "2")
    assert BASE_KEYS | {"roofline", "clocks", "variants", "parity"} <= set(d) and "impl" not in d
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["unit"] == "GB/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 4000000 * 24 and d["e2e"]["d2h_bytes_per_step"] == 1028 * 1028 * 8
    assert d["e2e"]["value"] > 0 and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] > 0
    assert d["parity"]["mass_conservation_rel_err"] < 1e-12
    assert set(d["variants"]) == {"x_sorted", "cell_sorted"}
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
