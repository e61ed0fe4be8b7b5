"""The benchmark contract on the CPU side: the committed JSON lines of the round carry every key the driver reads, the
reference arm prints the GPU arm's config object, and a non-zero rank of the reference arm exits quietly."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = os.path.join(ROOT, "profiles", "r2_bench")
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"}


def _line(name):
    with open(os.path.join(LINES, name)) as f:
        return json.loads(f.read())


@pytest.mark.parametrize("name", ["bench_c1.json", "bench_c2.json", "bench_c3.json", "bench_c4.json", "bench_c5.json"])
def test_committed_bench_lines_carry_the_contract(name):
    d = _line(name)
    assert KEYS <= set(d), KEYS - set(d)
    assert d["metric"] == "frames_per_s" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "int16"
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["steps"] >= 1 and d["warmup"] >= 3
    frames = d["config"].get("job_frames") or d["config"]["frames_per_step_per_gpu"] * d["n_gpus"]   # c5: the whole job is a step
    assert abs(d["value"] - frames / d["ms_per_step"] * 1e3) <= 1e-6 * d["value"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - r["algorithmic_bytes_per_run"] / (r["ms_per_run"] * 1e-3) / 1e9) <= 1e-9 * r["achieved"]
    assert abs(sum(r["groups_ms_per_run"].values()) - r["ms_per_run"]) <= 1e-9
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["scaling"] == ("strong" if name == "bench_c5.json" else "weak")
    if name != "bench_c5.json":   # (the c5 line is taken with --no-cpu)
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["parity_check"]["ok"] is True


def test_reference_arm_line_matches_the_gpu_arm():
    ref, gpu = _line("bench_c3_reference_arm.json"), _line("bench_c3.json")
    assert ref["impl"] == "reference" and ref["config"] == gpu["config"]
    for k in ("metric", "unit", "higher_is_better", "dtype", "scaling"):
        assert ref[k] == gpu[k], k
    assert ref["e2e"] == {"value": ref["value"], "unit": ref["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert ref["cpu_baseline"]["value"] == ref["value"] and ref["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
