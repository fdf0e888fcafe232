"""CPU checks of the bench.py contract: the reference arm (`--impl reference`, the oracle port on the
host cores) prints ONE JSON line with the agreed keys, alone and under a two-rank torchrun launch
(rank 0 prints, the other rank exits 0 without work); the committed bench line of the round carries
every key the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


def _check_reference_line(d, n_gpus):
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["n_gpus"] == n_gpus and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "scenes/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0                    # nothing of the product runs on this arm


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 1)


def test_reference_arm_under_torchrun_only_rank0_reports():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", "bench.py", "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 2)


def test_committed_bench_line_has_every_contract_key():
    d = _json_lines(open(os.path.join(ROOT, "profiles", "bench_r01.json")).read())[-1]
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline"} <= set(d)
    assert d["metric"].startswith("scenes/sec Pointnet2Backbone fwd") and d["n_gpus"] == 1
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert d["roofline"]["bound"] in ("hbm", "tensor")
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 40000 * 4 * 4      # 8 scenes x 40k points x (xyz + height) fp32
    assert d["e2e"]["value"] < d["value"]                             # the copies are inside the timed region
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
