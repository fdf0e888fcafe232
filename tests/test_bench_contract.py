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


def test_final_bench_line_of_round_2_has_every_contract_key():
    d = _json_lines(open(os.path.join(ROOT, "profiles", "bench_r02_final.json")).read())[-1]
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline"} <= set(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "fp32-alu/latency"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["clocks"]["samples_before_region_end"] >= 1 and not d["clocks"]["reasons"]
    assert d["e2e"]["value"] < d["value"] and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["train_step"]["graphed"]["ms_per_step"] < d["train_step"]["ms_per_step"]


class _FakeNvml:
    NVML_CLOCK_SM = 0

    def __init__(self):
        self.calls = 0

    def nvmlDeviceGetClockInfo(self, h, which):
        self.calls += 1
        return 1965

    def nvmlDeviceGetCurrentClocksEventReasons(self, h):
        return 0x4   # sw_power_cap


class _FakeEvent:
    """query() turns True on its third call: the timed region 'ends' after two in-region samples."""

    def __init__(self):
        self.n = 0

    def query(self):
        self.n += 1
        return self.n >= 3


def test_clock_sampler_counts_only_samples_taken_inside_the_region():
    """bench.ClockSampler.sample_until: the main thread samples NVML between the last enqueue and the closing
    synchronisation; a sample counts as in-region only while the closing event is still pending."""
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.nv, s.h, s.max_mhz = _FakeNvml(), object(), 1965.0
    s.sample_until(_FakeEvent(), period_s=0.0)
    out = s.stop()
    assert out["samples"] == 3 and out["samples_before_region_end"] == 2
    assert out["sm_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]


def test_graphed_train_step_refuses_a_non_capturable_optimizer():
    """graph.GraphedTrainStep checks the optimizer before it touches the device."""
    import importlib
    import pytest
    import torch
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("3dvlp_b200")
    lin = torch.nn.Linear(2, 2)
    with pytest.raises(RuntimeError, match="capturable"):
        pkg.GraphedTrainStep(lambda: lin(torch.zeros(1, 2)).sum(), torch.optim.Adam(lin.parameters(), lr=1e-3))
