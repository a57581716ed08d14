"""The committed bench lines (profiles/bench_r1_*.json, written by `python bench.py` on a B200) carry every key of
the measurement contract: a regression guard for bench.py's output format that runs without a GPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def load(name):
    p = os.path.join(PROFILES, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not committed")
    return json.load(open(p))


def test_default_bench_line_has_the_contract_keys():
    d = load("bench_r1_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "train_samples_per_s" and d["unit"] == "samples/s" and d["n_gpus"] == 1
    assert d["warmup"] >= 3 and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] is None or r["traffic"] >= 0.9 * r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] <= d["value"] * 1.02           # the end-to-end number includes the copies
    assert d["gpu_launches"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert abs(d["value"] - d["config"]["global_batch"] * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]


def test_reference_arm_line():
    d = load("bench_r1_reference_arm.json")
    ours = load("bench_r1_final.json")
    assert d["impl"] == "reference" and d["metric"] == ours["metric"] and d["unit"] == ours["unit"]
    assert d["config"]["workload"] == ours["config"]["workload"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_lines_are_weak_scaling_aggregates(n):
    d = load(f"bench_r1_n{n}.json")
    one = load("bench_r1_final.json")
    assert d["n_gpus"] == n and d["scaling"] == "weak"
    assert d["config"]["global_batch"] == n * one["config"]["global_batch"]
    assert one["value"] < d["value"] <= n * one["value"] * 1.05


def test_only_the_result_line_reaches_stdout(tmp_path):
    """bench.py hands file descriptor 1 to stderr before anything runs (NCCL prints a banner to stdout) and writes
    the JSON line to the original stdout."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('noise from a library'); "
            "bench.emit({'metric': 'x', 'value': 1.0})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    assert r.stdout == json.dumps({"metric": "x", "value": 1.0}) + "\n"
    assert "NCCL version" in r.stderr and "noise from a library" in r.stderr


def test_round2_default_line_is_cfg3_with_the_contract_keys():
    d = load("bench_r2_cfg3_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks",
              "pll_eval", "vq_assign", "cfg2"):
        assert k in d, k
    assert d["config"]["workload"].startswith("cfg3") and d["dtype"] == "bf16" and d["n_gpus"] == 1
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["kernel"].startswith("dense_") and 0.3 < r["frac"] < 1
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < d["e2e"]["value"] <= d["value"] * 1.02 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


@pytest.mark.parametrize("n", [2, 4, 8])
def test_round2_multi_gpu_lines_carry_parity_of_the_sharded_exchange(n):
    """Every multi-GPU line of the sharded peer-to-peer exchange carries its own correctness evidence: the
    data-parallel run agrees with one GPU on the same global batches and the replicas are bit-identical."""
    d = load(f"bench_r2_cfg3_n{n}_sharded.json")
    one = load("bench_r2_cfg3_n1.json")
    assert d["n_gpus"] == n and d["scaling"] == "weak"
    assert d["config"]["global_batch"] == n * one["config"]["global_batch"]
    assert 0.85 * n * one["value"] < d["value"] <= n * one["value"] * 1.05
    p = d["dp_parity"]
    assert p["world"] == n and p["sharded_exchange"] is True and p["replicas_identical"] is True and p["failed"] == []
    for k in ("loss_rel", "weight_rel", "moment_rel", "count_l1", "pll_rel_sample_sharded", "pll_rel_variable_sharded"):
        assert p[k] <= 1e-3, (k, p[k])
    nccl = load(f"bench_r2_cfg3_n{n}.json")               # the NCCL exchange of the same round
    assert d["value"] > nccl["value"]
