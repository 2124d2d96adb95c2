"""The committed evidence stays parseable: the final bench record carries every key of the bench contract, the
ablation record follows the reference's summary_statistics.json schema, and the launch-list aggregator reads the
committed ncu CSVs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def test_final_bench_record_has_the_contract_keys():
    d = json.load(open(os.path.join(PROFILES, "r01_bench_final3.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "decode_tokens_per_s" and d["unit"] == "tokens/s" and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] == d["launches_per_step"] * d["steps"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["vision_encode"]["frac_of_bf16_burst_peak"] > 0.5            # north-star target for the vision tower
    assert d["roofline"]["step"]["frac"] > 0.6                            # and for the batch-1 decode step


def test_ablation_record_follows_the_reference_schema():
    d = json.load(open(os.path.join(PROFILES, "r01_ablation_kv_on_off.json")))
    for L in (16, 32, 64, 128, 256):
        for key, cached in (("kv_cache_%d" % L, True), ("no_kv_cache_%d" % L, False)):
            e = d[key]
            assert e["sequence_length"] == L and e["kv_cache_enabled"] is cached
            for field in ("steady_state_tps", "steady_state_ms_per_token", "peak_memory_mb"):
                assert set(e[field]) == {"mean", "ci_95", "std"}
            assert e["tokens_generated"]["mean"] == float(L)
        assert d["kv_cache_%d" % L]["steady_state_tps"]["mean"] > d["no_kv_cache_%d" % L]["steady_state_tps"]["mean"]


def test_launch_list_aggregator_reads_the_committed_csv():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "agg_launches.py"),
                          os.path.join(PROFILES, "r01_launches_batch32_final.csv"), "5"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "total" in out.stdout and "gemm_tc" in out.stdout
