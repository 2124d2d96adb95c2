"""CPU checks of the report tooling: tools/visualize_results.py draws the four ablation figures (the reference's
visualize_results.py:38-113) as well-formed SVG from JSON in the schema tools/ablation_sweep.py writes."""
import json
import os
import subprocess
import sys
import xml.dom.minidom

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _summary(n_gpus):
    out = {"_meta": {"n_gpus": n_gpus}}
    for L in (16, 64, 256):
        for on in (True, False):
            ms = (1.2 / n_gpus ** 0.5) if on else 6.0 + L / 100
            out[("kv_cache_%d" if on else "no_kv_cache_%d") % L] = {
                "sequence_length": L, "kv_cache_enabled": on, "num_samples": 2,
                "steady_state_tps": {"mean": 1e3 / ms, "ci_95": 0.5, "std": 0.3},
                "steady_state_ms_per_token": {"mean": ms, "ci_95": 0.01, "std": 0.01},
                "peak_memory_mb": {"mean": 5800 + (L * 0.02 if on else 40), "ci_95": 0.0, "std": 0.0}}
    return out


def test_visualize_results_writes_the_four_figures(tmp_path):
    files = []
    for n in (1, 8):
        p = tmp_path / f"ablation_n{n}.json"
        p.write_text(json.dumps(_summary(n)))
        files.append(str(p))
    out = tmp_path / "figures"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "visualize_results.py"), *files, "--out", str(out)], check=True)
    names = sorted(os.listdir(out))
    assert names == ["fig1_latency.svg", "fig2_throughput.svg", "fig3_memory.svg", "fig4_speedup.svg"]
    for n in names:
        doc = xml.dom.minidom.parse(str(out / n))
        assert len(doc.getElementsByTagName("polyline")) >= 2          # one curve per (mode, GPU count)
    speed = (out / "fig4_speedup.svg").read_text()
    assert "1 GPU" in speed and "8 GPUs" in speed


def test_ncu_compact_feeds_the_bench_traffic_reader(tmp_path, monkeypatch):
    """tools/ncu_compact.py keeps the columns bench.py::ncu_traffic_per_launch reads (dram read + write per launch of
    the dominant decode kernel) out of a wide `ncu --page raw --csv` export."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, ROOT)
    import csv
    import io
    import ncu_compact
    head = ["ID", "Kernel Name", "junk__metric.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__time_duration.sum", "launch__grid_size", "x1", "x2", "x3", "x4"]
    units = ["", "", "", "Mbyte", "Mbyte", "us", "", "", "", "", ""]
    rows = [head, units,
            ["0", "void pg::decode_gateup_kernel<__nv_bfloat16, 1, 0>(...)", "7", "134.25", "3.5", "27.1", "444", "", "", "", ""],
            ["1", "void pg::decode_gateup_kernel<__nv_bfloat16, 1, 0>(...)", "7", "134.25", "4.5", "26.9", "444", "", "", "", ""],
            ["2", "void pg::gemv_res_kernel<__nv_bfloat16, 1, 4, 0>(...)", "7", "67.16", "0.4", "20.9", "888", "", "", "", ""]]
    out = ncu_compact.compact(rows)
    assert out[0][:3] == ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum"] and "junk__metric.sum" not in out[0]
    assert out[1][1] == "Mbyte" and len(out) == 5
    prof = tmp_path / "profiles"
    prof.mkdir()
    buf = io.StringIO()
    csv.writer(buf, lineterminator="\n").writerows(out)
    (prof / "r02_ncu_full_gateup.csv").write_text(buf.getvalue())
    import bench
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    traffic, source = bench.ncu_traffic_per_launch()
    assert abs(traffic - (134.25 + 4.0) * 1e6) < 1.0 and "2 launches" in source


def test_bench_stdout_carries_only_the_json_line():
    """bench.py's contract: ONE JSON line on stdout.  Anything a library writes to file descriptor 1 during the run
    (NCCL prints its version banner there under torchrun) must land on stderr instead."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.quiet_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('chatter'); bench.emit({'metric': 'x', 'value': 1.0})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
    assert r.stdout.strip().splitlines() == [json.dumps({"metric": "x", "value": 1.0})]
    assert "NCCL version" in r.stderr and "chatter" in r.stderr
