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
