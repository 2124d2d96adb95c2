#!/usr/bin/env python
"""Figures of the KV-cache ablation (the reference's visualize_results.py:38-113: latency, throughput, memory and
speed-up against sequence length), drawn from the JSON files tools/ablation_sweep.py writes -- one per GPU count, so the
1/2/4/8-GPU curves share each plot.  matplotlib is not in this image: the plots are written as plain SVG.

  python tools/visualize_results.py profiles/r02_ablation_n1.json [profiles/r02_ablation_n2.json ...] --out profiles/figures
"""
import argparse
import json
import math
import os

COLORS = ["#1f77b4", "#d62728", "#2ca02c", "#9467bd", "#ff7f0e", "#8c564b", "#17becf", "#7f7f7f"]


def nice_ticks(lo, hi, n=5):
    if hi <= lo:
        hi = lo + 1
    step = 10 ** math.floor(math.log10((hi - lo) / n))
    for m in (1, 2, 5, 10):
        if (hi - lo) / (step * m) <= n:
            step *= m
            break
    t0 = math.floor(lo / step) * step
    return [t0 + i * step for i in range(int((hi - t0) / step) + 2)]


def line_chart(path, title, xlabel, ylabel, series, logy=False):
    """series: list of (label, xs, ys, errs).  x positions are categorical (sequence lengths)."""
    W, H, L, R, T, B = 520, 340, 64, 150, 34, 46
    xs_all = sorted({x for _, xs, _, _ in series for x in xs})
    ys_all = [y for _, _, ys, _ in series for y in ys]
    f = (lambda v: math.log10(max(v, 1e-9))) if logy else (lambda v: v)
    lo, hi = min(map(f, ys_all)), max(map(f, ys_all))
    pad = 0.08 * (hi - lo or 1.0)
    lo, hi = (lo - pad, hi + pad) if logy else (min(0.0, lo), hi + pad)
    px = lambda x: L + (W - L - R) * (xs_all.index(x) + 0.5) / len(xs_all)
    py = lambda y: T + (H - T - B) * (1 - (f(y) - lo) / (hi - lo))
    o = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="serif" font-size="11">',
         f'<rect width="{W}" height="{H}" fill="white"/>',
         f'<text x="{(L + W - R) / 2}" y="18" text-anchor="middle" font-size="13">{title}</text>',
         f'<line x1="{L}" y1="{H - B}" x2="{W - R}" y2="{H - B}" stroke="black"/>',
         f'<line x1="{L}" y1="{T}" x2="{L}" y2="{H - B}" stroke="black"/>',
         f'<text x="{(L + W - R) / 2}" y="{H - 10}" text-anchor="middle">{xlabel}</text>',
         f'<text x="14" y="{(T + H - B) / 2}" text-anchor="middle" transform="rotate(-90 14 {(T + H - B) / 2})">{ylabel}</text>']
    ticks = [10 ** t for t in range(math.floor(lo), math.ceil(hi) + 1)] if logy else nice_ticks(lo, hi)
    for t in ticks:
        if lo <= f(t) <= hi:
            y = py(t)
            o.append(f'<line x1="{L}" y1="{y:.1f}" x2="{W - R}" y2="{y:.1f}" stroke="#dddddd"/>')
            o.append(f'<text x="{L - 6}" y="{y + 4:.1f}" text-anchor="end">{t:g}</text>')
    for x in xs_all:
        o.append(f'<text x="{px(x):.1f}" y="{H - B + 16}" text-anchor="middle">{x}</text>')
    for i, (label, xs, ys, errs) in enumerate(series):
        c = COLORS[i % len(COLORS)]
        dash = ' stroke-dasharray="5,3"' if "off" in label.lower() else ""
        pts = " ".join(f"{px(x):.1f},{py(y):.1f}" for x, y in zip(xs, ys))
        o.append(f'<polyline points="{pts}" fill="none" stroke="{c}" stroke-width="1.6"{dash}/>')
        for x, y, e in zip(xs, ys, errs):
            o.append(f'<circle cx="{px(x):.1f}" cy="{py(y):.1f}" r="2.6" fill="{c}"/>')
            if e and not logy:
                o.append(f'<line x1="{px(x):.1f}" y1="{py(y - e):.1f}" x2="{px(x):.1f}" y2="{py(y + e):.1f}" stroke="{c}"/>')
        ly = T + 14 * i + 8
        o.append(f'<line x1="{W - R + 10}" y1="{ly}" x2="{W - R + 34}" y2="{ly}" stroke="{c}" stroke-width="1.6"{dash}/>')
        o.append(f'<text x="{W - R + 40}" y="{ly + 4}">{label}</text>')
    o.append("</svg>")
    with open(path, "w") as fh:
        fh.write("\n".join(o))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="+")
    ap.add_argument("--out", default="profiles/figures")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    runs = []
    for path in args.files:
        with open(path) as f:
            d = json.load(f)
        n = d.get("_meta", {}).get("n_gpus", 1)
        rows = [v for k, v in d.items() if not k.startswith("_")]
        runs.append((n, rows))
    runs.sort(key=lambda r: r[0])

    def series(metric, on):
        out = []
        for n, rows in runs:
            sel = sorted([r for r in rows if r["kv_cache_enabled"] == on], key=lambda r: r["sequence_length"])
            if sel:
                out.append((f"KV {'on' if on else 'off'}, {n} GPU" + ("s" if n > 1 else ""),
                            [r["sequence_length"] for r in sel], [r[metric]["mean"] for r in sel],
                            [r[metric].get("ci_95", 0.0) for r in sel]))
        return out

    line_chart(os.path.join(args.out, "fig1_latency.svg"), "Steady-state latency vs sequence length", "output tokens",
               "ms / token (log)", series("steady_state_ms_per_token", True) + series("steady_state_ms_per_token", False), logy=True)
    line_chart(os.path.join(args.out, "fig2_throughput.svg"), "Throughput vs sequence length", "output tokens",
               "tokens / s (log)", series("steady_state_tps", True) + series("steady_state_tps", False), logy=True)
    line_chart(os.path.join(args.out, "fig3_memory.svg"), "Peak decode-phase memory vs sequence length", "output tokens",
               "MB per GPU", series("peak_memory_mb", True) + series("peak_memory_mb", False))
    speed = []
    for n, rows in runs:
        on = {r["sequence_length"]: r for r in rows if r["kv_cache_enabled"]}
        off = {r["sequence_length"]: r for r in rows if not r["kv_cache_enabled"]}
        ls = sorted(set(on) & set(off))
        if ls:
            speed.append((f"{n} GPU" + ("s" if n > 1 else ""), ls,
                          [off[x]["steady_state_ms_per_token"]["mean"] / on[x]["steady_state_ms_per_token"]["mean"] for x in ls],
                          [0.0] * len(ls)))
    line_chart(os.path.join(args.out, "fig4_speedup.svg"), "KV-cache speed-up (latency off / on)", "output tokens", "x", speed)
    print("wrote", sorted(os.listdir(args.out)))


if __name__ == "__main__":
    main()
