#!/usr/bin/env python
"""KV-cache on/off ablation sweep in the reference's protocol and output schema (ablation_study_fixed.py:168-287,
473-517 -> ablation_results/summary_statistics.json): for every output length, per-token latency of the cached loop
and of the cache-off loop (SigLIP + projector + full-prefix recompute per token), steady state from token 32 on
(`:23,210`; the whole run for shorter sequences), peak device memory.  Random-init weights of the exact
PaliGemma-3B-pt-224 shapes, bf16, one synthetic image + 'caption en'.

  python tools/ablation_sweep.py [--lengths 16,32,64,128,256] [--runs 2] [--out gpurun_out/ablation.json]
"""
import argparse
import json
import math
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
from pg_b200 import synth  # noqa: E402
from pg_b200.engine import PaliGemmaEngine  # noqa: E402
from kernel_sweep import gpu_weights  # noqa: E402

STEADY_FROM = 32


def stats(xs):
    m = statistics.fmean(xs)
    sd = statistics.stdev(xs) if len(xs) > 1 else 0.0
    return {"mean": round(m, 3), "ci_95": round(1.96 * sd / math.sqrt(len(xs)), 3), "std": round(sd, 3)}


def timed_tokens(step_fn, n_tokens):
    """ms per token, each token bracketed by CUDA events (the harness synchronises around every token)."""
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_tokens + 1)]
    evs[0].record()
    for t in range(n_tokens):
        step_fn(t)
        evs[t + 1].record()
    torch.cuda.synchronize()
    return [evs[t].elapsed_time(evs[t + 1]) for t in range(n_tokens)]


@torch.no_grad()
def run_cached(eng, ids, pix, L):
    kv = eng.new_kv(1)
    try:
        N = ids.shape[1]
        kv.reserve(N + L + 1)
        logits = eng.text_forward(ids, eng.encode_images(pix), kv, logits="last")      # untimed prefill, as the harness
        ds = eng.decode_state(1)
        ds.bind(kv, logits[:, -1].argmax(-1), position=N + 1)
        ds.run_steps(kv, 1)                                                             # graph capture / warm-up token
        return timed_tokens(lambda t: ds.run_steps(kv, 1), L - 1)
    finally:
        kv.release()


@torch.no_grad()
def run_uncached(eng, ids, pix, L):
    cur = [ids]

    def step(t):
        lg = eng.text_forward(cur[0], eng.encode_images(pix), None, logits="last")
        cur[0] = torch.cat([cur[0], lg[:, -1].argmax(-1, keepdim=True)], 1)
    step(0)                                                                             # warm-up token
    return timed_tokens(step, L - 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lengths", default="16,32,64,128,256")
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ablation.json"))
    args = ap.parse_args()
    cfg = synth.CONFIGS["paligemma-3b-pt-224"]
    eng = PaliGemmaEngine(cfg, gpu_weights(cfg, torch.bfloat16))
    ids, pix = synth.synth_prompt_ids(cfg).cuda(), synth.synth_pixels(cfg).cuda()
    out = {}
    for L in [int(x) for x in args.lengths.split(",")]:
        for cached in (True, False):
            tps, mspt, mem = [], [], []
            for _ in range(args.runs):
                torch.cuda.synchronize()
                torch.cuda.reset_peak_memory_stats()
                lat = (run_cached if cached else run_uncached)(eng, ids, pix, L)
                steady = lat[STEADY_FROM:] if len(lat) > STEADY_FROM + 4 else lat
                ms = statistics.fmean(steady)
                mspt.append(ms)
                tps.append(1e3 / ms)
                mem.append(torch.cuda.max_memory_allocated() / 2 ** 20)
            out[("kv_cache_%d" if cached else "no_kv_cache_%d") % L] = {
                "sequence_length": L, "kv_cache_enabled": cached, "num_samples": args.runs,
                "steady_state_tps": stats(tps), "steady_state_ms_per_token": stats(mspt),
                "peak_memory_mb": stats(mem), "tokens_generated": {"mean": float(L)}}
            print(L, "kv on " if cached else "kv off", "%.1f tok/s  %.3f ms/token  %.0f MB" % (stats(tps)["mean"], stats(mspt)["mean"], stats(mem)["mean"]), flush=True)
    out["_meta"] = {"device": torch.cuda.get_device_name(0), "dtype": "bf16", "prompt_len": int(ids.shape[1]),
                    "weights": "random-init, exact PaliGemma-3B-pt-224 shapes", "steady_state_from_token": STEADY_FROM,
                    "memory": "torch.cuda.max_memory_allocated (weights 5.8 GB + 65536-token paged KV pool 1.2 GB + activations)"}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
