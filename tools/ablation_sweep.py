#!/usr/bin/env python
"""KV-cache on/off ablation sweep in the reference's protocol and output schema (ablation_study_fixed.py:168-287,
473-517 -> ablation_results/summary_statistics.json): for every output length, per-token latency of the cached loop
and of the cache-off loop (SigLIP + projector + full-prefix recompute per token), steady state from token 32 on
(`:23,210`; the whole run for shorter sequences), peak decode-phase device memory (statistics reset after the prefill,
`:202,255`).  Random-init weights of the exact PaliGemma-3B-pt-224 shapes, bf16, one synthetic image + 'caption en'.

  python tools/ablation_sweep.py [--lengths 16,32,64,128,256] [--runs 2] [--out gpurun_out/ablation_n1.json]
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ablation_sweep.py ...      (N = 2, 4, 8)

N > 1: the cache-on loop runs ONE sequence on the tensor-parallel decoder (latency falls with N); the cache-off loop has
no cross-GPU dependency worth sharding at batch 1 (the whole prefix is recomputed: compute-bound prefill kernels), so
every GPU recomputes its own sequence and the job's throughput is N x the per-GPU rate at the same latency.

Memory axis: the reference's KVCache grows by torch.cat, so its allocator peak tracks the cache; this engine
pre-allocates a paged pool.  `peak_memory_mb` is therefore the allocator peak of the decode phase MINUS the part of the
pool no live sequence holds pages in (`kv_pool_idle_mb`): weights + activations + the pages in use, the quantity the
reference's number measures.  Raw allocator peak and the pool size are reported beside it.
"""
import argparse
import json
import math
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pg_b200 import synth  # noqa: E402
from pg_b200.dist import TP  # noqa: E402
from pg_b200.engine import PaliGemmaEngine  # noqa: E402
import bench  # noqa: E402

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


class Peak:
    """Decode-phase memory the way the reference measures it (reset after the prefill), pool-aware."""

    def __init__(self, eng):
        self.eng, self.kv_peak = eng, 0

    def reset(self):
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        self.kv_peak = self.eng.kv_bytes_in_use()

    def touch(self):
        self.kv_peak = max(self.kv_peak, self.eng.kv_bytes_in_use())

    def result(self):
        raw = torch.cuda.max_memory_allocated()
        idle = self.eng.kv_pool_bytes() - self.kv_peak
        return {"peak_memory_mb": (raw - idle) / 2 ** 20, "allocator_peak_mb": raw / 2 ** 20,
                "kv_cache_mb": self.kv_peak / 2 ** 20, "kv_pool_idle_mb": idle / 2 ** 20}


@torch.no_grad()
def run_cached(eng, ids, pix, L, peak):
    kv = eng.new_kv(1)
    try:
        N = ids.shape[1]
        kv.reserve(N + L + 1)
        logits = eng.text_forward(ids, eng.encode_images(pix), kv, logits="last")      # untimed prefill, as the harness
        ds = eng.decode_state(1)
        ds.want_full_logits = False
        ds.bind(kv, logits[:, -1].argmax(-1), position=N + 1)
        ds.run_steps(kv, 1)                                                             # graph capture / warm-up token
        peak.reset()
        lat = timed_tokens(lambda t: ds.run_steps(kv, 1), L - 1)
        peak.touch()
        return lat
    finally:
        kv.release()


@torch.no_grad()
def run_uncached(eng, ids, pix, L, peak):
    cur = [ids]

    def step(t):
        kv = eng.new_kv(cur[0].shape[0])               # the cache-off forward still needs K/V for its own attention
        try:
            lg = eng.text_forward(cur[0], eng.encode_images(pix), kv, logits="last")
            peak.touch()
        finally:
            kv.release()
        cur[0] = torch.cat([cur[0], lg[:, -1].argmax(-1, keepdim=True)], 1)
    step(0)                                                                             # warm-up token
    peak.reset()
    return timed_tokens(step, L - 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lengths", default="16,32,64,128,256")
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = bench.dist_setup()
    cfg = synth.CONFIGS["paligemma-3b-pt-224"]
    sd = bench.build_weights_gpu(cfg, torch.bfloat16)
    pool = dict(kv_pool_tokens=8192)                                # one sequence of <= 516 + 256 tokens at a time
    eng = PaliGemmaEngine(cfg, sd, **pool)                          # every rank's own copy (cache-off loop)
    eng_tp = PaliGemmaEngine(cfg, sd, tp=TP(rank, world, None), **pool) if world > 1 else eng
    del sd                                                          # the engines hold (fused / sharded) copies of what they use
    torch.cuda.empty_cache()
    ids, pix = synth.synth_prompt_ids(cfg).cuda(), synth.synth_pixels(cfg).cuda()
    weights_mb = eng_tp.weight_bytes_per_decode_step() / 2 ** 20
    out = {}
    for L in [int(x) for x in args.lengths.split(",")]:
        for cached in (True, False):
            e = eng_tp if cached else eng
            tps, mspt, mem = [], [], []
            for _ in range(args.runs):
                peak = Peak(e)
                bench.barrier(world)
                lat = (run_cached if cached else run_uncached)(e, ids, pix, L, peak)
                steady = lat[STEADY_FROM:] if len(lat) > STEADY_FROM + 4 else lat
                ms = bench.max_over_ranks(statistics.fmean(steady), world)
                mspt.append(ms)
                tps.append((1 if cached else world) * 1e3 / ms)
                mem.append(peak.result())
            key = ("kv_cache_%d" if cached else "no_kv_cache_%d") % L
            out[key] = {
                "sequence_length": L, "kv_cache_enabled": cached, "num_samples": args.runs,
                "steady_state_tps": stats(tps), "steady_state_ms_per_token": stats(mspt),
                "peak_memory_mb": stats([m["peak_memory_mb"] for m in mem]),
                "kv_cache_mb": stats([m["kv_cache_mb"] for m in mem]),
                "allocator_peak_mb": stats([m["allocator_peak_mb"] for m in mem]),
                "tokens_generated": {"mean": float(L)},
                "parallelism": (f"tp{world}" if cached else f"replicas x{world}") if world > 1 else "single GPU"}
            if rank == 0:
                print(L, "kv on " if cached else "kv off", "%.1f tok/s  %.3f ms/token  %.0f MB (kv %.1f MB)" % (
                    stats(tps)["mean"], stats(mspt)["mean"], out[key]["peak_memory_mb"]["mean"], out[key]["kv_cache_mb"]["mean"]),
                    flush=True)
    out["_meta"] = {"device": torch.cuda.get_device_name(0), "n_gpus": world, "dtype": "bf16", "prompt_len": int(ids.shape[1]),
                    "weights": "random-init, exact PaliGemma-3B-pt-224 shapes", "steady_state_from_token": STEADY_FROM,
                    "weights_mb_per_rank_decode": round(weights_mb, 1), "kv_pool_mb": round(eng.kv_pool_bytes() / 2 ** 20, 1),
                    "memory": "peak_memory_mb = torch.cuda.max_memory_allocated of the decode phase minus the idle part of the "
                              "pre-allocated paged KV pool (see the module docstring); with N > 1 each rank holds the "
                              "replicated model for the cache-off loop AND its tensor-parallel shard"}
    if rank == 0:
        path = args.out or os.path.join(ROOT, "gpurun_out", f"ablation_n{world}.json")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
    if world > 1:
        sys.stdout.flush()
        bench.barrier(world)
        os._exit(0)


if __name__ == "__main__":
    main()
