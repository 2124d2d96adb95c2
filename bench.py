#!/usr/bin/env python
"""Headline benchmark: PaliGemma-3B-pt-224 decode tokens/s + p50 per-token latency, KV cache on/off.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

One "step" = one decode step (one new token per sequence) of the cached loop of inference.py:55-78 on random-init
weights of the exact PaliGemma-3B-pt-224 shapes, one synthetic 224x224 image + 'caption en' prompt (N=260), bf16,
greedy.  `value` is measured with everything resident in HBM (CUDA-graph replays, CUDA events); `e2e` goes through the
reference-facing API (`model(input_ids=..., kv_cache=...)` per token, ids copied from pinned host memory and the chosen
token read back every step, as inference.py:72 does).  The weight stream (5.0 GB/step) is far larger than the 126 MB
L2, so no explicit L2 flush is needed between steps.

N > 1 (default --parallel tp): ONE sequence decoded by the tensor-parallel decoder of the north star (heads / MLP /
vocabulary sharded, the 36 partial sums per token exchanged inside the decode kernels over NVLink peer memory:
csrc/tp_exchange.cuh) -- strong scaling; independent replicas (weak scaling, no data-path collective) are measured too
and reported under `replicas`.  Every N > 1 run checks the tensor-parallel path against the reference's golden greedy
tokens (`tensor_parallel.golden_tokens_match`) and against the single-GPU engine on the benchmarked model.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

MODEL = "paligemma-3b-pt-224"
METRIC = "decode_tokens_per_s"
UNIT = "tokens/s"
FLOP_PER_IMAGE = 2.202e11   # SigLIP So400m/14 + projector per 224x224 image


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tc": float(p["bf16_tflops"]),
                "tc_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "tc": 1590.0, "tc_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic_per_launch():
    """dram read + write bytes per launch of the dominant kernel, from the committed `ncu --set full` capture."""
    for name in ("r02_ncu_full_gateup.csv", "r01_ncu_full_final_gateup.csv"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            with open(path) as f:
                rows = list(csv.reader(f))
            head = rows[0]
            r, w = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
            unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[rows[1][r]]
            vals = [(float(x[r]) + float(x[w])) * unit for x in rows[2:] if x and "decode_gateup" in x[0]]
            if vals:
                return statistics.fmean(vals), f"profiles/{name} ({len(vals)} launches)"
        except Exception:
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during a timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for the one JSON line
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def max_over_ranks(ms: float, world: int) -> float:
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world: int):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def spread(xs):
    return {"min": min(xs), "median": statistics.median(xs), "max": max(xs), "n": len(xs)}


def build_weights_gpu(cfg, dtype):
    """Same-seed random weights generated on each rank's GPU (timing runs with N > 1: identical on every
    rank, no 12 GB CPU checkpoint per process)."""
    from pg_b200 import synth
    g = torch.Generator(device="cuda").manual_seed(1234)
    sd = {}
    for key, shape, kind in synth.state_dict_spec(cfg):
        t = torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * synth._STD[kind]
        if kind == "ln_w":
            t += 1
        sd[key] = t.to(dtype)
    return sd


def build_model(cfg, sd, dtype, tp=None):
    import modeling_gemma as MG
    opts = {} if tp is None else {"tp": tp}
    model = MG.PaliGemmaForConditionalGeneration(MG.PaliGemmaConfig(**cfg), init_weights=False, **opts)
    for key, t in sd.items():
        if key.endswith("lm_head.weight"):
            continue
        mod, _, leaf = key.rpartition(".")
        getattr(model.get_submodule(mod), leaf).data = t.to(device="cuda", dtype=dtype)
    model.tie_weights()
    return model.eval()


# ----------------------------------------------------------------------------------------- CPU arm
def host_threads() -> int:
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm uses every core the process may run on
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    return torch.get_num_threads()


def cpu_reference_decode(cfg, sd_cpu, ids, pix, warmup: int, steps: int):
    """The reference's CPU implementation of the path, fp32, all host threads: prefill once, then time `steps` cached
    greedy steps of the loop of inference.py:55-78 (pixel_values=None after the first call, as
    ablation_study_fixed.py:243).  Runs the reference's OWN modeling_gemma.py / modeling_siglip.py when oracle/_ref holds
    them (staged by oracle/build_ref.py; kind "reference"), else the oracle restatement (kind "port")."""
    from oracle import build_ref
    lat = []
    if build_ref.available():
        ref_gemma, _ = build_ref.import_reference()
        config = ref_gemma.PaliGemmaConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()})
        with torch.device("meta"):
            model = ref_gemma.PaliGemmaForConditionalGeneration(config)
        res = model.load_state_dict({k: v for k, v in sd_cpu.items() if "lm_head" not in k}, strict=False, assign=True)
        assert not res.unexpected_keys and all("lm_head" in k for k in res.missing_keys), res
        model.tie_weights()
        for m in model.modules():       # non-persistent buffers are not in any checkpoint
            if hasattr(m, "inv_freq") and m.inv_freq.is_meta:
                m.inv_freq = 1.0 / (m.base ** (torch.arange(0, m.dim, 2, dtype=torch.int64).float() / m.dim))
            if hasattr(m, "position_ids") and m.position_ids.is_meta:
                m.position_ids = torch.arange(m.num_positions).expand((1, -1))
        model = model.eval()
        kv = ref_gemma.KVCache()
        mask = torch.ones_like(ids)
        with torch.no_grad():
            out = model(input_ids=ids, pixel_values=pix, attention_mask=mask, kv_cache=kv)
            cur = out["logits"][:, -1].argmax(-1, keepdim=True)
            for i in range(warmup + steps):
                mask = torch.cat([mask, torch.ones((ids.shape[0], 1))], -1)
                t0 = time.perf_counter()
                out = model(input_ids=cur, pixel_values=None, attention_mask=mask, kv_cache=out["kv_cache"])
                cur = out["logits"][:, -1].argmax(-1, keepdim=True)
                if i >= warmup:
                    lat.append(time.perf_counter() - t0)
        return lat, "reference", "the reference's own modeling_gemma.py / modeling_siglip.py (oracle/_ref), fp32, torch CPU ops"
    from oracle import paligemma_oracle as O
    kv = O.OracleKV()
    mask = torch.ones_like(ids)
    with torch.no_grad():
        lg = O.forward(sd_cpu, cfg, ids, pix, mask, kv, True)[:, -1]
        cur = lg.argmax(-1, keepdim=True)
        for i in range(warmup + steps):
            mask = torch.cat([mask.float(), torch.ones((ids.shape[0], 1))], -1)
            t0 = time.perf_counter()
            lg = O.forward(sd_cpu, cfg, cur, None, mask, kv, True)[:, -1]
            cur = lg.argmax(-1, keepdim=True)
            if i >= warmup:
                lat.append(time.perf_counter() - t0)
    return lat, "port", "oracle/paligemma_oracle.py (restatement of the reference), fp32, torch CPU ops"


def workload_string(B, N):
    return (f"{MODEL} random-init, batch {B}, 1 synthetic 224x224 image + 'caption en' prompt (N={N}), greedy cached decode "
            f"(BASELINE.json configs[0])")


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, same workload."""
    if rank != 0:
        return
    cores = host_threads()
    from pg_b200 import synth
    cfg = synth.CONFIGS[MODEL]
    sd = synth.synth_state_dict(cfg, tie=True)
    ids, pix = synth.synth_prompt_ids(cfg, batch=args.batch), synth.synth_pixels(cfg, batch=args.batch)
    lat, kind, what = cpu_reference_decode(cfg, sd, ids, pix, args.warmup, args.steps)
    total = sum(lat)
    val = args.batch * len(lat) / total
    sample = f"{len(lat)} cached decode steps after a {ids.shape[1]}-token prefill: {what}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(lat), "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.batch, ids.shape[1]), "kv_cache": True},
        "p50_ms_per_token": 1e3 * statistics.median(lat),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"cpu_count": os.cpu_count(), "torch_threads": cores},
    }
    emit(line)


# ----------------------------------------------------------------------------------------- GPU measurements
def time_dominant_kernel(eng, reps: int = 3):
    """Average duration of one decode_gateup launch (RMSNorm + gate/up GEMV + GeGLU, 134 MB of bf16
    weights per launch = 48 % of the step's bytes), cycling through all layers' weights so no launch
    finds its weights in L2.  Back-to-back launches overlap their ramps through programmatic dependent launch, as they
    do inside the step: this is the kernel's THROUGHPUT; the isolated duration is the ncu figure in profiles/."""
    from pg_b200 import _cabi as cabi
    d = eng.dims
    x = torch.randn(1, d.D, device="cuda").to(eng.dtype)
    F_l = eng.F_l
    out = torch.empty(1, F_l, dtype=eng.dtype, device="cuda")
    L, st = cabi.lib(), cabi.stream()

    def sweep():
        for w in eng.t_layers:
            cabi.check(L.pg_decode_gateup(out.data_ptr(), x.data_ptr(), w["ln2"].data_ptr(), w["gu"].data_ptr(), 1,
                                          d.D, F_l, d.eps, None, None, eng.dt, st))
    sweep()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sweep()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * len(eng.t_layers))
    # isolated launches: a synchronisation between launches, no overlap with a neighbour
    iso = []
    for w in eng.t_layers[:8]:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        cabi.check(L.pg_decode_gateup(out.data_ptr(), x.data_ptr(), w["ln2"].data_ptr(), w["gu"].data_ptr(), 1,
                                      d.D, F_l, d.eps, None, None, eng.dt, st))
        b.record()
        torch.cuda.synchronize()
        iso.append(a.elapsed_time(b))
    esize = torch.tensor([], dtype=eng.dtype).element_size()
    bytes_per_launch = esize * (2 * F_l * d.D + 2 * d.D) + esize * F_l
    return ms, bytes_per_launch, statistics.median(iso)


def timed_decode(eng, ids_d, pix_d, B, K, W, world, local, sample=None, dp_vision=False):
    """Prefill (untimed), W warm-up graph replays, then K timed replays with one CUDA event per step.
    Returns (total_ms max over ranks, per-step ms list of this rank, launches per step, clocks, tokens)."""
    from pg_b200 import _cabi as cabi
    N = ids_d.shape[1]
    kv = eng.new_kv(B)
    try:
        kv.reserve(N + W + K + 8)
        with torch.no_grad():
            feats = eng.encode_images_dp(pix_d) if dp_vision else eng.encode_images(pix_d)
            logits = eng.text_forward(ids_d, feats, kv, logits="last")
        first = logits[:, -1].argmax(-1)
        ds = eng.decode_state(B)
        ds.want_full_logits = False
        ds.bind(kv, first, position=N + 1)
        c0 = cabi.launch_count()
        ds.run_steps(kv, 1, sample=sample)                      # captures the graph (+1 warm-up launch set)
        launches_per_step = (cabi.launch_count() - c0) // 2
        for _ in range(max(W - 1, 0)):
            ds.run_steps(kv, 1, sample=sample)
        barrier(world)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        g = ds.graphs[(kv.page_table.data_ptr(), kv.kv_len.data_ptr(), kv.max_pages, sample, False)]
        with ClockSampler(local) as clk:
            evs[0].record()
            for i in range(K):
                g.replay()
                evs[i + 1].record()
            torch.cuda.synchronize()
        kv.length += K
        barrier(world)
        total_ms = max_over_ranks(evs[0].elapsed_time(evs[-1]), world)
        per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
        cols = (torch.arange(max(W, 1) + K, device="cuda")) % ds.max_hist
        toks = torch.cat([first[:, None], ds.history[:, cols]], 1).cpu()
        eng.check_errors(sync=True)
        return total_ms, per_step, launches_per_step, clk.summary(), toks
    finally:
        kv.release()


def batched_rows(eng, cfg, quick: bool, world: int, local: int, which=(3, 4)):
    """BASELINE.json configs[3] and [4], device-resident (prefill untimed, then CUDA-graph replays timed with CUDA
    events): batch 32 x 256 greedy tokens, and batch 8 with a 320-token prefix x 1024 tokens of top-p sampling
    (temperature 0.8, top_p 0.9: inference.py:92-93) over the paged KV cache.  Tensor parallel when the engine is."""
    from pg_b200 import synth
    rows = {}
    specs = {3: ("configs[3] batch 32 x 256 tokens, greedy", 32, None, 64 if quick else 256, None),
             4: ("configs[4] batch 8, 256 image + 64 prefix tokens, 1024 tokens, paged KV, top-p", 8, 64,
                 128 if quick else 1024, (0.8, 0.9, 1234))}
    for c in which:
        name, B, prefix_len, new_tokens, sample = specs[c]
        ids = synth.synth_prompt_ids(cfg, batch=B, prefix_len=prefix_len).cuda()
        pix = synth.synth_pixels(cfg, batch=B).cuda()
        N = ids.shape[1]
        n = new_tokens - 4
        total_ms, per_step, lps, clk, toks = timed_decode(eng, ids, pix, B, n, 4, world, local, sample=sample,
                                                          dp_vision=world > 1)
        esize = 2
        ctx_mid = N + 4 + n // 2
        step_bytes = eng.weight_bytes_per_decode_step() + B * eng.dims.L * 2 * eng.dims.nkv * eng.dims.hd * esize * ctx_mid
        rows[name] = {"batch": B, "prompt_len": N, "timed_steps": n, "tokens_per_s": B * n / (total_ms / 1e3),
                      "ms_per_step": total_ms / n, "p50_ms_per_step": statistics.median(per_step),
                      "context": f"{N + 4} -> {N + 4 + n}", "launches_per_step": lps,
                      "bytes_per_rank_per_step": step_bytes,
                      "hbm_frac_per_rank": step_bytes / (total_ms / n * 1e-3) / 1e9 / measured_peaks()["hbm"],
                      "distinct_tokens_sampled": int(toks.unique().numel()),
                      "parallelism": f"tp{world}" if world > 1 else "single GPU",
                      "sampling": "greedy" if sample is None else f"temperature {sample[0]}, top_p {sample[1]} (pg_top_p_sample)"}
    return rows


def tp_golden_check(tp, rank):
    """Driver-visible parity of the tensor-parallel path on the real GPUs of this run: fp32 verification mode, the
    `small` / `tiny` shapes, greedy tokens through the graph-captured decode loop must equal the golden tokens the
    unmodified reference produced (tests/golden/*.npz)."""
    import numpy as np
    import modeling_gemma as MG
    from pg_b200 import synth
    out = {}
    for name in ("tiny", "small"):
        cfg = synth.CONFIGS[name]
        path = os.path.join(ROOT, "tests", "golden", f"{name}_fp32.npz")
        if cfg["text_config"]["num_attention_heads"] % tp.size or not os.path.exists(path):
            continue
        g = np.load(path)
        model = MG.PaliGemmaForConditionalGeneration(MG.PaliGemmaConfig(**cfg), init_weights=False, tp=tp)
        model.load_state_dict(synth.synth_state_dict(cfg, tie=False), strict=False)
        model.tie_weights()
        model = model.to("cuda").eval()
        ids, pix = synth.synth_prompt_ids(cfg).cuda(), synth.synth_pixels(cfg).cuda()
        steps = g["cached_tokens"].shape[1]
        toks = model.generate(ids, pix, steps).cpu().tolist()
        eng = model._engine_ready()
        out[name] = {"match": toks == g["cached_tokens"].tolist(), "tokens": steps,
                     "exchange": "peer-memory" if eng.fabric is not None else "nccl"}
        del model, eng
    flag = torch.tensor([1 if out and all(v["match"] for v in out.values()) else 0], device="cuda")
    import torch.distributed as dist
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item()), out


def run_ours(args, rank, world, local):
    from pg_b200 import synth
    from pg_b200.dist import TP
    import modeling_gemma as MG
    cfg = synth.CONFIGS[MODEL]
    dtype = torch.bfloat16
    B, K, W = args.batch, args.steps, args.warmup
    pk = measured_peaks()
    use_tp = world > 1 and args.parallel == "tp"
    t0 = time.time()
    if world > 1:
        sd = build_weights_gpu(cfg, dtype)
        sd_cpu = None
    else:
        sd_cpu = synth.synth_state_dict(cfg, tie=True)
        sd = sd_cpu
    model = build_model(cfg, sd, dtype)                          # single-GPU engine (every rank: the replica)
    t_weights = time.time() - t0
    eng = model._engine_ready()
    d = eng.dims
    ids = synth.synth_prompt_ids(cfg, batch=B)
    pix = synth.synth_pixels(cfg, batch=B)
    N = ids.shape[1]
    ids_d, pix_d = ids.cuda(), pix.cuda()

    # ---------------- replicas: device-resident decode on every rank's own copy of the model
    rep_ms, rep_steps, rep_lps, rep_clk, rep_toks = timed_decode(eng, ids_d, pix_d, B, K, W, world, local)
    rep_value = world * B * K / (rep_ms / 1e3)

    # ---------------- tensor parallel (N > 1): ONE sequence over all GPUs
    tp_block = model_tp = eng_tp = None
    if world > 1:
        golden_ok, golden_detail = tp_golden_check(TP(rank, world, None), rank)
        tp = TP(rank, world, None)
        model_tp = build_model(cfg, {k: v.data for k, v in model.state_dict().items()}, dtype, tp)
        eng_tp = model_tp._engine_ready()
        tp_ms, tp_steps, tp_lps, tp_clk, tp_toks = timed_decode(eng_tp, ids_d, pix_d, B, K, W, world, local, dp_vision=True)
        # greedy tokens of the tensor-parallel run against the single-GPU engine on the benchmarked model (bf16: the
        # sums over ranks are fp32 in a different order, so a near-tie may flip; the prefix that agrees is reported)
        same = (tp_toks == rep_toks)[0].tolist()
        agree = same.index(False) if False in same else len(same)
        tp_block = {"value": B * K / (tp_ms / 1e3), "unit": UNIT, "ms_per_step": tp_ms / K,
                    "p50_ms_per_token": statistics.median(tp_steps), "scaling": "strong", "launches_per_step": tp_lps,
                    "parallelism": (f"tp{world}: q/o by head, gate/up/down by feature, lm_head by vocab, K/V replicated; the 36 "
                                    "partial sums per token travel inside the GEMV kernels over NVLink peer memory "
                                    "(no collective launch), (max,index) pairs through the same exchange")
                    if eng_tp.fabric is not None else f"tp{world} over NCCL collectives (PG_TP_EXCHANGE=nccl)",
                    "bytes_per_rank_per_step": eng_tp.weight_bytes_per_decode_step(),
                    "hbm_frac_per_rank": eng_tp.weight_bytes_per_decode_step() / (tp_ms / K * 1e-3) / 1e9 / pk["hbm"],
                    "golden_tokens_match": golden_ok, "golden_detail": golden_detail,
                    "greedy_prefix_equal_to_single_gpu": f"{agree}/{len(same)} tokens",
                    "lost_peer_flag": bool(eng_tp.fabric.lost_peer()) if eng_tp.fabric is not None else None,
                    "clocks": tp_clk}

    if use_tp:
        total_ms, per_step, launches_per_step, clocks = tp_ms, tp_steps, tp_lps, tp_clk
        value, head_eng, head_model = tp_block["value"], eng_tp, model_tp
    else:
        total_ms, per_step, launches_per_step, clocks = rep_ms, rep_steps, rep_lps, rep_clk
        value, head_eng, head_model = rep_value, eng, model
    streams = 1 if use_tp else world
    ctx_mid = N + W + K // 2

    # ---------------- e2e through the reference-facing API with host buffers
    kvc = MG.KVCache()
    mask = torch.ones((B, N), dtype=torch.int64, device="cuda")
    host_ids = torch.empty((B, 1), dtype=torch.int64).pin_memory()
    host_tok = torch.empty((B, 1), dtype=torch.int64).pin_memory()
    with torch.no_grad():
        out = head_model(input_ids=ids_d, pixel_values=pix_d, attention_mask=mask, kv_cache=kvc)
        nxt = out["logits"][:, -1].argmax(-1, keepdim=True)
        host_ids.copy_(nxt)
        torch.cuda.synchronize()
        e2e_steps = []
        for i in range(W + K):
            if i == W:
                barrier(world)
                t0 = time.perf_counter()
            ts = time.perf_counter()
            cur = host_ids.to("cuda", non_blocking=True)                   # H2D: this step's input ids
            mask = torch.cat([mask, torch.ones((B, 1), dtype=mask.dtype, device="cuda")], -1)
            out = head_model(input_ids=cur, pixel_values=None, attention_mask=mask, kv_cache=kvc)
            nxt = out["logits"][:, -1].argmax(-1, keepdim=True)
            host_tok.copy_(nxt, non_blocking=False)                          # D2H: the step's result
            host_ids.copy_(host_tok)
            if i >= W:
                e2e_steps.append((time.perf_counter() - ts) * 1e3)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(e2e_ms, world)
    e2e = {"value": streams * B * K / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": 8 * B,
           "d2h_bytes_per_step": 8 * B, "ms_per_step": e2e_ms / K, "ms_per_step_spread": spread(e2e_steps),
           "api": "PaliGemmaForConditionalGeneration.forward(input_ids, pixel_values, attention_mask, kv_cache) per token"}
    kvc._paged.release()

    # ---------------- KV-cache-off ablation (configs[1]): full-prefix recompute incl. the vision tower, every rank its own
    kv_off = None
    if args.kv_off_steps > 0:
        with torch.no_grad():
            cur = ids_d
            model.generate(cur, pix_d, 1, use_kv_cache=False)  # warm-up
            barrier(world)
            lat = []
            for t in range(args.kv_off_steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                f = eng.encode_images(pix_d)
                lg = eng.text_forward(cur, f, None, logits="last")
                nxt = lg[:, -1].argmax(-1, keepdim=True)
                e1.record()
                torch.cuda.synchronize()
                lat.append(e0.elapsed_time(e1))
                cur = torch.cat([cur, nxt], 1)
        p50 = max_over_ranks(statistics.median(lat), world)
        kv_off = {"tokens_per_s": world * B / (p50 / 1e3), "p50_ms_per_token": p50, "ms_per_token": spread(lat),
                  "steps": len(lat), "prefix_len": N, "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                  "note": "each step = SigLIP + projector + unmasked recompute of the whole prefix "
                          "(ablation_study_fixed.py:245-251); the rate is 1 / p50"}

    # ---------------- SigLIP + projector batch encode (configs[2]): data parallel over the ranks; 260-token prefill
    vision = prefill = None
    if args.vision_batch > 0:
        with torch.no_grad():
            # prefill first: the batch-64 encode below runs into the power cap and would depress the clocks
            f1 = eng.encode_images(pix_d)
            eng.text_forward(ids_d, f1, None, logits="last")
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.text_forward(ids_d, f1, None, logits="last"); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            pms = statistics.median(ts)
            pflop = 2 * N * 1.98e9 + 4 * d.nq * N * N * d.hd * d.L + 2 * d.V * d.D
            prefill = {"tokens": N, "ms": pms, "ms_spread": spread(ts), "tflops": pflop / (pms / 1e3) / 1e12,
                       "frac_of_bf16_burst_peak": pflop / (pms / 1e3) / 1e12 / pk["tc"],
                       "hbm_floor_ms": eng.weight_bytes_per_decode_step() / (pk["hbm"] * 1e9) * 1e3,
                       "note": "text decoder over the 260-token prompt, last-position logits (weights read once: at the HBM/tensor ridge)"}
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.encode_images(pix_d); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            v1 = statistics.median(ts)
        vb = args.vision_batch
        g = torch.Generator(device="cuda").manual_seed(99)
        pixb = torch.rand(vb, 3, d.S, d.S, device="cuda", generator=g) * 2 - 1      # the same batch on every rank
        enc = (head_eng.encode_images_dp if world > 1 else eng.encode_images)
        with torch.no_grad():
            for _ in range(3):
                enc(pixb)
            barrier(world)
            ts = []
            with ClockSampler(local) as vclk:
                for _ in range(12):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    barrier(world)
                    e0.record(); enc(pixb); e1.record(); torch.cuda.synchronize()
                    ts.append(max_over_ranks(e0.elapsed_time(e1), world))
            vms = statistics.median(ts)
            tf = vb * FLOP_PER_IMAGE / (vms / 1e3) / 1e12
            vision = {"batch": vb, "ms": vms, "ms_spread": spread(ts), "images_per_s": vb / (vms / 1e3), "tflops": tf,
                      "frac_of_bf16_burst_peak_per_gpu": tf / world / pk["tc"],
                      "frac_of_bf16_sustained_peak_per_gpu": tf / world / pk["tc_sustained"],
                      "flop_per_image": FLOP_PER_IMAGE, "peak_tflops": {"burst": pk["tc"], "sustained": pk["tc_sustained"]},
                      "batch1_ms": v1, "clocks": vclk.summary(),
                      "parallelism": (f"data parallel: {vb}/{world} images per GPU + NCCL all-gather of the (B,256,2048) "
                                      "features" if world > 1 else "single GPU")}

    # ---------------- batched decode rows (configs[3], configs[4]): tensor parallel at N > 1
    batched = None
    if args.batched:
        batched = batched_rows(head_eng if world > 1 else eng, cfg, quick=args.batched == 2, world=world, local=local,
                               which=(3, 4) if world == 1 else (3,))

    if rank != 0:
        return
    # ---------------- roofline of the dominant kernel + whole-step accounting
    k_ms, k_bytes, k_iso_ms = time_dominant_kernel(eng)
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic_per_launch()
    esize = 2
    step_bytes = head_eng.weight_bytes_per_decode_step() + B * d.L * 2 * d.nkv * d.hd * esize * ctx_mid
    step_gbs = step_bytes / ((total_ms / K) * 1e-3) / 1e9

    # ---------------- CPU baseline (bounded sample)
    cpu = None
    if args.cpu_steps > 0 and world == 1:
        cores = host_threads()
        lat, kind, what = cpu_reference_decode(cfg, sd_cpu, ids, pix, 1, args.cpu_steps)
        cpu = {"value": B * len(lat) / sum(lat), "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{len(lat)} cached decode steps after a {N}-token prefill: {what}",
               "p50_ms_per_token": 1e3 * statistics.median(lat)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if use_tp else "weak",
        "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_string(B, N) + ", bf16",
                   "kv_cache": True, "context_at_mid_run": ctx_mid,
                   "parallelism": (f"tp{world}: one sequence, tensor-parallel decoder, partial sums exchanged inside the decode "
                                   f"kernels over NVLink peer memory" if use_tp
                                   else (f"replicas x{world}" if world > 1 else "single GPU")),
                   "l2": "inputs (5.0 GB weight stream per step) larger than the 126 MB L2; no flush needed"},
        "p50_ms_per_token": statistics.median(per_step),
        "ms_per_step_spread": spread(per_step),
        "e2e": e2e,
        "gpu_launches": launches_per_step * K,
        "launches_per_step": launches_per_step,
        "clocks": clocks,
        "parity": ("tests/ (pytest -m gpu): fp32 greedy tokens bit-exact vs the reference at full size; bf16 full-size cached "
                   "decode teacher-forced 16 steps + single layers with reference inputs at rtol 2e-2; batched step, vision "
                   "batch 64 and emulated tp 2/4/8 vs the oracle"
                   + ("; this run: tensor_parallel.golden_tokens_match" if world > 1 else "")),
        "roofline": {"bound": "hbm", "kernel": "decode_gateup_kernel (RMSNorm + gate/up GEMV + GeGLU)",
                     "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                     "traffic": traffic, "traffic_source": traffic_src,
                     "bytes_per_launch": k_bytes, "ms_per_launch": k_ms, "peak_source": pk["source"],
                     "isolated": {"ms_per_launch": k_iso_ms, "achieved": k_bytes / (k_iso_ms * 1e-3) / 1e9,
                                  "frac": k_bytes / (k_iso_ms * 1e-3) / 1e9 / pk["hbm"],
                                  "note": "one launch between two synchronisations (no PDL overlap with a neighbour)"},
                     "step": {"algorithmic_bytes": step_bytes, "achieved": step_gbs, "frac": step_gbs / pk["hbm"],
                              "note": "per rank" if use_tp else "whole step"}},
        "cpu_baseline": cpu,
        "kv_off": kv_off,
        "tensor_parallel": tp_block,
        "replicas": {"value": rep_value, "unit": UNIT, "ms_per_step": rep_ms / K, "scaling": "weak",
                     "note": "one independent sequence per GPU, no data-path collective"} if world > 1 else None,
        "vision_encode": vision,
        "prefill": prefill,
        "batched_decode": batched,
        "setup_s": {"synthetic_weights": round(t_weights, 1)},
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE line, the JSON record: whatever else a library writes to file descriptor 1 during the
    run (NCCL's version banner under torchrun, for one) is sent to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--kv-off-steps", type=int, default=16)
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--vision-batch", type=int, default=64)
    ap.add_argument("--batched", type=int, default=1, help="0 skip, 1 full configs[3]/[4] rows, 2 shortened")
    ap.add_argument("--parallel", default="tp", choices=["tp", "replicas"],
                    help="N > 1: what `value` measures. tp = ONE sequence, tensor-parallel decoder (strong scaling, the north "
                         "star's sharding); replicas = one independent sequence per GPU (weak scaling, no data-path "
                         "collective).  Both are measured in every N > 1 run; the other one is reported beside the headline.")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    quiet_stdout()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world, local = dist_setup()
    run_ours(args, rank, world, local)
    if world > 1:
        # graphs that captured NCCL kernels are still alive; communicator teardown can block under them
        sys.stdout.flush()
        barrier(world)
        os._exit(0)


if __name__ == "__main__":
    main()
