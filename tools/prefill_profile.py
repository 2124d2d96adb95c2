#!/usr/bin/env python
"""Timing of one cache-off step (SigLIP + projector + full-prefix recompute) and of the batched
vision encode, full-size bf16 model with GPU-generated random weights (timing only)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
from pg_b200 import synth  # noqa: E402
from pg_b200.engine import PaliGemmaEngine  # noqa: E402
from kernel_sweep import gpu_weights  # noqa: E402


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    gpu, wall = [], []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3); gpu.append(e0.elapsed_time(e1))
    return round(sorted(gpu)[len(gpu) // 2], 3), round(sorted(wall)[len(wall) // 2], 3)


def main():
    cfg = synth.CONFIGS["paligemma-3b-pt-224"]
    eng = PaliGemmaEngine(cfg, gpu_weights(cfg, torch.bfloat16))
    ids = synth.synth_prompt_ids(cfg).cuda()
    pix1 = synth.synth_pixels(cfg, 1).cuda()
    vb = int(os.environ.get("VISION_BATCH", "64"))
    pixb = torch.rand(vb, 3, 224, 224, device="cuda") * 2 - 1
    res = {}
    if os.environ.get("ONLY_PREFILL", "0") == "1":
        feats = torch.randn(1, 256, 2048, device="cuda", dtype=torch.bfloat16)
        res["text_prefill_260_last_ms"] = timed(lambda: eng.text_forward(ids, feats, None, logits="last"))
        print(json.dumps(res))
        return
    if os.environ.get("ONLY_VISION", "0") == "0":
        feats = eng.encode_images(pix1)
        res["vision_b1_ms(gpu,wall)"] = timed(lambda: eng.encode_images(pix1))
        res["text_prefill_260_last_ms"] = timed(lambda: eng.text_forward(ids, feats, None, logits="last"))
        res["text_prefill_260_all_logits_ms"] = timed(lambda: eng.text_forward(ids, feats, None, logits="all"))
    g, w = timed(lambda: eng.encode_images(pixb), reps=3)
    res[f"vision_b{vb}_ms(gpu,wall)"] = (g, w)
    res[f"vision_b{vb}_img_per_s"] = round(vb / (g / 1e3), 1)
    res["vision_tflops"] = round(vb * 2.202e11 / (g / 1e3) / 1e12, 1)
    if os.environ.get("PG_OP_TIMING", "0") == "1":
        eng.op_times()
        eng.encode_images(pixb)
        res["vision_op_ms"] = {k: round(v, 3) for k, v in eng.op_times().items()}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
