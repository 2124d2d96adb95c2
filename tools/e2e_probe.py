#!/usr/bin/env python
"""Host-side profile of the API decode loop (bench.py's e2e region): cProfile over N cached steps through
model(input_ids=..., kv_cache=...) plus wall-clock per phase.  python tools/e2e_probe.py [small|paligemma-3b-pt-224]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from pg_b200 import synth  # noqa: E402
import modeling_gemma as MG  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "paligemma-3b-pt-224"
    cfg = synth.CONFIGS[name]
    torch.cuda.set_device(0)
    model = bench.build_model(cfg, bench.build_weights_gpu(cfg, torch.bfloat16), torch.bfloat16)
    ids, pix = synth.synth_prompt_ids(cfg).cuda(), synth.synth_pixels(cfg).cuda()
    B, N = ids.shape
    kvc = MG.KVCache()
    mask = torch.ones((B, N), dtype=torch.int64, device="cuda")
    host_ids = torch.empty((B, 1), dtype=torch.int64).pin_memory()
    host_tok = torch.empty((B, 1), dtype=torch.int64).pin_memory()
    steps = 48
    with torch.no_grad():
        out = model(input_ids=ids, pixel_values=pix, attention_mask=mask, kv_cache=kvc)
        host_ids.copy_(out["logits"][:, -1].argmax(-1, keepdim=True))
        torch.cuda.synchronize()
        phases = {"h2d+cat": 0.0, "forward": 0.0, "argmax+d2h": 0.0}
        pr = cProfile.Profile()
        for i in range(8 + steps):
            if i == 8:
                torch.cuda.synchronize()
                t_all = time.perf_counter()
                pr.enable()
            t0 = time.perf_counter()
            cur = host_ids.to("cuda", non_blocking=True)
            mask = torch.cat([mask, torch.ones((B, 1), dtype=mask.dtype, device="cuda")], -1)
            t1 = time.perf_counter()
            out = model(input_ids=cur, pixel_values=None, attention_mask=mask, kv_cache=kvc)
            t2 = time.perf_counter()
            nxt = out["logits"][:, -1].argmax(-1, keepdim=True)
            host_tok.copy_(nxt, non_blocking=False)
            host_ids.copy_(host_tok)
            t3 = time.perf_counter()
            if i >= 8:
                phases["h2d+cat"] += t1 - t0; phases["forward"] += t2 - t1; phases["argmax+d2h"] += t3 - t2
        pr.disable()
        torch.cuda.synchronize()
        total = time.perf_counter() - t_all
    print(f"{name}: {1e3 * total / steps:.3f} ms/step; host phases (ms/step): "
          + ", ".join(f"{k} {1e3 * v / steps:.3f}" for k, v in phases.items()))
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
    print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
