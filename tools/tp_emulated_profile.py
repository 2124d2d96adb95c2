#!/usr/bin/env python
"""The tensor-parallel decode step on ONE GPU for profiling: N rank engines of the full-size model share the device and
advance one kernel at a time (dist.LockstepGroup), so every kernel runs with its real shard shapes and the exchange
fused in (producer stores into all N exchange buffers, consumer prologues poll and sum them) -- the peer buffers are
local here, so durations exclude the NVLink hop; ncu cannot wrap a multi-rank job.

  TP=8 STEPS=3 python tools/tp_emulated_profile.py            (wrap in ncu for the launch list / --set full captures)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from pg_b200 import synth  # noqa: E402
from pg_b200.dist import TP, Fabric, LockstepGroup  # noqa: E402
from pg_b200.engine import PaliGemmaEngine  # noqa: E402


@torch.no_grad()
def main():
    tp, steps, B = int(os.environ.get("TP", "8")), int(os.environ.get("STEPS", "3")), int(os.environ.get("BATCH", "1"))
    cfg = synth.CONFIGS["paligemma-3b-pt-224"]
    torch.cuda.set_device(0)
    sd = bench.build_weights_gpu(cfg, torch.bfloat16)
    fabs = Fabric.emulated(tp, cfg["text_config"]["hidden_size"], "cuda")
    engines = [PaliGemmaEngine(cfg, sd, tp=TP(r, tp, fabric=fabs[r], emulated=True), kv_pool_tokens=4096) for r in range(tp)]
    del sd
    group = LockstepGroup()
    ids = synth.synth_prompt_ids(cfg, batch=B).cuda()
    pix = synth.synth_pixels(cfg, batch=B).cuda()
    N = ids.shape[1]
    feats = engines[0].encode_images(pix)
    kvs = [e.new_kv(B) for e in engines]
    for kv in kvs:
        kv.reserve(N + steps + 2)
    lg = group.run([e.text_forward_gen(ids, feats, kv, logits="last") for e, kv in zip(engines, kvs)])
    first = lg[0][:, -1].argmax(-1)
    dss = [e.decode_state(B) for e in engines]
    for ds, kv in zip(dss, kvs):
        ds.bind(kv, first, position=N + 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        group.run([ds.step_gen(kv) for ds, kv in zip(dss, kvs)])
        for kv in kvs:
            kv.length += 1
    torch.cuda.synchronize()
    toks = dss[0].history[:, :steps].tolist()
    assert all(ds.history[:, :steps].tolist() == toks for ds in dss), "ranks disagree"
    assert not any(e.fabric.lost_peer() for e in engines)
    print(f"tp{tp} emulated on one GPU, batch {B}: {steps} steps, tokens {toks}, "
          f"{(time.perf_counter() - t0) * 1e3 / steps:.2f} ms/step wall (launch-bound: {tp} ranks, no graph)")


if __name__ == "__main__":
    main()
