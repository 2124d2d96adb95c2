"""Tensor-parallel parity on real GPUs (run under torchrun, one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/tp_check.py
Every rank must reproduce the reference's greedy tokens (golden vectors, fp32) through both the
public forward API and the graph-captured generate loop; logits within 1e-4 of the scale."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
from pg_b200 import synth  # noqa: E402
from pg_b200.dist import TP  # noqa: E402
import modeling_gemma as MG  # noqa: E402


def main():
    os.environ.setdefault("PG_TP_ALLREDUCE", "oneshot")  # exercise the peer-memory all-reduce too (falls back to NCCL)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tp = TP(rank, world, None)
    ok = True
    for name in ("tiny", "small"):
        cfg = synth.CONFIGS[name]
        if cfg["text_config"]["num_attention_heads"] % world:
            continue
        g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}_fp32.npz"))
        for dtype in (torch.float32,):
            model = MG.PaliGemmaForConditionalGeneration(MG.PaliGemmaConfig(**cfg), init_weights=False, tp=tp)
            sd = synth.synth_state_dict(cfg, tie=False)
            model.load_state_dict(sd, strict=False)
            model.tie_weights()
            model = model.to("cuda").eval()
            ids, pix = synth.synth_prompt_ids(cfg).cuda(), synth.synth_pixels(cfg).cuda()
            steps = g["cached_tokens"].shape[1]
            with torch.no_grad():
                out = model(input_ids=ids, pixel_values=pix, attention_mask=torch.ones_like(ids), kv_cache=None)
            lg, want = out["logits"].cpu(), torch.from_numpy(g["prefill_logits_all"])
            err = float((lg - want).abs().max())
            tol = 1e-4 * float(want.abs().max()) * 3
            toks = model.generate(ids, pix, steps).cpu().tolist()
            # public API loop (inference.py)
            kv, cur, mask, api = MG.KVCache(), ids, torch.ones_like(ids), []
            with torch.no_grad():
                for _ in range(steps):
                    o = model(input_ids=cur, pixel_values=pix, attention_mask=mask, kv_cache=kv)
                    cur = o["logits"][:, -1].argmax(-1, keepdim=True)
                    api.append(int(cur))
                    mask = torch.cat([mask, torch.ones((1, 1), dtype=mask.dtype, device="cuda")], -1)
            good = toks == g["cached_tokens"].tolist() and [api] == g["cached_tokens"].tolist() and err <= tol
            ok &= good
            print(f"[rank {rank}] {name} tp={world}: tokens {'OK' if toks == g['cached_tokens'].tolist() else toks} "
                  f"api {'OK' if [api] == g['cached_tokens'].tolist() else api} max|dlogit| {err:.2e} (tol {tol:.2e})", flush=True)
    if tp.oneshot is not None:
        ok &= int(tp.oneshot.err.item()) == 0
        print(f"[rank {rank}] one-shot all-reduce used, {int(tp.oneshot.step.item())} calls, err flag {int(tp.oneshot.err.item())}", flush=True)
    else:
        print(f"[rank {rank}] NCCL all-reduce path", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    passed = int(flag.item()) == 1
    if rank == 0 and passed:
        print("TP_CHECK_PASS", flush=True)
    sys.stdout.flush()
    # CUDA graphs that captured NCCL kernels are still alive: tearing the communicator down under them
    # can block, and there is nothing left to clean up
    os._exit(0 if passed else 1)


if __name__ == "__main__":
    main()
