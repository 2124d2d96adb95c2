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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
from pg_b200 import synth  # noqa: E402
from pg_b200.dist import TP  # noqa: E402
import modeling_gemma as MG  # noqa: E402


def main():
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
    # bf16, small shapes: GEMV step (batch 2) and tensor-core step (batch 6), teacher-forced against the CPU oracle
    from oracle import paligemma_oracle as O
    from pg_b200.generate import generate
    cfg = synth.CONFIGS["small"]
    if cfg["text_config"]["num_attention_heads"] % world == 0:
        sd32 = synth.synth_state_dict(cfg)
        sdb = {k: v.to(torch.bfloat16) for k, v in sd32.items()}
        model = MG.PaliGemmaForConditionalGeneration(MG.PaliGemmaConfig(**cfg), init_weights=False, tp=tp)
        model.load_state_dict({k: v for k, v in sdb.items() if "lm_head" not in k}, strict=False)
        model.tie_weights()
        model = model.to("cuda").eval()
        eng = model._engine_ready()
        for batch in (2, 6):
            ids = synth.synth_prompt_ids(cfg, batch=batch, prefix_len=7)
            pix = synth.synth_pixels(cfg, batch=batch)
            steps = 5
            ref_t, ref_lg = O.generate_cached(sdb, cfg, ids, pix.to(torch.bfloat16), steps + 1, patched=True, return_logits=True)
            _, truth = O.generate_cached(sd32, cfg, ids, pix, steps + 1, patched=True, return_logits=True, teacher=ref_t)
            kv = eng.new_kv(batch)
            kv.reserve(ids.shape[1] + steps + 1)
            with torch.no_grad():
                got = [eng.text_forward(ids.cuda(), eng.encode_images_dp(pix.cuda()), kv, logits="last")[:, -1].cpu()]
            ds = eng.decode_state(batch)
            ds.bind(kv, ref_t[:, 0].cuda(), position=ids.shape[1] + 1)
            ds.want_full_logits = True
            for t in range(steps):
                ds.ids.copy_(ref_t[:, t].cuda())
                ds.run_steps(kv, 1)
                got.append(ds.logits.cpu().clone())
            kv.release()
            got = torch.stack(got, 1)
            rms = lambda x: float(x.float().pow(2).mean().sqrt())
            ref_noise, our_noise, delta = rms(ref_lg - truth), rms(got - truth), rms(got - ref_lg)
            good = our_noise <= 1.5 * ref_noise + 1e-3 * rms(truth) and delta <= 2.5 * ref_noise + 1e-3 * rms(truth)
            ok &= good
            print(f"[rank {rank}] small bf16 batch {batch} tp={world}: ours-vs-fp32 {our_noise:.3g}, reference-bf16-vs-fp32 "
                  f"{ref_noise:.3g}, ours-vs-reference {delta:.3g} -> {'OK' if good else 'FAIL'}", flush=True)
    fab = eng.fabric if "eng" in dir() else None
    if fab is not None:
        torch.cuda.synchronize()
        ok &= not fab.lost_peer()
        print(f"[rank {rank}] peer-memory exchange used: {int(fab.epoch.item())} decode steps, lost-peer flag "
              f"{int(fab.lost_peer())}", flush=True)
    else:
        print(f"[rank {rank}] NCCL collective path", flush=True)
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
