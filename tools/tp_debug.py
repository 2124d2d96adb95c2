import os, sys, time
import torch, torch.distributed as dist
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200"))
from pg_b200 import synth
from pg_b200.dist import TP
from pg_b200.generate import generate
import modeling_gemma as MG
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def log(*a): print(f"[r{rank} {time.time()%1000:.1f}]", *a, flush=True)
torch.cuda.set_device(local)
log("init pg")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t=torch.ones(4,device="cuda"); dist.all_reduce(t); torch.cuda.synchronize(); log("allreduce ok", t[0].item())
tp = TP(rank, world, None)
cfg = synth.CONFIGS["tiny"]
model = MG.PaliGemmaForConditionalGeneration(MG.PaliGemmaConfig(**cfg), init_weights=False, tp=tp)
model.load_state_dict(synth.synth_state_dict(cfg, tie=False), strict=False); model.tie_weights(); model = model.to("cuda").eval()
ids, pix = synth.synth_prompt_ids(cfg).cuda(), synth.synth_pixels(cfg).cuda()
log("prefill")
with torch.no_grad():
    out = model(input_ids=ids, pixel_values=pix, attention_mask=torch.ones_like(ids), kv_cache=None)
torch.cuda.synchronize(); log("prefill ok", out["logits"].shape)
eng = model._engine_ready()
log("generate no graph")
toks = generate(eng, ids, pix, 4, use_graph=False); torch.cuda.synchronize(); log("nograph ok", toks.tolist())
log("generate graph")
toks = generate(eng, ids, pix, 4, use_graph=True); torch.cuda.synchronize(); log("graph ok", toks.tolist())
dist.destroy_process_group()
