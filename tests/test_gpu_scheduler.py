"""Continuous batching (pg_b200/scheduler.py): requests of different prompt lengths and budgets share one decode
graph through slot rows of a static page table; every request's tokens must equal the CPU ORACLE's greedy loop on that
request alone (the reference loop of inference.py:50-78 run once per request)."""
import pytest
import torch

from oracle import paligemma_oracle as O
from pg_b200 import synth
from pg_b200.generate import generate
from pg_b200.scheduler import ContinuousBatcher
import modeling_gemma as MG

pytestmark = pytest.mark.gpu


def _model(name, dtype):
    cfg = synth.CONFIGS[name]
    model = MG.PaliGemmaForConditionalGeneration(MG.PaliGemmaConfig(**cfg), init_weights=False)
    for key, shape, kind in synth.state_dict_spec(cfg):
        mod, _, leaf = key.rpartition(".")
        getattr(model.get_submodule(mod), leaf).data = synth.synth_tensor(key, shape, kind, w_std=cfg.get("synth_w_std")).to(dtype)
    model.tie_weights()
    return model.to("cuda").eval(), cfg


def _requests(cfg):
    spec = [(None, 6), (5, 11), (9, 3), (14, 9), (7, 1), (6, 13)]
    out = []
    for i, (prefix_len, budget) in enumerate(spec):
        ids = synth.synth_prompt_ids(cfg, batch=1, prefix_len=prefix_len, seed=100 + i)
        pix = synth.synth_pixels(cfg, batch=1, seed=200 + i)
        out.append((ids, pix, budget))
    return out


def test_continuous_batching_equals_per_request_generation_fp32():
    model, cfg = _model("tiny", torch.float32)
    eng = model._engine_ready()
    reqs = _requests(cfg)
    sd = synth.synth_state_dict(cfg)
    want = [O.generate_cached(sd, cfg, ids, pix, budget, patched=True)[0].tolist() for ids, pix, budget in reqs]
    assert want == [generate(eng, ids.cuda(), pix.cuda(), budget).cpu()[0].tolist() for ids, pix, budget in reqs]
    free_before = len(eng._free)
    cb = ContinuousBatcher(eng, slots=3, max_tokens=256, chunk=4)
    rids = [cb.submit(ids, pix, budget) for ids, pix, budget in reqs]
    done = cb.run()
    for rid, w in zip(rids, want):
        assert done[rid].tokens == w, (rid, done[rid].tokens, w)
    assert not cb.running and not cb.queue
    cb.close()
    assert len(eng._free) == free_before                      # every page went back to the pool
    # EOS stops a sequence early and frees its slot for the next request
    eos = want[1][3]
    stop_at = want[1].index(eos) + 1
    cb = ContinuousBatcher(eng, slots=2, max_tokens=256, chunk=3)
    r_eos = cb.submit(reqs[1][0], reqs[1][1], reqs[1][2], eos_token_id=eos)
    r_other = [cb.submit(ids, pix, budget) for ids, pix, budget in (reqs[0], reqs[3], reqs[5])]
    done = cb.run()
    assert done[r_eos].tokens == want[1][:stop_at]
    for rid, i in zip(r_other, (0, 3, 5)):
        assert done[rid].tokens == want[i]
    cb.close()


def test_continuous_batching_on_the_tensor_core_step_bf16():
    """4 slots in bf16 take the batched (skinny-GEMM) decode step: budgets and bookkeeping hold, and the first token
    (from the shared prefill path) equals the per-request run."""
    model, cfg = _model("small", torch.bfloat16)
    eng = model._engine_ready()
    reqs = _requests(cfg)
    cb = ContinuousBatcher(eng, slots=4, max_tokens=512, chunk=5, do_sample=False)
    rids = [cb.submit(ids, pix, budget) for ids, pix, budget in reqs]
    done = cb.run()
    for rid, (ids, pix, budget) in zip(rids, reqs):
        assert len(done[rid].tokens) == budget
        first = generate(eng, ids.cuda(), pix.cuda(), 1).cpu()[0].tolist()
        assert done[rid].tokens[0] == first[0]
        assert all(0 <= t < cfg["vocab_size"] for t in done[rid].tokens)
    cb.close()


def test_batched_prefill_of_equal_length_prompts_fp32():
    """Requests with the same prompt length are prefilled together (one vision batch + one text_forward per group of free
    slots): tokens must still equal the CPU oracle's per-request greedy loop, and fewer prefill calls are spent."""
    model, cfg = _model("tiny", torch.float32)
    eng = model._engine_ready()
    sd = synth.synth_state_dict(cfg)
    reqs = []
    for i, budget in enumerate((7, 4, 9, 5, 6, 3, 8)):
        ids = synth.synth_prompt_ids(cfg, batch=1, prefix_len=8 if i != 3 else 11, seed=300 + i)
        pix = synth.synth_pixels(cfg, batch=1, seed=400 + i)
        reqs.append((ids, pix, budget))
    want = [O.generate_cached(sd, cfg, ids, pix, budget, patched=True)[0].tolist() for ids, pix, budget in reqs]
    free_before = len(eng._free)
    cb = ContinuousBatcher(eng, slots=3, max_tokens=256, chunk=4)
    rids = [cb.submit(ids, pix, budget) for ids, pix, budget in reqs]
    done = cb.run()
    for rid, w in zip(rids, want):
        assert done[rid].tokens == w, (rid, done[rid].tokens, w)
    assert cb.prefill_calls < len(reqs)                       # the first admission alone seats three requests in one call
    cb.close()
    assert len(eng._free) == free_before
