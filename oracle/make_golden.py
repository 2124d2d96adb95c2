"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference) on the seeded synthetic weights/inputs of pg_b200.synth.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no
/root/reference):   python oracle/make_golden.py [tiny] [small] [full] [full_bf16]

The reference ships no golden vectors; these are the vectors the
oracle restatement and the CUDA path are pinned to.
"""
from __future__ import annotations

import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multimodal-financial-analysis-tool-using-paligemma_b200")
REF = os.environ.get("PG_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF)
    import modeling_gemma as ref_gemma  # noqa: the reference's own module
    import modeling_siglip as ref_siglip  # noqa
    import ablation_study_fixed as ref_abl  # noqa  (patched merge / rotary, _sample_top_p)
    sys.path.remove(REF)
    for m in ("modeling_gemma", "modeling_siglip", "processing_paligemma", "ablation_study_fixed"):
        sys.modules.pop(m, None)
    return ref_gemma, ref_siglip, ref_abl


ref_gemma, ref_siglip, ref_abl = _import_reference()
sys.path.insert(0, PKG)
sys.path.insert(0, ROOT)
from pg_b200 import synth  # noqa: E402


def build_reference_model(cfg: dict, dtype, patched: bool):
    config = ref_gemma.PaliGemmaConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()})
    torch.manual_seed(0)
    model = ref_gemma.PaliGemmaForConditionalGeneration(config)
    sd = synth.synth_state_dict(cfg, tie=False)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("lm_head" in m for m in missing), missing
    model.tie_weights()
    model = model.to(dtype).eval()
    if patched:  # exactly what ablation_study_fixed.py:335-342 does
        model._merge_input_ids_with_image_features = types.MethodType(
            ref_abl.patched_merge_input_ids_with_image_features, model)
        for layer in model.language_model.model.layers:
            layer.self_attn.rotary_emb.forward = types.MethodType(
                ref_abl.patched_rotary_forward, layer.self_attn.rotary_emb)
    return model


@torch.no_grad()
def ref_generate_cached(model, ids, pix, steps, refeed=False):
    """The loop of inference.py:50-78 (pixel_values re-passed every step, as :58 does)."""
    mask = torch.ones_like(ids)
    kv = ref_gemma.KVCache()
    if refeed:  # ablation_study_fixed.py:193-199
        model(input_ids=ids, pixel_values=pix, attention_mask=mask, kv_cache=kv)
    toks, logits = [], []
    for _ in range(steps):
        out = model(input_ids=ids, pixel_values=pix, attention_mask=mask, kv_cache=kv)
        kv = out["kv_cache"]
        lg = out["logits"][:, -1, :]
        nxt = torch.argmax(lg, dim=-1, keepdim=True)
        toks.append(nxt)
        logits.append(lg)
        ids = nxt
        mask = torch.cat([mask, torch.ones((mask.shape[0], 1))], dim=-1)
        if refeed:
            pix = None  # ablation_study_fixed.py:243
    return torch.cat(toks, -1), torch.stack(logits, 1), kv


@torch.no_grad()
def ref_generate_uncached(model, ids0, pix, steps):
    """ablation_study_fixed.py:209-251 with kv_cache=None."""
    ids, toks, logits = ids0, [], []
    for _ in range(steps):
        mask = torch.ones_like(ids)
        out = model(input_ids=ids, pixel_values=pix, attention_mask=mask, kv_cache=None)
        assert "kv_cache" not in out
        lg = out["logits"][:, -1, :]
        nxt = torch.argmax(lg, dim=-1, keepdim=True)
        toks.append(nxt)
        logits.append(lg)
        ids = torch.cat([ids, nxt], dim=-1)
    return torch.cat(toks, -1), torch.stack(logits, 1)


def _np(x):
    return x.detach().float().cpu().numpy() if x.is_floating_point() else x.detach().cpu().numpy()


def topk_summary(logits: torch.Tensor, k=8):
    v, i = torch.topk(logits.float(), k, dim=-1)
    return _np(v), _np(i)


def golden_small_model(name: str, steps: int):
    cfg = synth.CONFIGS[name]
    out = {}
    ids = synth.synth_prompt_ids(cfg)
    pix = synth.synth_pixels(cfg)
    # --- unpatched model (inference.py semantics)
    m = build_reference_model(cfg, torch.float32, patched=False)
    with torch.no_grad():
        feats = m.vision_tower(pix)
        out["vision_features"] = _np(feats)
        out["projected"] = _np(m.multi_modal_projector(feats))
        full = m(input_ids=ids, pixel_values=pix, attention_mask=torch.ones_like(ids), kv_cache=None)
        out["prefill_logits_all"] = _np(full["logits"])
    toks, lg, kv = ref_generate_cached(m, ids, pix, steps)
    out["cached_tokens"], out["cached_logits"] = _np(toks), _np(lg)
    out["cached_kv_len"] = np.array(kv.num_items())
    out["cached_k_layer0"] = _np(kv.key_cache[0])
    out["cached_v_last"] = _np(kv.value_cache[-1])
    toks, lg = ref_generate_uncached(m, ids, pix, steps)
    out["uncached_tokens"], out["uncached_logits"] = _np(toks), _np(lg)
    # --- patched model (ablation harness semantics): refeed quirk Q6 and batch>1 decode Q7
    mp = build_reference_model(cfg, torch.float32, patched=True)
    toks, lg, kv = ref_generate_cached(mp, ids, pix, steps, refeed=True)
    out["refeed_tokens"], out["refeed_logits"] = _np(toks), _np(lg)
    out["refeed_kv_len"] = np.array(kv.num_items())
    ids2 = synth.synth_prompt_ids(cfg, batch=3, prefix_len=6)
    pix2 = synth.synth_pixels(cfg, batch=3)
    toks, lg, _ = ref_generate_cached(mp, ids2, pix2, steps)
    out["batch3_tokens"], out["batch3_logits"] = _np(toks), _np(lg)
    # --- pad / image-token feedback (Q8): pad id embeds to zeros
    ids_pad = ids.clone()
    ids_pad[0, -2] = cfg["pad_token_id"]
    with torch.no_grad():
        o = m(input_ids=ids_pad, pixel_values=pix, attention_mask=torch.ones_like(ids), kv_cache=None)
    out["pad_logits_last"] = _np(o["logits"][:, -1])
    # --- nucleus distribution (inference.py:15-24,65-66), temperature 0.8, top_p 0.9
    lg0 = torch.from_numpy(out["cached_logits"][:, 0])
    probs = torch.softmax(lg0 / 0.8, dim=-1)
    ps, idx = torch.sort(probs, dim=-1, descending=True)
    cum = torch.cumsum(ps, dim=-1)
    keep = ~(cum - ps > 0.9)
    out["topp_keep_count"] = _np(keep.sum(-1))
    torch.manual_seed(7)
    draws = torch.cat([ref_abl._sample_top_p(probs.clone(), 0.9) for _ in range(64)], -1)
    nucleus = torch.zeros_like(probs, dtype=torch.bool).scatter_(-1, idx, keep)
    assert bool(nucleus.gather(-1, draws).all())
    out["topp_nucleus_ids"] = _np(idx[0, : int(keep.sum())])
    # --- bf16 run of the same model (rounding-point check for the oracle)
    mb = build_reference_model(cfg, torch.bfloat16, patched=False)
    toks, lg, _ = ref_generate_cached(mb, ids, pix.to(torch.bfloat16), steps)
    out["bf16_cached_tokens"], out["bf16_cached_logits"] = _np(toks), _np(lg)
    np.savez_compressed(os.path.join(OUT, f"{name}_fp32.npz"), **out)
    print(name, "cached", out["cached_tokens"].tolist(), "uncached", out["uncached_tokens"].tolist(),
          "refeed", out["refeed_tokens"].tolist())


def golden_full(dtype, tag: str, steps: int, unc_steps: int):
    cfg = synth.CONFIGS["paligemma-3b-pt-224"]
    t0 = time.time()
    m = build_reference_model(cfg, dtype, patched=False)
    print(f"built full model in {time.time() - t0:.0f}s", flush=True)
    ids = synth.synth_prompt_ids(cfg)
    pix = synth.synth_pixels(cfg).to(dtype)
    out = {}
    with torch.no_grad():
        t0 = time.time()
        feats = m.vision_tower(pix)
        proj = m.multi_modal_projector(feats)
        print(f"vision {time.time() - t0:.1f}s", flush=True)
        out["vision_features_sub"] = _np(feats[0, ::17, ::13])
        out["projected_sub"] = _np(proj[0, ::17, ::13])
        out["vision_features_absmean"] = np.array(float(feats.float().abs().mean()))
    t0 = time.time()
    toks, lg, kv = ref_generate_cached(m, ids, pix, steps)
    print(f"cached {steps} steps {time.time() - t0:.1f}s tokens {toks.tolist()}", flush=True)
    out["cached_tokens"] = _np(toks)
    out["cached_topv"], out["cached_topi"] = topk_summary(lg)
    out["cached_logits_sub"] = _np(lg[:, :, ::97])
    out["cached_logits_step0"] = _np(lg[:, 0])
    out["cached_k_layer0_sub"] = _np(kv.key_cache[0][0, 0, ::7, ::5])
    if unc_steps:
        t0 = time.time()
        toks, lg = ref_generate_uncached(m, ids, pix, unc_steps)
        print(f"uncached {unc_steps} steps {time.time() - t0:.1f}s tokens {toks.tolist()}", flush=True)
        out["uncached_tokens"] = _np(toks)
        out["uncached_topv"], out["uncached_topi"] = topk_summary(lg)
        out["uncached_logits_sub"] = _np(lg[:, :, ::97])
    np.savez_compressed(os.path.join(OUT, f"full_{tag}.npz"), **out)


def _bits(x: torch.Tensor) -> np.ndarray:
    """Exact bit pattern of a bf16 tensor (npz has no bf16)."""
    return x.detach().contiguous().view(torch.int16).cpu().numpy()


@torch.no_grad()
def golden_full_decode_bf16(steps: int = 16, probe_step: int = 4, probe_layers=(0, 9, 17)):
    """Full-size bf16 CACHED DECODE, the configuration bench.py times: the reference model in bf16 decodes `steps`
    greedy tokens; every step's logits are summarised (top-8 + a 1/31 subsample) and at `probe_step` the input and
    output hidden state of a few decoder layers plus those layers' K/V caches are dumped bit-exactly, so a single
    layer of the CUDA decode path can be checked with the reference's own inputs (no error amplification through the
    stack).  The fp32 reference model then replays the SAME tokens (teacher forced): its logits are the truth both
    bf16 implementations are measured against."""
    cfg = synth.CONFIGS["paligemma-3b-pt-224"]
    ids = synth.synth_prompt_ids(cfg)
    pix = synth.synth_pixels(cfg)
    out = {}
    m = build_reference_model(cfg, torch.bfloat16, patched=False)
    layers = m.language_model.model.layers
    grabbed = {}

    def pre_hook(k):
        def fn(mod, args, kwargs):
            if grabbed.get("on"):
                hs = kwargs.get("hidden_states", args[0] if args else None)
                grabbed[f"in_{k}"] = hs.detach().clone()
                kvc = kwargs["kv_cache"]
                grabbed[f"k_{k}"] = kvc.key_cache[k].detach().clone()
                grabbed[f"v_{k}"] = kvc.value_cache[k].detach().clone()
        return fn

    def post_hook(k):
        def fn(mod, args, kwargs, output):
            if grabbed.get("on"):
                grabbed[f"out_{k}"] = (output[0] if isinstance(output, tuple) else output).detach().clone()
        return fn

    for k in probe_layers:
        layers[k].register_forward_pre_hook(pre_hook(k), with_kwargs=True)
        layers[k].register_forward_hook(post_hook(k), with_kwargs=True)
    mask = torch.ones_like(ids)
    kv = ref_gemma.KVCache()
    t0 = time.time()
    o = m(input_ids=ids, pixel_values=pix.to(torch.bfloat16), attention_mask=mask, kv_cache=kv)
    lg = o["logits"][:, -1, :]
    toks, logits = [torch.argmax(lg, -1, keepdim=True)], [lg]
    for s in range(steps):
        mask = torch.cat([mask, torch.ones((1, 1))], dim=-1)
        grabbed["on"] = s == probe_step
        o = m(input_ids=toks[-1], pixel_values=None, attention_mask=mask, kv_cache=kv)
        grabbed["on"] = False
        lg = o["logits"][:, -1, :]
        logits.append(lg)
        toks.append(torch.argmax(lg, -1, keepdim=True))
    print(f"bf16 decode {steps} steps {time.time() - t0:.0f}s tokens {torch.cat(toks, -1).tolist()}", flush=True)
    lg = torch.stack(logits, 1)                       # (1, steps+1, V): index 0 = prefill, s+1 = cached step s
    out["tokens"] = _np(torch.cat(toks, -1))          # tokens[:, s] is FED at cached step s; tokens[:, s+1] is its argmax
    out["logits_sub"] = _np(lg[:, :, ::31])
    out["topv"], out["topi"] = topk_summary(lg)
    out["probe_step"] = np.array(probe_step)
    out["probe_layers"] = np.array(probe_layers)
    out["probe_position"] = np.array(ids.shape[1] + probe_step + 1)   # attention-mask length at that step (Q3)
    for k in probe_layers:
        out[f"in_{k}"], out[f"out_{k}"] = _bits(grabbed[f"in_{k}"]), _bits(grabbed[f"out_{k}"])
        out[f"k_{k}"], out[f"v_{k}"] = _bits(grabbed[f"k_{k}"]), _bits(grabbed[f"v_{k}"])
    del m, kv
    # fp32 truth, teacher forced with the bf16 run's tokens
    m32 = build_reference_model(cfg, torch.float32, patched=False)
    mask = torch.ones_like(ids)
    kv = ref_gemma.KVCache()
    o = m32(input_ids=ids, pixel_values=pix, attention_mask=mask, kv_cache=kv)
    truth = [o["logits"][:, -1, :]]
    for s in range(steps):
        mask = torch.cat([mask, torch.ones((1, 1))], dim=-1)
        o = m32(input_ids=toks[s], pixel_values=None, attention_mask=mask, kv_cache=kv)
        truth.append(o["logits"][:, -1, :])
    tr = torch.stack(truth, 1)
    out["truth_logits_sub"] = _np(tr[:, :, ::31])
    out["truth_at_topi"] = _np(torch.gather(tr, -1, torch.from_numpy(out["topi"])))
    out["ref_noise_rms"] = np.array(float((lg.float() - tr).pow(2).mean().sqrt()))
    out["logit_rms"] = np.array(float(tr.pow(2).mean().sqrt()))
    print(f"reference bf16 vs fp32 truth: rms {float(out['ref_noise_rms']):.4f} (logit rms {float(out['logit_rms']):.3f})", flush=True)
    np.savez_compressed(os.path.join(OUT, "full_bf16_decode.npz"), **out)


@torch.no_grad()
def golden_vision_batch_bf16(batch: int = 64):
    """BASELINE configs[2]: SigLIP tower + projector over a batch of 64 synthetic images in bf16 through the
    reference modules; a strided subsample of every image's projected features (and the fp32 reference run of the
    first 4 images as the truth for the noise yardstick)."""
    cfg = synth.CONFIGS["paligemma-3b-pt-224"]
    pix = synth.synth_pixels(cfg, batch=batch)
    out = {}
    m = build_reference_model(cfg, torch.bfloat16, patched=False)
    t0 = time.time()
    feats = torch.cat([m.vision_tower(pix[i:i + 8].to(torch.bfloat16)) for i in range(0, batch, 8)])
    proj = m.multi_modal_projector(feats)
    print(f"vision bf16 batch {batch}: {time.time() - t0:.0f}s", flush=True)
    out["features_sub"] = _np(feats[:, ::17, ::13])
    out["projected_sub"] = _np(proj[:, ::17, ::13])
    del m
    m32 = build_reference_model(cfg, torch.float32, patched=False)
    f32 = m32.vision_tower(pix[:4])
    p32 = m32.multi_modal_projector(f32)
    out["truth_features_sub"] = _np(f32[:, ::17, ::13])
    out["truth_projected_sub"] = _np(p32[:, ::17, ::13])
    np.savez_compressed(os.path.join(OUT, "vision_b64_bf16.npz"), **out)


def golden_processor():
    """The reference's own PaliGemmaProcessor (processing_paligemma.py:52-117) on a synthetic RGB image and the
    stub tokenizer: pins resize / rescale / normalise / prompt construction of the drop-in processor."""
    sys.path.insert(0, REF)
    import processing_paligemma as ref_proc
    sys.path.remove(REF)
    sys.modules.pop("processing_paligemma", None)
    from PIL import Image
    yy, xx = np.mgrid[0:480, 0:640]
    img = Image.fromarray(np.stack([(xx * 255 // 639), (yy * 255 // 479), ((xx + yy) % 256)], -1).astype(np.uint8))
    proc = ref_proc.PaliGemmaProcessor(synth.StubTokenizer(), 256, 224)
    out = proc(text=["caption en"], images=[img])
    np.savez_compressed(os.path.join(OUT, "processor.npz"), pixel_values=out["pixel_values"].numpy(),
                        input_ids=out["input_ids"].numpy(), attention_mask=out["attention_mask"].numpy())
    print("processor", out["pixel_values"].shape, out["input_ids"].shape)


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    os.makedirs(OUT, exist_ok=True)
    what = sys.argv[1:] or ["tiny", "small"]
    if "processor" in what:
        golden_processor()
    if "tiny" in what:
        golden_small_model("tiny", steps=8)
    if "small" in what:
        golden_small_model("small", steps=6)
    if "full" in what:
        golden_full(torch.float32, "fp32", steps=32, unc_steps=4)
    if "full_bf16" in what:
        golden_full(torch.bfloat16, "bf16", steps=32, unc_steps=2)
    if "full_bf16_decode" in what:
        golden_full_decode_bf16()
    if "vision_b64" in what:
        golden_vision_batch_bf16()
