"""Parity of the BENCHMARKED bf16 paths against the reference run in bf16 (golden vectors made by
oracle/make_golden.py from the unmodified reference) -- full PaliGemma-3B-pt-224 shapes:

  * cached decode through the CUDA-graph step (what bench.py's `value` times), teacher-forced for 16 steps,
  * single decoder layers of the GEMV step and of the tensor-core (batched) step fed the REFERENCE's own layer inputs
    and K/V caches: elementwise rtol 2e-2 (no amplification through the stack to hide a kernel bug),
  * SigLIP + projector at batch 64 (BASELINE configs[2]),
and on the `small` shapes the batched decode step of configs[3] / configs[4] (batch 32 greedy, batch 8 top-p)
against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import paligemma_oracle as O  # noqa: E402
from pg_b200 import synth  # noqa: E402
import modeling_gemma as MG  # noqa: E402
from _decode_util import (assert_argmax_outside_band, build_engines, compare_reduced_precision,  # noqa: E402
                          decode_through_engines, rms)
from test_gpu_model import build_model, golden  # noqa: E402

FULL = "paligemma-3b-pt-224"


def bf16_from_bits(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.copy()).view(torch.bfloat16)


@pytest.fixture(scope="module")
def full_bf16():
    model, cfg = build_model(FULL, torch.bfloat16)
    return model, cfg, model._engine_ready()


# ------------------------------------------------------------------------------- full size, cached decode
def test_full_size_bf16_cached_decode_teacher_forced(full_bf16, golden_dir):
    """16 cached steps through DecodeState's CUDA graph fed the reference's bf16 tokens: every step's logits against the
    reference's bf16 logits (1/31 subsample + its top-8) with the fp32 reference as truth."""
    model, cfg, eng = full_bf16
    g = golden(golden_dir, "full_bf16_decode.npz")
    ids, pix = synth.synth_prompt_ids(cfg).cuda(), synth.synth_pixels(cfg).cuda()
    tokens = torch.from_numpy(g["tokens"]).cuda()            # tokens[:, s] is fed at cached step s
    steps = tokens.shape[1] - 1
    N = ids.shape[1]
    kv = eng.new_kv(1)
    got = []
    try:
        kv.reserve(N + steps + 2)
        with torch.no_grad():
            lg0 = eng.text_forward(ids, eng.encode_images(pix), kv, logits="last")
        got.append(lg0[:, -1].clone())
        ds = eng.decode_state(1)
        ds.bind(kv, tokens[:, 0], position=N + 1)
        for s in range(steps):
            ds.ids.copy_(tokens[:, s])
            ds.run_steps(kv, 1)                               # graph replay (captured at the first call)
            got.append(ds.logits.clone())
        assert len(ds.graphs) > 0, "the decode step did not run as a CUDA graph"
    finally:
        kv.release()
    got = torch.stack(got, 1).cpu()                           # (1, steps+1, V)
    ref_sub, truth_sub = torch.from_numpy(g["logits_sub"]), torch.from_numpy(g["truth_logits_sub"])
    stats = compare_reduced_precision(got[:, :, ::31], ref_sub, truth_sub, what="full-size bf16 cached decode, 16 steps")
    # decode steps only (index 0 is the prefill): the kernels bench.py times
    compare_reduced_precision(got[:, 1:, ::31], ref_sub[:, 1:], truth_sub[:, 1:], what="decode steps only")
    # the reference's top-8 candidates: same values within the band, and the same winner when its margin is clear
    topi, topv = torch.from_numpy(g["topi"]), torch.from_numpy(g["topv"])
    ours_at_top = torch.gather(got, -1, topi)
    scale = float(ref_sub.abs().max())
    ref_noise = float(g["ref_noise_rms"])
    assert float((ours_at_top - topv).abs().max()) <= 6 * ref_noise + 2e-2 * scale, stats
    margin = topv[..., 0] - topv[..., 1]
    band = 2 * (2e-2 * topv[..., 0].abs() + 2e-2 * scale)
    clear = margin > band
    agree = got.argmax(-1) == topi[..., 0]
    print(f"argmax agrees on {int(agree.sum())}/{agree.numel()} steps; {int(clear.sum())} have a margin above the band")
    assert bool(agree[clear].all())
    # whatever we pick is a token the reference's bf16 run rates within the bf16 noise of its own best
    ours_pick = got.argmax(-1)                                # (1, steps+1)
    in_top8 = (topi == ours_pick[..., None]).any(-1)
    print(f"our argmax is among the reference's top-8 on {int(in_top8.sum())}/{in_top8.numel()} steps")
    assert float(in_top8.float().mean()) >= 0.8


def _fill_layer_cache(eng, kv, li, k, v):
    """Reference K/V of one layer, (1, n_kv, T, hd) bf16, into the pages of every sequence of `kv`."""
    T = k.shape[2]
    tok = torch.arange(T, device="cuda")
    pages = kv.page_table[:, : (T + eng.page_size - 1) // eng.page_size]
    pg = pages[:, tok // eng.page_size].long()
    off = (tok % eng.page_size).expand_as(pg)
    eng.k_pool[li][pg, off] = k.cuda().permute(0, 2, 1, 3).reshape(1, T, -1).expand(kv.batch, -1, -1)
    eng.v_pool[li][pg, off] = v.cuda().permute(0, 2, 1, 3).reshape(1, T, -1).expand(kv.batch, -1, -1)
    kv.kv_len.fill_(T)
    kv.length = T


@pytest.mark.parametrize("batch", [1, 4])
def test_full_size_bf16_single_layers_with_reference_inputs(full_bf16, golden_dir, batch):
    """One decoder layer at a time, fed the reference's own hidden state and K/V cache (cached step 4 of the bf16 run):
    input RMSNorm + q/k/v + RoPE + append + MQA attention + o_proj + residual + post norm + GeGLU MLP + residual.
    batch 1 = the GEMV kernels of the bs-1 step; batch 4 = the tensor-core (skinny tcgen05 GEMM) step, four copies of
    the row.  Elementwise: rtol 2e-2 plus 2e-2 of the tensor's scale (the north star's bf16 tolerance)."""
    model, cfg, eng = full_bf16
    g = golden(golden_dir, "full_bf16_decode.npz")
    pos = int(g["probe_position"])
    ds = eng.decode_state(batch)
    worst = 0.0
    for li in [int(x) for x in g["probe_layers"]]:
        k, v = bf16_from_bits(g[f"k_{li}"]), bf16_from_bits(g[f"v_{li}"])
        x_in, want = bf16_from_bits(g[f"in_{li}"]).view(1, -1), bf16_from_bits(g[f"out_{li}"]).view(1, -1).float()
        kv = eng.new_kv(batch)
        try:
            kv.reserve(k.shape[2] + 1)
            _fill_layer_cache(eng, kv, li, k, v)
            ds.x.copy_(x_in.cuda().expand(batch, -1))
            ds.pos.fill_(pos)
            gen = ds.layer_gemv_gen(li, kv, None, first=True) if batch < eng.batched_min else \
                ds.layer_batched_gen(li, kv, None, first=True)
            for _ in gen:
                pass
            got = ds.x.float().cpu()
            # the layer appended the new token's K/V at slot T: compare them with the reference's next-step cache? not
            # dumped -- the hidden state covers them (attention reads the appended entry)
        finally:
            kv.release()
        scale = float(want.abs().max())
        err = (got - want.expand(batch, -1)).abs()
        tol = 2e-2 * want.abs().expand(batch, -1) + 2e-2 * scale
        worst = max(worst, float((err / tol).max()))
        print(f"layer {li} batch {batch}: max|err| {float(err.max()):.4g} (scale {scale:.4g}), worst err/tol {float((err / tol).max()):.3f}, "
              f"rms err {rms(err):.4g}")
        assert bool((err <= tol).all()), (li, float(err.max()), scale)
        assert float((got - got[:1]).abs().max()) == 0.0        # identical rows -> identical results
    assert worst < 1.0


# ------------------------------------------------------------------------------- vision tower, batch 64
def test_vision_batch64_bf16_features(full_bf16, golden_dir):
    """BASELINE configs[2]: SigLIP So400m/14 + projector over 64 synthetic images in bf16, against the reference modules
    run in bf16 (strided subsample of every image's features); fp32 reference of the first four images as truth."""
    model, cfg, eng = full_bf16
    g = golden(golden_dir, "vision_b64_bf16.npz")
    pix = synth.synth_pixels(cfg, batch=64).cuda()
    with torch.no_grad():
        feats = eng.vision_features(pix)
        proj = eng.project(feats)
    f_sub, p_sub = feats[:, ::17, ::13].float().cpu(), proj[:, ::17, ::13].float().cpu()
    ref_f, ref_p = torch.from_numpy(g["features_sub"]), torch.from_numpy(g["projected_sub"])
    tr_f, tr_p = torch.from_numpy(g["truth_features_sub"]), torch.from_numpy(g["truth_projected_sub"])
    for name, ours, ref, truth in (("features", f_sub, ref_f, tr_f), ("projected", p_sub, ref_p, tr_p)):
        ref_noise, our_noise = rms(ref[:4] - truth), rms(ours[:4] - truth)
        delta = rms(ours - ref)
        print(f"vision b64 {name}: reference-bf16 vs fp32 rms {ref_noise:.4g}, ours vs fp32 rms {our_noise:.4g}, "
              f"ours vs reference-bf16 rms {delta:.4g} (all 64 images), value rms {rms(ref):.4g}")
        assert our_noise <= 1.5 * ref_noise + 1e-3 * rms(truth)
        assert delta <= 2.5 * ref_noise + 1e-3 * rms(truth)
    # every image of the batch is as close to its reference as the first four (no tail / tile-edge damage)
    per_image = (p_sub - ref_p).pow(2).mean(dim=(1, 2)).sqrt()
    assert float(per_image.max()) <= 2.0 * float(per_image.median()) + 1e-3
    # batch-of-64 launch == batch-of-1 launches (different kernels: 2-CTA GEMMs vs the persistent single-CTA ones)
    with torch.no_grad():
        one = eng.project(eng.vision_features(pix[5:6]))
    d1 = rms(one[0, ::17, ::13].float().cpu() - p_sub[5])
    assert d1 <= 2.5 * rms(ref_p[:4] - tr_p), d1


# ------------------------------------------------------------------------------- batched decode step (configs[3], configs[4])
@pytest.mark.parametrize("batch", [8, 32])
def test_batched_step_bf16_against_oracle(batch):
    """`_step_batched_gen` (skinny tcgen05 GEMMs) on the small shapes, teacher-forced with the oracle's bf16 greedy tokens
    (patched batch>1 semantics, ablation_study_fixed.py:99-142): logits of every step against the oracle."""
    dtype = torch.bfloat16
    engines, cfg = build_engines("small", dtype, 1)
    assert batch >= engines[0].batched_min
    sd32 = synth.synth_state_dict(cfg)
    sdb = {k: v.to(dtype) for k, v in sd32.items()}
    ids = synth.synth_prompt_ids(cfg, batch=batch, prefix_len=9)
    pix = synth.synth_pixels(cfg, batch=batch)
    steps = 6
    ref_t, ref_lg = O.generate_cached(sdb, cfg, ids, pix.to(dtype), steps + 1, patched=True, return_logits=True)
    _, truth = O.generate_cached(sd32, cfg, ids, pix, steps + 1, patched=True, return_logits=True, teacher=ref_t)
    _, lg = decode_through_engines(engines, ids, pix, steps, teacher=ref_t)
    compare_reduced_precision(lg, ref_lg, truth, what=f"small bf16 batched step, batch {batch}")
    assert_argmax_outside_band(lg, ref_lg)


def test_batched_step_fp32_tokens_exact():
    """The batched step's launch sequence (norm, GEMM, RoPE/append, cluster attention, argmax, advance) forced to run in
    fp32 on the tiny shapes: greedy tokens bit-identical to the oracle, logits at 1e-4."""
    engines, cfg = build_engines("tiny", torch.float32, 1, batched_min=4)
    sd = synth.synth_state_dict(cfg)
    ids = synth.synth_prompt_ids(cfg, batch=12, prefix_len=5)
    pix = synth.synth_pixels(cfg, batch=12)
    want, want_lg = O.generate_cached(sd, cfg, ids, pix, 7, patched=True, return_logits=True)
    toks, lg = decode_through_engines(engines, ids, pix, 6)
    assert toks.tolist() == want.tolist()
    torch.testing.assert_close(lg, want_lg, rtol=1e-4, atol=1e-4 * float(want_lg.abs().max()))


def test_batched_top_p_draws_from_the_reference_nucleus():
    """configs[4]'s sampling path on the batched step (batch 8, temperature 0.8, top_p 0.9): teacher-forced with the
    oracle's tokens, every token the engine draws lies in the nucleus the reference rule (inference.py:15-24) builds from
    the ORACLE's logits widened by the bf16 band, and logits match as in the greedy case."""
    dtype = torch.bfloat16
    engines, cfg = build_engines("small", dtype, 1)
    sd32 = synth.synth_state_dict(cfg)
    sdb = {k: v.to(dtype) for k, v in sd32.items()}
    ids = synth.synth_prompt_ids(cfg, batch=8, prefix_len=9)
    pix = synth.synth_pixels(cfg, batch=8)
    steps = 6
    ref_t, ref_lg = O.generate_cached(sdb, cfg, ids, pix.to(dtype), steps + 1, patched=True, return_logits=True)
    _, truth = O.generate_cached(sd32, cfg, ids, pix, steps + 1, patched=True, return_logits=True, teacher=ref_t)
    toks, lg = decode_through_engines(engines, ids, pix, steps, teacher=ref_t, sample=(0.8, 0.9, 1234))
    compare_reduced_precision(lg, ref_lg, truth, what="small bf16 batched step with top-p, batch 8")
    # the draw at step t comes from OUR logits of step t: it must be inside the nucleus of those logits (reference rule)
    for t in range(1, steps + 1):
        dist = O.top_p_distribution(lg[:, t], 0.8, 0.9)
        p = torch.gather(dist, -1, toks[:, t:t + 1])
        assert bool((p > 0).all()), (t, p.tolist())
