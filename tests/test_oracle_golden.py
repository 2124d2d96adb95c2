"""The oracle restatement against vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only; the reference itself is not needed here."""
import os

import numpy as np
import pytest
import torch

from oracle import paligemma_oracle as O
from pg_b200 import synth


def _load(golden_dir, name):
    p = os.path.join(golden_dir, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated")
    return np.load(p)


@pytest.fixture(scope="module", params=["tiny", "small"])
def case(request, golden_dir):
    cfg = synth.CONFIGS[request.param]
    g = _load(golden_dir, f"{request.param}_fp32.npz")
    sd = synth.synth_state_dict(cfg)
    return cfg, sd, g, synth.synth_prompt_ids(cfg), synth.synth_pixels(cfg)


def _close(a, b, rtol=1e-4, atol=1e-4):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


def test_vision_and_projector(case):
    cfg, sd, g, ids, pix = case
    f = O.siglip_forward(sd, cfg, pix)
    _close(f, g["vision_features"], 1e-4, 1e-4)
    _close(O.projector(sd, f), g["projected"], 1e-4, 1e-4)


def test_prefill_logits_all_positions(case):
    cfg, sd, g, ids, pix = case
    lg = O.forward(sd, cfg, ids, pix, torch.ones_like(ids), None, patched=False)
    assert lg.dtype == torch.float32 and tuple(lg.shape) == g["prefill_logits_all"].shape
    _close(lg, g["prefill_logits_all"], 1e-4, 2e-4)


def test_cached_greedy_tokens_and_kv(case):
    cfg, sd, g, ids, pix = case
    steps = g["cached_tokens"].shape[1]
    toks, lg = O.generate_cached(sd, cfg, ids, pix, steps, patched=False, return_logits=True)
    assert toks.tolist() == g["cached_tokens"].tolist()
    _close(lg, g["cached_logits"], 1e-4, 2e-4)


def test_uncached_greedy_differs_by_construction(case):
    cfg, sd, g, ids, pix = case
    steps = g["uncached_tokens"].shape[1]
    toks, lg = O.generate_uncached(sd, cfg, ids, pix, steps, return_logits=True)
    assert toks.tolist() == g["uncached_tokens"].tolist()
    _close(lg, g["uncached_logits"], 1e-4, 2e-4)
    # Q4: cache-off is not equivalent to cache-on in the reference
    assert g["uncached_tokens"].tolist() != g["cached_tokens"].tolist()


def test_harness_refeed_quirk(case):
    cfg, sd, g, ids, pix = case
    steps = g["refeed_tokens"].shape[1]
    toks, lg = O.generate_cached(sd, cfg, ids, pix, steps, patched=True, refeed_prompt=True,
                                 return_logits=True)
    assert toks.tolist() == g["refeed_tokens"].tolist()
    _close(lg, g["refeed_logits"], 1e-4, 2e-4)
    assert int(g["refeed_kv_len"]) == 2 * ids.shape[1] + steps - 1


def test_batched_patched_decode(case):
    cfg, sd, g, _, _ = case
    ids = synth.synth_prompt_ids(cfg, batch=3, prefix_len=6)
    pix = synth.synth_pixels(cfg, batch=3)
    steps = g["batch3_tokens"].shape[1]
    toks, lg = O.generate_cached(sd, cfg, ids, pix, steps, patched=True, return_logits=True)
    assert toks.tolist() == g["batch3_tokens"].tolist()
    _close(lg, g["batch3_logits"], 1e-4, 2e-4)


def test_pad_token_embeds_to_zero(case):
    cfg, sd, g, ids, pix = case
    ids = ids.clone()
    ids[0, -2] = cfg["pad_token_id"]
    lg = O.forward(sd, cfg, ids, pix, torch.ones_like(ids), None, patched=False)
    _close(lg[:, -1], g["pad_logits_last"], 1e-4, 2e-4)


def test_top_p_nucleus(case):
    cfg, sd, g, ids, pix = case
    lg0 = torch.from_numpy(g["cached_logits"][:, 0])
    dist = O.top_p_distribution(lg0, 0.8, 0.9)
    nz = torch.nonzero(dist[0]).flatten().tolist()
    assert sorted(nz) == sorted(g["topp_nucleus_ids"].tolist())
    assert len(nz) == int(g["topp_keep_count"][0])
    assert abs(float(dist.sum()) - 1.0) < 1e-5
    gen = torch.Generator().manual_seed(3)
    probs = torch.softmax(lg0 / 0.8, -1)
    for _ in range(16):
        assert int(O.sample_top_p(probs, 0.9, gen)) in nz


def test_bf16_rounding_points(case):
    """Same weights cast to bf16: the oracle must round where the reference rounds."""
    cfg, sd, g, ids, pix = case
    sdb = {k: v.to(torch.bfloat16) for k, v in sd.items()}
    steps = g["bf16_cached_tokens"].shape[1]
    toks, lg = O.generate_cached(sdb, cfg, ids, pix.to(torch.bfloat16), steps, patched=False,
                                 return_logits=True)
    assert toks.tolist() == g["bf16_cached_tokens"].tolist()
    np.testing.assert_array_equal(lg.numpy(), g["bf16_cached_logits"])
