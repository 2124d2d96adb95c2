"""Tensor-parallel decoder parity on ONE GPU: N rank engines share the device and are advanced one kernel at a time
(dist.LockstepGroup), so the peer-memory exchange fused into the decode kernels (csrc/tp_exchange.cuh) runs exactly as
it does across GPUs -- every rank's producer stores into every rank's buffer, every consumer sums what it finds -- and
its results are checked against the reference's golden vectors and the CPU oracle.  The driver's 1-GPU test box
therefore sees the tensor-parallel path too; tests/tp_check.py repeats the fp32 cases on real GPUs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import paligemma_oracle as O  # noqa: E402
from pg_b200 import synth  # noqa: E402
from _decode_util import (assert_argmax_outside_band, build_engines, compare_reduced_precision,  # noqa: E402
                          decode_through_engines)


@pytest.mark.parametrize("name,tp", [("tiny", 2), ("tiny", 4), ("small", 2), ("small", 8)])
def test_tp_fp32_tokens_match_the_reference(name, tp, golden_dir):
    """fp32 verification mode, GEMV step with the fused exchange: greedy tokens identical to the reference's, logits
    within 1e-4 of the scale (the sum over ranks is fp32 in rank order: only the summation order differs)."""
    g = np.load(os.path.join(golden_dir, f"{name}_fp32.npz"))
    engines, cfg = build_engines(name, torch.float32, tp)
    assert all(e.fabric is not None for e in engines)
    ids, pix = synth.synth_prompt_ids(cfg), synth.synth_pixels(cfg)
    steps = g["cached_tokens"].shape[1]
    toks, lg = decode_through_engines(engines, ids, pix, steps - 1)
    assert toks.tolist() == g["cached_tokens"].tolist()
    want = torch.from_numpy(g["cached_logits"])
    torch.testing.assert_close(lg, want, rtol=1e-4, atol=3e-4 * float(want.abs().max()))
    assert not any(e.fabric.lost_peer() for e in engines)


def test_tp_fp32_batch3_patched_semantics(golden_dir):
    """Batch 3 (patched batch>1 semantics) through the exchange: words of three rows per slot."""
    g = np.load(os.path.join(golden_dir, "tiny_fp32.npz"))
    engines, cfg = build_engines("tiny", torch.float32, 2)
    ids = synth.synth_prompt_ids(cfg, batch=3, prefix_len=6)
    pix = synth.synth_pixels(cfg, batch=3)
    steps = g["batch3_tokens"].shape[1]
    toks, lg = decode_through_engines(engines, ids, pix, steps - 1)
    assert toks.tolist() == g["batch3_tokens"].tolist()
    want = torch.from_numpy(g["batch3_logits"])
    torch.testing.assert_close(lg, want, rtol=1e-4, atol=3e-4 * float(want.abs().max()))


@pytest.mark.parametrize("tp", [2, 4])
def test_tp_batched_step_fp32(tp):
    """The batched (tensor-core) step's launch sequence under tensor parallelism -- partial GEMM, pg_tp_push,
    pg_rmsnorm_reduce, key exchange -- forced to run in fp32 (SIMT GEMMs) so tokens must equal the oracle's exactly."""
    engines, cfg = build_engines("tiny", torch.float32, tp, batched_min=4)
    sd = synth.synth_state_dict(cfg)
    ids = synth.synth_prompt_ids(cfg, batch=6, prefix_len=5)
    pix = synth.synth_pixels(cfg, batch=6)
    steps = 5
    want, want_lg = O.generate_cached(sd, cfg, ids, pix, steps + 1, patched=True, return_logits=True)
    toks, lg = decode_through_engines(engines, ids, pix, steps)
    assert toks.tolist() == want.tolist()
    torch.testing.assert_close(lg, want_lg, rtol=1e-4, atol=3e-4 * float(want_lg.abs().max()))
    assert not any(e.fabric.lost_peer() for e in engines)


@pytest.mark.parametrize("batch,tp", [(1, 2), (2, 8), (6, 2), (32, 4)])
def test_tp_bf16_teacher_forced(batch, tp):
    """bf16 on the `small` shapes (real head dims): batch 1-2 runs the GEMV step, batch 6 / 32 the tcgen05 step, both
    with the exchange.  Teacher-forced with the oracle's bf16 tokens; logits against the oracle's bf16 run with the
    fp32 oracle as truth (rule: _decode_util.compare_reduced_precision)."""
    dtype = torch.bfloat16
    engines, cfg = build_engines("small", dtype, tp)
    sd32 = synth.synth_state_dict(cfg)
    sdb = {k: v.to(dtype) for k, v in sd32.items()}
    ids = synth.synth_prompt_ids(cfg, batch=batch, prefix_len=None if batch == 1 else 7)
    pix = synth.synth_pixels(cfg, batch=batch)
    steps = 5
    ref_t, ref_lg = O.generate_cached(sdb, cfg, ids, pix.to(dtype), steps + 1, patched=True, return_logits=True)
    _, truth = O.generate_cached(sd32, cfg, ids, pix, steps + 1, patched=True, return_logits=True, teacher=ref_t)
    _, lg = decode_through_engines(engines, ids, pix, steps, teacher=ref_t)
    compare_reduced_precision(lg, ref_lg, truth, what=f"small bf16 batch {batch} tp{tp}")
    assert_argmax_outside_band(lg, ref_lg)
    assert not any(e.fabric.lost_peer() for e in engines)


def test_tp_top_p_sampling_is_rank_consistent():
    """Nucleus sampling under tensor parallelism: the vocabulary shards are all-gathered, every rank draws with the same
    counter-based RNG, so every rank feeds the same token (checked inside decode_through_engines)."""
    engines, cfg = build_engines("tiny", torch.float32, 2)
    ids = synth.synth_prompt_ids(cfg, batch=2, prefix_len=5)
    pix = synth.synth_pixels(cfg, batch=2)
    toks, _ = decode_through_engines(engines, ids, pix, 6, sample=(0.8, 0.9, 11))
    single, _ = build_engines("tiny", torch.float32, 1)
    toks1, _ = decode_through_engines(single, ids, pix, 6, sample=(0.8, 0.9, 11))
    assert toks[:, 1:].tolist() == toks1[:, 1:].tolist()      # same logits (1e-4) -> same nucleus -> same draw
