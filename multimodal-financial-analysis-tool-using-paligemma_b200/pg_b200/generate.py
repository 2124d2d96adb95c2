"""Engine-native generation loops: the reference's decode loops (inference.py:50-78,
ablation_study_fixed.py:209-251) with the per-token host round trip removed.

cache on : one prefill (vision tower + projector + L layers over the N prompt tokens), then one
           CUDA-graph replay per token; sampled ids, positions, KV lengths advance on the device.
cache off: every step re-runs the vision tower and the whole prefix (positions 0..N+t-1,
           unmasked) with the prefill kernels, exactly what `kv_cache=None` means in the reference.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi as cabi
from .engine import PaliGemmaEngine


@torch.no_grad()
def generate(eng: PaliGemmaEngine, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor],
             max_new_tokens: int, do_sample: bool = False, temperature: float = 0.8, top_p: float = 0.9,
             seed: int = 0, use_kv_cache: bool = True, refeed_prompt: bool = False,
             use_graph: bool = True, return_last_logits: bool = False):
    """Returns int64 (B, max_new_tokens) generated ids on the device (never stops at EOS, like the
    ablation harness; the caller trims).  refeed_prompt reproduces the harness quirk of caching the
    prompt twice (ablation_study_fixed.py:193-199,216-221)."""
    d = eng.dims
    ids = input_ids.to(eng.device)
    B, N = ids.shape
    sample = (float(temperature), float(top_p), int(seed)) if do_sample else None
    if not use_kv_cache:
        return _generate_uncached(eng, ids, pixel_values, max_new_tokens, sample)
    kv = eng.new_kv(B)
    try:
        kv.reserve(N * (2 if refeed_prompt else 1) + max_new_tokens)
        feats = eng.encode_images(pixel_values.to(eng.device)) if pixel_values is not None else None
        if refeed_prompt:
            eng.text_forward(ids, feats, kv, logits="last")
            logits = eng.text_forward(ids, feats, kv, position_value=N, logits="last")
        else:
            logits = eng.text_forward(ids, feats, kv, logits="last")
        first = _pick(eng, logits[:, -1, :].contiguous(), sample, step=0)
        out = torch.empty((B, max_new_tokens), dtype=torch.int64, device=eng.device)
        out[:, 0] = first
        if max_new_tokens > 1:
            ds = eng.decode_state(B)
            # after t generated tokens the mask length is N+t; the next token is fed at position N+t (Q3)
            ds.bind(kv, first, position=N + 1)
            hist0 = 1 if sample is not None else 0     # RNG offset continues from the prefill draw
            ds.step.fill_(hist0)
            ds.want_full_logits = bool(return_last_logits)
            ds.run_steps(kv, max_new_tokens - 1, sample=sample, use_graph=use_graph)
            cols = (torch.arange(max_new_tokens - 1, device=eng.device) + hist0) % ds.max_hist   # ring buffer
            out[:, 1:] = ds.history[:, cols]
            if return_last_logits:
                eng.check_errors(sync=True)
                return out, ds.logits.clone()
        eng.check_errors(sync=True)
        return out
    finally:
        kv.release()


def _pick(eng: PaliGemmaEngine, logits: torch.Tensor, sample, step: int) -> torch.Tensor:
    """argmax (inference.py:68) or temperature + top-p draw (inference.py:65-66) of fp32 (B,V) logits."""
    B, V = logits.shape
    out = torch.empty(B, dtype=torch.int64, device=eng.device)
    L, st = cabi.lib(), cabi.stream()
    if sample is None:
        keys = torch.zeros(B, dtype=torch.int64, device=eng.device)
        cabi.check(L.pg_argmax(out.data_ptr(), logits.data_ptr(), keys.data_ptr(), B, V, st), "argmax")
    else:
        temperature, top_p, seed = sample
        probs = torch.empty_like(logits)
        off = torch.full((1,), step, dtype=torch.int32, device=eng.device)
        cabi.check(L.pg_top_p_sample(out.data_ptr(), logits.data_ptr(), probs.data_ptr(), B, V, temperature, top_p,
                                     seed, off.data_ptr(), None, st), "top_p")
    return out


def _generate_uncached(eng, ids, pixel_values, max_new_tokens, sample):
    B = ids.shape[0]
    out = torch.empty((B, max_new_tokens), dtype=torch.int64, device=eng.device)
    cur = ids
    for t in range(max_new_tokens):
        feats = eng.encode_images(pixel_values.to(eng.device)) if pixel_values is not None else None
        logits = eng.text_forward(cur, feats, None, logits="last")
        nxt = _pick(eng, logits[:, -1, :].contiguous(), sample, step=t)
        out[:, t] = nxt
        cur = torch.cat([cur, nxt[:, None]], dim=1)
    eng.check_errors(sync=True)
    return out
