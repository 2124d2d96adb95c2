"""CPU oracle for the PaliGemma hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product package never
does (it fails loudly when its CUDA library is missing instead).

This is a functional restatement (plain tensors + a state dict, no nn.Module)
of the reference's algorithm, each function citing the reference file:line it
follows.  Arithmetic runs through torch CPU ops in the model dtype so the
rounding points are the reference's: fp32 islands inside RMSNorm and the two
softmaxes, RoPE cos/sin rounded to the model dtype before the multiply, fp32
logits on return.

Pinning: the reference has no tests and ships no golden vectors,
so this oracle is pinned against outputs of the reference itself, run in the
build container by `oracle/make_golden.py` (which imports /root/reference) on
the same seeded synthetic weights; the vectors live in `tests/golden/` and
`tests/test_oracle_golden.py` checks them without the reference present.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- cache
@dataclass
class OracleKV:
    """Growing per-layer K/V lists — reference `KVCache` (modeling_gemma.py:10-36)."""
    k: List[torch.Tensor] = field(default_factory=list)
    v: List[torch.Tensor] = field(default_factory=list)

    def num_items(self) -> int:  # modeling_gemma.py:16-21
        return 0 if not self.k else self.k[0].shape[-2]

    def update(self, k, v, layer):  # modeling_gemma.py:23-36
        if len(self.k) <= layer:
            self.k.append(k)
            self.v.append(v)
        else:
            self.k[layer] = torch.cat([self.k[layer], k], dim=-2)
            self.v[layer] = torch.cat([self.v[layer], v], dim=-2)
        return self.k[layer], self.v[layer]


# --------------------------------------------------------------------------- vision
def siglip_embeddings(sd: SD, cfg: dict, pixels: torch.Tensor) -> torch.Tensor:
    """Patch conv (stride == kernel, 'valid') + learned positions — modeling_siglip.py:62-79."""
    v = cfg["vision_config"]
    pre = "vision_tower.vision_model.embeddings."
    x = F.conv2d(pixels, sd[pre + "patch_embedding.weight"], sd[pre + "patch_embedding.bias"],
                 stride=v["patch_size"])
    x = x.flatten(2).transpose(1, 2)
    return x + sd[pre + "position_embedding.weight"][None]


def siglip_attention(sd: SD, pre: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """Unmasked MHA, scale applied after QK^T, softmax in fp32 — modeling_siglip.py:97-147."""
    b, s, e = x.shape
    hd = e // heads

    def proj(nm):
        y = F.linear(x, sd[pre + nm + ".weight"], sd[pre + nm + ".bias"])
        return y.view(b, s, heads, hd).transpose(1, 2)

    q, k, v = proj("q_proj"), proj("k_proj"), proj("v_proj")
    w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v).transpose(1, 2).reshape(b, s, e)
    return F.linear(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def siglip_layer(sd: SD, cfg: dict, i: int, x: torch.Tensor) -> torch.Tensor:
    """Pre-LN block — modeling_siglip.py:179-204 (MLP :157-167, GELU tanh :162)."""
    v = cfg["vision_config"]
    e, eps = v["hidden_size"], v.get("layer_norm_eps", 1e-6)
    L = f"vision_tower.vision_model.encoder.layers.{i}."
    h = F.layer_norm(x, (e,), sd[L + "layer_norm1.weight"], sd[L + "layer_norm1.bias"], eps)
    x = siglip_attention(sd, L + "self_attn.", h, v["num_attention_heads"]) + x
    h = F.layer_norm(x, (e,), sd[L + "layer_norm2.weight"], sd[L + "layer_norm2.bias"], eps)
    h = F.linear(h, sd[L + "mlp.fc1.weight"], sd[L + "mlp.fc1.bias"])
    h = F.gelu(h, approximate="tanh")
    h = F.linear(h, sd[L + "mlp.fc2.weight"], sd[L + "mlp.fc2.bias"])
    return h + x


def siglip_forward(sd: SD, cfg: dict, pixels: torch.Tensor, upto: Optional[int] = None) -> torch.Tensor:
    """(B,3,S,S) -> (B,P,Hv) — modeling_siglip.py:236-255."""
    v = cfg["vision_config"]
    x = siglip_embeddings(sd, cfg, pixels)
    n = v["num_hidden_layers"] if upto is None else upto
    for i in range(n):
        x = siglip_layer(sd, cfg, i, x)
    if upto is not None:
        return x
    pre = "vision_tower.vision_model.post_layernorm."
    return F.layer_norm(x, (v["hidden_size"],), sd[pre + "weight"], sd[pre + "bias"],
                        v.get("layer_norm_eps", 1e-6))


def projector(sd: SD, feats: torch.Tensor) -> torch.Tensor:
    """Linear Hv -> D with bias — modeling_gemma.py:435-438."""
    return F.linear(feats, sd["multi_modal_projector.linear.weight"],
                    sd["multi_modal_projector.linear.bias"])


# --------------------------------------------------------------------------- merge
def merge_embeddings(cfg: dict, image_features: torch.Tensor, text_embeds: torch.Tensor,
                     input_ids: torch.Tensor) -> torch.Tensor:
    """Text / image / pad merge — modeling_gemma.py:476-500.

    Image features are divided by the exact python float sqrt(D) (:481); image rows are
    consumed in row-major order by masked_scatter (:498); pad ids give zero rows (:500).
    """
    d = text_embeds.shape[-1]
    pad = cfg["pad_token_id"] if cfg.get("pad_token_id") is not None else -1
    scaled = image_features / (cfg["hidden_size"] ** 0.5)
    out = torch.zeros_like(text_embeds)
    is_img = input_ids == cfg["image_token_index"]
    is_pad = input_ids == pad
    is_txt = ~is_img & ~is_pad
    out = torch.where(is_txt[..., None], text_embeds, out)
    out = out.masked_scatter(is_img[..., None].expand(-1, -1, d), scaled)
    return torch.where(is_pad[..., None], torch.zeros_like(out), out)


def position_ids(attention_mask: torch.Tensor, cached: int, q_len: int, patched: bool) -> torch.Tensor:
    """Positions the reference feeds RoPE — modeling_gemma.py:524-535 (unpatched) and
    ablation_study_fixed.py:130-140 (patched).  Empty cache: 0..L-1.  Non-empty cache:
    the mask length itself (N+t: position N is skipped), shaped (1,B) unpatched or
    (B,1) patched and broadcast over q.  Returned as (B|1, q) before the RoPE clamp."""
    if cached > 0:
        pos = attention_mask.cumsum(-1)
        pos = pos[:, -1:] if patched else pos[:, -1].unsqueeze(0)
        return pos
    L = attention_mask.shape[1]
    pos = torch.arange(L).unsqueeze(0).expand(attention_mask.shape[0], -1)
    return pos.masked_fill(attention_mask == 0, 0)


# --------------------------------------------------------------------------- text
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """fp32 x*rsqrt(mean x^2+eps)*(1+w), cast back — modeling_gemma.py:114-120."""
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (y * (1.0 + w.float())).type_as(x)


def inv_freq(head_dim: int, theta: float, dtype=torch.float32) -> torch.Tensor:
    """modeling_gemma.py:151.  `inv_freq` is a (non-persistent) floating buffer, so the
    `model.to(dtype)` every loader of the reference performs (utils.py:41,
    ablation_study_fixed.py:182,330) rounds it to the model dtype; forward() then
    upcasts the rounded values (:168)."""
    f = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    return f.to(dtype).float()


def rope_cos_sin(pos: torch.Tensor, head_dim: int, theta: float, max_pos: int, dtype):
    """Clamp to [0,max_pos-1]; fp32 angles; cos/sin rounded to model dtype — modeling_gemma.py:155-185."""
    pos = torch.clamp(pos, 0, max_pos - 1)
    ang = pos[:, :, None].float() * inv_freq(head_dim, theta, dtype)[None, None, :]
    emb = torch.cat((ang, ang), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rot_half(x):  # modeling_gemma.py:187-191
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def gemma_attention(sd: SD, t: dict, i: int, x: torch.Tensor, pos: torch.Tensor,
                    kv: Optional[OracleKV]) -> torch.Tensor:
    """MQA with zero additive mask, /sqrt(hd) after QK^T, fp32 softmax — modeling_gemma.py:231-293."""
    b, q_len, _ = x.shape
    nq, nkv, hd = t["num_attention_heads"], t["num_key_value_heads"], t.get("head_dim", 256)
    L = f"language_model.model.layers.{i}.self_attn."
    q = F.linear(x, sd[L + "q_proj.weight"]).view(b, q_len, nq, hd).transpose(1, 2)
    k = F.linear(x, sd[L + "k_proj.weight"]).view(b, q_len, nkv, hd).transpose(1, 2)
    v = F.linear(x, sd[L + "v_proj.weight"]).view(b, q_len, nkv, hd).transpose(1, 2)
    cos, sin = rope_cos_sin(pos, hd, t.get("rope_theta", 10000.0),
                            t.get("max_position_embeddings", 8192), x.dtype)
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    q = (q * cos) + (_rot_half(q) * sin)
    k = (k * cos) + (_rot_half(k) * sin)
    if kv is not None:
        k, v = kv.update(k, v, i)
    rep = nq // nkv
    if rep > 1:  # repeat_kv, modeling_gemma.py:136-141
        k = k[:, :, None].expand(b, nkv, rep, k.shape[-2], hd).reshape(b, nq, k.shape[-2], hd)
        v = v[:, :, None].expand(b, nkv, rep, v.shape[-2], hd).reshape(b, nq, v.shape[-2], hd)
    w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd)
    w = w + torch.zeros((), dtype=x.dtype)  # the reference's mask is all zeros (:506-514)
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v).transpose(1, 2).contiguous().view(b, q_len, -1)
    return F.linear(o, sd[L + "o_proj.weight"])


def gemma_layer(sd: SD, t: dict, i: int, x: torch.Tensor, pos, kv) -> torch.Tensor:
    """modeling_gemma.py:307-338 (MLP :133-134)."""
    eps = t.get("rms_norm_eps", 1e-6)
    L = f"language_model.model.layers.{i}."
    h = rms_norm(x, sd[L + "input_layernorm.weight"], eps)
    x = x + gemma_attention(sd, t, i, h, pos, kv)
    h = rms_norm(x, sd[L + "post_attention_layernorm.weight"], eps)
    g = F.gelu(F.linear(h, sd[L + "mlp.gate_proj.weight"]), approximate="tanh")
    u = F.linear(h, sd[L + "mlp.up_proj.weight"])
    return x + F.linear(g * u, sd[L + "mlp.down_proj.weight"])


def gemma_hidden(sd: SD, cfg: dict, embeds: torch.Tensor, pos: torch.Tensor,
                 kv: Optional[OracleKV], upto: Optional[int] = None) -> torch.Tensor:
    """x*sqrt(D) with the normaliser rounded to model dtype, layers, final norm — modeling_gemma.py:357-382."""
    t = cfg["text_config"]
    x = embeds * torch.tensor(t["hidden_size"] ** 0.5, dtype=embeds.dtype)
    n = t["num_hidden_layers"] if upto is None else upto
    for i in range(n):
        x = gemma_layer(sd, t, i, x, pos, kv)
    if upto is not None:
        return x
    return rms_norm(x, sd["language_model.model.norm.weight"], t.get("rms_norm_eps", 1e-6))


def lm_head(sd: SD, h: torch.Tensor) -> torch.Tensor:
    """Tied lm_head, logits returned fp32 — modeling_gemma.py:417-418."""
    w = sd.get("language_model.lm_head.weight", sd["language_model.model.embed_tokens.weight"])
    return F.linear(h, w).float()


# --------------------------------------------------------------------------- top level
def forward(sd: SD, cfg: dict, input_ids: torch.Tensor, pixel_values: Optional[torch.Tensor],
            attention_mask: torch.Tensor, kv: Optional[OracleKV] = None, patched: bool = True,
            image_features: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PaliGemmaForConditionalGeneration.forward — modeling_gemma.py:539-617.  Returns fp32
    logits (B,q,V).  `image_features` lets a caller reuse projector output across cache-off
    steps (identical result: the tower is a pure function of the pixels)."""
    if attention_mask is None:
        raise ValueError("attention_mask must be provided")
    assert bool(torch.all(attention_mask == 1)), "The input cannot be padded"
    emb_w = sd["language_model.model.embed_tokens.weight"]
    dtype = emb_w.dtype
    pad = cfg.get("pad_token_id")
    text = F.embedding(input_ids, emb_w, padding_idx=pad)
    if image_features is None:
        if pixel_values is not None:
            image_features = projector(sd, siglip_forward(sd, cfg, pixel_values.to(dtype)))
        else:
            image_features = torch.zeros(text.shape[0], 0, text.shape[-1], dtype=dtype)
    cached = 0 if kv is None else kv.num_items()
    if cached > 0 and not patched:
        assert input_ids.shape[1] == 1  # modeling_gemma.py:509
    embeds = merge_embeddings(cfg, image_features, text, input_ids)
    pos = position_ids(attention_mask, cached, input_ids.shape[1], patched)
    h = gemma_hidden(sd, cfg, embeds, pos, kv)
    return lm_head(sd, h)


def sample_top_p(probs: torch.Tensor, p: float, generator=None) -> torch.Tensor:
    """Nucleus sampling — inference.py:15-24."""
    ps, idx = torch.sort(probs, dim=-1, descending=True)
    cum = torch.cumsum(ps, dim=-1)
    ps = ps.masked_fill(cum - ps > p, 0.0)
    ps = ps / ps.sum(dim=-1, keepdim=True)
    nxt = torch.multinomial(ps, num_samples=1, generator=generator)
    return torch.gather(idx, -1, nxt)


def top_p_distribution(logits: torch.Tensor, temperature: float, p: float) -> torch.Tensor:
    """The renormalised nucleus distribution in vocab order (what multinomial draws from)."""
    probs = torch.softmax(logits / temperature, dim=-1)
    ps, idx = torch.sort(probs, dim=-1, descending=True)
    cum = torch.cumsum(ps, dim=-1)
    ps = ps.masked_fill(cum - ps > p, 0.0)
    ps = ps / ps.sum(dim=-1, keepdim=True)
    return torch.zeros_like(probs).scatter_(-1, idx, ps)


@torch.no_grad()
def generate_cached(sd: SD, cfg: dict, input_ids, pixel_values, max_tokens: int,
                    patched: bool = True, refeed_prompt: bool = False, return_logits: bool = False,
                    teacher: Optional[torch.Tensor] = None):
    """Greedy cache-on loop — inference.py:50-78 (with pixel_values dropped after the first
    call as ablation_study_fixed.py:243; Q5 shows the output is unchanged).  refeed_prompt
    reproduces the ablation harness: an extra prefill before the loop (:193-199) so the
    prompt is cached twice (Q6).  teacher (B, >= max_tokens-1): feed teacher[:, t] instead of the step's own argmax
    (parity tests compare logits of reduced-precision runs on identical token streams)."""
    b, n = input_ids.shape
    mask = torch.ones((b, n), dtype=torch.int64)
    kv = OracleKV()
    ids, pix = input_ids, pixel_values
    if refeed_prompt:
        forward(sd, cfg, ids, pix, mask, kv, patched)
    toks, all_logits = [], []
    for t in range(max_tokens):
        logits = forward(sd, cfg, ids, pix, mask, kv, patched)[:, -1, :]
        nxt = torch.argmax(logits, dim=-1, keepdim=True)
        toks.append(nxt)
        if return_logits:
            all_logits.append(logits)
        ids, pix = (nxt if teacher is None or t >= teacher.shape[1] else teacher[:, t:t + 1]), None
        mask = torch.cat([mask.to(torch.float32) if mask.dtype != torch.float32 else mask,
                          torch.ones((b, 1))], dim=-1)
    out = torch.cat(toks, dim=-1)
    return (out, torch.stack(all_logits, 1)) if return_logits else out


@torch.no_grad()
def generate_uncached(sd: SD, cfg: dict, input_ids, pixel_values, max_tokens: int,
                      return_logits: bool = False):
    """Greedy cache-off loop — ablation_study_fixed.py:209-251: every step recomputes the whole
    prefix (positions 0..N+t-1, unmasked) and the vision tower."""
    b, n = input_ids.shape
    dtype = sd["language_model.model.embed_tokens.weight"].dtype
    feats = projector(sd, siglip_forward(sd, cfg, pixel_values.to(dtype)))
    ids = input_ids
    toks, all_logits = [], []
    for _ in range(max_tokens):
        mask = torch.ones(ids.shape, dtype=torch.int64)
        logits = forward(sd, cfg, ids, None, mask, None, True, image_features=feats)[:, -1, :]
        nxt = torch.argmax(logits, dim=-1, keepdim=True)
        toks.append(nxt)
        if return_logits:
            all_logits.append(logits)
        ids = torch.cat([ids, nxt], dim=-1)
    out = torch.cat(toks, dim=-1)
    return (out, torch.stack(all_logits, 1)) if return_logits else out
