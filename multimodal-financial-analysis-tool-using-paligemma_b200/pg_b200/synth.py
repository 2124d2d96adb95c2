"""Synthetic PaliGemma checkpoints and inputs (no network, no real weights).

Every tensor is drawn from its own seeded CPU generator keyed by the HF
state-dict name, so any subset of the checkpoint can be produced on any box
and is bit-identical between this container, the GPU box, the oracle and the
golden-vector script.  Key names follow the reference's `state_dict()`
(SURVEY.md §2.3 census; reference `modeling_gemma.py:429-456`,
`modeling_siglip.py:36-255`).

Distribution: Linear/Conv weight ~ N(0, 0.05), bias ~ N(0, 0.02);
token embedding ~ N(0, 0.05); position embedding ~ N(0, 0.02); LayerNorm
w = 1 + N(0, 0.1), b ~ N(0, 0.02); RMSNorm w ~ N(0, 0.1).  The default torch init
makes the tied lm_head echo its input token, which would make greedy-token
parity vacuous; this recipe gives varied tokens.
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterator, Tuple

import torch

PALIGEMMA_3B_224 = {
    "image_token_index": 257152,
    "vocab_size": 257216,
    "projection_dim": 2048,
    "hidden_size": 2048,
    "pad_token_id": 0,
    "vision_config": {
        "hidden_size": 1152,
        "intermediate_size": 4304,
        "num_hidden_layers": 27,
        "num_attention_heads": 16,
        "num_channels": 3,
        "image_size": 224,
        "patch_size": 14,
        "num_image_tokens": 256,
    },
    "text_config": {
        "vocab_size": 257216,
        "hidden_size": 2048,
        "intermediate_size": 16384,
        "num_hidden_layers": 18,
        "num_attention_heads": 8,
        "num_key_value_heads": 1,
        "head_dim": 256,
    },
}

# Small shape set for tests the CPU oracle finishes in milliseconds.  It keeps the
# awkward properties of the real model: vision head_dim 72, an intermediate size
# that is not a multiple of 64, MQA with one KV head, vocab rows above the
# image-token id, image-token id != pad id.
TINY = {
    "synth_w_std": 0.2,
    "image_token_index": 1216,
    "vocab_size": 1280,
    "projection_dim": 128,
    "hidden_size": 128,
    "pad_token_id": 0,
    "vision_config": {
        "hidden_size": 144,
        "intermediate_size": 272,
        "num_hidden_layers": 2,
        "num_attention_heads": 2,
        "num_channels": 3,
        "image_size": 56,
        "patch_size": 14,
        "num_image_tokens": 16,
    },
    "text_config": {
        "vocab_size": 1280,
        "hidden_size": 128,
        "intermediate_size": 512,
        "num_hidden_layers": 2,
        "num_attention_heads": 4,
        "num_key_value_heads": 1,
        "head_dim": 32,
    },
}

# A middle size: real head dims (256 / 72) with few layers, for kernel-shape coverage.
SMALL = {
    "synth_w_std": 0.1,
    "image_token_index": 8000,
    "vocab_size": 8064,
    "projection_dim": 512,
    "hidden_size": 512,
    "pad_token_id": 0,
    "vision_config": {
        "hidden_size": 288,
        "intermediate_size": 1072,
        "num_hidden_layers": 3,
        "num_attention_heads": 4,
        "num_channels": 3,
        "image_size": 112,
        "patch_size": 14,
        "num_image_tokens": 64,
    },
    "text_config": {
        "vocab_size": 8064,
        "hidden_size": 512,
        "intermediate_size": 2048,
        "num_hidden_layers": 3,
        "num_attention_heads": 8,
        "num_key_value_heads": 1,
        "head_dim": 256,
    },
}

CONFIGS = {"paligemma-3b-pt-224": PALIGEMMA_3B_224, "small": SMALL, "tiny": TINY}

# Fixed stub token ids (no tokenizer offline): BOS=2, EOS=1, pad=0, "\n"=108.
BOS_ID, EOS_ID, PAD_ID, NEWLINE_ID = 2, 1, 0, 108
CAPTION_EN_IDS = (7907, 659)  # stand-in ids for "caption", " en"


def state_dict_spec(cfg: dict) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (hf_key, shape, kind) for every persistent tensor, in checkpoint order."""
    v, t = cfg["vision_config"], cfg["text_config"]
    hv, iv, p, c = v["hidden_size"], v["intermediate_size"], v["patch_size"], v.get("num_channels", 3)
    npos = (v["image_size"] // p) ** 2
    vm = "vision_tower.vision_model."
    yield vm + "embeddings.patch_embedding.weight", (hv, c, p, p), "w"
    yield vm + "embeddings.patch_embedding.bias", (hv,), "b"
    yield vm + "embeddings.position_embedding.weight", (npos, hv), "pos"
    for i in range(v["num_hidden_layers"]):
        L = f"{vm}encoder.layers.{i}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            yield L + f"self_attn.{nm}.weight", (hv, hv), "w"
            yield L + f"self_attn.{nm}.bias", (hv,), "b"
        yield L + "layer_norm1.weight", (hv,), "ln_w"
        yield L + "layer_norm1.bias", (hv,), "b"
        yield L + "mlp.fc1.weight", (iv, hv), "w"
        yield L + "mlp.fc1.bias", (iv,), "b"
        yield L + "mlp.fc2.weight", (hv, iv), "w"
        yield L + "mlp.fc2.bias", (hv,), "b"
        yield L + "layer_norm2.weight", (hv,), "ln_w"
        yield L + "layer_norm2.bias", (hv,), "b"
    yield vm + "post_layernorm.weight", (hv,), "ln_w"
    yield vm + "post_layernorm.bias", (hv,), "b"
    yield "multi_modal_projector.linear.weight", (cfg["projection_dim"], hv), "w"
    yield "multi_modal_projector.linear.bias", (cfg["projection_dim"],), "b"
    d, f = t["hidden_size"], t["intermediate_size"]
    nq, nkv, hd = t["num_attention_heads"], t["num_key_value_heads"], t.get("head_dim", 256)
    lm = "language_model.model."
    yield lm + "embed_tokens.weight", (t["vocab_size"], d), "emb"
    for i in range(t["num_hidden_layers"]):
        L = f"{lm}layers.{i}."
        yield L + "self_attn.q_proj.weight", (nq * hd, d), "w"
        yield L + "self_attn.k_proj.weight", (nkv * hd, d), "w"
        yield L + "self_attn.v_proj.weight", (nkv * hd, d), "w"
        yield L + "self_attn.o_proj.weight", (d, nq * hd), "w"
        yield L + "mlp.gate_proj.weight", (f, d), "w"
        yield L + "mlp.up_proj.weight", (f, d), "w"
        yield L + "mlp.down_proj.weight", (d, f), "w"
        yield L + "input_layernorm.weight", (d,), "rms_w"
        yield L + "post_attention_layernorm.weight", (d,), "rms_w"
    yield lm + "norm.weight", (d,), "rms_w"
    # language_model.lm_head.weight is tied to embed_tokens (reference modeling_gemma.py:396-397)


_STD = {"w": 0.05, "b": 0.02, "pos": 0.02, "emb": 0.05, "ln_w": 0.1, "rms_w": 0.1}


def _key_seed(seed: int, key: str) -> int:
    return (seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFF


def synth_tensor(key: str, shape, kind: str, seed: int = 1234, w_std: float | None = None) -> torch.Tensor:
    """fp32 CPU tensor for one checkpoint entry (deterministic in (seed, key)).

    w_std overrides the Linear/Conv weight std: the narrow test configs need a larger
    one for the layers (not the tied embedding) to decide the next token."""
    g = torch.Generator(device="cpu").manual_seed(_key_seed(seed, key))
    std = w_std if (kind == "w" and w_std is not None) else _STD[kind]
    x = torch.randn(shape, generator=g, dtype=torch.float32).mul_(std)
    if kind == "ln_w":
        x.add_(1.0)
    return x


def synth_state_dict(cfg: dict, seed: int = 1234, dtype=torch.float32, device="cpu",
                     tie: bool = True) -> Dict[str, torch.Tensor]:
    """Full synthetic checkpoint with HF key names.  `tie` adds lm_head as an alias."""
    sd = {}
    for key, shape, kind in state_dict_spec(cfg):
        sd[key] = synth_tensor(key, shape, kind, seed, cfg.get("synth_w_std")).to(device=device, dtype=dtype)
    if tie:
        sd["language_model.lm_head.weight"] = sd["language_model.model.embed_tokens.weight"]
    return sd


def synth_pixels(cfg: dict, batch: int = 1, seed: int = 1234) -> torch.Tensor:
    """(B,3,S,S) fp32 in [-1,1): the range PaliGemmaProcessor produces."""
    s = cfg["vision_config"]["image_size"]
    g = torch.Generator(device="cpu").manual_seed(_key_seed(seed, "pixels"))
    return torch.rand((batch, 3, s, s), generator=g, dtype=torch.float32) * 2 - 1


def synth_prompt_ids(cfg: dict, batch: int = 1, prefix_len: int | None = None,
                     seed: int = 1234) -> torch.Tensor:
    """int64 (B, N) prompt: image tokens, BOS, prefix ids, newline.

    prefix_len=None gives the 'caption en' stand-in of config 1; otherwise BOS +
    (prefix_len-2) pseudo-random text ids + newline, distinct per batch row
    (never pad id 0 and never the image-token id).
    """
    nimg = cfg["vision_config"]["num_image_tokens"]
    img = cfg["image_token_index"]
    rows = []
    g = torch.Generator(device="cpu").manual_seed(_key_seed(seed, "prompt"))
    for _ in range(batch):
        if prefix_len is None:
            body = [min(i, img - 1) for i in CAPTION_EN_IDS]
        else:
            body = torch.randint(3, img, (prefix_len - 2,), generator=g).tolist()
        rows.append([img] * nimg + [BOS_ID] + body + [NEWLINE_ID])
    return torch.tensor(rows, dtype=torch.int64)


class StubTokenizer:
    """Offline stand-in for the HF Gemma tokenizer (no tokenizer.model on disk): whitespace-free
    greedy matching of the special strings the processor emits, hashed ids for everything else.
    Implements exactly the surface PaliGemmaProcessor touches (reference
    processing_paligemma.py:63-75,108-113) plus decode() for the generation loop."""

    def __init__(self, vocab_size: int = 257216, image_token_id: int = 257152):
        self.vocab_size, self._image_id = vocab_size, image_token_id
        self.bos_token, self.eos_token = "<bos>", "<eos>"
        self.bos_token_id, self.eos_token_id, self.pad_token_id = BOS_ID, EOS_ID, PAD_ID
        self.add_bos_token = self.add_eos_token = True
        self._special = {"<bos>": BOS_ID, "<eos>": EOS_ID, "<pad>": PAD_ID, "\n": NEWLINE_ID}
        self.added_tokens = []

    def add_special_tokens(self, mapping):
        for tok in mapping.get("additional_special_tokens", []):
            self._special[tok] = self._image_id if tok == "<image>" else self._hash(tok)
        return len(mapping.get("additional_special_tokens", []))

    def add_tokens(self, tokens):
        self.added_tokens.extend(tokens)
        return len(tokens)

    def convert_tokens_to_ids(self, tok):
        return self._special.get(tok, self._hash(tok))

    def _hash(self, word: str) -> int:
        return 3 + zlib.crc32(word.encode()) % (min(self._image_id, self.vocab_size) - 3)

    def _encode(self, text: str):
        ids, i = [], 0
        specials = sorted(self._special, key=len, reverse=True)
        word = ""
        while i < len(text):
            for s in specials:
                if text.startswith(s, i):
                    if word:
                        ids.append(self._hash(word)); word = ""
                    ids.append(self._special[s]); i += len(s)
                    break
            else:
                if text[i] == " ":
                    if word:
                        ids.append(self._hash(word)); word = ""
                else:
                    word += text[i]
                i += 1
        if word:
            ids.append(self._hash(word))
        return ids

    def __call__(self, texts, return_tensors="pt", padding="longest", truncation=True):
        rows = [self._encode(t) for t in texts]
        n = max(len(r) for r in rows)
        ids = torch.tensor([r + [PAD_ID] * (n - len(r)) for r in rows], dtype=torch.int64)
        mask = torch.tensor([[1] * len(r) + [0] * (n - len(r)) for r in rows], dtype=torch.int64)
        return {"input_ids": ids, "attention_mask": mask}

    def decode(self, ids, skip_special_tokens=True):
        ids = ids.tolist() if hasattr(ids, "tolist") else list(ids)
        drop = set(self._special.values()) if skip_special_tokens else set()
        return " ".join(f"<{i}>" for i in ids if i not in drop)
