"""Drop-in for the reference's `modeling_gemma.py`: same public classes, constructor arguments,
state-dict keys and call signatures (reference modeling_gemma.py:10-617), so `inference.py` and
`ablation_study_fixed.py` run unchanged with this directory first on `sys.path`.

The nn.Modules only hold parameters (HF checkpoint names).  All arithmetic of
`PaliGemmaForConditionalGeneration.forward` runs in hand-written sm_100a kernels through the
C ABI of `include/pg_b200.h` (pg_b200.engine).  There is no CPU path: calling the model on a
non-CUDA device raises.
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Tuple

import torch
from torch import nn

from modeling_siglip import SiglipVisionConfig, SiglipVisionModel, _EngineOnly
from pg_b200 import _cabi as cabi
from pg_b200._cabi import MASK_KIND
from pg_b200.engine import PagedKV, PaliGemmaEngine


# --------------------------------------------------------------------------------------- KVCache
class _LayerView:
    """List-like view of per-layer (B, n_kv, T, hd) tensors gathered from the paged pool."""

    def __init__(self, cache: "KVCache", which: str):
        self._cache, self._which = cache, which

    def __len__(self):
        c = self._cache
        if c._paged is not None:
            return c._paged.engine.dims.L if c._paged.length > 0 else 0
        return len(c._k_list if self._which == "k" else c._v_list)

    def __getitem__(self, i):
        c = self._cache
        if c._paged is not None:
            n = len(self)
            if not -n <= i < n:
                raise IndexError(i)
            return c._paged.gather(i % n, self._which)
        return (c._k_list if self._which == "k" else c._v_list)[i]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class KVCache:
    """Reference `KVCache` (modeling_gemma.py:10-36): `KVCache()`, `num_items()`,
    `update(k, v, layer_idx)`, `.key_cache` / `.value_cache`.

    Storage is the engine's paged bf16/fp16/fp32 pool: the model appends in place from the
    RoPE epilogue instead of `torch.cat`-ing a fresh tensor per layer per token.  A cache that has
    never met a model behaves as the reference's plain lists (`update` stores tensors); it is
    imported into pages the first time it is passed to `forward`."""

    def __init__(self) -> None:
        self._paged: Optional[PagedKV] = None
        self._k_list: List[torch.Tensor] = []
        self._v_list: List[torch.Tensor] = []

    @property
    def key_cache(self):
        return _LayerView(self, "k")

    @property
    def value_cache(self):
        return _LayerView(self, "v")

    def num_items(self) -> int:
        if self._paged is not None:
            return self._paged.length
        return 0 if not self._k_list else self._k_list[0].shape[-2]

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor, layer_idx: int
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._paged is not None:
            raise RuntimeError("this KVCache is owned by the B200 engine; the model appends to it in place")
        if len(self._k_list) <= layer_idx:
            self._k_list.append(key_states)
            self._v_list.append(value_states)
        else:
            self._k_list[layer_idx] = torch.cat([self._k_list[layer_idx], key_states], dim=-2)
            self._v_list[layer_idx] = torch.cat([self._v_list[layer_idx], value_states], dim=-2)
        return self._k_list[layer_idx], self._v_list[layer_idx]

    # -- engine side
    def _bind(self, engine: PaliGemmaEngine, batch: int) -> PagedKV:
        if self._paged is not None:
            if self._paged.engine is not engine:
                raise RuntimeError("KVCache was filled by a different model/engine")
            return self._paged
        paged = engine.new_kv(batch)
        if self._k_list:  # import list contents (B, n_kv, T, hd) into pages
            d, T = engine.dims, self._k_list[0].shape[-2]
            if len(self._k_list) != d.L:
                raise ValueError("KVCache lists do not cover every layer")
            paged.reserve(T)
            tok = torch.arange(T, device=engine.device)
            pages = paged.page_table[:, : (T + engine.page_size - 1) // engine.page_size]
            pg = pages[:, tok // engine.page_size].long()            # (B, T)
            off = (tok % engine.page_size).expand_as(pg)
            for li in range(d.L):
                k = self._k_list[li].to(device=engine.device, dtype=engine.dtype).permute(0, 2, 1, 3).reshape(batch, T, -1)
                v = self._v_list[li].to(device=engine.device, dtype=engine.dtype).permute(0, 2, 1, 3).reshape(batch, T, -1)
                engine.k_pool[li][pg, off] = k
                engine.v_pool[li][pg, off] = v
            paged.length = T
            paged.kv_len.fill_(T)
            self._k_list, self._v_list = [], []
        self._paged = paged
        return paged


# --------------------------------------------------------------------------------------- configs
class GemmaConfig:
    """Reference `GemmaConfig` (modeling_gemma.py:39-71)."""

    def __init__(self, vocab_size, hidden_size, intermediate_size, num_hidden_layers, num_attention_heads,
                 num_key_value_heads, head_dim=256, max_position_embeddings=8192, rms_norm_eps=1e-6,
                 rope_theta=10000.0, attention_bias=False, attention_dropout=0.0, pad_token_id=None, **kwargs):
        self.vocab_size = vocab_size
        self.max_position_embeddings = max_position_embeddings
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.head_dim = head_dim
        self.num_key_value_heads = num_key_value_heads
        self.rms_norm_eps = rms_norm_eps
        self.rope_theta = rope_theta
        self.attention_bias = attention_bias
        self.attention_dropout = attention_dropout
        self.pad_token_id = pad_token_id


class PaliGemmaConfig:
    """Reference `PaliGemmaConfig` (modeling_gemma.py:74-105): builds the two sub-configs from
    dicts, overrides vocab_size from the text config and derives num_image_tokens."""

    def __init__(self, vision_config=None, text_config=None, ignore_index=-100, image_token_index=256000,
                 vocab_size=257152, projection_dim=2048, hidden_size=2048, pad_token_id=None, **kwargs):
        self.ignore_index = ignore_index
        self.image_token_index = image_token_index
        self.projection_dim = projection_dim
        self.hidden_size = hidden_size
        self.is_encoder_decoder = False
        self.pad_token_id = pad_token_id
        self.vision_config = SiglipVisionConfig(**vision_config)
        self.text_config = GemmaConfig(**text_config, pad_token_id=pad_token_id)
        self.vocab_size = self.text_config.vocab_size
        self.text_config.num_image_tokens = (self.vision_config.image_size // self.vision_config.patch_size) ** 2
        self.vision_config.projection_dim = projection_dim


# --------------------------------------------------------------------------------------- parameter holders
class GemmaRMSNorm(_EngineOnly):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.zeros(dim))


class GemmaMLP(_EngineOnly):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.hidden_size = config.hidden_size
        self.intermediate_size = config.intermediate_size
        self.gate_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
        self.up_proj = nn.Linear(self.hidden_size, self.intermediate_size, bias=False)
        self.down_proj = nn.Linear(self.intermediate_size, self.hidden_size, bias=False)


class GemmaRotaryEmbedding(_EngineOnly):
    """Holds `inv_freq` (non-persistent, modeling_gemma.py:151-152).  `forward` is a monkey-patch
    target of ablation_study_fixed.py:339-342; the engine computes RoPE natively with the patched
    semantics (position clamp, fp32 angles, cos/sin rounded to the model dtype)."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000, device=None):
        super().__init__()
        self.dim = dim
        self.max_position_embeddings = max_position_embeddings
        self.base = base
        inv_freq = 1.0 / (self.base ** (torch.arange(0, self.dim, 2, dtype=torch.int64).float() / self.dim))
        self.register_buffer("inv_freq", tensor=inv_freq, persistent=False)


class GemmaAttention(_EngineOnly):
    def __init__(self, config: GemmaConfig, layer_idx: Optional[int] = None):
        super().__init__()
        self.config = config
        self.layer_idx = layer_idx
        self.attention_dropout = config.attention_dropout
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = config.head_dim
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.max_position_embeddings = config.max_position_embeddings
        self.rope_theta = config.rope_theta
        self.is_causal = True
        assert self.hidden_size % self.num_heads == 0
        self.q_proj = nn.Linear(self.hidden_size, self.num_heads * self.head_dim, bias=config.attention_bias)
        self.k_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=config.attention_bias)
        self.v_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=config.attention_bias)
        self.o_proj = nn.Linear(self.num_heads * self.head_dim, self.hidden_size, bias=config.attention_bias)
        self.rotary_emb = GemmaRotaryEmbedding(self.head_dim, max_position_embeddings=self.max_position_embeddings,
                                               base=self.rope_theta)


class GemmaDecoderLayer(_EngineOnly):
    def __init__(self, config: GemmaConfig, layer_idx: int):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = GemmaAttention(config=config, layer_idx=layer_idx)
        self.mlp = GemmaMLP(config)
        self.input_layernorm = GemmaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = GemmaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)


class GemmaModel(_EngineOnly):
    def __init__(self, config: GemmaConfig):
        super().__init__()
        self.config = config
        self.padding_idx = config.pad_token_id
        self.vocab_size = config.vocab_size
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, self.padding_idx)
        self.layers = nn.ModuleList([GemmaDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = GemmaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)

    def get_input_embeddings(self):
        return self.embed_tokens


class GemmaForCausalLM(_EngineOnly):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.model = GemmaModel(config)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def tie_weights(self):
        self.lm_head.weight = self.model.embed_tokens.weight


class PaliGemmaMultiModalProjector(nn.Module):
    def __init__(self, config: PaliGemmaConfig):
        super().__init__()
        self.linear = nn.Linear(config.vision_config.hidden_size, config.vision_config.projection_dim, bias=True)
        self._owner = None

    def forward(self, image_features):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise RuntimeError("the projector runs through PaliGemmaForConditionalGeneration's B200 engine")
        return owner._engine_ready().project(image_features)


# --------------------------------------------------------------------------------------- top level
class PaliGemmaForConditionalGeneration(nn.Module):
    """Reference `PaliGemmaForConditionalGeneration` (modeling_gemma.py:440-617).

    `init_weights=False` skips the (slow, CPU) default initialisation when a checkpoint will be
    loaded right after; `engine_options` are forwarded to PaliGemmaEngine (page_size,
    kv_pool_tokens, gemm_impl)."""

    def __init__(self, config: PaliGemmaConfig, init_weights: bool = True, **engine_options):
        super().__init__()
        self.config = config
        self._engine: Optional[PaliGemmaEngine] = None
        self._engine_key = None
        self._engine_options = engine_options
        self._mask_flag = None      # pinned int32 written by pg_decode_inputs when a decode-step mask is padded
        self._mask_flag_np = None
        if init_weights:
            self._build(config)
        else:
            with torch.device("meta"):
                self._build(config)
            self.to_empty(device="cpu")
            # buffers are not part of any checkpoint: recompute them
            for m in self.modules():
                if isinstance(m, GemmaRotaryEmbedding):
                    m.inv_freq = 1.0 / (m.base ** (torch.arange(0, m.dim, 2, dtype=torch.int64).float() / m.dim))
            emb = self.vision_tower.vision_model.embeddings
            emb.position_ids = torch.arange(emb.num_positions).expand((1, -1))
        self.vocab_size = config.vocab_size
        self.pad_token_id = self.config.pad_token_id if self.config.pad_token_id is not None else -1
        self.vision_tower._owner = weakref.ref(self)
        self.multi_modal_projector._owner = weakref.ref(self)

    def _build(self, config):
        self.vision_tower = SiglipVisionModel(config.vision_config)
        self.multi_modal_projector = PaliGemmaMultiModalProjector(config)
        self.language_model = GemmaForCausalLM(config.text_config)

    # ---- reference API
    def tie_weights(self):
        self._engine = None
        return self.language_model.tie_weights()

    def get_output_embeddings(self):
        return self.language_model.lm_head

    def prepare_inputs_for_generation(self, input_ids=None, **kwargs):
        return {"input_ids": input_ids, **kwargs}

    # ---- engine lifetime
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if self._engine is not None and self._fingerprint() != self._engine_key:
            self._engine = None
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kwargs):
        res = super().load_state_dict(state_dict, strict=strict, **kwargs)
        self._engine = None
        return res

    def _fingerprint_modules(self):
        ms = getattr(self, "_fp_modules", None)
        if ms is None:      # module objects are stable (only their parameters' storage moves); look them up once
            ms = self._fp_modules = (self.language_model.model.embed_tokens, self.language_model.lm_head,
                                     self.language_model.model.layers[-1].mlp.down_proj,
                                     self.vision_tower.vision_model.embeddings.patch_embedding,
                                     self.multi_modal_projector.linear)
        return ms

    def _fingerprint(self):
        # storage address + dtype + device of five parameters spread over the model: any .to() / load / re-tie moves them
        return tuple((m.weight.data_ptr(), m.weight.dtype, m.weight.device) for m in self._fingerprint_modules())

    def _engine_ready(self) -> PaliGemmaEngine:
        if self._engine is None or self._fingerprint() != self._engine_key:
            params = dict(self.named_parameters(remove_duplicate=False))
            dev = self.language_model.model.embed_tokens.weight.device
            if dev.type != "cuda":
                raise RuntimeError("PaliGemma (B200 build) has no CPU path: move the model to a CUDA device first")

            def adopt(key, view):
                params[key].data = view

            self._engine = PaliGemmaEngine(self.config, {k: v.data for k, v in params.items()},
                                           adopt=adopt, **self._engine_options)
            self._engine_key = self._fingerprint()
        return self._engine

    # ---- monkey-patch target kept for API parity (ablation_study_fixed.py:335-337)
    def _merge_input_ids_with_image_features(self, image_features, inputs_embeds, input_ids, attention_mask,
                                             kv_cache: Optional[KVCache] = None):
        """Same triple as the reference (modeling_gemma.py:468-537, with the patched position
        shape of ablation_study_fixed.py:130-133): merged embeddings (before the sqrt(D)
        normaliser), the all-zero additive mask, position ids.  forward() does not call this:
        the merge is fused into the embedding kernel."""
        eng = self._engine_ready()
        d = eng.dims
        B, q = input_ids.shape
        ids = input_ids.to(eng.device).contiguous().view(-1)
        img = None if image_features is None or image_features.numel() == 0 else \
            image_features.reshape(-1, d.D).to(eng.dtype).contiguous()
        out = torch.empty((B * q, d.D), dtype=eng.dtype, device=eng.device)
        cabi.check(cabi.lib().pg_embed_merge(out.data_ptr(), ids.data_ptr(), eng.emb.data_ptr(),
                                             None if img is None else img.data_ptr(), B * q, d.D, d.V,
                                             d.image_token_index, d.pad_token_id, 0 if img is None else img.shape[0],
                                             eng.img_div, 1.0, eng.err_flag.data_ptr(), eng.dt, cabi.stream()),
                   "embed_merge")
        cached = 0 if kv_cache is None else kv_cache.num_items()
        mask = torch.zeros((B, 1, q, cached + q), dtype=eng.dtype, device=eng.device)
        if cached > 0:
            pos = attention_mask.cumsum(-1)[:, -1:]
        else:
            pos = torch.arange(attention_mask.shape[1], device=eng.device).unsqueeze(0).expand(B, -1)
        return out.view(B, q, d.D), mask, pos

    # ---- forward
    def forward(self, input_ids: Optional[torch.LongTensor] = None, pixel_values: Optional[torch.FloatTensor] = None,
                attention_mask: Optional[torch.Tensor] = None, inputs_embeds: Optional[torch.FloatTensor] = None,
                kv_cache: Optional[KVCache] = None, labels: Optional[torch.LongTensor] = None,
                return_dict: bool = True, **kwargs):
        if attention_mask is None:
            raise ValueError("attention_mask must be provided")
        self._raise_deferred_mask_error()
        if self._engine is not None:
            self._engine.check_errors()      # bad ids met by an earlier cached step (pinned flag, no sync)
        # cached single-token steps check the mask on the device (pg_decode_inputs) and report it at the next call
        # instead of paying the reference's host synchronisation (modeling_gemma.py:559) on every token
        fast = (kv_cache is not None and input_ids is not None and inputs_embeds is None and labels is None
                and input_ids.dim() == 2 and input_ids.shape[1] == 1 and input_ids.is_cuda
                and input_ids.dtype == torch.int64 and kv_cache._paged is not None and kv_cache._paged.length > 0
                and attention_mask.is_cuda and attention_mask.is_contiguous() and attention_mask.dtype in MASK_KIND)
        if not fast:
            assert bool(torch.all(attention_mask == 1)), "The input cannot be padded"
        if inputs_embeds is not None:
            raise NotImplementedError("inputs_embeds (PEFT) is outside the accelerated inference path")
        if input_ids is None:
            raise ValueError("You must provide either input_ids or inputs_embeds")
        if labels is not None:
            raise NotImplementedError("the loss branch (modeling_gemma.py:596-603) is training-only")
        eng = self._engine_ready()
        B, q = input_ids.shape
        paged = None if kv_cache is None else kv_cache._bind(eng, B)
        cached = 0 if paged is None else paged.length
        if paged is not None and cached > 0 and q == 1:
            logits = self._decode_one(eng, paged, input_ids, int(attention_mask.shape[1]), attention_mask if fast else None)
        else:
            feats = None
            if pixel_values is not None:
                feats = eng.encode_images(pixel_values.to(eng.device))
            logits = eng.text_forward(input_ids, feats, paged, position_value=int(attention_mask.shape[1]))
            eng.check_errors(sync=True)      # this path already synchronised for the mask check above
        if return_dict:
            out = {"logits": logits}
            if kv_cache is not None:
                out["kv_cache"] = kv_cache
            return out
        return (logits,)

    def _decode_one(self, eng: PaliGemmaEngine, paged: PagedKV, input_ids, position: int,
                    mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One cached step through the graph-captured decode kernels (vision tower skipped: a
        single new token has no image slot to fill).  `mask` (already known to be a
        contiguous CUDA tensor of a supported dtype) is checked for padding by the same launch
        that stages ids / positions; a violation raises AssertionError at the next forward()."""
        ds = eng.decode_state(paged.batch)
        if mask is not None and input_ids.is_contiguous():
            if self._mask_flag is None:
                self._mask_flag = torch.zeros(1, dtype=torch.int32).pin_memory()
                self._mask_flag_np = self._mask_flag.numpy()
            cabi.check(cabi.lib().pg_decode_inputs(ds.ids.data_ptr(), input_ids.data_ptr(), ds.pos.data_ptr(), position,
                                                   mask.data_ptr(), MASK_KIND[mask.dtype], mask.numel(),
                                                   self._mask_flag.data_ptr(), paged.batch, cabi.stream()),
                       "decode_inputs")
        else:
            ds.ids.copy_(input_ids.reshape(-1), non_blocking=True)
            ds.pos.fill_(position)
        ds.want_full_logits = True       # tensor parallel: forward() returns the whole vocabulary row
        ds.run_steps(paged, 1)
        # a fresh tensor per call, as the reference returns (callers keep earlier steps' logits)
        return ds.logits.clone().view(paged.batch, 1, -1)

    def _raise_deferred_mask_error(self):
        if self._mask_flag_np is not None and self._mask_flag_np[0] != 0:
            torch.cuda.current_stream().synchronize()
            self._mask_flag_np[0] = 0
            raise AssertionError("The input cannot be padded (attention mask of an earlier cached decode step)")

    # ---- engine-native generation loop (not in the reference; used by bench.py)
    @torch.no_grad()
    def generate(self, input_ids, pixel_values, max_new_tokens: int, do_sample: bool = False,
                 temperature: float = 0.8, top_p: float = 0.9, seed: int = 0, use_kv_cache: bool = True):
        from pg_b200.generate import generate
        return generate(self._engine_ready(), input_ids, pixel_values, max_new_tokens, do_sample=do_sample,
                        temperature=temperature, top_p=top_p, seed=seed, use_kv_cache=use_kv_cache)
