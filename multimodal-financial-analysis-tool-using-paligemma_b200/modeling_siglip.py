"""Drop-in for the reference's `modeling_siglip.py` (SigLIP So400m/14 vision tower).

Same public names, constructor arguments and state-dict keys as the reference
(`modeling_siglip.py:7-255`); the modules hold parameters only — the arithmetic runs in the
B200 engine (pg_b200.engine.PaliGemmaEngine.vision_features), never in torch ops.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


class SiglipVisionConfig:
    """Reference `SiglipVisionConfig` (modeling_siglip.py:7-34): same keywords and defaults;
    unknown keys are accepted and ignored."""

    def __init__(self, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                 num_attention_heads=12, num_channels=3, image_size=224, patch_size=16,
                 layer_norm_eps=1e-6, attention_dropout=0.0, num_image_tokens: Optional[int] = None,
                 **kwargs):
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.num_channels = num_channels
        self.patch_size = patch_size
        self.image_size = image_size
        self.attention_dropout = attention_dropout
        self.layer_norm_eps = layer_norm_eps
        self.num_image_tokens = num_image_tokens


class _EngineOnly(nn.Module):
    """Parameter holder: its math lives in the CUDA engine."""

    def forward(self, *args, **kwargs):
        raise RuntimeError(
            f"{type(self).__name__} only stores parameters in the B200 build; call "
            "PaliGemmaForConditionalGeneration.forward / SiglipVisionModel.forward instead")


class SiglipVisionEmbeddings(_EngineOnly):
    """Keys: patch_embedding.{weight,bias}, position_embedding.weight (modeling_siglip.py:36-60)."""

    def __init__(self, config: SiglipVisionConfig):
        super().__init__()
        self.config = config
        self.embed_dim = config.hidden_size
        self.image_size = config.image_size
        self.patch_size = config.patch_size
        self.patch_embedding = nn.Conv2d(config.num_channels, self.embed_dim, kernel_size=self.patch_size,
                                         stride=self.patch_size, padding="valid")
        self.num_patches = (self.image_size // self.patch_size) ** 2
        self.num_positions = self.num_patches
        self.position_embedding = nn.Embedding(self.num_positions, self.embed_dim)
        self.register_buffer("position_ids", torch.arange(self.num_positions).expand((1, -1)), persistent=False)


class SiglipAttention(_EngineOnly):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embed_dim = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.embed_dim // self.num_heads
        self.scale = self.head_dim ** -0.5
        self.dropout = config.attention_dropout
        self.k_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.v_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.q_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.out_proj = nn.Linear(self.embed_dim, self.embed_dim)


class SiglipMLP(_EngineOnly):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.fc1 = nn.Linear(config.hidden_size, config.intermediate_size)
        self.fc2 = nn.Linear(config.intermediate_size, config.hidden_size)


class SiglipEncoderLayer(_EngineOnly):
    def __init__(self, config: SiglipVisionConfig):
        super().__init__()
        self.embed_dim = config.hidden_size
        self.self_attn = SiglipAttention(config)
        self.layer_norm1 = nn.LayerNorm(self.embed_dim, eps=config.layer_norm_eps)
        self.mlp = SiglipMLP(config)
        self.layer_norm2 = nn.LayerNorm(self.embed_dim, eps=config.layer_norm_eps)


class SiglipEncoder(_EngineOnly):
    def __init__(self, config: SiglipVisionConfig):
        super().__init__()
        self.config = config
        self.layers = nn.ModuleList([SiglipEncoderLayer(config) for _ in range(config.num_hidden_layers)])


class SiglipVisionTransformer(_EngineOnly):
    def __init__(self, config: SiglipVisionConfig):
        super().__init__()
        self.config = config
        self.embeddings = SiglipVisionEmbeddings(config)
        self.encoder = SiglipEncoder(config)
        self.post_layernorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class SiglipVisionModel(nn.Module):
    """`SiglipVisionModel(pixel_values (B,3,S,S)) -> (B, P, hidden)` (modeling_siglip.py:246-255).

    When owned by PaliGemmaForConditionalGeneration the owner's engine is used; standalone use is
    not part of the accelerated path."""

    def __init__(self, config: SiglipVisionConfig = SiglipVisionConfig):
        super().__init__()
        self.config = config
        self.vision_model = SiglipVisionTransformer(config)
        self._owner = None  # set by PaliGemmaForConditionalGeneration (not a submodule registration)

    def forward(self, pixel_values) -> torch.Tensor:
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise RuntimeError("SiglipVisionModel runs through PaliGemmaForConditionalGeneration's B200 engine")
        return owner._engine_ready().vision_features(pixel_values)
