"""Drop-in for the reference's `utils.py`: `load_hf_model(model_path, device) -> (model, tokenizer)`
(reference utils.py:6-46).  Loads config.json + every *.safetensors shard straight into the
parameter holders in fp16 and ties the head; without shards the weights stay random (the
reference's no-accelerate branch, :39-41)."""
from __future__ import annotations

import glob
import json
import os

import torch

from modeling_gemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration


def load_hf_model(model_path: str, device: str = "cuda", dtype=torch.float16):
    from transformers import AutoTokenizer

    tokenizer = AutoTokenizer.from_pretrained(model_path, padding_side="right")
    with open(os.path.join(model_path, "config.json"), "r") as f:
        config = PaliGemmaConfig(**json.load(f))
    shards = sorted(glob.glob(os.path.join(model_path, "*.safetensors")))
    model = PaliGemmaForConditionalGeneration(config, init_weights=not shards)
    if shards:
        from safetensors.torch import load_file
        # init_weights=False leaves uninitialised storage: every parameter must come from some shard
        loaded, unexpected = set(), set()
        for path in shards:
            sd = load_file(path)
            res = model.load_state_dict(sd, strict=False)
            unexpected.update(res.unexpected_keys)
            loaded.update(k for k in sd if k not in res.unexpected_keys)
        never = [k for k, _ in model.named_parameters(remove_duplicate=False)
                 if k not in loaded and k != "language_model.lm_head.weight"]     # the head is tied below
        if never or unexpected:
            raise RuntimeError(f"checkpoint {model_path} does not match the model: {len(never)} parameter(s) never loaded "
                               f"{never[:5]}, {len(unexpected)} unexpected key(s) {sorted(unexpected)[:5]}")
    else:
        print("No *.safetensors found: keeping random-init weights.")
    model.to(device=device, dtype=dtype)
    model.tie_weights()
    return model, tokenizer
