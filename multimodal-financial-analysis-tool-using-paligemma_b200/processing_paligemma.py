"""Drop-in for the reference's `processing_paligemma.py` (host-side pre-processing, once per
request): same `PaliGemmaProcessor(tokenizer, num_image_tokens, image_size)` and
`processor(text=[str], images=[PIL.Image]) -> {"pixel_values", "input_ids", "attention_mask"}`
contract (reference processing_paligemma.py:52-117)."""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np
import torch
from PIL import Image

IMAGENET_STANDARD_MEAN = [0.5, 0.5, 0.5]
IMAGENET_STANDARD_STD = [0.5, 0.5, 0.5]


def add_image_tokens_to_prompt(prefix_prompt, bos_token, image_seq_len, image_token):
    """`<image>`*n + BOS + prompt + newline (reference :10-11)."""
    return "".join([image_token * image_seq_len, str(bos_token), prefix_prompt, "\n"])


def image_to_chw(image: Image.Image, size: int, mean: Sequence[float], std: Sequence[float],
                 rescale_factor: float = 1 / 255.0) -> np.ndarray:
    """PIL bicubic resize -> float32 /255 -> (x-mean)/std -> CHW (reference :13-50).  Like the
    reference there is no RGB conversion: a grayscale or RGBA image fails the channel transpose
    / normalisation the same way."""
    arr = np.array(image.resize((size, size), resample=Image.Resampling.BICUBIC))
    arr = (arr * rescale_factor).astype(np.float32)
    arr = (arr - np.array(mean, dtype=arr.dtype)) / np.array(std, dtype=arr.dtype)
    return arr.transpose(2, 0, 1)


def process_images(images: Iterable[Image.Image], size, resample=None, rescale_factor=None,
                   image_mean=None, image_std=None) -> List[np.ndarray]:
    """Signature-compatible wrapper (reference :31-50); `size` is (height, width), square here."""
    assert size[0] == size[1], "PaliGemma pre-processing is square"
    return [image_to_chw(im, size[0], image_mean, image_std,
                         1 / 255.0 if rescale_factor is None else rescale_factor) for im in images]


class PaliGemmaProcessor:
    IMAGE_TOKEN = "<image>"

    def __init__(self, tokenizer, num_image_tokens: int, image_size: int):
        self.image_seq_length = num_image_tokens
        self.image_size = image_size
        # same tokenizer mutations as the reference (:63-75): <image>, 1024 <locXXXX>, 128 <segXXX>,
        # and BOS/EOS handled by the prompt builder
        tokenizer.add_special_tokens({"additional_special_tokens": [self.IMAGE_TOKEN]})
        extra = ["<loc%04d>" % i for i in range(1024)] + ["<seg%03d>" % i for i in range(128)]
        tokenizer.add_tokens(extra)
        self.image_token_id = tokenizer.convert_tokens_to_ids(self.IMAGE_TOKEN)
        tokenizer.add_bos_token = False
        tokenizer.add_eos_token = False
        self.tokenizer = tokenizer

    def __call__(self, text: List[str], images: List[Image.Image], padding: str = "longest",
                 truncation: bool = True) -> dict:
        assert len(images) == 1 and len(text) == 1, f"Received {len(images)} images for {len(text)} prompts."
        chw = process_images(images, size=(self.image_size, self.image_size), resample=Image.Resampling.BICUBIC,
                             rescale_factor=1 / 255.0, image_mean=IMAGENET_STANDARD_MEAN,
                             image_std=IMAGENET_STANDARD_STD)
        pixel_values = torch.tensor(np.stack(chw, axis=0))
        prompts = [add_image_tokens_to_prompt(p, self.tokenizer.bos_token, self.image_seq_length, self.IMAGE_TOKEN)
                   for p in text]
        tokens = self.tokenizer(prompts, return_tensors="pt", padding=padding, truncation=truncation)
        return {"pixel_values": pixel_values, **tokens}
