"""Image pre-processing: the host restatement of Pillow's 8-bit bicubic resize (coefficient tables + integer passes)
against PIL itself on the CPU, and the CUDA path (pg_resample_u8 + pg_u8_to_chw through the C ABI) against the
reference's `process_images` recipe (processing_paligemma.py:13-50) on the GPU -- both bit-exact."""
import numpy as np
import pytest
import torch
from PIL import Image

from pg_b200 import preprocess as P
import processing_paligemma as PP

SIZES = [(37, 61), (224, 224), (448, 300), (1000, 750), (225, 223), (100, 224), (224, 500), (17, 900)]


def _image(h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("h,w", SIZES)
def test_resample_tables_reproduce_pillow(h, w):
    a = _image(h, w)
    ref = np.asarray(Image.fromarray(a, "RGB").resize((224, 224), resample=Image.Resampling.BICUBIC))
    np.testing.assert_array_equal(P.resample_u8_numpy(a, 224, 224), ref)


def test_value_table_matches_reference_dtypes():
    b = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    ref = PP.process_images([Image.fromarray(b, "RGB")], size=(16, 16), resample=Image.Resampling.BICUBIC,
                            rescale_factor=1 / 255.0, image_mean=PP.IMAGENET_STANDARD_MEAN,
                            image_std=PP.IMAGENET_STANDARD_STD)[0]
    np.testing.assert_array_equal(P.value_table()[b[:, :, 0]], ref[0])


@pytest.mark.gpu
@pytest.mark.parametrize("size", [224, 56])
def test_cuda_preprocess_is_bit_exact(size):
    imgs = [_image(h, w, seed=i) for i, (h, w) in enumerate(SIZES)]
    want = np.stack(PP.process_images([Image.fromarray(a, "RGB") for a in imgs], size=(size, size),
                                      resample=Image.Resampling.BICUBIC, rescale_factor=1 / 255.0,
                                      image_mean=PP.IMAGENET_STANDARD_MEAN, image_std=PP.IMAGENET_STANDARD_STD))
    got = P.preprocess_images_cuda([torch.from_numpy(a).cuda() for a in imgs], size)
    assert got.dtype == torch.float32 and tuple(got.shape) == (len(imgs), 3, size, size)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    got16 = P.preprocess_images_cuda([torch.from_numpy(a).cuda() for a in imgs[:2]], size, dtype=torch.bfloat16)
    torch.testing.assert_close(got16.cpu(), torch.from_numpy(want[:2]).to(torch.bfloat16), rtol=0, atol=0)
    with pytest.raises(RuntimeError):
        P.preprocess_images_cuda([torch.from_numpy(imgs[0])], size)   # host tensors are refused: no CPU path here
