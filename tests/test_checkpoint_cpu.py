"""CPU checks of the checkpoint path (reference utils.py:6-46, inference.py:104-112): a synthetic
PaliGemma checkpoint written to disk the way the hub stores it (config.json, two *.safetensors shards with HF key
names, a `tokenizers` fast tokenizer) goes through the drop-in `utils.load_hf_model`, and the drop-in
`PaliGemmaProcessor` is driven by that real HF tokenizer object instead of the offline stub."""
import json
import os

import numpy as np
import pytest
import torch

from pg_b200 import synth
import processing_paligemma as PP
import utils as U


def _write_tokenizer(path):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    vocab = {"<pad>": 0, "<eos>": 1, "<bos>": 2, "<unk>": 3, "caption": 4, "en": 5, "\n": 6, "describe": 7, "chart": 8}
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(" ", behavior="removed"),
                                                 pre_tokenizers.Split("\n", behavior="isolated")])
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<bos>", eos_token="<eos>", pad_token="<pad>",
                                   unk_token="<unk>")
    fast.save_pretrained(path)
    return vocab


@pytest.fixture(scope="module")
def checkpoint_dir(tmp_path_factory):
    from safetensors.torch import save_file
    d = str(tmp_path_factory.mktemp("ckpt"))
    cfg = synth.TINY
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    sd = {k: v.contiguous() for k, v in synth.synth_state_dict(cfg).items() if "lm_head" not in k}
    keys = sorted(sd)
    half = len(keys) // 2
    save_file({k: sd[k] for k in keys[:half]}, os.path.join(d, "model-00001-of-00002.safetensors"))
    save_file({k: sd[k] for k in keys[half:]}, os.path.join(d, "model-00002-of-00002.safetensors"))
    vocab = _write_tokenizer(d)
    return d, cfg, sd, vocab


def test_load_hf_model_reads_every_shard_and_ties_the_head(checkpoint_dir):
    d, cfg, sd, _ = checkpoint_dir
    model, tokenizer = U.load_hf_model(d, device="cpu", dtype=torch.float32)
    got = model.state_dict()
    for k, v in sd.items():
        torch.testing.assert_close(got[k], v, rtol=0, atol=0, msg=k)
    emb = model.language_model.model.embed_tokens.weight
    assert model.language_model.lm_head.weight.data_ptr() == emb.data_ptr()        # tie_weights (utils.py:44)
    assert model.config.vision_config.num_image_tokens == cfg["vision_config"]["num_image_tokens"]
    assert tokenizer.bos_token == "<bos>" and tokenizer.padding_side == "right"
    # no CPU path: the forward refuses to run off the GPU instead of silently falling back
    with pytest.raises(RuntimeError):
        model(input_ids=torch.zeros((1, 4), dtype=torch.int64), pixel_values=None,
              attention_mask=torch.ones((1, 4), dtype=torch.int64))


def test_processor_with_a_real_hf_tokenizer(checkpoint_dir):
    from PIL import Image
    d, cfg, _, vocab = checkpoint_dir
    _, tokenizer = U.load_hf_model(d, device="cpu", dtype=torch.float32)
    n_img, size = cfg["vision_config"]["num_image_tokens"], cfg["vision_config"]["image_size"]
    n_before = len(tokenizer)
    proc = PP.PaliGemmaProcessor(tokenizer, n_img, size)
    # <image> + 1024 <loc> + 128 <seg> tokens appended after the base vocabulary (processing_paligemma.py:63-72)
    assert len(tokenizer) == n_before + 1 + 1024 + 128
    assert proc.image_token_id == n_before and tokenizer.add_bos_token is False and tokenizer.add_eos_token is False
    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 256, size=(37, 61, 3), dtype=np.uint8), "RGB")
    out = proc(text=["caption en"], images=[img])
    ids = out["input_ids"]
    assert ids.dtype == torch.int64 and tuple(ids.shape) == (1, n_img + 4)
    assert ids[0, :n_img].tolist() == [proc.image_token_id] * n_img                 # image slots first
    assert ids[0, n_img:].tolist() == [vocab["<bos>"], vocab["caption"], vocab["en"], vocab["\n"]]
    assert out["attention_mask"].tolist() == [[1] * (n_img + 4)]
    px = out["pixel_values"]
    assert tuple(px.shape) == (1, 3, size, size) and px.dtype == torch.float32
    assert float(px.min()) >= -1.0 and float(px.max()) <= 1.0
    # the image branch equals the reference recipe: bicubic resize, /255, (x - 0.5) / 0.5, CHW
    ref = np.asarray(img.resize((size, size), resample=Image.Resampling.BICUBIC)).astype(np.float32) / 255.0
    ref = ((ref - 0.5) / 0.5).transpose(2, 0, 1)
    np.testing.assert_allclose(px[0].numpy(), ref, rtol=0, atol=1e-6)


def write_sentencepiece_tokenizer(path, vocab_size=200):
    """A REAL SentencePiece model (trained here on a synthetic corpus: no network) stored the way the hub stores
    Gemma's: tokenizer.model + tokenizer_config.json; `AutoTokenizer.from_pretrained` (reference utils.py:8) loads it."""
    import random
    import sentencepiece as spm
    words = ("caption en describe the chart revenue grew by ten percent in q3 what is shown table figure axis value "
             "total net income margin year over quarter").split()
    rng = random.Random(0)
    corpus = os.path.join(path, "corpus.txt")
    with open(corpus, "w") as f:
        for _ in range(2000):
            f.write(" ".join(rng.choices(words, k=8)) + "\n")
    spm.SentencePieceTrainer.train(input=corpus, model_prefix=os.path.join(path, "tokenizer"), vocab_size=vocab_size,
                                   model_type="bpe", pad_id=0, eos_id=1, bos_id=2, unk_id=3, pad_piece="<pad>",
                                   eos_piece="<eos>", bos_piece="<bos>", unk_piece="<unk>", user_defined_symbols=["\n"],
                                   minloglevel=2)
    os.remove(corpus)
    os.remove(os.path.join(path, "tokenizer.vocab"))
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"tokenizer_class": "GemmaTokenizer", "bos_token": "<bos>", "eos_token": "<eos>", "pad_token": "<pad>",
                   "unk_token": "<unk>"}, f)


def test_processor_with_a_sentencepiece_tokenizer(tmp_path):
    """The processor driven by a SentencePiece-backed Gemma tokenizer: `<image>` is appended right after the base
    vocabulary (processing_paligemma.py:63-66), the prompt is image tokens + <bos> + text + newline (:10-11)."""
    from PIL import Image
    from transformers import AutoTokenizer
    d = str(tmp_path)
    write_sentencepiece_tokenizer(d)
    tok = AutoTokenizer.from_pretrained(d, padding_side="right")
    assert type(tok).__name__.startswith("GemmaTokenizer")
    n_before = len(tok)
    proc = PP.PaliGemmaProcessor(tok, 16, 56)
    assert proc.image_token_id == n_before
    img = Image.fromarray(np.zeros((40, 52, 3), dtype=np.uint8), "RGB")
    out = proc(text=["caption en"], images=[img])
    ids = out["input_ids"][0].tolist()
    assert ids[:16] == [n_before] * 16 and ids[16] == tok.bos_token_id
    text_ids = ids[17:]
    assert tok.decode(text_ids, skip_special_tokens=True).strip() == "caption en"
    assert tok.decode(text_ids).endswith("\n")                    # the newline the processor appends survives the round trip


def test_load_hf_model_refuses_an_incomplete_checkpoint(checkpoint_dir, tmp_path):
    """A shard set that misses parameters must not leave uninitialised weights behind silently."""
    import shutil
    from safetensors.torch import load_file, save_file
    d, cfg, sd, _ = checkpoint_dir
    bad = str(tmp_path / "bad")
    shutil.copytree(d, bad)
    shard = os.path.join(bad, "model-00002-of-00002.safetensors")
    part = load_file(shard)
    dropped = sorted(part)[0]
    del part[dropped]
    save_file(part, shard)
    with pytest.raises(RuntimeError, match="never loaded"):
        U.load_hf_model(bad, device="cpu", dtype=torch.float32)
    part["not.a.parameter"] = torch.zeros(3)
    part[dropped] = sd[dropped]
    save_file(part, shard)
    with pytest.raises(RuntimeError, match="unexpected"):
        U.load_hf_model(bad, device="cpu", dtype=torch.float32)
