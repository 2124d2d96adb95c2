"""Boundary evidence: the reference's OWN driver files run UNMODIFIED against the drop-in modules.

`oracle/_ref/` holds `ablation_study_fixed.py` and `inference.py` exactly as the reference ships them (staged by
oracle/build_ref.py, never committed).  With this repo's package directory first on sys.path their
`from modeling_gemma import ...`, `from processing_paligemma import ...`, `from utils import load_hf_model` resolve to the
drop-ins, so the harness code -- load_model_simple (config.json, safetensors shards, `.half()`, `.to(device, dtype)`,
`tie_weights`, the two monkey-patches on the live model), run_inference (model.to(dtype) per run, prefill + refeed loop,
cache on and off), reset_model_state, and inference.py's main()/test_inference -- executes line for line on the B200
engine.  Tokens must equal the CPU oracle's for the same checkpoint, image and prompt (fp32 run of fp16-rounded weights,
as the harness produces them)."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import build_ref  # noqa: E402
from oracle import paligemma_oracle as O  # noqa: E402
from pg_b200 import synth  # noqa: E402
import modeling_gemma as MG  # noqa: E402
import processing_paligemma as PP  # noqa: E402

if not build_ref.available():
    pytest.skip("oracle/_ref not staged (python oracle/build_ref.py in the build container)", allow_module_level=True)


def _write_tokenizer(path, n_vocab):
    """A `tokenizers` fast tokenizer with exactly n_vocab entries, so the `<image>` token the processor appends gets id
    n_vocab == config.image_token_index (as 257152 in the real checkpoint)."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    vocab = {"<pad>": 0, "<eos>": 1, "<bos>": 2, "<unk>": 3, "caption": 4, "en": 5, "\n": 6, "describe": 7, "chart": 8}
    for i in range(len(vocab), n_vocab):
        vocab[f"w{i}"] = i
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(" ", behavior="removed"),
                                                 pre_tokenizers.Split("\n", behavior="isolated")])
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<bos>", eos_token="<eos>", pad_token="<pad>",
                                   unk_token="<unk>")
    fast.save_pretrained(path)


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    from PIL import Image
    from safetensors.torch import save_file
    d = str(tmp_path_factory.mktemp("ckpt"))
    cfg = synth.TINY
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    sd = {k: v.contiguous() for k, v in synth.synth_state_dict(cfg, tie=False).items()}
    keys = sorted(sd)
    save_file({k: sd[k] for k in keys[:len(keys) // 2]}, os.path.join(d, "model-00001-of-00002.safetensors"))
    save_file({k: sd[k] for k in keys[len(keys) // 2:]}, os.path.join(d, "model-00002-of-00002.safetensors"))
    _write_tokenizer(d, cfg["image_token_index"])
    rng = np.random.default_rng(3)
    img_path = os.path.join(d, "chart.png")
    Image.fromarray(rng.integers(0, 256, size=(90, 130, 3), dtype=np.uint8), "RGB").save(img_path)
    # reference drivers: imported from oracle/_ref AFTER the package directory, so the model / processor / utils modules
    # they import by name are the drop-ins
    sys.path.append(build_ref.OUT)
    sys.modules.setdefault("fire", types.SimpleNamespace(Fire=lambda fn: None))     # inference.py:5 (CLI only)
    import ablation_study_fixed as A
    import inference as I
    assert A.PaliGemmaForConditionalGeneration is MG.PaliGemmaForConditionalGeneration and A.KVCache is MG.KVCache
    assert I.PaliGemmaProcessor is PP.PaliGemmaProcessor
    assert os.path.dirname(os.path.abspath(A.__file__)) == build_ref.OUT
    # what the harness's checkpoint looks like to the model: fp32 values rounded through fp16 (load_model_simple :316)
    sd_h = {k: v.half().float() for k, v in sd.items()}
    sd_h["language_model.lm_head.weight"] = sd_h["language_model.model.embed_tokens.weight"]
    return d, cfg, sd_h, img_path, A, I


def _inputs(cfg, ckpt, img_path):
    from PIL import Image
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(ckpt, padding_side="right")
    v = cfg["vision_config"]
    proc = PP.PaliGemmaProcessor(tok, v["num_image_tokens"], v["image_size"])
    out = proc(text=["caption en"], images=[Image.open(img_path)])
    return out["input_ids"], out["pixel_values"], tok


def test_ablation_harness_runs_unmodified(setup):
    """ablation_study_fixed.py: load_model_simple + run_inference, cache on (prompt cached twice: :193-199,216-221) and
    cache off (:245-251), temperature 0 -> argmax; token ids against the oracle."""
    ckpt, cfg, sd_h, img_path, A, I = setup
    assert A.DEVICE == "cuda"
    model, tokenizer = A.load_model_simple(ckpt, "cuda")
    assert isinstance(model, MG.PaliGemmaForConditionalGeneration)
    # the harness monkey-patched the live model (:335-342): the attributes exist and took the patch
    assert model._merge_input_ids_with_image_features.__func__ is A.patched_merge_input_ids_with_image_features
    assert model.language_model.model.layers[0].self_attn.rotary_emb.forward.__func__ is A.patched_rotary_forward
    v = model.config.vision_config
    processor = A.PaliGemmaProcessor(tokenizer, v.num_image_tokens, v.image_size)
    ids, pix, _ = _inputs(cfg, ckpt, img_path)
    n = 7
    # warm-up call form of main() (:383-388)
    warm = A.move_inputs_to_device(processor(text=["warmup"], images=[__import__("PIL.Image").Image.open(img_path)]), "cuda")
    with torch.no_grad():
        model(**warm, kv_cache=None)
    for use_cache in (True, False):
        A.reset_model_state(model)
        run = {"name": "t", "kv_cache": use_cache, "dtype": torch.float32, "temperature": 0.0, "max_tokens": n}
        res = A.run_inference(model, processor, img_path, "caption en", run, return_tokens=True)
        if use_cache:
            want = O.generate_cached(sd_h, cfg, ids, pix, n, patched=True, refeed_prompt=True)[0].tolist()
        else:
            want = O.generate_uncached(sd_h, cfg, ids, pix, n)[0].tolist()
        assert res["token_ids"] == want, (use_cache, res["token_ids"], want)
        assert res["tokens_generated"] == n and res["peak_memory_mb"] > 0 and res["steady_state_tps"] > 0
    # and in the harness's own dtype (fp16) the run completes with in-vocabulary tokens
    run = {"name": "t", "kv_cache": True, "dtype": torch.float16, "temperature": 0.0, "max_tokens": n}
    res = A.run_inference(model, processor, img_path, "caption en", run, return_tokens=True)
    assert len(res["token_ids"]) == n and all(0 <= t < cfg["vocab_size"] for t in res["token_ids"])


def test_inference_driver_runs_unmodified(setup, capsys):
    """inference.py: test_inference() on an fp32 model (tokens against the oracle, EOS rule included) and main() end to
    end through the drop-in utils.load_hf_model (fp16, the reference's default)."""
    ckpt, cfg, sd_h, img_path, A, I = setup
    ids, pix, tok = _inputs(cfg, ckpt, img_path)
    import utils as U
    model, tokenizer = U.load_hf_model(ckpt, "cuda", dtype=torch.float32)
    model = model.to("cuda").eval()
    v = model.config.vision_config
    processor = PP.PaliGemmaProcessor(tokenizer, v.num_image_tokens, v.image_size)
    n = 8
    with torch.no_grad():
        text = I.test_inference(model, processor, "cuda", "caption en", img_path, n, 0.8, 0.9, False)
    sd32 = synth.synth_state_dict(cfg)
    want = O.generate_cached(sd32, cfg, ids, pix, n, patched=False)[0].tolist()
    if tokenizer.eos_token_id in want:
        want = want[:want.index(tokenizer.eos_token_id) + 1]
    assert text == "caption en" + tokenizer.decode(torch.tensor(want), skip_special_tokens=True)
    # sampling branch of the same loop (:65-66): runs, and only emits ids the tokenizer can decode
    with torch.no_grad():
        torch.manual_seed(0)
        text_s = I.test_inference(model, processor, "cuda", "caption en", img_path, 5, 0.8, 0.9, True)
    assert text_s.startswith("caption en")
    # main(): device selection, load_hf_model (fp16), processor construction, generation, print
    I.main(model_path=ckpt, prompt="caption en", image_file_path=img_path, max_tokens_to_generate=4, do_sample=False)
    out = capsys.readouterr().out
    assert "Device in use:  cuda" in out and "caption en" in out


def test_sentencepiece_checkpoint_decodes_on_the_gpu(setup, tmp_path):
    """Checkpoint path end to end with a REAL SentencePiece tokenizer (trained offline on a synthetic corpus): config.json
    + safetensors shards + tokenizer.model -> utils.load_hf_model -> PaliGemmaProcessor -> the reference's own
    test_inference loop on the B200 engine -> text.  Tokens against the CPU oracle."""
    from PIL import Image
    from safetensors.torch import save_file
    from test_checkpoint_cpu import write_sentencepiece_tokenizer
    import inference as I
    import utils as U
    d = str(tmp_path)
    write_sentencepiece_tokenizer(d)
    from transformers import AutoTokenizer
    n_tok = len(AutoTokenizer.from_pretrained(d))
    cfg = json.loads(json.dumps(synth.TINY))
    cfg["image_token_index"] = n_tok                               # where the processor will put <image>
    cfg["vocab_size"] = cfg["text_config"]["vocab_size"] = n_tok + 1 + 1024 + 128 + (n_tok + 1) % 2
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    sd = {k: v.contiguous() for k, v in synth.synth_state_dict(cfg, tie=False).items()}
    save_file(sd, os.path.join(d, "model.safetensors"))
    img_path = os.path.join(d, "chart.png")
    Image.fromarray(np.random.default_rng(5).integers(0, 256, size=(70, 90, 3), dtype=np.uint8), "RGB").save(img_path)
    model, tokenizer = U.load_hf_model(d, "cuda", dtype=torch.float32)
    model = model.to("cuda").eval()
    v = model.config.vision_config
    processor = PP.PaliGemmaProcessor(tokenizer, v.num_image_tokens, v.image_size)
    inputs = processor(text=["describe the chart"], images=[Image.open(img_path)])
    assert int((inputs["input_ids"] == n_tok).sum()) == v.num_image_tokens
    n = 6
    with torch.no_grad():
        text = I.test_inference(model, processor, "cuda", "describe the chart", img_path, n, 0.8, 0.9, False)
    sd["language_model.lm_head.weight"] = sd["language_model.model.embed_tokens.weight"]
    want = O.generate_cached(sd, cfg, inputs["input_ids"], inputs["pixel_values"], n, patched=False)[0].tolist()
    if tokenizer.eos_token_id in want:
        want = want[:want.index(tokenizer.eos_token_id) + 1]
    assert text == "describe the chart" + tokenizer.decode(torch.tensor(want), skip_special_tokens=True)
