"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
ctypes table agrees with the header, and the drop-in host classes keep the reference's surface."""
import os
import re

import pytest
import torch

from pg_b200 import _cabi, synth
import modeling_gemma as MG
import modeling_siglip as MS
import processing_paligemma as PP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(pg_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = _cabi.lib()
    decl = header_functions()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), name
    assert lib.pg_abi_version() == 1


def test_ctypes_table_matches_header():
    decl = header_functions()
    assert set(_cabi.SIGNATURES) == set(decl)
    for name, args in _cabi.SIGNATURES.items():
        assert len(args) == decl[name], (name, len(args), decl[name])


def test_config_surface_matches_reference_derivations():
    cfg = MG.PaliGemmaConfig(**synth.PALIGEMMA_3B_224, some_unknown_key=1)
    assert cfg.vocab_size == 257216 and cfg.text_config.num_image_tokens == 256
    assert cfg.vision_config.projection_dim == 2048 and cfg.text_config.pad_token_id == 0
    assert cfg.text_config.head_dim == 256 and cfg.text_config.max_position_embeddings == 8192
    assert cfg.vision_config.layer_norm_eps == 1e-6 and cfg.is_encoder_decoder is False
    v = MS.SiglipVisionConfig()
    assert (v.hidden_size, v.patch_size, v.num_hidden_layers) == (768, 16, 12)


def test_state_dict_keys_and_tying():
    cfg = synth.TINY
    m = MG.PaliGemmaForConditionalGeneration(MG.PaliGemmaConfig(**cfg), init_weights=False)
    spec = {k: tuple(s) for k, s, _ in synth.state_dict_spec(cfg)}
    sd = m.state_dict()
    assert set(sd) == set(spec) | {"language_model.lm_head.weight"}
    for k, s in spec.items():
        assert tuple(sd[k].shape) == s, k
    assert m.language_model.lm_head.weight is not m.language_model.model.embed_tokens.weight
    m.tie_weights()
    assert m.language_model.lm_head.weight is m.language_model.model.embed_tokens.weight
    res = m.load_state_dict(synth.synth_state_dict(cfg, tie=False), strict=False)
    assert res.unexpected_keys == []
    # monkey-patch targets of ablation_study_fixed.py:335-342 exist
    assert callable(m._merge_input_ids_with_image_features)
    assert all(hasattr(l.self_attn.rotary_emb, "forward") for l in m.language_model.model.layers)
    assert m.config.vision_config.num_image_tokens == 16 and m.config.vision_config.image_size == 56


def test_no_cpu_fallback():
    m = MG.PaliGemmaForConditionalGeneration(MG.PaliGemmaConfig(**synth.TINY), init_weights=False)
    ids = torch.zeros((1, 3), dtype=torch.int64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(input_ids=ids, attention_mask=torch.ones_like(ids))
    with pytest.raises(ValueError):
        m(input_ids=ids, attention_mask=None)
    with pytest.raises(AssertionError):
        m(input_ids=ids, attention_mask=torch.zeros_like(ids))


def test_kvcache_list_semantics_before_binding():
    kv = MG.KVCache()
    assert kv.num_items() == 0 and len(kv.key_cache) == 0
    k = torch.randn(1, 1, 3, 8)
    kv.update(k, k + 1, 0)
    k2, v2 = kv.update(k[:, :, :1], k[:, :, :1], 0)
    assert kv.num_items() == 4 and tuple(k2.shape) == (1, 1, 4, 8) and tuple(kv.value_cache[0].shape) == (1, 1, 4, 8)


def test_processor_contract(golden_dir):
    from PIL import Image
    import numpy as np
    tok = synth.StubTokenizer()
    proc = PP.PaliGemmaProcessor(tok, 256, 224)
    assert tok.add_bos_token is False and tok.add_eos_token is False and len(tok.added_tokens) == 1152
    assert proc.image_token_id == 257152
    yy, xx = np.mgrid[0:480, 0:640]
    img = Image.fromarray(np.stack([(xx * 255 // 639), (yy * 255 // 479), ((xx + yy) % 256)], -1).astype(np.uint8))
    out = proc(text=["caption en"], images=[img])
    assert tuple(out["pixel_values"].shape) == (1, 3, 224, 224) and out["pixel_values"].dtype == torch.float32
    assert float(out["pixel_values"].min()) >= -1.0 and float(out["pixel_values"].max()) <= 1.0
    ids = out["input_ids"][0].tolist()
    assert ids[:256] == [257152] * 256 and ids[256] == synth.BOS_ID and ids[-1] == synth.NEWLINE_ID
    assert out["attention_mask"].tolist() == [[1] * len(ids)]
    g = os.path.join(golden_dir, "processor.npz")
    if os.path.exists(g):  # produced by the reference's own processor on the same image (make_golden.py)
        ref = np.load(g)
        np.testing.assert_array_equal(out["pixel_values"].numpy(), ref["pixel_values"])
        assert ids == ref["input_ids"][0].tolist()
    with pytest.raises(AssertionError):
        proc(text=["a", "b"], images=[img])
