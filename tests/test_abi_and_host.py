"""CPU: the C-ABI library loads and exports every symbol include/merlin_b200.h declares; host-side
logic of the mmgpt mirror (state-dict contract, packing, grammar validation, error behaviour)."""
import ctypes
import os

import pytest
import torch

from helpers import build_product_model, load_golden
from merlin_b200 import lib as mlib
from merlin_b200 import ops
from oracle import merlin_oracle as O


def test_library_exports_every_declared_symbol():
    L = mlib.load()
    names = mlib.declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/merlin_b200.h but not exported"


def test_prototypes_cover_the_header():
    covered = set(mlib._PROTOTYPES) | set(mlib._SIZE_FNS) | {"merlin_last_error"}
    assert set(mlib.declared_symbols()) <= covered


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("MERLIN_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(mlib, "_lib", None)
    with pytest.raises(mlib.MerlinError):
        mlib.load()


def test_invalid_arguments_return_error_codes_without_a_gpu():
    L = mlib.load()
    rc = L.merlin_gemm_bf16(None, None, 0, None, 0, None, 0, None, 0, None, 0, 0, 0, 0, 2)
    assert rc == -1 and b"gemm" in L.merlin_last_error()
    rc = L.merlin_flash_attn_fwd(None, 1, 0, 1, 0, None, None, 1, 1, 1, 96, 1, 0.0)
    assert rc == -1 and b"head_dim" in L.merlin_last_error()
    assert L.merlin_decode_workspace_bytes(4096, 11008, 32, 16) > 2 * (3 * 4096 + 11008)


def test_ops_reject_cpu_tensors():
    a = torch.zeros(4, 8, dtype=torch.bfloat16)
    with pytest.raises(TypeError):
        ops.gemm(a, a)


def test_swiglu_interleave_layout():
    g = torch.arange(256 * 4, dtype=torch.float32).view(256, 4)
    u = -g
    p = ops.swiglu_interleave(g, u)
    assert p.shape == (512, 4)
    assert torch.equal(p[:128], g[:128]) and torch.equal(p[128:256], u[:128])
    assert torch.equal(p[256:384], g[128:]) and torch.equal(p[384:], u[128:])


@pytest.mark.parametrize("name", ["tiny_conv", "tiny_mlp"])
def test_state_dict_contract_round_trip(name):
    """Reference key names / shapes in, packed storage inside, identical tensors out (Appendix A)."""
    g, cfg = load_golden(name)
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd, device=None)
    back = model.state_dict()
    assert set(back) == set(sd)
    for k in sd:
        assert torch.equal(back[k], sd[k]), k
    # packed layouts the kernels rely on
    D, FF = cfg.hidden_size, cfg.intermediate_size
    m = model.model
    assert torch.equal(m.qkv_w[0][D:2 * D], sd["model.layers.0.self_attn.k_proj.weight"])
    assert torch.equal(m.gate_up_w[1][128:256], sd["model.layers.1.mlp.up_proj.weight"][:128])
    assert torch.equal(m.gate_up_w[1][256:384], sd["model.layers.1.mlp.gate_proj.weight"][128:256])
    vt = m.vision_tower
    assert vt.patch_weight.shape[1] % 8 == 0 and bool((vt.patch_weight[:, vt.k_patch:] == 0).all())
    assert vt.num_patches == cfg.num_patches


def test_strict_loading_reports_unknown_keys():
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    sd["model.layers.0.bogus.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        build_product_model(cfg, sd, device=None)


def test_grammar_validation_matches_reference_errors():
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0), device=None)
    ids = g["input_ids"].clone()
    P = cfg.num_patches
    model._validate_image_grammar(ids, g["images"], P)          # valid batch passes
    bad = ids.clone()
    bad[0, 1 + P + 1] = 7                                       # <im_end> overwritten
    with pytest.raises(ValueError):
        model._validate_image_grammar(bad, g["images"], P)
    bad = ids.clone()
    bad[0, -2] = cfg.im_start_token                             # unmatched <im_start>
    with pytest.raises(ValueError):
        model._validate_image_grammar(bad, g["images"], P)
    bad = ids.clone()
    bad[0, -3] = cfg.im_patch_token                             # stray <im_patch> outside any image block
    with pytest.raises(ValueError, match="may only appear between"):
        model._validate_image_grammar(bad, g["images"], P)
    model.use_im_start_end = False
    with pytest.raises(NotImplementedError):
        model._validate_image_grammar(ids, g["images"], P)


def test_prepare_inputs_for_generation_contract():
    from merlin_b200.mmgpt.model.mmgpt.llama_mmgpt import MerlinKVCache
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0), device=None)
    ids = g["input_ids"][:1]
    first = model.prepare_inputs_for_generation(ids, None, None, use_cache=True, images=g["images"][:1])
    assert first["input_ids"].shape == ids.shape and first["past_key_values"] is None
    empty = MerlinKVCache(cfg.num_layers, 1, 8, cfg.hidden_size, "cpu")
    assert not empty                                            # empty cache == no past
    assert model.prepare_inputs_for_generation(ids, empty)["input_ids"].shape == ids.shape
    empty.seq_len = ids.shape[1]
    step = model.prepare_inputs_for_generation(ids, empty, images=g["images"][:1])
    assert step["input_ids"].shape == (1, 1) and step["images"] is not None


def test_forward_without_cuda_fails_loudly():
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0), device=None)
    with pytest.raises(RuntimeError):
        model(input_ids=g["input_ids"], images=g["images"])


def test_build_model_tokenizer_from_checkpoint(tmp_path):
    """builder.py contract: config + checkpoint + tokenizer in, assembled model out; tower / projector
    slices are picked out of the same pytorch_model.bin by their key prefixes; freezing policy applied."""
    import json
    from types import SimpleNamespace
    from helpers import FakeTokenizer
    from merlin_b200.mmgpt.model.builder import build_model_tokenizer
    from merlin_b200.mmgpt.utils.arguments import DataArguments, ModelArguments
    from merlin_b200.mmgpt.utils.constants import (DEFAULT_IM_END_TOKEN, DEFAULT_IM_PATCH_TOKEN,
                                                   DEFAULT_IM_START_TOKEN)
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    torch.save(sd, tmp_path / "pytorch_model.bin")
    json.dump(dict(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                   num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                   num_key_value_heads=cfg.num_heads, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
                   max_position_embeddings=4096, tie_word_embeddings=False), open(tmp_path / "config.json", "w"))
    vcfg = SimpleNamespace(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_intermediate,
                           num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                           image_size=cfg.image_size, patch_size=cfg.patch_size, hidden_act="quick_gelu",
                           layer_norm_eps=cfg.vit_ln_eps, num_channels=3)
    margs = ModelArguments(model_name_or_path=str(tmp_path), vision_tower="clip-synthetic", projector="conv",
                           conv_stride=2, freeze_lm_model=True, freeze_vision_tower=False, freeze_projector=False)
    dargs = DataArguments()
    targs = SimpleNamespace(model_max_length=2048, bf16=True, fp16=False, lora_enable=False, device="cpu",
                            cache_dir=None, gradient_checkpointing=False)
    tok = FakeTokenizer(256, extra=(DEFAULT_IM_PATCH_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN))
    model, tok2, dargs2 = build_model_tokenizer(margs, dargs, targs, tokenizer=tok, vision_config=vcfg)
    assert tok2.pad_token == tok2.unk_token and len(tok2) == cfg.vocab_size
    assert (model.im_patch_token, model.im_start_token, model.im_end_token) == (256, 257, 258)
    assert dargs2.image_token_len == cfg.num_patches and dargs2.use_im_start_end
    assert model.dtype == torch.bfloat16
    back = model.state_dict()
    for k, v in sd.items():                     # weights (LLM + tower + projector) came from the checkpoint
        assert torch.equal(back[k].float(), v), k
    m = model.model
    assert not m.qkv_w[0].requires_grad and m.embed_tokens.requires_grad and not model.lm_head_w.requires_grad
    assert m.vision_tower.qkv_w[0].requires_grad and not m.vision_tower.qkv_w[cfg.vit_layers - 1].requires_grad
    assert m.projector.weight.requires_grad
    # a fresh tokenizer (image tokens not yet present) grows the vocabulary and mean-initialises the new rows
    tok3 = FakeTokenizer(256)
    model3, _, _ = build_model_tokenizer(margs, DataArguments(), targs, tokenizer=tok3, vision_config=vcfg)
    e = model3.model.embed_tokens.float()
    assert model3.num_new_tokens == 2 and e.shape[0] == 259
    assert torch.allclose(e[-1], e[:-2].mean(0), atol=1e-2)
    with pytest.raises(NotImplementedError):
        build_model_tokenizer(SimpleNamespace(**{**margs.__dict__, "model_name_or_path": "x/baichuan-7b"}), dargs, targs)


def test_clip_tower_loads_pretrained_clip_then_overlays_the_mmgpt_slice(tmp_path):
    """clip_encoder.py:21-24: CLIPVisionModel.from_pretrained(args.vision_tower) first, then the
    `model.vision_tower.*` keys of model_name_or_path on top; a tower with neither source warns (its parameters are
    all zero)."""
    import json
    from types import SimpleNamespace
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from merlin_b200.mmgpt.model.vision_encoder.clip_encoder import CLIPVisionTower
    vcfg = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
                            image_size=56, patch_size=14, hidden_act="quick_gelu")
    torch.manual_seed(0)
    hf = CLIPVisionModel(vcfg)
    clip_dir = tmp_path / "clip-vit-tiny"
    hf.save_pretrained(str(clip_dir))
    assert ((clip_dir / "model.safetensors").exists() or (clip_dir / "pytorch_model.bin").exists()) \
        and (clip_dir / "config.json").exists()
    args = SimpleNamespace(vision_tower=str(clip_dir), vision_select_layer=-2, vision_select_feature="patch",
                           freeze_vision_tower=True, conv_stride=2, model_name_or_path=None)
    tower = CLIPVisionTower(args)
    assert tower.weight_sources == [f"clip:{clip_dir}"] and tower.config.hidden_size == 128
    want = {"vision_tower." + k: v for k, v in hf.state_dict().items() if "position_ids" not in k}
    got = tower.state_dict()
    assert set(got) == set(want) and all(torch.equal(got[k], want[k]) for k in want)
    assert float(tower.qkv_w[0].abs().sum()) > 0 and float(tower.ln1_w[0].abs().sum()) > 0
    # overlay: an mmgpt checkpoint whose tower slice differs wins over the CLIP weights
    ck = tmp_path / "mmgpt-ckpt"
    ck.mkdir()
    over = {"model.vision_tower." + k: v + 1.0 for k, v in want.items()}
    over["model.layers.0.mlp.down_proj.weight"] = torch.zeros(2, 2)
    torch.save(over, ck / "pytorch_model.bin")
    args.model_name_or_path = str(ck)
    tower2 = CLIPVisionTower(args)
    assert tower2.weight_sources == [f"clip:{clip_dir}", f"mmgpt:{ck}"]
    assert all(torch.equal(tower2.state_dict()[k], want[k] + 1.0) for k in want)
    # neither source: loud warning, zero parameters
    args2 = SimpleNamespace(vision_tower="clip-does-not-exist", vision_select_layer=-2, vision_select_feature="patch",
                            freeze_vision_tower=True, conv_stride=2, model_name_or_path=None)
    with pytest.warns(RuntimeWarning, match="no pretrained weights"):
        empty = CLIPVisionTower(args2, config=SimpleNamespace(
            hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=56,
            patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5, num_channels=3))
    assert empty.weight_sources == [] and float(empty.qkv_w[0].abs().sum()) == 0
