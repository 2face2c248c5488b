"""CPU: the oracle reproduces the golden vectors that oracle/make_golden.py recorded from the REAL
reference (MMGPTLlamaForCausalLM + CLIPVisionTower + Conv/MLPProjector run on CPU fp32)."""
import pytest
import torch

from helpers import load_golden, weights_sha
from oracle import merlin_oracle as O

CASES = ["tiny_conv", "tiny_mlp", "tiny_pad"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g, cfg = load_golden(name)
    sd = O.init_state_dict(cfg, seed=g["seed"])
    assert weights_sha(sd) == g["weights_sha256"], "seeded weights drifted from the fixture"
    out = O.forward(sd, cfg, g["input_ids"], g["attention_mask"], g["labels"], g["images"])
    assert torch.equal(out["inputs_embeds"], g["ref_inputs_embeds"])          # index/gather: bit-exact
    for a, b in zip(out["image_features"], g["ref_image_features"]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
    valid = g["attention_mask"].bool()
    torch.testing.assert_close(out["logits"][valid], g["ref_logits"][valid], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(out["loss"], g["ref_loss"], rtol=1e-5, atol=1e-5)


def test_oracle_greedy_matches_reference_loop():
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    n_new = g["ref_generate"].shape[1] - g["input_ids"].shape[1]
    ids = O.greedy_generate(sd, cfg, g["input_ids"][:1], g["images"][:1], n_new, eos_token_id=None)
    assert torch.equal(ids, g["ref_generate"])


def test_splice_is_a_scatter_on_im_patch_positions():
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    feats = O.encode_images(sd, cfg, g["images"])
    emb = O.prepare_inputs_embeds(sd, cfg, g["input_ids"], feats)
    ref = torch.nn.functional.embedding(g["input_ids"], sd["model.embed_tokens.weight"])
    for b in range(ref.shape[0]):
        ref[b][g["input_ids"][b] == cfg.im_patch_token] = feats[b].reshape(-1, ref.shape[-1])
    assert torch.equal(emb, ref)


def test_splice_errors_follow_reference():
    cfg = O.OracleConfig.tiny()
    sd = O.init_state_dict(cfg, seed=0)
    batch = O.synthetic_batch(cfg, 1, 32, [1])
    feats = O.encode_images(sd, cfg, batch["images"])
    bad = batch["input_ids"].clone()
    bad[0, 1 + cfg.num_patches + 1] = 5  # overwrite <im_end>
    with pytest.raises(ValueError):
        O.prepare_inputs_embeds(sd, cfg, bad, feats)
    cfg2 = O.OracleConfig.tiny(use_im_start_end=False)
    with pytest.raises(NotImplementedError):
        O.prepare_inputs_embeds(sd, cfg2, batch["input_ids"], feats)


def test_text_only_sample_keeps_token_embeddings():
    cfg = O.OracleConfig.tiny()
    sd = O.init_state_dict(cfg, seed=0)
    ids = torch.randint(3, 200, (1, 16))
    feats = O.encode_images(sd, cfg, [torch.zeros(1, 3, cfg.image_size, cfg.image_size)])
    emb = O.prepare_inputs_embeds(sd, cfg, ids, feats)
    assert torch.equal(emb, torch.nn.functional.embedding(ids, sd["model.embed_tokens.weight"]))


def test_causality_and_padding_properties():
    """Size-independent properties of the path: logits of a prefix do not depend on the suffix, and
    right padding does not change valid-row logits."""
    cfg = O.OracleConfig.tiny()
    sd = O.init_state_dict(cfg, seed=0)
    b = O.synthetic_batch(cfg, 1, 40, [1])
    full = O.forward(sd, cfg, b["input_ids"], None, None, b["images"])["logits"]
    pre = O.forward(sd, cfg, b["input_ids"][:, :24], None, None, b["images"])["logits"]
    torch.testing.assert_close(full[:, :24], pre, rtol=1e-4, atol=1e-5)
    padded = torch.cat([b["input_ids"], torch.zeros(1, 8, dtype=torch.int64)], 1)
    mask = torch.cat([torch.ones(1, 40, dtype=torch.bool), torch.zeros(1, 8, dtype=torch.bool)], 1)
    pad = O.forward(sd, cfg, padded, mask, None, b["images"])["logits"]
    torch.testing.assert_close(pad[:, :40], full, rtol=1e-4, atol=1e-5)


def test_llrd_param_groups_follow_reference_rules():
    """llrd_utils.py:4-79: no weight decay on .bias / 1-D, ViT layer i scaled 0.9**(22-i), other
    vision_model params x0.1."""
    names = [("model.layers.0.self_attn.q_proj.weight", (8, 8)), ("model.norm.weight", (8,)),
             ("model.vision_tower.vision_tower.vision_model.encoder.layers.22.mlp.fc1.weight", (8, 8)),
             ("model.vision_tower.vision_tower.vision_model.encoder.layers.20.mlp.fc1.bias", (8,)),
             ("model.vision_tower.vision_tower.vision_model.pre_layrnorm.weight", (8,)),
             ("model.projector.projector.weight", (8, 8, 3, 3))]
    groups = O.param_groups(names, lr=1e-3, wd=0.05)
    by_name = {n: g for g in groups for n in g["names"]}
    assert by_name[names[0][0]]["weight_decay"] == 0.05 and by_name[names[0][0]]["lr"] == 1e-3
    assert by_name[names[1][0]]["weight_decay"] == 0.0
    assert by_name[names[2][0]]["lr"] == pytest.approx(1e-3)           # layer 22 -> 0.9**0 == 1 (unscaled group)
    assert by_name[names[3][0]]["lr"] == pytest.approx(1e-3 * 0.9 ** 2) and by_name[names[3][0]]["weight_decay"] == 0.0
    assert by_name[names[4][0]]["lr"] == pytest.approx(1e-4)
    assert by_name[names[5][0]]["weight_decay"] == 0.05


@pytest.mark.parametrize("name", CASES)
def test_oracle_autograd_matches_reference_golden(name):
    """The reference's own loss.backward() (recorded by make_golden.py) pins the oracle's autograd, which
    is what the CUDA backward schedule is checked against on the GPU."""
    g, cfg = load_golden(name)
    sd = O.init_state_dict(cfg, seed=0)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.forward(leaves, cfg, g["input_ids"], g["attention_mask"], g["labels"], g["images"])["loss"].backward()
    assert len(g["ref_grad_norms"]) >= 70
    for k, norm in g["ref_grad_norms"].items():
        assert leaves[k].grad is not None
        assert float(leaves[k].grad.norm()) == pytest.approx(norm, rel=1e-4, abs=1e-7), k
    for k, probe in g["ref_grad_probe"].items():
        torch.testing.assert_close(leaves[k].grad.flatten()[:64], probe, rtol=1e-4, atol=1e-7)


def test_llrd_param_groups_match_the_reference_function_outputs():
    """tests/golden/llrd_groups.json holds what the reference's OWN get_param_groups (+ vit_lr_scale_func /
    llm_lr_scale_func, mmgpt/utils/llrd_utils.py:4-79, called as MMGPTTrainer.create_optimizer does) returned for the
    414 parameters of an assembled reference model with the real 24-layer CLIP depth (oracle/make_golden.py).  The
    oracle's grouping, the product's get_param_groups mirror and the per-parameter (lr multiplier, weight decay)
    tables the fused AdamW kernel consumes must all reproduce it; when oracle/_ref is importable the reference
    function is also called live."""
    import json
    import os
    from helpers import GOLDEN_DIR
    from merlin_b200.mmgpt.utils import llrd_utils as P
    gold = json.load(open(os.path.join(GOLDEN_DIR, "llrd_groups.json")))
    named = [(n, tuple(s)) for n, s in gold["names"]]
    fns = {"llrd": P.vit_lr_scale_func, "llm_llrd": P.llm_lr_scale_func, "none": None}
    for tag, want in gold["cases"].items():
        got = P.get_param_groups(named, None, fns[tag], 1e-3, 0.05)
        assert [(g["lr"], g["weight_decay"], g["params"]) for g in got] == \
            [(g["lr"], g["weight_decay"], g["params"]) for g in want], tag
        by_name = {n: g for g in want for n in g["params"]}
        for n, shape in named:                       # what the AdamW kernel's per-chunk tables are built from
            mult, wd = P.param_hyper(n, shape, fns[tag], 0.05)
            assert wd == by_name[n]["weight_decay"] and 1e-3 * mult == pytest.approx(by_name[n]["lr"], rel=1e-12), (tag, n)
        if tag != "llm_llrd":
            mine = O.param_groups(named, 1e-3, 0.05, llrd=(tag == "llrd"))
            assert [(g["lr"], g["weight_decay"], g["names"]) for g in mine] == \
                [(g["lr"], g["weight_decay"], g["params"]) for g in want], tag
    from oracle import refmodel
    if refmodel.available():
        U = refmodel.import_reference().llrd_utils
        params = [(n, torch.nn.Parameter(torch.empty(s))) for n, s in named]
        live = U.get_param_groups(type("M", (), {"named_parameters": staticmethod(lambda: params)}), None,
                                  U.vit_lr_scale_func, 1e-3, 0.05)
        ids = {id(p): n for n, p in params}
        assert [(g["lr"], g["weight_decay"], [ids[id(p)] for p in g["params"]]) for g in live] == \
            [(g["lr"], g["weight_decay"], g["params"]) for g in gold["cases"]["llrd"]]
