"""GPU parity proper: the product path (mmgpt mirror -> C ABI -> sm_100a kernels) against the CPU
oracle and the golden vectors recorded from the real reference, on identical seeded inputs.

Tolerances (bf16 weights/activations vs the fp32 oracle): logits max|Δ| <= 1e-2 (north_star),
loss |Δ| <= 2e-3, image features / hidden states max|Δ| <= 2^-6·max|ref|; splice bit-exact;
greedy token ids identical."""
import pytest
import torch

from helpers import build_product_model, load_golden
from oracle import merlin_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def to_dev(images):
    return [im.to(DEV) for im in images]


@pytest.mark.parametrize("name", ["tiny_conv", "tiny_mlp", "tiny_pad"])
def test_forward_matches_reference_golden(name):
    g, cfg = load_golden(name)
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)
    out = model(input_ids=g["input_ids"].to(DEV), attention_mask=g["attention_mask"].to(DEV),
                labels=g["labels"].to(DEV), images=to_dev(g["images"]))
    valid = g["attention_mask"].bool()
    logits = out.logits.float().cpu()
    assert logits.shape == g["ref_logits"].shape
    err = (logits[valid] - g["ref_logits"][valid]).abs().max().item()
    assert err <= 1e-2, f"logits max|Δ| = {err}"
    assert abs(out.loss.item() - g["ref_loss"].item()) <= 2e-3
    # the oracle itself agrees with the golden on the same inputs (ties the three together)
    mine = O.forward(sd, cfg, g["input_ids"], g["attention_mask"], g["labels"], g["images"])
    assert (mine["logits"][valid] - g["ref_logits"][valid]).abs().max().item() < 1e-4


def test_vision_tower_and_projector_api_match_oracle():
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)
    tower, proj = model.get_model().vision_tower, model.get_model().projector
    feats = tower(to_dev(g["images"]))                       # tuple of [n_i, P, C]  (clip_encoder.py:74-82)
    ref = O.vision_tower_forward(sd, cfg, g["images"])
    assert [tuple(f.shape) for f in feats] == [tuple(r.shape) for r in ref]
    for f, r in zip(feats, ref):
        assert (f.float().cpu() - r).abs().max().item() <= 2.0 ** -6 * r.abs().max().item()
    pf = proj(feats)                                          # list of [n_i, P', D]  (conv_projector.py:23-39)
    for f, r in zip(pf, g["ref_image_features"]):
        assert tuple(f.shape) == tuple(r.shape)
        assert (f.float().cpu() - r).abs().max().item() <= 2.0 ** -6 * r.abs().max().item()
    assert tower.num_patches == cfg.num_patches and tower.hidden_size == cfg.vit_hidden
    assert tower.dummy_feature.shape == (1, cfg.vit_hidden)


def test_splice_bit_exact_against_oracle():
    """inputs_embeds: gather rows are bit-exact copies of bf16 table rows / projected features."""
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)
    ids = g["input_ids"].to(DEV)
    _, _, _, embeds, _ = model.prepare_inputs_labels_for_multimodal(ids, None, None, None, to_dev(g["images"]))
    feats = model.encode_images(to_dev(g["images"]))
    ref = model.get_model().embed_tokens[ids].clone()
    for b in range(ids.shape[0]):
        ref[b][ids[b] == cfg.im_patch_token] = feats[b].reshape(-1, ref.shape[-1])
    assert torch.equal(embeds, ref)
    # text rows equal the oracle's embedding rows exactly (weights are bf16-representable)
    text = (g["input_ids"] != cfg.im_patch_token)
    assert torch.equal(embeds.float().cpu()[text], g["ref_inputs_embeds"][text])


def test_hidden_states_per_layer():
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)
    out = model(input_ids=g["input_ids"].to(DEV), images=to_dev(g["images"]), output_hidden_states=True)
    feats = O.encode_images(sd, cfg, g["images"])
    emb = O.prepare_inputs_embeds(sd, cfg, g["input_ids"], feats)
    final, layers = O.decoder_forward(sd, cfg, emb, None, return_layers=True)
    refs = [emb] + layers[:-1] + [final]
    assert len(out.hidden_states) == len(refs)
    for h, r in zip(out.hidden_states, refs):
        assert (h.float().cpu() - r).abs().max().item() <= 2.0 ** -6 * r.abs().max().item()


def test_error_behaviour_on_device():
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0))
    bad = g["input_ids"].clone()
    bad[0, 1 + cfg.num_patches + 1] = 7
    with pytest.raises(ValueError):
        model(input_ids=bad.to(DEV), images=to_dev(g["images"]))
    with pytest.raises(ValueError):
        model(input_ids=g["input_ids"].to(DEV), images=to_dev(g["images"][:1]))
    with pytest.raises(AssertionError):
        model(input_ids=g["input_ids"].to(DEV), images=to_dev(g["images"]), output_attentions=True)


def test_greedy_generate_matches_reference_tokens():
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0))
    ids = g["input_ids"][:1]
    n_new = g["ref_generate"].shape[1] - ids.shape[1]
    out = model.generate(ids.to(DEV), images=to_dev(g["images"][:1]), max_new_tokens=n_new, eos_token_id=None)
    assert torch.equal(out.cpu(), g["ref_generate"])
    # the step-by-step forward(past_key_values=...) path gives the same tokens as the graph-replayed loop
    out2 = model.generate(ids.to(DEV), images=to_dev(g["images"][:1]), max_new_tokens=n_new, eos_token_id=None,
                          stopping_criteria=[lambda i, s: False])
    assert torch.equal(out2, out)


def test_decode_step_logits_match_full_forward():
    """Decode property: logits of token t from the cached single-token path == row t of a full forward."""
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0))
    ids = g["input_ids"][:1].to(DEV)
    imgs = to_dev(g["images"][:1])
    full = model(input_ids=ids, images=imgs).logits
    pre = model(input_ids=ids[:, :-1], images=imgs, use_cache=True)
    step = model(input_ids=ids[:, -1:], images=imgs, past_key_values=pre.past_key_values,
                 attention_mask=torch.ones_like(ids), use_cache=True)
    assert (step.logits[0, 0].float() - full[0, -1].float()).abs().max().item() <= 1e-2


def test_full_width_single_layer_properties():
    """At Vicuna-7B width (D=4096, FF=11008, S=2048) the oracle is too slow for every test run, so use
    size-independent properties: causality (prefix logits unchanged by the suffix) and determinism."""
    from merlin_b200.mmgpt.model.mmgpt.llama_mmgpt import MMGPTLlamaForCausalLM, vicuna_7b_config
    torch.manual_seed(0)
    model = MMGPTLlamaForCausalLM(vicuna_7b_config(num_hidden_layers=1, vocab_size=2048)).to(device=DEV, dtype=torch.bfloat16)
    model.requires_grad_(False)
    ids = torch.randint(3, 2000, (1, 2048), device=DEV)
    a = model(input_ids=ids).logits
    b = model(input_ids=ids).logits
    assert torch.equal(a, b)
    c = model(input_ids=ids[:, :1024]).logits
    assert (a[:, :1024].float() - c.float()).abs().max().item() <= 2.0 ** -6 * a.float().abs().max().item()
    # against torch fp32 math of the same layer on the GPU
    sd = {k: v.float() for k, v in model.state_dict().items()}
    ocfg = O.OracleConfig(vocab_size=2048, num_layers=1)
    emb = torch.nn.functional.embedding(ids[:, :256], sd["model.embed_tokens.weight"])
    ref = torch.nn.functional.linear(O.decoder_forward(sd, ocfg, emb, None, positions=torch.arange(256, device=DEV)),
                                     sd["lm_head.weight"])
    got = model(input_ids=ids[:, :256]).logits.float()
    # |logits| reach ~5 here, where one bf16 ulp of the OUTPUT is already 3e-2: use the per-op bound
    assert (got - ref).abs().max().item() <= 2.0 ** -6 * ref.abs().max().item()


def test_operator_seam_in_hf_llama_matches_eager():
    """replace_llama_attn_with_flash_attn(): HF LlamaModel dispatching to merlin_flash_attn_fwd through
    the attention registry equals HF eager attention (valid rows), incl. a right-padded sample."""
    from transformers import LlamaConfig, LlamaModel
    from merlin_b200.mmgpt.utils.llama_flash_attn_monkey_patch import replace_llama_attn_with_flash_attn
    name = replace_llama_attn_with_flash_attn()
    cfg = LlamaConfig(vocab_size=300, hidden_size=256, intermediate_size=384, num_hidden_layers=2,
                      num_attention_heads=2, num_key_value_heads=2, attn_implementation="eager")
    torch.manual_seed(0)
    m = LlamaModel(cfg).to(DEV, torch.bfloat16).eval()
    ids = torch.randint(0, 300, (2, 200), device=DEV)
    mask = torch.ones(2, 200, dtype=torch.long, device=DEV)
    mask[1, 150:] = 0
    with torch.no_grad():
        ref = m(input_ids=ids, attention_mask=mask).last_hidden_state.float()
        m.config._attn_implementation = name
        got = m(input_ids=ids, attention_mask=mask).last_hidden_state.float()
    valid = mask.bool()
    assert (got[valid] - ref[valid]).abs().max().item() <= 2.0 ** -5 * ref[valid].abs().max().item()


def test_full_width_c1_sample_matches_oracle():
    """BASELINE config[0] shape at full WIDTH: one 336 px image through the real CLIP-L/14-336 dims (24 layers,
    23 used), ConvProjector(1024 -> 4096, stride 2) -> 144 visual tokens, 32-token caption, Vicuna-7B width
    (D=4096, FF=11008, H=32, V=32003) with the decoder depth cut to 2 layers so the fp32 CPU oracle runs in
    seconds.  Every kernel runs at its production shape; logits (|ref| up to ~3, where one bf16 ulp of the
    output is 1.6e-2) must agree within 2^-6·max|ref|, the loss within 2e-3, image features within 2^-6."""
    cfg = O.OracleConfig(num_layers=2)
    sd = O.init_state_dict(cfg, seed=0)
    batch = O.synthetic_batch(cfg, 1, 181, [1], seed=1234)
    ref = O.forward(sd, cfg, batch["input_ids"], batch["attention_mask"], batch["labels"], batch["images"])
    model = build_product_model(cfg, sd)
    out = model(input_ids=batch["input_ids"].to(DEV), attention_mask=batch["attention_mask"].to(DEV),
                labels=batch["labels"].to(DEV), images=to_dev(batch["images"]))
    logits = out.logits.float().cpu()
    scale = ref["logits"].abs().max().item()
    err = (logits - ref["logits"]).abs().max().item()
    assert err <= 2.0 ** -6 * scale, f"logits max|Δ|={err:.4g} vs max|ref|={scale:.4g}"
    assert abs(out.loss.item() - ref["loss"].item()) <= 2e-3
    feats = model.encode_images(to_dev(batch["images"]))[0].float().cpu()
    rf = ref["image_features"][0]
    assert (feats - rf).abs().max().item() <= 2.0 ** -6 * rf.abs().max().item()
    # greedy continuation: same tokens as the oracle for 4 steps
    gen = model.generate(batch["input_ids"].to(DEV), images=to_dev(batch["images"]), max_new_tokens=4, eos_token_id=None)
    want = O.greedy_generate(sd, cfg, batch["input_ids"], batch["images"], 4, eos_token_id=None)
    assert torch.equal(gen.cpu(), want)


def test_beam_search_generate_matches_oracle_driven_beam_search():
    """generate(num_beams=3) over the kernels == the same beam bookkeeping (merlin_b200/beam.py, pinned against HF
    on CPU) driven by the fp32 oracle's logits.  bf16 noise may reorder near-tied beams, so a different sequence
    is accepted only if the oracle scores it as well as its own winner."""
    from merlin_b200.beam import BeamSearch
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)
    ids = g["input_ids"][:1]
    images = g["images"][:1]
    nb, max_new = 3, 6
    got = model.generate(ids.to(DEV), images=to_dev(images), num_beams=nb, max_new_tokens=max_new,
                         eos_token_id=None).cpu()
    assert got.shape == (1, ids.shape[1] + max_new) and torch.equal(got[:, :ids.shape[1]], ids)

    def oracle_logits(seqs):
        rows = []
        for s in seqs:
            out = O.forward(sd, cfg, s.view(1, -1), None, None, images)
            rows.append(out["logits"][0])
        return rows

    bs = BeamSearch(ids, nb, None, max_new)
    while True:
        rows = oracle_logits(list(bs.running))
        if bs.step(torch.stack([r[-1] for r in rows])) is None:
            break
    want = bs.finalize()

    def score(seq):
        lp = torch.log_softmax(oracle_logits([seq[0]])[0], dim=-1)
        idx = torch.arange(ids.shape[1] - 1, seq.shape[1] - 1)
        return float(lp[idx, seq[0, ids.shape[1]:]].sum()) / max_new

    if not torch.equal(got, want):
        assert abs(score(got) - score(want)) <= 2e-2, (got.tolist(), want.tolist())


def test_sampling_with_top_p_runs_on_the_kernels_and_respects_the_nucleus():
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0))
    ids = g["input_ids"][:1].to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(0)
    a = model.generate(ids, images=to_dev(g["images"][:1]), do_sample=True, temperature=0.2, top_p=0.7,
                       max_new_tokens=5, eos_token_id=None, generator=gen)
    # top_p -> 0 keeps only the most likely token: sampling degenerates to greedy decoding
    b = model.generate(ids, images=to_dev(g["images"][:1]), do_sample=True, temperature=1.0, top_p=1e-6,
                       max_new_tokens=5, eos_token_id=None, generator=gen)
    c = model.generate(ids, images=to_dev(g["images"][:1]), max_new_tokens=5, eos_token_id=None)
    assert a.shape == b.shape == c.shape and torch.equal(b, c)
