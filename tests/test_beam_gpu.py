"""Batched beam decode (merlin_decode_step_batched: all beams of a beam search advance one position in ONE pass over
the weights, shared prompt cache + per-beam suffix caches) against the single-sequence decode kernels
(merlin_decode_step, itself pinned to the oracle in test_model_gpu.py) on the same token histories.

Tolerance: both paths accumulate in fp32 and round to bf16 at the same places, only the summation order differs
(mma.sync K blocks vs per-lane FMA chains; prefix + suffix softmax merge vs one pass), so logits agree to
2^-6 · max|logits| and the argmax of every row with a clear margin is identical."""
import dataclasses

import pytest
import torch

from helpers import build_product_model, load_golden
from merlin_b200 import decode as D
from merlin_b200.mmgpt.model.mmgpt.llama_mmgpt import MerlinKVCache
from oracle import merlin_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(hidden=256, heads=2, inter=384, layers=2, vocab=None):
    g, cfg = load_golden("tiny_conv")
    cfg = dataclasses.replace(cfg, hidden_size=hidden, num_heads=heads, intermediate_size=inter, num_layers=layers,
                              **({} if vocab is None else {"vocab_size": vocab}))
    sd = O.init_state_dict(cfg, seed=3)
    return g, cfg, build_product_model(cfg, sd)


def _prefill(model, ids, images, cache):
    cache.seq_len = 0
    inputs = model.prepare_inputs_for_generation(ids, None, None, use_cache=True, images=images)
    inputs["past_key_values"] = cache
    with torch.no_grad():
        return model(**inputs)


@pytest.mark.parametrize("nb,graph", [(1, False), (3, False), (5, True), (8, True)])
def test_batched_step_matches_single_sequence_decode(nb, graph, monkeypatch):
    monkeypatch.setenv("MERLIN_BEAM_GRAPH", "1" if graph else "0")
    monkeypatch.setenv("MERLIN_DECODE", "kernels")
    g, cfg, model = _model(hidden=512, heads=4, inter=640, layers=3)
    ids = g["input_ids"][:1].to(DEV)
    images = [g["images"][0].to(DEV)]
    S = ids.shape[1]
    prefix = MerlinKVCache(cfg.num_layers, 1, S + 8, cfg.hidden_size, DEV)
    _prefill(model, ids, images, prefix)
    runner = D.BeamDecodeRunner(model, prefix, nb, 64)
    assert (runner.graph is not None) == graph
    runner.begin(S)
    # the single-sequence references: one cache + runner per beam, re-gathered by parent like HF's _reorder_cache
    singles = []
    for _ in range(nb):
        c = MerlinKVCache(cfg.num_layers, 1, S + 72, cfg.hidden_size, DEV)
        _prefill(model, ids, images, c)
        singles.append(c)
    gen = torch.Generator().manual_seed(nb)
    steps = 37                                    # crosses a 32-token suffix block
    for t in range(steps):
        tokens = torch.randint(3, 30000 if cfg.vocab_size > 30000 else cfg.vocab_size - 4, (nb,), generator=gen).tolist()
        parents = list(range(nb)) if t % 3 == 0 else torch.randint(0, nb, (nb,), generator=gen).tolist()
        got = runner.step(tokens, parents).float().cpu()
        if t > 0:
            new = []
            for j in range(nb):
                c = MerlinKVCache(cfg.num_layers, 1, S + 72, cfg.hidden_size, DEV)
                src = singles[parents[j]]
                for a, b in zip(c.kv, src.kv):
                    a.copy_(b)
                c.seq_len = src.seq_len
                new.append(c)
            singles = new
        emb = model.model.embed_tokens[torch.tensor(tokens, device=DEV)]
        for j in range(nb):
            ref = D.DecodeRunner(model, singles[j]).step_logits(emb[j:j + 1])[0].float().cpu()
            scale = ref.abs().max().item()
            err = (got[j] - ref).abs().max().item()
            assert err <= 2.0 ** -6 * scale, f"step {t} beam {j}: max|Δ| = {err} (scale {scale})"
            top2 = ref.topk(2).values
            if float(top2[0] - top2[1]) > 2.0 ** -4 * scale:
                assert int(got[j].argmax()) == int(ref.argmax())


@pytest.mark.parametrize("extra,nb,dims", [(1437, 5, (640, 5, 896, 2)), (2000, 8, (640, 5, 896, 2)), (331, 2, (640, 5, 896, 2)),
                                           (4049, 5, (4096, 32, 11008, 1))])
def test_batched_step_long_prompt_matches_single_sequence_decode(extra, nb, dims, monkeypatch):
    """Prompt lengths that give every prefix CTA several 16-key groups per warp and a ragged last group (the tiny
    prompts above keep one group per CTA); an odd head count; and one Vicuna-7B-width layer at the bench's context."""
    monkeypatch.setenv("MERLIN_DECODE", "kernels")
    g, cfg, model = _model(hidden=dims[0], heads=dims[1], inter=dims[2], layers=dims[3])
    gen = torch.Generator().manual_seed(extra)
    hi = min(30000, cfg.vocab_size - 4)          # below the image tokens
    ids = torch.cat([g["input_ids"][:1], torch.randint(3, hi, (1, extra), generator=gen)], dim=1).to(DEV)
    images = [g["images"][0].to(DEV)]
    S = ids.shape[1]
    prefix = MerlinKVCache(cfg.num_layers, 1, S + 8, cfg.hidden_size, DEV)
    _prefill(model, ids, images, prefix)
    runner = D.BeamDecodeRunner(model, prefix, nb, 64)
    runner.begin(S)
    singles = []
    for _ in range(nb):
        c = MerlinKVCache(cfg.num_layers, 1, S + 40, cfg.hidden_size, DEV)
        _prefill(model, ids, images, c)
        singles.append(c)
    for t in range(5):
        tokens = torch.randint(3, hi, (nb,), generator=gen).tolist()
        got = runner.step(tokens, None).float().cpu()
        emb = model.model.embed_tokens[torch.tensor(tokens, device=DEV)]
        for j in range(nb):
            ref = D.DecodeRunner(model, singles[j]).step_logits(emb[j:j + 1])[0].float().cpu()
            scale = ref.abs().max().item()
            err = (got[j] - ref).abs().max().item()
            assert err <= 2.0 ** -6 * scale, f"S={S} step {t} beam {j}: max|Δ| = {err} (scale {scale})"


def test_batched_step_handles_an_odd_vocabulary_and_an_empty_prefix_split():
    """vocab 32003 (the last row pair has one row) and a prompt shorter than the number of key splits."""
    monkeypatch_vocab = 32003
    g, cfg, model = _model(hidden=256, heads=2, inter=384, layers=1, vocab=monkeypatch_vocab)
    ids = g["input_ids"][:1].to(DEV)
    images = [g["images"][0].to(DEV)]
    S = ids.shape[1]
    prefix = MerlinKVCache(cfg.num_layers, 1, S + 8, cfg.hidden_size, DEV)
    out = _prefill(model, ids, images, prefix)
    runner = D.BeamDecodeRunner(model, prefix, 2, 32)
    runner.begin(S)
    tok = int(out.logits[0, -1].argmax())
    got = runner.step([tok, tok], None).float().cpu()
    assert got.shape == (2, cfg.vocab_size) and torch.isfinite(got).all()
    assert torch.equal(got[0], got[1])                 # identical beams -> identical rows (deterministic reductions)
    single = MerlinKVCache(cfg.num_layers, 1, S + 8, cfg.hidden_size, DEV)
    _prefill(model, ids, images, single)
    ref = D.DecodeRunner(model, single).step_logits(model.model.embed_tokens[tok].view(1, -1))[0].float().cpu()
    assert (got[0] - ref).abs().max().item() <= 2.0 ** -6 * ref.abs().max().item()


def test_generate_num_beams_reuses_one_runner_and_is_deterministic():
    g, cfg, model = _model()
    ids = g["input_ids"][:1].to(DEV)
    images = [g["images"][0].to(DEV)]
    a = model.generate(ids, images=images, num_beams=4, max_new_tokens=9, eos_token_id=None)
    b = model.generate(ids, images=images, num_beams=4, max_new_tokens=9, eos_token_id=None)
    assert torch.equal(a, b) and a.shape == (1, ids.shape[1] + 9)
    assert len(model._beam_runners) == 1
    # more beams than the batched kernel's 8 columns: the sequential path (one cache per beam) still answers
    c = model.generate(ids, images=images, num_beams=9, max_new_tokens=3, eos_token_id=None)
    assert c.shape == (1, ids.shape[1] + 3)
