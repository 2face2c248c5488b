"""N4: host-side generation logic, pinned on CPU against the third-party library the reference
calls (transformers `generate`): beam search bookkeeping (merlin_b200/beam.py), top-p / temperature sampling
filter and KeywordsStoppingCriteria (mmgpt/utils/mm_utils.py:62-85)."""
import pytest
import torch

from merlin_b200.beam import BeamSearch

transformers = pytest.importorskip("transformers")


def _tiny_llama(seed, vocab=37):
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(seed)
    cfg = LlamaConfig(vocab_size=vocab, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                      num_attention_heads=2, num_key_value_heads=2, max_position_embeddings=256,
                      attn_implementation="eager")
    m = LlamaForCausalLM(cfg).eval()
    with torch.no_grad():           # sharpen the distribution a little so that EOS / ties are not degenerate
        m.lm_head.weight.mul_(8.0)
    return m


def _our_beam(model, ids, nb, eos, max_new, **kw):
    bs = BeamSearch(ids, nb, eos, max_new, **kw)
    with torch.no_grad():
        while True:
            logits = model(input_ids=bs.running).logits[:, -1, :]
            if bs.step(logits) is None:
                break
    return bs.finalize()


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("nb", [2, 5])
@pytest.mark.parametrize("eos", [None, 3, 11])
def test_beam_search_matches_hf_generate(seed, nb, eos):
    model = _tiny_llama(seed)
    torch.manual_seed(100 + seed)
    ids = torch.randint(4, 37, (1, 9))
    max_new = 12
    with torch.no_grad():
        ref = model.generate(ids, num_beams=nb, do_sample=False, max_new_tokens=max_new, eos_token_id=eos,
                             pad_token_id=0, use_cache=False)
    got = _our_beam(model, ids, nb, eos, max_new)
    ref = ref[0]
    if eos is not None and (ref[ids.shape[1]:] == eos).any():      # HF pads after EOS in its fixed-size buffer
        ref = ref[: ids.shape[1] + int((ref[ids.shape[1]:] == eos).nonzero()[0]) + 1]
    assert got[0].tolist() == ref.tolist()


@pytest.mark.parametrize("lp,es", [(0.0, False), (2.0, False), (1.0, True), (1.0, "never")])
def test_beam_search_length_penalty_and_early_stopping_match_hf(lp, es):
    model = _tiny_llama(7)
    ids = torch.randint(4, 37, (1, 6), generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = model.generate(ids, num_beams=4, do_sample=False, max_new_tokens=10, eos_token_id=5, pad_token_id=0,
                             length_penalty=lp, early_stopping=es, use_cache=False)[0]
    got = _our_beam(model, ids, 4, 5, 10, length_penalty=lp, early_stopping=es)[0]
    if (ref[6:] == 5).any():
        ref = ref[: 6 + int((ref[6:] == 5).nonzero()[0]) + 1]
    assert got.tolist() == ref.tolist()


def test_beam_search_total_length_normalisation_changes_only_the_denominator():
    """transformers 4.31 (pinned by the reference) divides by the total length; with length_penalty 0 both
    variants must coincide, and with a long prompt the total-length variant prefers what penalty 0 prefers."""
    model = _tiny_llama(2)
    ids = torch.randint(4, 37, (1, 40), generator=torch.Generator().manual_seed(9))
    a = _our_beam(model, ids, 3, 7, 8, length_penalty=0.0, length_norm="generated")
    b = _our_beam(model, ids, 3, 7, 8, length_penalty=0.0, length_norm="total")
    assert torch.equal(a, b)
    c = _our_beam(model, ids, 3, 7, 8, length_penalty=1.0, length_norm="total")
    assert c.shape[1] >= ids.shape[1] + 1


def test_sampling_distribution_matches_hf_warpers():
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopPLogitsWarper
    from merlin_b200.mmgpt.model.mmgpt.llama_mmgpt import sampling_distribution
    torch.manual_seed(0)
    logits = torch.randn(3, 101) * 3
    for temperature, top_p in ((0.2, 1.0), (0.2, 0.7), (1.0, 0.9), (0.7, 0.05)):
        ref = TemperatureLogitsWarper(temperature)(None, logits.clone())
        if top_p < 1.0:
            ref = TopPLogitsWarper(top_p)(None, ref)
        ref = torch.softmax(ref, dim=-1)
        got = sampling_distribution(logits, temperature, top_p)
        assert torch.allclose(got, ref, atol=1e-7)
        assert torch.equal(got > 0, ref > 0)


class _WordTokenizer:
    """whitespace tokenizer with a BOS, enough of the HF API for KeywordsStoppingCriteria"""
    bos_token_id = 1

    def __init__(self):
        self.vocab = {"<s>": 1}

    def _id(self, w):
        return self.vocab.setdefault(w, len(self.vocab) + 1)

    def __call__(self, text):
        from types import SimpleNamespace
        return SimpleNamespace(input_ids=[1] + [self._id(w) for w in text.split()])

    def batch_decode(self, ids, skip_special_tokens=True):
        inv = {v: k for k, v in self.vocab.items()}
        return [" ".join(inv.get(int(t), "?") for t in row if not (skip_special_tokens and int(t) == 1)) for row in ids]


def test_keywords_stopping_criteria_matches_the_reference_class():
    import importlib.util
    import os
    from merlin_b200.mmgpt.utils.mm_utils import KeywordsStoppingCriteria
    tok = _WordTokenizer()
    prompt = torch.tensor([tok("a b c").input_ids])
    ours = KeywordsStoppingCriteria(["</s>", "stop now"], tok, prompt)
    theirs = None
    ref_path = "/root/reference/mmgpt/utils/mm_utils.py"
    if os.path.exists(ref_path):     # build container only: the reference's own class on the same inputs
        spec = importlib.util.spec_from_file_location("ref_mm_utils", ref_path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        theirs = mod.KeywordsStoppingCriteria(["</s>", "stop now"], tok, prompt)
    end = tok("</s>").input_ids[1]
    cases = [
        (prompt, False),
        (torch.cat([prompt, torch.tensor([[tok._id("d")]])], 1), False),
        (torch.cat([prompt, torch.tensor([[tok._id("d"), end]])], 1), True),                       # single-token keyword
        (torch.cat([prompt, torch.tensor([[tok._id("x"), tok._id("stop"), tok._id("now")]])], 1), True),  # decoded text
        (torch.cat([prompt, torch.tensor([[tok._id("stop"), tok._id("x"), tok._id("y"), tok._id("z")]])], 1), False),
    ]
    for ids, want in cases:
        assert ours(ids, None) is want
        if theirs is not None and ids.shape[1] > prompt.shape[1]:
            try:
                ref = bool(theirs(ids, None))
            except RuntimeError:      # the reference's `if tensor == tensor` is ambiguous for multi-token keywords
                continue
            assert ref is want


def test_mm_utils_helpers():
    from PIL import Image
    from merlin_b200.mmgpt.utils.mm_utils import expand2square, get_model_name_from_path
    im = Image.new("RGB", (6, 4), (255, 0, 0))
    sq = expand2square(im, (1, 2, 3))
    assert sq.size == (6, 6) and sq.getpixel((0, 0)) == (255, 0, 0) and sq.getpixel((0, 5)) == (1, 2, 3)
    assert expand2square(sq, (0, 0, 0)) is sq
    assert get_model_name_from_path("/a/b/merlin-7b/checkpoint-100/") == "merlin-7b_checkpoint-100"
    assert get_model_name_from_path("/a/b/merlin-7b") == "merlin-7b"
