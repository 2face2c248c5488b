"""GPU: the drop-in boundary of SURVEY.md §8b as a reference user would exercise it.

  * `MMGPTLlamaForCausalLM` is an HF `PreTrainedModel` + `GenerationMixin` registered as "mmgpt"
    (llama_mmgpt.py:38,137-138): `AutoModelForCausalLM.from_config`, `save_pretrained` / `from_pretrained`;
  * `model(**batch).loss.backward()` (what HF Trainer.training_step does, train.py:49-59) fills `param.grad` with the
    gradients `MMGPTTrainer`'s explicit schedule computes and that the oracle's autograd (pinned against the
    reference's own backward) predicts — also with `gradient_checkpointing_enable()` (pretrain.sh:35);
  * the attention seam (`replace_llama_attn_with_flash_attn`, train_flash_attn.py:6-8) trains: gradients of an HF
    LlamaModel through `merlin_flash_attn_fwd / _bwd` against HF eager attention;
  * `generate(**hf_kwargs)`: GenerationConfig / keyword handling, HF's own generation loop over our forward,
    the streamer hook (model_worker.py:162-187), thread-safety, and reuse of one decode state across prompt lengths.
"""
import threading

import pytest
import torch

from helpers import build_product_model, load_golden
from oracle import merlin_oracle as O
from test_train_gpu import oracle_grads, unpack_grads

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _to_dev(images):
    return [im.to(DEV) for im in images]


def _trainable(model, tower: bool):
    model.requires_grad_(True)
    for p in model.get_model().vision_tower.parameters():
        p.requires_grad_(tower)
    for p in (model.get_model().vision_tower.post_ln_w, model.get_model().vision_tower.post_ln_b):
        p.requires_grad_(False)
    return model


def _batch(g):
    return dict(input_ids=g["input_ids"].to(DEV), attention_mask=g["attention_mask"].to(DEV),
                labels=g["labels"].to(DEV), images=_to_dev(g["images"]))


def test_auto_classes_resolve_mmgpt_and_checkpoint_round_trip(tmp_path):
    from transformers import AutoConfig, AutoModelForCausalLM, GenerationMixin, PreTrainedModel
    from merlin_b200.mmgpt.model.mmgpt.llama_mmgpt import MMGPTConfig, MMGPTLlamaForCausalLM
    cfg = MMGPTConfig(vocab_size=300, hidden_size=256, intermediate_size=384, num_hidden_layers=2,
                      num_attention_heads=2, num_key_value_heads=2, tie_word_embeddings=False)
    assert isinstance(AutoConfig.for_model("mmgpt"), MMGPTConfig) or AutoConfig.for_model("mmgpt").model_type == "mmgpt"
    model = AutoModelForCausalLM.from_config(cfg)
    assert isinstance(model, MMGPTLlamaForCausalLM) and isinstance(model, PreTrainedModel)
    assert isinstance(model, GenerationMixin) and model.can_generate()
    model = model.to(device=DEV, dtype=torch.bfloat16).requires_grad_(False)
    ids = torch.randint(3, 290, (1, 40), device=DEV)
    a = model(input_ids=ids).logits
    model.save_pretrained(str(tmp_path))
    again = MMGPTLlamaForCausalLM.from_pretrained(str(tmp_path), torch_dtype=torch.bfloat16).to(DEV).requires_grad_(False)
    assert torch.equal(again(input_ids=ids).logits, a)
    model.save_pretrained(str(tmp_path / "st"), safe_serialization=True)
    third = MMGPTLlamaForCausalLM.from_pretrained(str(tmp_path / "st")).to(device=DEV, dtype=torch.bfloat16)
    assert torch.equal(third.requires_grad_(False)(input_ids=ids).logits, a)


@pytest.mark.parametrize("checkpointing", [False, True])
def test_loss_backward_fills_param_grad_like_the_trainer_and_the_oracle(checkpointing):
    from merlin_b200.mmgpt.engine.train.trainer import MMGPTTrainer, TrainingArguments
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    ref_loss, ref = oracle_grads(cfg, sd, g)

    model = _trainable(build_product_model(cfg, sd), tower=True)
    model.train()
    if checkpointing:
        model.gradient_checkpointing_enable()
    out = model(**_batch(g))
    assert out.logits.shape == (2, g["input_ids"].shape[1], cfg.vocab_size) and out.loss.requires_grad
    assert abs(out.loss.item() - ref_loss.item()) <= 2e-3
    out.loss.backward()
    named = dict(model.named_parameters_reference(trainable_only=True))
    assert all(p.grad is not None and p.grad.shape == p.shape for p in named.values()), \
        [n for n, p in named.items() if p.grad is None]

    # the same gradients out of MMGPTTrainer's explicit schedule (second model: the trainer re-homes parameters)
    twin = _trainable(build_product_model(cfg, sd), tower=True)
    trainer = MMGPTTrainer(twin, TrainingArguments(learning_rate=0.0, max_steps=10))
    trainer.args.max_grad_norm = 0.0
    batch = {k: g[k] for k in ("input_ids", "labels", "attention_mask")}
    batch["images"] = _to_dev(g["images"])
    with torch.no_grad():
        trainer._last_micro, trainer._fold_ptr = True, 0
        trainer.reducer.start_step()
        trainer._forward_backward(batch, accumulate=False, grad_scale=1.0)
    worst = 0.0
    for n, p in named.items():
        a, b = p.grad.float().flatten(), trainer.arena.grad(n).float().flatten()
        scale = b.abs().max().item()
        worst = max(worst, (a - b).abs().max().item() / (scale + 1e-12))
        assert (a - b).abs().max().item() <= 2.0 ** -6 * scale + 1e-9, (n, (a - b).abs().max().item(), scale)
    print(f"\n[boundary] loss.backward() vs MMGPTTrainer: worst rel max|Δ| over {len(named)} tensors = {worst:.2e}")

    # and against the oracle's autograd, under the reference's parameter names
    class _View:
        names = [n for n in named]
        arena = type("A", (), {"grad": staticmethod(lambda n: named[n].grad)})()
    got = unpack_grads(_View, cfg)
    checked = 0
    for k, gr in ref.items():
        if k not in got or gr.abs().max().item() < 1e-7:
            continue
        a, b = got[k].float().flatten().cpu(), gr.flatten()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        assert cos >= 0.999, (k, cos)
        checked += 1
    assert checked >= 6 * cfg.num_layers + 10


def test_attention_seam_trains_gradients_match_hf_eager():
    from transformers import LlamaConfig, LlamaModel
    from merlin_b200.mmgpt.utils.llama_flash_attn_monkey_patch import replace_llama_attn_with_flash_attn
    name = replace_llama_attn_with_flash_attn()
    cfg = LlamaConfig(vocab_size=300, hidden_size=256, intermediate_size=384, num_hidden_layers=2,
                      num_attention_heads=2, num_key_value_heads=2, attn_implementation="eager")
    torch.manual_seed(0)
    m = LlamaModel(cfg).to(DEV, torch.bfloat16).train()
    ids = torch.randint(0, 300, (2, 200), device=DEV)
    mask = torch.ones(2, 200, dtype=torch.long, device=DEV)
    mask[1, 150:] = 0
    probe = torch.randn(2, 200, 256, device=DEV, dtype=torch.bfloat16) * mask[..., None]

    def grads(impl):
        m.config._attn_implementation = impl
        m.zero_grad(set_to_none=True)
        (m(input_ids=ids, attention_mask=mask).last_hidden_state * probe).sum().backward()
        return {k: p.grad.float().clone() for k, p in m.named_parameters() if p.grad is not None}

    ref, got = grads("eager"), grads(name)
    assert set(ref) == set(got)
    for k in ref:
        if ref[k].abs().max().item() < 1e-6:
            continue
        cos = torch.nn.functional.cosine_similarity(got[k].flatten(), ref[k].flatten(), dim=0).item()
        assert cos >= 0.995, (k, cos)


def test_generate_accepts_hf_arguments_and_rejects_unknown_ones():
    from transformers import GenerationConfig
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0))
    ids, imgs = g["input_ids"][:1].to(DEV), _to_dev(g["images"][:1])
    n_new = g["ref_generate"].shape[1] - ids.shape[1]
    want = g["ref_generate"]
    a = model.generate(ids, images=imgs, max_new_tokens=n_new, eos_token_id=None)
    assert torch.equal(a.cpu(), want)
    b = model.generate(input_ids=ids, images=imgs, generation_config=GenerationConfig(max_new_tokens=n_new, eos_token_id=None,
                                                                                  do_sample=False))
    assert torch.equal(b.cpu(), want)
    c = model.generate(ids, images=imgs, max_length=ids.shape[1] + n_new, eos_token_id=None, return_dict_in_generate=True)
    assert torch.equal(c.sequences.cpu(), want)
    with pytest.raises(ValueError):
        model.generate(ids, images=imgs, max_new_tokens=2, penalty_alpha=0.6)
    # repetition penalty / top-k / custom processors run through the host loop on the same decode kernels
    seen = []
    out = model.generate(ids, images=imgs, max_new_tokens=6, eos_token_id=None, repetition_penalty=1.3, do_sample=True,
                         top_k=5, temperature=0.7, generator=torch.Generator(DEV).manual_seed(0),
                         logits_processor=[lambda i, s: (seen.append(i.shape[1]), s)[1]])
    assert out.shape[1] == ids.shape[1] + 6 and seen == list(range(ids.shape[1], ids.shape[1] + 6))


def test_hf_generation_loop_runs_over_our_forward():
    """use_hf_loop=True: transformers' GenerationMixin.generate drives forward() / prepare_inputs_for_generation with
    its own DynamicCache bookkeeping (SURVEY.md §0 fact 9 is what breaks the reference here) and reproduces the
    reference's greedy tokens."""
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0))
    ids, imgs = g["input_ids"][:1].to(DEV), _to_dev(g["images"][:1])
    n_new = g["ref_generate"].shape[1] - ids.shape[1]
    out = model.generate(ids, images=imgs, max_new_tokens=n_new, do_sample=False, eos_token_id=None, pad_token_id=0,
                         use_hf_loop=True)
    assert torch.equal(out.cpu(), g["ref_generate"])


class _Streamer:
    def __init__(self):
        self.tokens, self.ended = [], False

    def put(self, value):
        self.tokens += value.flatten().tolist()

    def end(self):
        self.ended = True


def test_streamer_receives_every_new_token_then_end():
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0))
    ids, imgs = g["input_ids"][:1].to(DEV), _to_dev(g["images"][:1])
    n_new = g["ref_generate"].shape[1] - ids.shape[1]
    st = _Streamer()
    out = model.generate(ids, images=imgs, max_new_tokens=n_new, eos_token_id=None, streamer=st)
    assert st.ended and st.tokens == out[0, ids.shape[1]:].tolist() == g["ref_generate"][0, ids.shape[1]:].tolist()
    # the serve worker's pattern (model_worker.py:162-187): generate in a background thread, consume the streamer
    st2, box = _Streamer(), {}
    t = threading.Thread(target=lambda: box.update(out=model.generate(ids, images=imgs, max_new_tokens=n_new,
                                                                      eos_token_id=None, streamer=st2)))
    t.start()
    t.join(120)
    assert st2.ended and st2.tokens == st.tokens and torch.equal(box["out"], out)


def test_concurrent_generate_calls_do_not_corrupt_each_other():
    g, cfg = load_golden("tiny_conv")
    model = build_product_model(cfg, O.init_state_dict(cfg, seed=0))
    imgs = _to_dev(g["images"][:1])
    prompts = [g["input_ids"][:1, :n].to(DEV) for n in (g["input_ids"].shape[1], g["input_ids"].shape[1] - 5,
                                                        g["input_ids"].shape[1] - 9)]
    serial = [model.generate(p, images=imgs, max_new_tokens=10, eos_token_id=None).cpu() for p in prompts]
    results = [None] * 6

    def work(i):
        torch.cuda.set_device(0)
        results[i] = model.generate(prompts[i % 3], images=imgs, max_new_tokens=10, eos_token_id=None)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(300)
    for i in range(6):
        assert results[i] is not None and torch.equal(results[i].cpu(), serial[i % 3]), i
    # every prompt length shared ONE decode state (KV cache + runner + per-token graph) of the 1024-token bucket
    assert len(model._decode_states) == 1 and len(model._graphed) == 3
