"""Pins oracle/merlin_oracle.py against the REAL reference and writes tests/golden/*.pt.

Runs only in the build container (needs /root/reference).  What it does:
  1. builds oracle/_ref (oracle/build_ref.py: a copy of /root/reference/mmgpt with the one missing ')' of
     mmgpt/utils/constants.py:72 restored — the tree as scraped has a SyntaxError at :25);
  2. imports the reference's own MMGPTLlamaForCausalLM / CLIPVisionTower / ConvProjector /
     MLPProjector (HF transformers 5.5.0 modules underneath, eager attention, fp32 RoPE),
     assembles them exactly like build_vision_tokenizer does (base_mmgpt.py:23-79) and loads the
     oracle's seeded weights through load_state_dict(strict=True) — which also pins the
     state-dict key contract of SURVEY.md Appendix A;
  3. runs the reference forward (loss, logits, inputs_embeds via a hook, image features) and a
     cache-free greedy loop on CPU fp32, asserts the oracle reproduces them, and saves the reference
     outputs as golden fixtures.

Usage: python oracle/make_golden.py            (writes tests/golden/{tiny_conv,tiny_mlp,tiny_pad}.pt)
"""
from __future__ import annotations

import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import merlin_oracle as O  # noqa: E402

REFERENCE = "/root/reference"


def import_reference():
    """The real reference classes, from the oracle/_ref copy oracle/build_ref.py makes (one missing ')' restored)."""
    from oracle import refmodel
    return refmodel.import_reference()


def build_reference_model(ref, cfg: O.OracleConfig, sd):
    from oracle import refmodel
    return refmodel.build_reference_model(cfg, sd, device="cpu", dtype=torch.float32, attn="eager", ref=ref)


def sha(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def make_case(ref, name: str, cfg: O.OracleConfig, batch, gen_tokens: int):
    sd = O.init_state_dict(cfg, seed=0)
    model = build_reference_model(ref, cfg, sd)
    captured = {}

    def hook(_m, args, kwargs):
        captured["inputs_embeds"] = kwargs.get("inputs_embeds")

    h = model.model.register_forward_pre_hook(hook, with_kwargs=True)
    with torch.no_grad():
        out = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                    labels=batch["labels"], images=batch["images"], use_cache=False)
        feats = model.encode_images(batch["images"])
    h.remove()
    ours = O.forward(sd, cfg, batch["input_ids"], batch["attention_mask"], batch["labels"], batch["images"])

    def close(a, b, what, tol=2e-5):
        err = (a - b).abs().max().item()
        scale = b.abs().max().item() + 1e-9
        print(f"  {name}: {what}: max|oracle-ref| = {err:.3e} (ref max {scale:.3e})")
        assert err <= tol * max(1.0, scale), f"oracle deviates from the reference on {what}"

    assert torch.equal(ours["inputs_embeds"], captured["inputs_embeds"]), "splice must be bit-exact"
    print(f"  {name}: inputs_embeds bit-exact vs reference")
    for a, b in zip(ours["image_features"], feats):
        close(a, b, "image_features")
    mask = batch["attention_mask"]
    valid = mask.bool()
    # padded query rows: flash-attn path returns zeros, eager returns finite garbage;
    # compare logits on valid rows only.
    close(ours["logits"][valid], out.logits[valid], "logits")
    close(ours["loss"], out.loss, "loss", tol=1e-5)

    # ---- backward: the reference's own autograd (loss.backward() through the real modules) pins the
    # oracle's autograd, which in turn is what the CUDA backward is tested against
    model.zero_grad()
    out_b = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], labels=batch["labels"],
                  images=batch["images"], use_cache=False)
    out_b.loss.backward()
    ref_grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.forward(leaves, cfg, batch["input_ids"], batch["attention_mask"], batch["labels"], batch["images"])["loss"].backward()
    worst = 0.0
    for k, gr in ref_grads.items():
        go = leaves[k].grad
        assert go is not None, f"oracle has no gradient for {k}"
        scale = gr.abs().max().item()
        err = (go - gr).abs().max().item()
        worst = max(worst, err / (scale + 1e-12) if scale > 1e-7 else 0.0)
        assert err <= 1e-4 * max(scale, 1e-6) + 1e-8, f"oracle autograd deviates from the reference on {k}: {err} vs {scale}"
    print(f"  {name}: {len(ref_grads)} parameter gradients: oracle autograd == reference autograd (worst rel {worst:.2e})")
    grad_probe = {k: ref_grads[k].flatten()[:64].clone() for k in sorted(ref_grads)[::7]}
    grad_norms = {k: float(g.norm()) for k, g in ref_grads.items()}

    gold = dict(name=name, cfg=cfg.__dict__, weights_sha256=sha(sd), seed=0, ref_grad_probe=grad_probe,
                ref_grad_norms=grad_norms,
                input_ids=batch["input_ids"], labels=batch["labels"], attention_mask=batch["attention_mask"],
                images=[im.clone() for im in batch["images"]],
                ref_loss=out.loss.detach().clone(), ref_logits=out.logits.detach().clone(),
                ref_inputs_embeds=captured["inputs_embeds"].detach().clone(),
                ref_image_features=[f.detach().clone() for f in feats])
    if gen_tokens:
        ids = batch["input_ids"][:1]
        imgs = batch["images"][:1]
        cur = ids.clone()
        with torch.no_grad():
            for _ in range(gen_tokens):  # cache-free greedy loop over the real reference forward
                lg = model(input_ids=cur, attention_mask=torch.ones_like(cur, dtype=torch.bool), images=imgs,
                           use_cache=False).logits
                cur = torch.cat([cur, lg[:, -1].argmax(-1, keepdim=True)], dim=1)
        mine = O.greedy_generate(sd, cfg, ids, imgs, gen_tokens, eos_token_id=None, use_cache=True)
        mine_nc = O.greedy_generate(sd, cfg, ids, imgs, 4, eos_token_id=None, use_cache=False)
        assert torch.equal(mine, cur), "greedy tokens (kv-cache oracle) differ from the reference loop"
        assert torch.equal(mine_nc, cur[:, : ids.shape[1] + 4])
        print(f"  {name}: {gen_tokens} greedy tokens identical to the reference loop")
        gold["ref_generate"] = cur
    path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(gold, path)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def make_llrd_golden(ref):
    """Pins the optimizer grouping (N1) against the reference's OWN `get_param_groups` + `vit_lr_scale_func` /
    `llm_lr_scale_func` (mmgpt/utils/llrd_utils.py:4-79), called the way MMGPTTrainer.create_optimizer does
    (mmgpt/engine/train/trainer.py:53-66), on the parameter names / shapes of an assembled reference model."""
    import json
    cfg = O.OracleConfig.tiny(vit_layers=24)          # real CLIP depth: the LLRD exponent depends on the layer index
    model = build_reference_model(ref, cfg, O.init_state_dict(cfg, seed=0))
    named = [(n, p) for n, p in model.named_parameters()]
    out = {"names": [[n, list(p.shape)] for n, p in named], "cases": {}}
    U = ref.llrd_utils
    for tag, fn in (("llrd", U.vit_lr_scale_func), ("llm_llrd", U.llm_lr_scale_func), ("none", None)):
        fake = type("M", (), {"named_parameters": staticmethod(lambda: named)})
        groups = U.get_param_groups(fake, None, fn, 1e-3, 0.05)
        ids = {id(p): n for n, p in named}
        out["cases"][tag] = [{"lr": g["lr"], "weight_decay": g["weight_decay"], "params": [ids[id(p)] for p in g["params"]]}
                             for g in groups]
        mine = O.param_groups([(n, tuple(p.shape)) for n, p in named], 1e-3, 0.05, llrd=(tag == "llrd")) if tag != "llm_llrd" else None
        if mine is not None:
            assert [(g["lr"], g["weight_decay"], g["names"]) for g in mine] == \
                [(g["lr"], g["weight_decay"], g["params"]) for g in out["cases"][tag]], f"oracle param_groups != reference ({tag})"
    path = os.path.join(ROOT, "tests", "golden", "llrd_groups.json")
    json.dump(out, open(path, "w"))
    print(f"  llrd: {len(named)} parameters, {[len(v) for v in out['cases'].values()]} groups -> {path}")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = import_reference()
    make_llrd_golden(ref)
    cfg = O.OracleConfig.tiny()
    # two samples, 1 and 2 images, S=48 (P' = 4 visual tokens per image at 56 px / stride 2)
    make_case(ref, "tiny_conv", cfg, O.synthetic_batch(cfg, 2, 48, [1, 2], seed=1234), gen_tokens=12)
    cfg_mlp = O.OracleConfig.tiny(projector="mlp", conv_stride=1)
    make_case(ref, "tiny_mlp", cfg_mlp, O.synthetic_batch(cfg_mlp, 1, 64, [2], seed=99), gen_tokens=0)
    # right-padded batch: exercises the key-padding mask and ignored labels
    b = O.synthetic_batch(cfg, 2, 40, [1, 1], seed=7, pad_to=56)
    b2 = O.synthetic_batch(cfg, 2, 56, [1, 1], seed=8)
    for k in ("input_ids", "labels", "attention_mask"):
        b[k][1] = b2[k][1]
    b["images"][1] = b2["images"][1]
    make_case(ref, "tiny_pad", cfg, b, gen_tokens=0)


if __name__ == "__main__":
    main()
