"""Parity at BASELINE.json's configs, at FULL depth and width, against the reference itself run on the GPU.

C2 (7B, 1 image, S=2048), C3 (7B, 5 frames, S=4096 — exactly what bench.py times), C4 (7B, 8 images, S=6144),
C5 (13B, 40 layers, 8 frames, S=8192).  The checker is the REAL `MMGPTLlamaForCausalLM` of the reference
(oracle/_ref, built by oracle/build_ref.py; HF transformers modules underneath, eager attention, fp32 RoPE
angles) holding the same seeded weights and run on this GPU in fp32 with TF32 off.  When oracle/_ref is absent
the oracle restatement (oracle/merlin_oracle.py, pinned against the reference on CPU) is run on the GPU instead.

What is compared, per config (SURVEY.md §8c tolerances; north_star: "matching reference logits to 1e-2 bf16"):
  * logits of every row, ours (bf16 kernels) vs reference fp32: max, p99.99 and rms of |Δ|;
  * the same statistics for the reference's OWN bf16 path (the same HF modules cast to bf16, cuBLAS GEMMs) vs
    its fp32 run — the precision floor of a bf16 residual stream at this depth;
  * per-row argmax agreement wherever the fp32 top-1 / top-2 margin exceeds twice the error bound;
  * C3 and C5: 32 decode tokens through the streaming decode kernel at context S .. S+31, teacher-forced with the
    reference's greedy tokens, against the reference's KV-cache decode in fp32; the free-running greedy ids too.

Bounds (measured distributions: DESIGN.md §4, profiles/r2_config_parity.json).  north_star's max|Δ| <= 1e-2 is
enforced on the 2-layer goldens (tests/test_model_gpu.py); a 32-40 layer bf16 residual stream with random-init
weights cannot hold it on logits of magnitude 7 — the reference's OWN bf16 path measures rms 0.069 / max 0.51
against its fp32 run at C2 (~260 bf16 roundings of 2^-9 each in sequence) — so at full depth the enforced bar is
  (a) ours is at least as close to the fp32 reference as the reference's bf16 path is: rms and max no larger;
  (b) absolute: rms|Δ| <= 2^-7 · max|ref| and max|Δ| <= 2^-4 · max|ref|;
  (c) per-row argmax identical wherever the fp32 top-1/top-2 margin exceeds twice the max error;
  (d) decode steps: same bounds per step, argmax identical or inside the margin.
"""
import json
import os

import pytest
import torch

from helpers import build_product_model
from oracle import merlin_oracle as O
from oracle import refmodel

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = {
    "C2": dict(dims="7b", n_img=1, S=2048, decode=0),
    "C3": dict(dims="7b", n_img=5, S=4096, decode=32),
    "C4": dict(dims="7b", n_img=8, S=6144, decode=0),
    "C5": dict(dims="13b", n_img=8, S=8192, decode=32),
}
RMS_REL, MAX_REL = 2.0 ** -7, 2.0 ** -4   # of max|reference logit|
REL_TO_REF_BF16 = 1.0                     # ours vs the reference's own bf16 path, both against fp32


def _cfg(dims):
    if dims == "13b":
        return O.OracleConfig(hidden_size=5120, intermediate_size=13824, num_layers=40, num_heads=40)
    return O.OracleConfig()


def _stats(a, b):
    """max / p99.99 / rms of |a - b| over all logits, computed in chunks (no second full-size temporary)."""
    mx, sq, n = 0.0, 0.0, 0
    samples = []
    g = torch.Generator(device=a.device).manual_seed(0)
    for i in range(0, a.shape[0], 512):
        d = (a[i:i + 512].float() - b[i:i + 512].float()).abs()
        mx = max(mx, d.max().item())
        sq += d.double().pow(2).sum().item()
        n += d.numel()
        flat = d.flatten()
        samples.append(flat[torch.randint(flat.numel(), (65536,), device=a.device, generator=g)])
    s = torch.cat(samples)
    k = max(1, int(s.numel() * 1e-4))
    p9999 = s.topk(k).values[-1].item()
    return dict(max=mx, p9999=p9999, rms=(sq / n) ** 0.5)


def _report(name, payload):
    os.makedirs("gpurun_out", exist_ok=True)
    path = os.path.join("gpurun_out", "config_parity.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[name] = payload
    json.dump(data, open(path, "w"), indent=1)
    print(f"\n[config parity] {name}: {json.dumps(payload)}", flush=True)


class _Ref:
    """The reference (or, without oracle/_ref, the oracle restatement) on the GPU."""

    def __init__(self, cfg, sd, S):
        self.cfg, self.real = cfg, refmodel.available()
        self.max_pos = max(4096, S + 64)
        if self.real:
            self.model = refmodel.build_reference_model(cfg, sd, device=DEV, dtype=torch.float32, attn="eager",
                                                        max_position_embeddings=self.max_pos)
        else:
            self.sd = {k: v.float() for k, v in sd.items()}
        self.cache = None

    def to_bf16(self):
        if self.real:
            self.model.to(torch.bfloat16)
            refmodel._fix_buffers(self.model, self.cfg, DEV)
        else:
            self.sd = {k: v.to(torch.bfloat16) for k, v in self.sd.items()}

    @torch.no_grad()
    def prefill(self, ids, images, use_cache=False):
        """Logits [S, V] of MMGPTLlamaForCausalLM.forward (llama_mmgpt.py:53-112); with use_cache the HF cache
        the decode steps continue from is kept (generate()'s step 0)."""
        if not self.real:
            dt = next(iter(self.sd.values())).dtype
            imgs = [im.to(dt) for im in images]
            feats = O.encode_images(self.sd, self.cfg, imgs)
            emb = O.prepare_inputs_embeds(self.sd, self.cfg, ids, feats)
            self.cache = [None] * self.cfg.num_layers if use_cache else None
            h = O.decoder_forward(self.sd, self.cfg, emb, None, torch.arange(ids.shape[1], device=DEV), self.cache)
            return torch.nn.functional.linear(h, self.sd["lm_head.weight"])[0]
        m = self.model
        dt = m.lm_head.weight.dtype
        out = m(input_ids=ids, attention_mask=torch.ones_like(ids, dtype=torch.bool),
                images=[im.to(dt) for im in images], use_cache=use_cache)
        self.cache = out.past_key_values if use_cache else None
        return out.logits[0]

    @torch.no_grad()
    def decode(self, token, pos):
        """One cached decode step of the HF path generate() drives (llama_mmgpt.py:114-134: last token only,
        vision path skipped, base_mmgpt.py:85-89)."""
        if not self.real:
            e = torch.nn.functional.embedding(token.view(1, 1), self.sd["model.embed_tokens.weight"])
            h = O.decoder_forward(self.sd, self.cfg, e, None, torch.tensor([pos], device=DEV), self.cache)
            return torch.nn.functional.linear(h[0, -1], self.sd["lm_head.weight"])
        m = self.model
        out = m.model(input_ids=token.view(1, 1), past_key_values=self.cache, use_cache=True,
                      position_ids=torch.tensor([[pos]], device=DEV))
        self.cache = out.past_key_values
        return m.lm_head(out.last_hidden_state[0, -1])


@pytest.mark.parametrize("name", list(CASES))
def test_config_matches_reference_on_gpu(name):
    case = CASES[name]
    if name == "C5" and torch.cuda.get_device_properties(0).total_memory < 150e9:
        pytest.skip("C5 (13B, S=8192) needs the 180 GB of a B200")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg, S, n_img = _cfg(case["dims"]), case["S"], case["n_img"]
    sd = O.init_state_dict(cfg, seed=0, dtype=torch.bfloat16, device=DEV)
    batch = O.synthetic_batch(cfg, 1, S, [n_img], seed=1234)
    ids = batch["input_ids"].to(DEV)
    images = [batch["images"][0].to(DEV).to(torch.bfloat16)]     # bf16-exact pixels for every arm
    model = build_product_model(cfg, sd, device=DEV, max_position_embeddings=max(4096, S + 64))
    ref = _Ref(cfg, sd, S)
    del sd
    torch.cuda.empty_cache()
    n_dec = case["decode"]

    # ---- ours: prefill (+ cache), then the free-running greedy ids through generate()
    with torch.no_grad():
        out = model(input_ids=ids, images=images, use_cache=bool(n_dec))
    ours = out.logits[0]
    assert ours.shape == (S, cfg.vocab_size) and torch.isfinite(ours).all()

    # ---- reference fp32: prefill, then n_dec teacher-forced decode steps on both sides
    ref32 = ref.prefill(ids, [im.float() for im in images], use_cache=bool(n_dec))
    st = _stats(ours, ref32)
    top2 = ref32.float().topk(2, dim=-1)
    margin = top2.values[:, 0] - top2.values[:, 1]
    agree = ours.float().argmax(-1) == top2.indices[:, 0]
    clear = margin > 2 * st["max"]
    payload = dict(case=case, reference="oracle/_ref (real reference, fp32, eager)" if ref.real else "oracle restatement fp32",
                   logit_absmax=ref32.abs().max().item(), ours_vs_fp32=st,
                   argmax_agree=agree.float().mean().item(), rows_with_clear_margin=int(clear.sum()),
                   argmax_agree_clear=bool(agree[clear].all()))
    dec = []
    if n_dec:
        runner = model._get_decoder(out.past_key_values)
        assert runner.mode == "stream"
        tok = top2.indices[-1, 0]
        ref_ids = [int(tok)]
        for t in range(n_dec):
            r = ref.decode(tok, S + t).float()
            o = runner.step_logits(model.model.embed_tokens[tok.view(1)]).float()[0].clone()
            err = (o - r).abs()
            tv, ti = r.topk(2)
            dec.append(dict(pos=S + t, max=err.max().item(), rms=err.pow(2).mean().sqrt().item(),
                            margin=(tv[0] - tv[1]).item(), same_argmax=bool(o.argmax() == ti[0])))
            tok = ti[0]
            ref_ids.append(int(tok))
        payload["decode"] = dict(tokens=n_dec, ctx=[S, S + n_dec - 1], max=max(d["max"] for d in dec),
                                 rms=max(d["rms"] for d in dec), argmax_same=sum(d["same_argmax"] for d in dec),
                                 steps=dec)
        gen = model.generate(ids, images=images, max_new_tokens=n_dec, eos_token_id=None)[0, S:].tolist()
        same = 0
        while same < n_dec and gen[same] == ref_ids[same]:
            same += 1
        payload["decode"]["free_running_greedy_prefix_equal"] = same
        payload["decode"]["ref_ids"], payload["decode"]["our_ids"] = ref_ids[:n_dec], gen

    # ---- the reference's own bf16 path against its fp32 run (precision floor)
    ref.cache = None
    torch.cuda.empty_cache()
    ref.to_bf16()
    ref16 = ref.prefill(ids, images)
    st16 = _stats(ref16, ref32)
    st_vs16 = _stats(ours, ref16)
    payload.update(ref_bf16_vs_fp32=st16, ours_vs_ref_bf16=st_vs16)
    _report(name, payload)

    scale = payload["logit_absmax"]
    assert st["rms"] <= RMS_REL * scale and st["max"] <= MAX_REL * scale, (name, st, scale)
    assert st["rms"] <= REL_TO_REF_BF16 * st16["rms"] and st["max"] <= REL_TO_REF_BF16 * st16["max"], (st, st16)
    assert payload["argmax_agree_clear"], f"{name}: argmax differs on a row whose fp32 margin exceeds 2x the error"
    for d in dec:
        assert d["rms"] <= RMS_REL * scale and d["max"] <= MAX_REL * scale, d
        assert d["same_argmax"] or d["margin"] <= 2 * d["max"], d
    del ref, model
    torch.cuda.empty_cache()
