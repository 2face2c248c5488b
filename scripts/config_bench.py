#!/usr/bin/env python
"""Inference timing of every BASELINE.json config on ONE GPU (one data-parallel rank's share): multimodal prefill +
32 greedy decode tokens through `model.generate` (CUDA graphs), random-init bf16 weights, synthetic inputs.
C2: 7B, 1 image, S=2048 | C3: 7B, 5 frames, S=4096 (= bench.py) | C4: 7B, 8 images, S=6144 | C5: 13B, 8 frames, S=8192.
Writes gpurun_out/r2_config_bench.json.   usage: python scripts/config_bench.py [C2 C3 C4 C5]"""
import json
import os
import sys
import warnings
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from merlin_b200.mmgpt.model.mmgpt.llama_mmgpt import MMGPTLlamaForCausalLM, vicuna_7b_config, vicuna_13b_config  # noqa: E402
from merlin_b200.mmgpt.model.projector.builder import build_projector  # noqa: E402
from merlin_b200.mmgpt.model.vision_encoder.builder import build_vision_tower  # noqa: E402

CASES = {"C2": ("7b", 1, 2048), "C3": ("7b", 5, 4096), "C4": ("7b", 8, 6144), "C5": ("13b", 8, 8192)}
DECODE = 32


def dims_of(dims):
    return (5120, 13824, 40, 32003) if dims == "13b" else (4096, 11008, 32, 32003)


def build(dims, S, dev):
    torch.manual_seed(0)
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            cfg = (vicuna_13b_config if dims == "13b" else vicuna_7b_config)(max_position_embeddings=max(4096, S + 64))
            model = MMGPTLlamaForCausalLM(cfg)
            margs = SimpleNamespace(vision_tower="clip-vit-large-patch14-336", vision_select_layer=-2,
                                    vision_select_feature="patch", freeze_vision_tower=True, freeze_projector=True,
                                    conv_stride=2, projector="conv", model_name_or_path=None)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                tower = build_vision_tower(margs)
            for prm in tower.parameters():
                if prm.dim() >= 2:
                    prm.data.normal_(0.0, 0.02)
            for lst in (tower.ln1_w, tower.ln2_w):
                for prm in lst:
                    prm.data.fill_(1.0)
            tower.pre_ln_w.data.fill_(1.0)
            model.get_model().vision_tower = tower
            model.get_model().projector = build_projector(margs, tower.hidden_size, cfg.hidden_size)
    finally:
        torch.set_default_dtype(torch.float32)
    return model.eval().requires_grad_(False)


def flops(dims, S, n_img):
    D, FF, L, V = dims_of(dims)
    dec = S * (L * (8 * D * D + 6 * D * FF + 2 * S * D) + 2 * D * V)
    vit = n_img * (577 * 23 * (8 * 1024 ** 2 + 4 * 1024 * 4096 + 4 * 577 * 1024) + 576 * 2 * 588 * 1024)
    return dec + vit + n_img * 144 * 2 * 9216 * D


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(pk)) if os.path.exists(pk) else {}
    out = {}
    for name in (sys.argv[1:] or list(CASES)):
        dims, n_img, S = CASES[name]
        model = build(dims, S, dev)
        g = torch.Generator().manual_seed(1234)
        ids = [1]
        for _ in range(n_img):
            ids += [32001] + [32000] * 144 + [32002, 13]
        ids += torch.randint(3, 32000, (S - len(ids) - 1,), generator=g).tolist() + [2]
        ids = torch.tensor([ids], device=dev)
        img = torch.empty(n_img, 3, 336, 336).normal_(0, 1, generator=g).to(device=dev, dtype=torch.bfloat16)

        def step(n_new):
            return model.generate(ids, images=[img], max_new_tokens=n_new, eos_token_id=None)

        def timed(n_new, reps=3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                step(n_new)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        for _ in range(3):
            step(DECODE)
        full = timed(DECODE)
        for _ in range(2):
            step(1)
        pre = timed(1)
        dec = (full - pre) / (DECODE - 1)
        fl = flops(dims, S, n_img)
        D, FF, L, V = dims_of(dims)
        wbytes = 2.0 * (L * (4 * D * D + 3 * D * FF + 2 * D) + D + V * D)
        kv = 4.0 * L * D * (S + DECODE / 2)
        out[name] = {"dims": dims, "images": n_img, "seq_len": S, "ms_per_step": full,
                     "tokens_per_s": (S + DECODE) / (full / 1e3), "prefill_ms": pre,
                     "prefill_tflops": fl / (pre / 1e3) / 1e12,
                     "prefill_frac_of_sustained": fl / (pre / 1e3) / 1e12 / peaks.get("bf16_tflops_sustained", 1404.0),
                     "decode_ms_per_token": dec, "decode_gb_per_token": (wbytes + kv) / 1e9,
                     "decode_hbm_frac": (wbytes + kv) / (dec / 1e3) / 1e9 / peaks.get("hbm_gbs", 6546.6),
                     "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9}
        print(name, json.dumps(out[name]), flush=True)
        del model
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_config_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
