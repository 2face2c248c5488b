#!/usr/bin/env python
"""BASELINE config C5 on ONE GPU: Merlin-13B (D=5120, 40 layers, FF=13824), 8 frames, S=8192 — one data-parallel
rank's training step (fwd + bwd + AdamW; LLM + projector train, CLIP tower frozen) with the activation-recompute
policy the reference gets from --gradient_checkpointing True (pretrain.sh:35), plus the same step at C3.
Prints peak memory and step time (DESIGN.md §3 records them).   usage: python scripts/c5_feasibility.py [c5|c3]"""
import json
import os
import sys
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from merlin_b200.mmgpt.engine.train.trainer import MMGPTTrainer, TrainingArguments  # noqa: E402
from merlin_b200.mmgpt.model.mmgpt.llama_mmgpt import MMGPTLlamaForCausalLM, vicuna_7b_config, vicuna_13b_config  # noqa: E402
from merlin_b200.mmgpt.model.projector.builder import build_projector  # noqa: E402
from merlin_b200.mmgpt.model.vision_encoder.builder import build_vision_tower  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c5"
recompute = (sys.argv[2] != "norecompute") if len(sys.argv) > 2 else True
S, n_img, cfg_fn = (8192, 8, vicuna_13b_config) if which == "c5" else (4096, 5, vicuna_7b_config)
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    model = MMGPTLlamaForCausalLM(cfg_fn(max_position_embeddings=S))
    margs = SimpleNamespace(vision_tower="clip-vit-large-patch14-336", vision_select_layer=-2, vision_select_feature="patch",
                            freeze_vision_tower=True, freeze_projector=False, conv_stride=2, projector="conv",
                            model_name_or_path=None)
    import warnings
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
    model.get_model().projector = build_projector(margs, tower.hidden_size, model.config.hidden_size)
torch.set_default_dtype(torch.float32)
g = torch.Generator().manual_seed(1234)
ids = [1]
for _ in range(n_img):
    ids += [32001] + [32000] * 144 + [32002, 13]
ids += torch.randint(3, 32000, (S - len(ids) - 1,), generator=g).tolist() + [2]
input_ids = torch.tensor([ids])
labels = input_ids.clone()
labels[:, : S // 2] = -100
images = torch.empty(n_img, 3, 336, 336).normal_(0, 1, generator=g).to(device=dev, dtype=torch.bfloat16)
batch = {"input_ids": input_ids, "labels": labels, "attention_mask": None, "images": [images]}
trainer = MMGPTTrainer(model, TrainingArguments(learning_rate=1e-5, max_steps=1000, gradient_checkpointing=recompute))
torch.cuda.reset_peak_memory_stats()
losses = [trainer.training_step(batch).item() for _ in range(2)]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    losses.append(trainer.training_step(batch).item())
e1.record()
torch.cuda.synchronize()
out = {"config": which, "seq_len": S, "images": n_img, "layers": model.config.num_hidden_layers, "hidden": model.config.hidden_size,
       "gradient_checkpointing": recompute, "trainable_params": trainer.arena.numel, "ms_per_step": e0.elapsed_time(e1) / 3,
       "tokens_per_s": S / (e0.elapsed_time(e1) / 3e3), "peak_memory_gb": torch.cuda.max_memory_allocated() / 1e9,
       "reserved_gb": torch.cuda.memory_reserved() / 1e9, "losses": losses}
print(json.dumps(out), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"r2_{which}_train_{'recompute' if recompute else 'full'}.json"), "w") as f:
    json.dump(out, f, indent=1)
