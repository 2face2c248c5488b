#!/usr/bin/env python
"""ONE training step of bench.py's `detail.train` workload (Merlin-7B, 5 frames, S=4096) on one GPU — the
target of the ncu captures of the training kernels (profiles/r2_train_*):
    ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -c 1 -o out python scripts/train_one_step.py
MERLIN_STEPS=n runs n steps (default 1, no warm-up: every launch of the step is a first launch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from merlin_b200.mmgpt.engine.train.trainer import MMGPTTrainer, TrainingArguments  # noqa: E402

torch.cuda.set_device(0)
device = torch.device("cuda", 0)
model = bench.build_model(device)
for p in model.parameters():
    p.requires_grad_(True)
for p in model.get_model().vision_tower.parameters():
    p.requires_grad_(False)
ids_h, img_h = bench.make_inputs(0)
img_d = img_h.to(device=device, dtype=torch.bfloat16)
trainer = MMGPTTrainer(model, TrainingArguments(learning_rate=1e-5, max_steps=1000))
labels = ids_h.clone()
labels[:, : bench.SEQ_LEN // 2] = -100
batch = {"input_ids": ids_h, "labels": labels, "attention_mask": None, "images": [img_d]}
n_steps = int(os.environ.get("MERLIN_STEPS", "1"))
timed = os.environ.get("MERLIN_TIME") == "1" and n_steps > 3      # MERLIN_TIME=1: 3 warm-up steps, the rest timed
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n_steps):
    if timed and i == 3:
        torch.cuda.synchronize()
        e0.record()
    loss = trainer.training_step(batch)
e1.record()
torch.cuda.synchronize()
if timed:
    print("train ms/step", e0.elapsed_time(e1) / (n_steps - 3))
print("loss", float(loss), "peak GB", torch.cuda.max_memory_allocated() / 1e9)
