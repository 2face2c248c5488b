"""One eager training step at the bench shape (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from merlin_b200.mmgpt.engine.train.trainer import MMGPTTrainer, TrainingArguments

dev = torch.device("cuda", 0)
model = bench.build_model(dev)
for p in model.parameters():
    p.requires_grad_(True)
if os.environ.get("TRAIN_TOWER", "0") != "1":
    for p in model.get_model().vision_tower.parameters():
        p.requires_grad_(False)
ids_h, img_h = bench.make_inputs(0)
labels = ids_h.clone(); labels[:, :2048] = -100
trainer = MMGPTTrainer(model, TrainingArguments(learning_rate=1e-5, max_steps=100))
batch = {"input_ids": ids_h, "labels": labels, "attention_mask": None, "images": [img_h.to(dev, torch.bfloat16)]}
n = int(os.environ.get("N_STEPS", "1"))
for _ in range(n):
    loss = trainer.training_step(batch)
torch.cuda.synchronize()
print("loss", loss.item())
