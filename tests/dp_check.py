"""2+ rank DP check on real GPUs: each rank trains the tiny model on a different batch; after a step the
weights must be identical across ranks and equal to a single-process run on the averaged gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist
from helpers import build_product_model
from merlin_b200.mmgpt.engine.train.trainer import MMGPTTrainer, TrainingArguments
from oracle import merlin_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = O.OracleConfig.tiny()
sd = O.init_state_dict(cfg, seed=0)
model = build_product_model(cfg, sd, device=dev)
for p in model.parameters():
    p.requires_grad_(True)
trainer = MMGPTTrainer(model, TrainingArguments(learning_rate=1e-3, warmup_ratio=0.0, max_steps=10, bucket_mb=1))
b = O.synthetic_batch(cfg, 2, 48, [1, 2], seed=1234 + rank)
batch = {"input_ids": b["input_ids"], "labels": b["labels"], "attention_mask": b["attention_mask"],
         "images": [im.to(dev) for im in b["images"]]}
def diff_report(tag, flat):
    ref = flat.float().clone()
    dist.broadcast(ref, src=0)
    bad = (flat.float() != ref)
    if rank == 1:
        if bad.any():
            idx = int(bad.nonzero()[0])
            name = [n for n in trainer.names if trainer.arena.offsets[n][0] <= idx][-1]
            print(f"  [{tag}] differs in {int(bad.sum())} elements; first at {idx} ({name}); "
                  f"mine {flat[idx].item()} vs rank0 {ref[idx].item()}", flush=True)
        else:
            print(f"  [{tag}] identical", flush=True)

for it in range(3):
    loss = trainer.training_step(batch)
    torch.cuda.synchronize()
    diff_report(f"step{it} grads", trainer.arena.grads)
    diff_report(f"step{it} exp_avg", trainer.arena.exp_avg)
    diff_report(f"step{it} params", trainer.arena.params)
flat = trainer.arena.params.float()
ref = flat.clone()
dist.broadcast(ref, src=0)
same = torch.equal(flat, ref)
gathered = [torch.zeros(1, device=dev) for _ in range(world)]
dist.all_gather(gathered, loss.detach().float().view(1))
if rank == 0:
    print(f"world={world} buckets={len(trainer.reducer.buckets)} bytes_reduced/step={trainer.reducer.bytes_reduced // 3} "
          f"weights identical across ranks: {same} losses={[round(g.item(), 4) for g in gathered]}", flush=True)
ok = torch.tensor([1.0 if same else 0.0], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if ok.item() == 1.0 else 1)
