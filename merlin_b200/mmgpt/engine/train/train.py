"""train() — mirror of mmgpt/engine/train/train.py:32-77 on the B200 step loop.

Builds model + tokenizer with `build_model_tokenizer`, wraps them in `MMGPTTrainer` and runs the data-parallel
loop over batches shaped like the reference collator's output (mmgpt/data/collator.py:12-34).  The data
layer itself (webdataset / S3 readers) is out of scope: `batches` is any iterable of
collator-format dicts; `synthetic_batches` yields the bench's synthetic interpair samples.
Launch with torchrun (one process per GPU), as pretrain.sh:2 does.
"""
from __future__ import annotations

import os

import torch

from ...model.builder import build_model_tokenizer
from .trainer import MMGPTTrainer, TrainingArguments


def init_distributed():
    """torchrun environment -> NCCL process group over NVLink (replaces dist_utils.py:68-84)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return local, world


def train(model_args, data_args, training_args, batches, step_args: TrainingArguments = None, tokenizer=None):
    local, world = init_distributed()
    training_args.device = torch.device("cuda", local)
    model, tokenizer, data_args = build_model_tokenizer(model_args, data_args, training_args, tokenizer=tokenizer)
    trainer = MMGPTTrainer(model, step_args or TrainingArguments(
        learning_rate=getattr(training_args, "learning_rate", 2e-5),
        weight_decay=getattr(training_args, "weight_decay", 0.0),
        max_steps=getattr(training_args, "max_steps", 1000),
        warmup_ratio=getattr(training_args, "warmup_ratio", 0.01),
        gradient_accumulation_steps=getattr(training_args, "gradient_accumulation_steps", 1),
        llrd=getattr(training_args, "llrd", False), llm_llrd=getattr(training_args, "llm_llrd", False),
        gradient_checkpointing=getattr(training_args, "gradient_checkpointing", False),
        output_dir=getattr(training_args, "output_dir", None), save_steps=getattr(training_args, "save_steps", 0) or 0,
        save_total_limit=getattr(training_args, "save_total_limit", 2) or 0))
    # train.py:56-59: trainer.train(resume_from_checkpoint=True) when output_dir already holds checkpoint-*
    if trainer.args.output_dir:
        resumed = trainer.resume_from_checkpoint(True)
        if resumed and int(os.environ.get("RANK", "0")) == 0:
            print(f"resumed from {resumed} at step {trainer.global_step}", flush=True)
    trainer.train(batches)
    out = getattr(training_args, "output_dir", None)
    if out and int(os.environ.get("RANK", "0")) == 0:
        model.save_pretrained(out)       # trainer._safe_save (trainer.py:29-43): full state dict on rank 0
    return trainer
