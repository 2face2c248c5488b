"""MMGPTTrainer on B200 kernels — the data-parallel training step loop.

Replaces the reference's `MMGPTTrainer(transformers.Trainer)` (mmgpt/engine/train/trainer.py:13-74) and
the HF `Trainer._inner_training_loop` + FSDP + autograd machinery underneath it:

  * `create_optimizer()`  — LLRD AdamW param groups (trainer.py:45-74 -> llrd_utils.get_param_groups) laid
    over a flat parameter arena; the update is ONE fused kernel (merlin_adamw_step) with per-chunk
    lr-multiplier / weight-decay tables, bf16 weights and state like the reference (builder.py:110-114,163).
  * `training_step(batch)` — explicit forward + backward of MMGPTLlamaForCausalLM.forward
    (llama_mmgpt.py:53-112) as a fixed schedule of kernel launches (no autograd graph): per decoder layer
    7 forward and 13 backward kernels, activations kept in HBM (17 GB at S=4096 for the 7B model — the reference
    needs gradient checkpointing for that, pretrain.sh:35).  `gradient_checkpointing=True` is that policy here:
    only each layer's input is kept (L x [tokens, D]) and the layer forward is re-run into ONE shared activation
    block right before its backward — what makes Merlin-13B at S=8192 (BASELINE config C5) fit next to 104 GB of
    weights + gradients + AdamW moments.
  * `save_checkpoint` / `resume_from_checkpoint` — `checkpoint-<step>` directories holding the weights (reference
    state-dict names), the AdamW moments, the step counters and the RNG-free schedule state; `train()` resumes
    from the newest one automatically like mmgpt/engine/train/train.py:56-59.
  * gradients are exchanged with bucketed NCCL all-reduces over slices of the flat gradient arena, launched
    while backward is still running (engine/train/dp.py); north_star: data-parallel only.
  * cosine schedule with warm-up ratio, gradient accumulation, global-norm clipping (HF defaults used by
    pretrain.sh: warmup_ratio 0.01, max_grad_norm 1.0, adam betas (0.9, 0.95) — pretrain.sh:23-31).

Everything on the path trains: language model, lm_head, embeddings, projector and (unless
`freeze_vision_tower`) the CLIP tower through its own explicit backward (clip_encoder.CLIPVisionTower.backward).
"""
from __future__ import annotations

import glob
import json
import math
import os
import re
import shutil
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from .... import ops
from .... import lib as _lib
from ....lib import EPI_RESIDUAL
from ...utils.constants import IGNORE_INDEX
from ...utils.llrd_utils import llm_lr_scale_func, param_hyper, vit_lr_scale_func
from .arena import ParamArena
from .dp import GradBucketReducer


@dataclass
class TrainingArguments:
    """The subset of the reference's TrainingArguments the step loop consumes (arguments.py:39-58 + the
    transformers.TrainingArguments fields set in pretrain.sh)."""
    learning_rate: float = 2e-5
    weight_decay: float = 0.05
    adam_beta1: float = 0.9
    adam_beta2: float = 0.95
    adam_epsilon: float = 1e-8
    max_grad_norm: float = 1.0
    warmup_ratio: float = 0.01
    max_steps: int = 1000
    gradient_accumulation_steps: int = 1
    llrd: bool = False
    llm_llrd: bool = False
    bucket_mb: int = 256
    gradient_checkpointing: bool = False   # pretrain.sh:35: keep only each layer's input, re-run the layer in backward
    output_dir: Optional[str] = None       # checkpoint-<step> directories (pretrain.sh:19-22: save_steps / save_total_limit)
    save_steps: int = 0                    # 0 = never
    save_total_limit: int = 2


def cosine_lr(step: int, base_lr: float, max_steps: int, warmup_ratio: float) -> float:
    """transformers.get_cosine_schedule_with_warmup (lr_scheduler_type cosine, pretrain.sh:27)."""
    warm = math.ceil(max_steps * warmup_ratio)
    if step < warm:
        return base_lr * step / max(1, warm)
    progress = (step - warm) / max(1, max_steps - warm)
    return base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * progress)))


class MMGPTTrainer:
    def __init__(self, model, args: TrainingArguments):
        self.model, self.args = model, args
        self.arena: Optional[ParamArena] = None
        self.reducer: Optional[GradBucketReducer] = None
        self.global_step = 0
        self.micro_step = 0
        self._f32: Dict[str, torch.Tensor] = {}
        self.create_optimizer()

    # ------------------------------------------------------------------ trainer.py:45-74
    def create_optimizer(self):
        a = self.args
        scale = vit_lr_scale_func if a.llrd else (llm_lr_scale_func if a.llm_llrd else None)
        named = self.model.named_parameters_reference(trainable_only=True)
        self.arena = ParamArena(named, lambda n, s: param_hyper(n, s, scale, a.weight_decay))
        self.names = [n for n, _ in named]
        self.reducer = GradBucketReducer(self.arena.grads, bucket_elems=a.bucket_mb * 1024 * 1024 // 2)
        self.reducer.adopt_registered_arena(self.arena)      # optional NCCL user-buffer registration of the gradients
        dev = self.arena.params.device
        self._sumsq = torch.zeros(4, dtype=torch.float32, device=dev)  # [0] = Σg², [2:4] = fixed-point accumulator
        # fp32 accumulators for 1-D gradients (norm weights, projector bias)
        for n, p in named:
            if p.dim() == 1 or n.endswith("position_embedding.weight"):
                self._f32[n] = torch.zeros(p.numel(), dtype=torch.float32, device=dev)
        return self.arena

    # ------------------------------------------------------------------ forward + backward
    def _forward_backward(self, batch, accumulate: bool, grad_scale: float):
        model = self.model
        m, cfg, g = model.model, model.config, self.arena.grad
        D, H, L = cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers
        dev = model.device
        input_ids, labels, images = batch["input_ids"], batch["labels"], batch["images"]
        am = batch.get("attention_mask")
        B, S = input_ids.shape
        M = B * S
        seqlens = None
        if am is not None:
            lens = am.sum(-1)
            if bool((lens != S).any()):
                seqlens = lens.to(device=dev, dtype=torch.int32)

        # ---- forward: vision (frozen tower) -> projector -> splice
        tower, proj = m.vision_tower, m.projector
        concat = torch.cat([im for im in images], dim=0)
        n_img = concat.shape[0]
        train_tower = any(n.startswith(tower.VM) for n in self.arena.offsets)
        if train_tower:
            hidden_v, tower_saved = tower.encode_train(concat)
        else:
            hidden_v = tower.encode(concat)
        grid = tower.config.image_size // tower.config.patch_size
        conv = hasattr(proj, "project") and tower.select_feature == "patch"
        if conv:
            col = ops.conv3x3_im2col(hidden_v, n_img, grid, proj.in_channels, proj.conv_stride, tower.tokens, 1)
        else:
            col = tower.feature_select(hidden_v, n_img).reshape(-1, hidden_v.shape[-1]).contiguous()
        feats = ops.gemm(col, proj.weight, bias=proj.bias)
        per_image = feats.shape[0] // n_img
        ids_host = input_ids if not input_ids.is_cuda else input_ids.cpu()
        model._validate_image_grammar(ids_host, images, per_image)
        rows = [im.shape[0] * per_image for im in images]
        offs = [sum(rows[:b]) for b in range(B)]
        feat_offset = torch.tensor(offs, dtype=torch.int32).to(dev)
        feat_limit = torch.tensor(rows[:B], dtype=torch.int32).to(dev)
        ids_dev = input_ids.to(dev).contiguous()
        labels_dev = labels.to(dev).contiguous()
        x, _, rank = ops.embed_splice(ids_dev, m.embed_tokens, feats.contiguous(), model.im_patch_token, feat_offset,
                                      feat_limit, return_rank=True)

        cos, sin = m.rope_tables(S)
        # one library call per layer (merlin_decoder_layer_forward_train); activations live in persistent
        # per-layer blocks whose first slot is x_in, so layer i writes its output straight into block i+1
        bufs = self._layer_buffers(B, S)
        x_final = bufs["x_final"]
        recompute = self.args.gradient_checkpointing
        if recompute:
            # forward keeps x_in of every layer only: layer i reads xin[i] and writes xin[i + 1]
            for i in range(L):
                x_dst = bufs["xin"][i + 1] if i + 1 < L else x_final
                if i == 0:
                    bufs["xin"][0].copy_(x)
                ops.decoder_layer_forward_train(self._layer_weights(i), bufs["xin"][i], x_dst, bufs["saved"][0], cos, sin,
                                                seqlens, B, S, D, H, cfg.intermediate_size, cfg.rms_norm_eps)
        else:
            for i in range(L):
                x_src = x if i == 0 else bufs["saved"][i]
                x_dst = bufs["saved"][i + 1] if i + 1 < L else x_final
                ops.decoder_layer_forward_train(self._layer_weights(i), x_src, x_dst, bufs["saved"][i], cos, sin, seqlens,
                                                B, S, D, H, cfg.intermediate_size, cfg.rms_norm_eps)
        hn = ops.rmsnorm(x_final, m.norm_w, cfg.rms_norm_eps)
        # lm_head + shifted CE, forward and backward, one library call each (merlin_lmhead_ce_forward / _backward)
        logits, loss, loss_acc = ops.lmhead_ce_forward(hn, model.lm_head_w, labels_dev, B, S, IGNORE_INDEX)

        # ---- backward
        acc = accumulate
        dhn = ops.lmhead_ce_backward(logits, labels_dev, loss_acc, grad_scale, hn, model.lm_head_w,
                                     g("lm_head.weight"), acc, B, S, IGNORE_INDEX)
        dx = ops.rmsnorm_bwd(x_final, m.norm_w, dhn, None, self._f32["model.norm.weight"], cfg.rms_norm_eps)
        self._ready_after("model.norm.weight")
        for i in reversed(range(L)):
            p = f"model.layers.{i}."
            grads = _lib.LayerGrads(g(p + "self_attn.qkv_proj.weight").data_ptr(), g(p + "self_attn.o_proj.weight").data_ptr(),
                                    g(p + "mlp.gate_up_proj.weight").data_ptr(), g(p + "mlp.down_proj.weight").data_ptr(),
                                    self._f32[p + "input_layernorm.weight"].data_ptr(),
                                    self._f32[p + "post_attention_layernorm.weight"].data_ptr())
            if recompute:
                if i != L - 1:   # the shared block still holds layer L-1's activations right after the forward
                    ops.decoder_layer_forward_train(self._layer_weights(i), bufs["xin"][i], bufs["scratch"],
                                                    bufs["saved"][0], cos, sin, seqlens, B, S, D, H,
                                                    cfg.intermediate_size, cfg.rms_norm_eps)
                block = bufs["saved"][0]
            else:
                block = bufs["saved"][i]
            ops.decoder_layer_backward(self._layer_weights(i), grads, block, dx, bufs["ws"], cos, sin,
                                       seqlens, B, S, D, H, cfg.intermediate_size, cfg.rms_norm_eps, acc)
            self._ready_after(p + "input_layernorm.weight")
        # embeddings + projector
        dtable = g("model.embed_tokens.weight") if "model.embed_tokens.weight" in self.arena.offsets else None
        train_proj = "model.projector.projector.weight" in self.arena.offsets
        dfeats = torch.zeros_like(feats) if (train_proj or train_tower) else None
        if dtable is not None and not acc:
            dtable.zero_()
        ops.embed_splice_bwd(ids_dev.view(-1), rank, dx, dtable, dfeats, cfg.vocab_size)
        if train_proj:
            ops.gemm_tn(dfeats.contiguous(), col, out=g("model.projector.projector.weight"), accumulate=acc)
            ops.colsum_accumulate(dfeats, self._f32["model.projector.projector.bias"])
            self._ready_after("model.projector.projector.bias")
        if train_tower:
            # projector input gradient -> gradient of the selected ViT hidden state (CLS rows get none)
            dhid = torch.zeros_like(hidden_v)
            if conv:
                dcol = ops.gemm_nt(dfeats.contiguous(), proj.weight)
                ops.conv3x3_col2im(dcol, dhid, n_img, grid, proj.in_channels, proj.conv_stride, tower.tokens, 1)
            else:
                dpatch = ops.gemm_nt(dfeats.contiguous(), proj.weight)
                dhid.view(n_img, tower.tokens, -1)[:, 1:].copy_(dpatch.view(n_img, tower.tokens - 1, -1))
            tower.backward(dhid, tower_saved, g, self._f32, acc, done=self._ready_after)
        self._ready_after(self.names[-1])
        return loss

    def _layer_weights(self, i: int):
        m = self.model.model
        return _lib.LayerWeights(m.qkv_w[i].data_ptr(), m.o_w[i].data_ptr(), m.gate_up_w[i].data_ptr(),
                                 m.down_w[i].data_ptr(), m.ln1_w[i].data_ptr(), m.ln2_w[i].data_ptr())

    def _layer_buffers(self, B: int, S: int):
        """Persistent activation blocks (one per decoder layer) + backward workspace for this batch shape."""
        key = (B, S, self.args.gradient_checkpointing)
        if getattr(self, "_buf_key", None) != key:
            self._bufs = None                      # release the old shape's blocks before allocating the new ones
            cfg = self.model.config
            D, FF, H, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads, cfg.num_hidden_layers
            L_ = _lib.load()
            dev = self.model.device
            nbytes = L_.merlin_decoder_layer_saved_bytes(B * S, D, FF, H, B, S)
            n_blocks = 1 if self.args.gradient_checkpointing else L
            self._bufs = dict(
                saved=[torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(n_blocks)],
                ws=torch.empty(L_.merlin_decoder_layer_backward_workspace_bytes(B * S, D, FF, H, B, S), dtype=torch.uint8,
                               device=dev),
                x_final=torch.empty((B * S, D), dtype=torch.bfloat16, device=dev))
            if self.args.gradient_checkpointing:
                self._bufs["xin"] = [torch.empty((B * S, D), dtype=torch.bfloat16, device=dev) for _ in range(L)]
                self._bufs["scratch"] = torch.empty((B * S, D), dtype=torch.bfloat16, device=dev)
            self._buf_key = key
        return self._bufs

    def _ready_after(self, name: str):
        """Gradients of every parameter up to and including `name` (arena order = backward completion
        order) are final: on the last micro-step fold their fp32 side accumulators into the bf16 arena
        and hand the finished range to the bucket reducer, which overlaps the all-reduce with the rest
        of backward."""
        if not self._last_micro:
            return
        idx = self.names.index(name)
        for k in range(self._fold_ptr, idx + 1):
            n = self.names[k]
            if n in self._f32:
                ops.f32_to_bf16(self._f32[n], self.arena.grad(n))
                self._f32[n].zero_()
        self._fold_ptr = max(self._fold_ptr, idx + 1)
        _, end = self.arena.span([name])
        self.reducer.ready(end)

    # ------------------------------------------------------------------ public step API
    def training_step(self, batch) -> torch.Tensor:
        """One micro-batch forward/backward; on the last micro-step of an accumulation window also the
        gradient all-reduce, clipping and the AdamW update.  Returns the (detached) loss."""
        a = self.args
        accumulate = self.micro_step % a.gradient_accumulation_steps != 0
        self._last_micro = (self.micro_step + 1) % a.gradient_accumulation_steps == 0
        world = self.reducer.world
        grad_scale = 1.0 / (a.gradient_accumulation_steps * world)
        if not accumulate:
            self.reducer.start_step()
        self._fold_ptr = 0
        with torch.no_grad():
            loss = self._forward_backward(batch, accumulate, grad_scale)
        self.micro_step += 1
        if self._last_micro:
            self.reducer.finish()
            self.optimizer_step()
        return loss

    def optimizer_step(self):
        a, ar = self.args, self.arena
        # HF Trainer steps the scheduler AFTER optimizer.step(): update k (1-based) runs at lr(k - 1), so the first
        # update of a warm-up schedule has lr 0 (it still moves the AdamW moments, as torch.optim.AdamW does)
        lr = cosine_lr(self.global_step, a.learning_rate, a.max_steps, a.warmup_ratio)
        self.global_step += 1
        ops.grad_sumsq(ar.grads, self._sumsq)
        ops.adamw_step(ar.params, ar.grads, ar.exp_avg, ar.exp_avg_sq, ar.lr_mult, ar.weight_decay, ar.chunk, lr,
                       a.adam_beta1, a.adam_beta2, a.adam_epsilon, self.global_step, 1.0, a.max_grad_norm, self._sumsq)
        self.last_lr = lr

    def grad_norm(self) -> float:
        return float(self._sumsq[0].sqrt())

    # ------------------------------------------------------------------ checkpoint / resume
    # Reference: HF Trainer `save_strategy steps` + `save_total_limit` (pretrain.sh:19-22) and the auto-resume of
    # mmgpt/engine/train/train.py:56-59 (`resume_from_checkpoint=True` when output_dir holds checkpoint-*).
    @staticmethod
    def checkpoints(output_dir: str) -> List[str]:
        """checkpoint-<step> directories under output_dir, oldest first."""
        found = [(int(m.group(1)), d) for d in glob.glob(os.path.join(output_dir, "checkpoint-*"))
                 if (m := re.search(r"checkpoint-(\d+)$", d)) and os.path.isdir(d)]
        return [d for _, d in sorted(found)]

    def _layout(self):
        return [[n, *self.arena.offsets[n][:2]] for n in self.arena.names]

    def save_checkpoint(self, output_dir: Optional[str] = None) -> Optional[str]:
        """Weights (reference state-dict names, `pytorch_model.bin` + config.json), AdamW moments (flat bf16 arenas +
        the layout they were saved with) and the step counters -> output_dir/checkpoint-<global_step>.  Rank 0
        writes (replicas are bit-identical); older checkpoints beyond save_total_limit are removed."""
        output_dir = output_dir or self.args.output_dir
        if not output_dir:
            raise ValueError("save_checkpoint: no output_dir")
        rank = int(os.environ.get("RANK", "0"))
        path = os.path.join(output_dir, f"checkpoint-{self.global_step}")
        if rank == 0:
            tmp = path + ".tmp"
            shutil.rmtree(tmp, ignore_errors=True)
            os.makedirs(tmp)
            self.model.save_pretrained(tmp)
            torch.save({"exp_avg": self.arena.exp_avg.cpu(), "exp_avg_sq": self.arena.exp_avg_sq.cpu(),
                        "layout": self._layout(), "chunk": self.arena.chunk}, os.path.join(tmp, "optimizer.pt"))
            a = self.args
            json.dump({"global_step": self.global_step, "micro_step": self.micro_step,
                       "last_lr": getattr(self, "last_lr", None),
                       "args": {k: getattr(a, k) for k in ("learning_rate", "weight_decay", "adam_beta1", "adam_beta2",
                                                          "adam_epsilon", "max_grad_norm", "warmup_ratio", "max_steps",
                                                          "gradient_accumulation_steps", "llrd", "llm_llrd")}},
                      open(os.path.join(tmp, "trainer_state.json"), "w"), indent=1)
            shutil.rmtree(path, ignore_errors=True)
            os.replace(tmp, path)                       # a checkpoint directory appears only when it is complete
            limit = self.args.save_total_limit
            if limit and limit > 0:
                for old in self.checkpoints(output_dir)[:-limit]:
                    shutil.rmtree(old, ignore_errors=True)
        if self.reducer.world > 1:
            import torch.distributed as dist
            dist.barrier()
        return path

    def resume_from_checkpoint(self, checkpoint=True) -> Optional[str]:
        """checkpoint: a checkpoint-<step> directory, or True = the newest one under args.output_dir (None if there
        is none).  Restores weights into the parameter arena, the AdamW moments and the counters; the LR schedule
        is a pure function of global_step, so nothing else needs saving."""
        if checkpoint is True:
            found = self.checkpoints(self.args.output_dir) if self.args.output_dir else []
            if not found:
                return None
            checkpoint = found[-1]
        state = json.load(open(os.path.join(checkpoint, "trainer_state.json")))
        opt = torch.load(os.path.join(checkpoint, "optimizer.pt"), map_location="cpu")
        if opt["layout"] != self._layout() or opt["chunk"] != self.arena.chunk:
            raise RuntimeError("resume_from_checkpoint: the optimizer state was saved for another set / order of "
                               "trainable parameters")
        sd = self.model._read_checkpoint(checkpoint)
        self.model.load_state_dict(sd, strict=False)    # parameters are views of the arena: this fills it
        self.arena.exp_avg.copy_(opt["exp_avg"])
        self.arena.exp_avg_sq.copy_(opt["exp_avg_sq"])
        self.global_step, self.micro_step = int(state["global_step"]), int(state["micro_step"])
        if state.get("last_lr") is not None:
            self.last_lr = state["last_lr"]
        return checkpoint

    def train(self, batches, log_every: int = 1):
        """Minimal loop with the reference's logging cadence (--logging_steps 1): loss, lr, tokens/s."""
        import time
        a = self.args
        for step, batch in enumerate(batches):
            if a.max_steps and self.global_step >= a.max_steps:
                break
            t0 = time.perf_counter()
            before = self.global_step
            loss = self.training_step(batch)
            if a.save_steps and a.output_dir and self.global_step != before and self.global_step % a.save_steps == 0:
                self.save_checkpoint()
            if (step + 1) % log_every == 0:
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                toks = batch["input_ids"].numel()
                print(f"step {self.global_step} loss {loss.item():.4f} lr {getattr(self, 'last_lr', 0):.3e} "
                      f"tokens/s {toks / dt:.0f}", flush=True)
