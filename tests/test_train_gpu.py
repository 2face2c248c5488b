"""GPU: the explicit forward/backward schedule of MMGPTTrainer against autograd through the fp32 oracle
(the oracle's forward is pinned to the real reference, so its autograd gradients are the reference's),
plus one optimizer step and a short loss-goes-down run.  Tolerance: cosine >= 0.999 and
max|Δ| <= 5% of max|ref| per parameter tensor (bf16 activations and gradients vs fp32)."""
import os

import pytest
import torch

from helpers import build_product_model, load_golden
from merlin_b200.mmgpt.engine.train.trainer import MMGPTTrainer, TrainingArguments
from oracle import merlin_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def oracle_grads(cfg, sd, batch):
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.forward(leaves, cfg, batch["input_ids"], batch["attention_mask"], batch["labels"], batch["images"])
    out["loss"].backward()
    return out["loss"].detach(), {k: v.grad for k, v in leaves.items() if v.grad is not None}


def unpack_grads(trainer, cfg):
    """Arena gradients under the reference's parameter names."""
    ar = trainer.arena
    D, FF = cfg.hidden_size, cfg.intermediate_size
    out = {}
    for name in trainer.names:
        g = ar.grad(name).float().cpu()
        if "vision_model" in name:
            d = cfg.vit_hidden
            if name.endswith("self_attn.qkv_proj.weight") or name.endswith("self_attn.qkv_proj.bias"):
                kind = name.rsplit(".", 1)[1]
                base = name[: -len("qkv_proj." + kind)]
                for j, n in enumerate("qkv"):
                    out[base + f"{n}_proj.{kind}"] = g[j * d:(j + 1) * d]
            elif name.endswith("patch_embedding.weight"):
                out[name] = g[:, : 3 * cfg.patch_size ** 2].reshape(d, 3, cfg.patch_size, cfg.patch_size)
            else:
                out[name] = g
        elif name.endswith("self_attn.qkv_proj.weight"):
            base = name[: -len("qkv_proj.weight")]
            for j, n in enumerate("qkv"):
                out[base + f"{n}_proj.weight"] = g[j * D:(j + 1) * D]
        elif name.endswith("mlp.gate_up_proj.weight"):
            base = name[: -len("gate_up_proj.weight")]
            gu = g.view(FF // 128, 2, 128, D)
            out[base + "gate_proj.weight"] = gu[:, 0].reshape(FF, D)
            out[base + "up_proj.weight"] = gu[:, 1].reshape(FF, D)
        elif name == "model.projector.projector.weight" and cfg.projector == "conv":
            out[name] = g.view(D, 3, 3, cfg.vit_hidden).permute(0, 3, 1, 2)
        else:
            out[name] = g
    return out


@pytest.mark.parametrize("name", ["tiny_conv", "tiny_mlp", "tiny_pad"])
def test_gradients_match_oracle_autograd(name):
    g, cfg = load_golden(name)
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)
    for p in model.parameters():
        p.requires_grad_(True)
    for p in model.get_model().vision_tower.parameters():
        p.requires_grad_(False)
    trainer = MMGPTTrainer(model, TrainingArguments(max_grad_norm=0.0))
    batch = {k: g[k] for k in ("input_ids", "labels", "attention_mask")}
    batch["images"] = [im.to(DEV) for im in g["images"]]
    trainer._last_micro = True
    trainer._fold_ptr = 0
    trainer.reducer.start_step()
    with torch.no_grad():
        loss = trainer._forward_backward(batch, accumulate=False, grad_scale=1.0)
    ref_loss, ref = oracle_grads(cfg, sd, {**batch, "images": g["images"]})
    assert abs(loss.item() - ref_loss.item()) <= 2e-3
    mine = unpack_grads(trainer, cfg)
    checked = 0
    for k, gm in mine.items():
        gr = ref[k]
        if gr.abs().max() == 0:
            assert gm.abs().max() == 0
            continue
        cos = torch.nn.functional.cosine_similarity(gm.flatten(), gr.flatten(), dim=0).item()
        err = (gm - gr).abs().max().item() / gr.abs().max().item()
        assert cos >= 0.999 and err <= 0.05, f"{k}: cos={cos:.5f} rel-max-err={err:.4f}"
        checked += 1
    assert checked >= 6 * cfg.num_layers + 4


def test_training_steps_reduce_the_loss_and_respect_frozen_tower():
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)
    for p in model.parameters():
        p.requires_grad_(True)
    tower = model.get_model().vision_tower
    for p in tower.parameters():
        p.requires_grad_(False)
    before = tower.qkv_w[0].clone()
    trainer = MMGPTTrainer(model, TrainingArguments(learning_rate=2e-3, warmup_ratio=0.0, max_steps=20,
                                                   gradient_accumulation_steps=2))
    batch = {k: g[k] for k in ("input_ids", "labels", "attention_mask")}
    batch["images"] = [im.to(DEV) for im in g["images"]]
    losses = [trainer.training_step(batch).item() for _ in range(12)]
    assert trainer.global_step == 6 and trainer.micro_step == 12
    assert losses[-1] < losses[0] - 0.5, losses
    assert torch.equal(tower.qkv_w[0], before)
    assert 0 < trainer.grad_norm() < 1e3
    # the model keeps working for inference after its weights moved into the arena
    out = model(input_ids=g["input_ids"].to(DEV), images=[im.to(DEV) for im in g["images"]], labels=g["labels"].to(DEV))
    assert out.loss.item() < losses[0]


@pytest.mark.parametrize("name", ["tiny_conv", "tiny_mlp"])
def test_vision_tower_gradients_match_oracle_autograd(name):
    """freeze_vision_tower=False (the reference's LLRD recipe): gradients of every CLIP parameter that can
    receive one (layers below select_layer, pre_layrnorm, embeddings) against oracle autograd."""
    g, cfg = load_golden(name)
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)
    for p in model.parameters():
        p.requires_grad_(True)
    trainer = MMGPTTrainer(model, TrainingArguments(max_grad_norm=0.0, llrd=True))
    assert any("vision_model" in n for n in trainer.names)
    batch = {k: g[k] for k in ("input_ids", "labels", "attention_mask")}
    batch["images"] = [im.to(DEV) for im in g["images"]]
    trainer._last_micro = True
    trainer._fold_ptr = 0
    trainer.reducer.start_step()
    with torch.no_grad():
        trainer._forward_backward(batch, accumulate=False, grad_scale=1.0)
    _, ref = oracle_grads(cfg, sd, {**batch, "images": g["images"]})
    mine = unpack_grads(trainer, cfg)
    vit = [k for k in mine if "vision_model" in k]
    assert len(vit) >= 16 * (cfg.vit_layers - 1) + 5
    for k in vit:
        gm, gr = mine[k], ref[k]
        if k.endswith("k_proj.bias"):
            # softmax is invariant to a constant added to every key, so this gradient is identically zero
            # (the oracle's is ~1e-9 rounding noise); ours must be rounding noise of bf16 size
            assert gr.abs().max().item() < 1e-6 and gm.abs().max().item() < 5e-3
            continue
        cos = torch.nn.functional.cosine_similarity(gm.flatten(), gr.flatten(), dim=0).item()
        err = (gm - gr).abs().max().item() / (gr.abs().max().item() + 1e-12)
        assert cos >= 0.995 and err <= 0.08, f"{k}: cos={cos:.5f} rel-max-err={err:.4f}"
    # LLRD tables: encoder layer j is scaled by 0.9 ** (22 - j), other vision_model params by 0.1
    ar = trainer.arena
    off = ar.offsets["model.vision_tower.vision_tower.vision_model.encoder.layers.1.mlp.fc1.weight"][0]
    assert ar.lr_mult[off // ar.chunk].item() == pytest.approx(0.9 ** 21)
    off = ar.offsets["model.vision_tower.vision_tower.vision_model.pre_layrnorm.bias"][0]
    assert ar.lr_mult[off // ar.chunk].item() == pytest.approx(0.1) and ar.weight_decay[off // ar.chunk].item() == 0.0


def test_data_parallel_replicas_stay_bit_identical():
    """2-rank NCCL run (needs >= 2 GPUs; the 1-GPU driver box skips it): different batches per rank, bucketed
    all-reduce overlapped with backward, AdamW — gradients, moments and weights must be bit-identical on
    both ranks after every step (tests/dp_check.py exits non-zero otherwise)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tests", "dp_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "weights identical across ranks: True" in r.stdout


def test_local_shards_to_training_step_end_to_end(tmp_path):
    """N2 <-> a11: tar shards on disk -> build_dataloader (InterPairWebDataset + merlin_image_preprocess on the
    GPU + DataCollatorForSupervisedDataset) -> MMGPTTrainer.training_step.  The first loss equals the oracle's
    loss on the same batch (pixels restated by oracle/image_oracle.py), and training reduces it."""
    import io
    import json
    import tarfile
    import types
    import numpy as np
    from PIL import Image
    from oracle import image_oracle as IO
    from merlin_b200.mmgpt.data.builder import build_dataloader

    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)

    class Tok:   # word-level tokenizer inside the tiny vocabulary (ids 3..255), HF call signature
        eos_token, pad_token_id, model_max_length = "</s>", 0, 96
        special = {"<im_patch>": cfg.im_patch_token, "<im_start>": cfg.im_start_token, "<im_end>": cfg.im_end_token,
                   "</s>": 2, "\n": 13}

        def convert_tokens_to_ids(self, toks):
            return [self.special[t] for t in toks]

        def __call__(self, text, padding=None, max_length=None, truncation=False):
            ids, i, word = [1], 0, ""
            flush = lambda w: ids.extend(3 + (sum(map(ord, x)) % 250) for x in w.split())
            while i < len(text):
                for sp, tid in self.special.items():
                    if text.startswith(sp, i):
                        flush(word)
                        word = ""
                        ids.append(tid)
                        i += len(sp)
                        break
                else:
                    word += text[i]
                    i += 1
            flush(word)
            return types.SimpleNamespace(input_ids=ids[:max_length] if truncation and max_length else ids)

    rng = np.random.default_rng(11)
    jpgs = {}
    shard = tmp_path / "interpair-000.tar"
    with tarfile.open(shard, "w") as tf:
        def add(name, data):
            ti = tarfile.TarInfo(name)
            ti.size = len(data)
            tf.addfile(ti, io.BytesIO(data))
        for r in range(4):
            meta = dict(prompt="Track the object <image> then <image>", txt=f"it moves {['left', 'right', 'up', 'down'][r]} fast",
                        image_name_list=[f"clip{r}-f0", f"clip{r}-f1"])
            add(f"rec{r}.json", json.dumps(meta).encode())
            for j in range(2):
                buf = io.BytesIO()
                Image.fromarray(rng.integers(0, 256, (70 + 9 * r, 96, 3), dtype=np.uint8)).save(buf, format="JPEG")
                jpgs[(r, j)] = buf.getvalue()
                add(f"rec{r}.f{j}.jpg", buf.getvalue())

    data_args = types.SimpleNamespace(image_token_len=cfg.num_patches, image_aspect_ratio="resize", use_im_start_end=True,
                                      image_processor=None, box_limit=0, image_size=cfg.image_size,
                                      interpair_webdatasets=str(tmp_path / "interpair-*.tar"))
    parts = build_dataloader(Tok(), data_args)
    ds, collate = parts["train_dataset"], parts["data_collator"]
    assert len(ds) == 4
    ds.shuffle = False        # file order, so that the records can be restated below (the seeded shuffles are covered on CPU)
    batch = collate([ds[0], ds[1]])
    assert batch["input_ids"].shape[0] == 2 and [tuple(i.shape) for i in batch["images"]] == [(2, 3, 56, 56)] * 2
    assert int((batch["input_ids"] == cfg.im_patch_token).sum()) == 2 * 2 * cfg.num_patches

    # oracle loss on the same batch, pixels restated on the CPU (bit-identical to the GPU image path)
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
    ref_images = [torch.stack([torch.from_numpy(IO.image_preprocess(
        np.asarray(Image.open(io.BytesIO(jpgs[(r, j)])).convert("RGB")), cfg.image_size, mean, std)) for j in range(2)])
        for r in range(2)]
    assert all(torch.equal(a, b) for a, b in zip(batch["images"], ref_images))
    ref = O.forward(sd, cfg, batch["input_ids"], batch["attention_mask"], batch["labels"], ref_images)

    for p in model.parameters():
        p.requires_grad_(True)
    for p in model.get_model().vision_tower.parameters():
        p.requires_grad_(False)
    trainer = MMGPTTrainer(model, TrainingArguments(learning_rate=2e-3, warmup_ratio=0.0, max_steps=20))
    dev_batch = dict(batch, images=[im.to(DEV) for im in batch["images"]])
    losses = [trainer.training_step(dev_batch).item() for _ in range(8)]
    assert abs(losses[0] - ref["loss"].item()) <= 2e-3, (losses[0], ref["loss"].item())
    assert losses[-1] < losses[0] - 0.5, losses


def _tiny_trainer(**kw):
    g, cfg = load_golden("tiny_conv")
    sd = O.init_state_dict(cfg, seed=0)
    model = build_product_model(cfg, sd)
    for p in model.parameters():
        p.requires_grad_(True)
    for p in model.get_model().vision_tower.parameters():
        p.requires_grad_(False)
    trainer = MMGPTTrainer(model, TrainingArguments(learning_rate=2e-3, warmup_ratio=0.0, max_steps=20, **kw))
    batch = {k: g[k] for k in ("input_ids", "labels", "attention_mask")}
    batch["images"] = [im.to(DEV) for im in g["images"]]
    return trainer, batch


def test_activation_recompute_gives_the_same_gradients():
    """gradient_checkpointing (pretrain.sh:35): only layer inputs are kept, each layer forward is re-run before its
    backward into one shared activation block — same kernels on the same inputs, so the gradients are identical
    up to the fp32 atomics of the attention dq reduction."""
    a, batch = _tiny_trainer()
    b, _ = _tiny_trainer(gradient_checkpointing=True)
    for t in (a, b):
        t.args.max_grad_norm = 0.0
        with torch.no_grad():
            t._last_micro, t._fold_ptr = True, 0
            t.reducer.start_step()
            t.loss = t._forward_backward(batch, accumulate=False, grad_scale=1.0)
    assert abs(a.loss.item() - b.loss.item()) <= 1e-6
    assert len(b._bufs["saved"]) == 1 and len(b._bufs["xin"]) == b.model.config.num_hidden_layers
    ga, gb = a.arena.grads.float(), b.arena.grads.float()
    assert (ga - gb).abs().max().item() <= 2.0 ** -6 * ga.abs().max().item()      # a bf16 ulp or two: the dq reduction order
    la = [a.training_step(batch).item() for _ in range(4)]
    lb = [b.training_step(batch).item() for _ in range(4)]
    assert max(abs(x - y) for x, y in zip(la, lb)) <= 2e-2 and la[-1] < la[0]


def test_checkpoint_resume_restores_weights_moments_and_counters(tmp_path):
    """HF Trainer save_steps / save_total_limit + auto-resume (pretrain.sh:19-22, train.py:56-59)."""
    out = str(tmp_path / "run")
    a, batch = _tiny_trainer(output_dir=out, save_steps=2, save_total_limit=2)
    a.train([batch] * 6, log_every=100)
    assert [os.path.basename(d) for d in MMGPTTrainer.checkpoints(out)] == ["checkpoint-4", "checkpoint-6"]
    assert {"config.json", "optimizer.pt", "pytorch_model.bin", "trainer_state.json"} <= set(
        os.listdir(os.path.join(out, "checkpoint-6")))
    b, _ = _tiny_trainer(output_dir=out, save_steps=0)
    assert b.resume_from_checkpoint(True).endswith("checkpoint-6")
    assert b.global_step == 6 and b.micro_step == 6
    assert torch.equal(b.arena.params, a.arena.params)
    assert torch.equal(b.arena.exp_avg, a.arena.exp_avg) and torch.equal(b.arena.exp_avg_sq, a.arena.exp_avg_sq)
    la, lb = a.training_step(batch).item(), b.training_step(batch).item()
    assert abs(la - lb) <= 1e-3 and b.global_step == 7 and abs(a.last_lr - b.last_lr) < 1e-12
    # a trainer over a different set of trainable parameters refuses the optimizer state
    g, cfg = load_golden("tiny_conv")
    other = build_product_model(cfg, O.init_state_dict(cfg, seed=0))
    other.lm_head_w.requires_grad_(True)
    c = MMGPTTrainer(other, TrainingArguments(output_dir=out))
    with pytest.raises(RuntimeError, match="another set"):
        c.resume_from_checkpoint(True)
    empty, _ = _tiny_trainer(output_dir=str(tmp_path / "nothing"))
    assert empty.resume_from_checkpoint(True) is None
