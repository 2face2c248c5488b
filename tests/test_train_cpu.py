"""CPU: host logic of the training engine — LLRD grouping, arena layout, cosine schedule, and the
data-parallel bucket reducer exercised with a world_size-2 gloo process group."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from merlin_b200.mmgpt.engine.train.arena import ParamArena
from merlin_b200.mmgpt.engine.train.dp import GradBucketReducer
from merlin_b200.mmgpt.engine.train.trainer import cosine_lr
from merlin_b200.mmgpt.utils import llrd_utils
from oracle import merlin_oracle as O


def test_llrd_rules_match_oracle_restatement():
    names = [("model.layers.3.self_attn.q_proj.weight", (8, 8)), ("model.norm.weight", (8,)),
             ("model.vision_tower.vision_tower.vision_model.encoder.layers.22.mlp.fc1.weight", (8, 8)),
             ("model.vision_tower.vision_tower.vision_model.encoder.layers.5.mlp.fc1.bias", (8,)),
             ("model.vision_tower.vision_tower.vision_model.pre_layrnorm.weight", (8,)),
             ("model.projector.projector.bias", (8,)), ("lm_head.weight", (8, 8))]
    mine = llrd_utils.get_param_groups(names, None, llrd_utils.vit_lr_scale_func, 1e-3, 0.05)
    ref = O.param_groups(names, 1e-3, 0.05, llrd=True)
    assert [(g["params"], g["weight_decay"], g["lr"]) for g in mine] == \
           [(g["names"], g["weight_decay"], g["lr"]) for g in ref]
    assert llrd_utils.param_hyper(names[3][0], names[3][1], llrd_utils.vit_lr_scale_func, 0.05) == \
           (pytest.approx(0.9 ** 17), 0.0)
    assert llrd_utils.llm_lr_scale_func("model.layers.31.mlp.up_proj.weight") == 1.0


def test_cosine_schedule_matches_hf_formula():
    from transformers.optimization import get_cosine_schedule_with_warmup
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=2e-5)
    sched = get_cosine_schedule_with_warmup(opt, num_warmup_steps=math.ceil(200 * 0.05), num_training_steps=200)
    for step in range(1, 200):
        opt.step()
        sched.step()
        assert cosine_lr(step, 2e-5, 200, 0.05) == pytest.approx(sched.get_last_lr()[0], rel=1e-6, abs=1e-12)


def test_arena_layout_and_views():
    ps = [("a.weight", torch.nn.Parameter(torch.arange(10000, dtype=torch.float32))),
          ("b.bias", torch.nn.Parameter(torch.ones(5))),
          ("c.weight", torch.nn.Parameter(torch.full((3, 7), 2.0)))]
    ar = ParamArena(ps, lambda n, s: (0.5 if n.startswith("c") else 1.0, 0.0 if n.endswith("bias") else 0.05),
                    chunk=8192)
    assert ar.numel == 8192 * 4 and ar.offsets["b.bias"][0] == 16384 and ar.offsets["c.weight"][0] == 24576
    assert ar.lr_mult.tolist() == [1.0, 1.0, 1.0, 0.5] and ar.weight_decay.tolist() == pytest.approx([0.05, 0.05, 0.0, 0.05])
    assert ps[0][1].data.data_ptr() == ar.params.data_ptr()            # parameters are views of the arena
    assert torch.equal(ps[2][1].data, torch.full((3, 7), 2.0))
    ar.grad("c.weight").fill_(3.0)
    assert ar.grads[24576:24576 + 21].eq(3.0).all() and ar.span(["b.bias", "c.weight"]) == (16384, 32768)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        n = 10_000
        local = torch.randn(n)
        flat = (local / world).clone()                       # loss pre-scaled by 1/world
        red = GradBucketReducer(flat, bucket_elems=3000)
        assert len(red.buckets) == 4 and red.buckets[-1] == (9000, 10000)
        red.start_step()
        red.ready(2999)                                      # nothing complete yet
        assert red.next_bucket == 0
        red.ready(6100)                                      # buckets 0 and 1 are final -> launched early
        assert red.next_bucket == 2
        red.finish()
        gathered = [torch.zeros(n) for _ in range(world)]
        dist.all_gather(gathered, local)
        want = torch.stack(gathered).mean(0)
        ok = torch.allclose(flat, want, atol=1e-6) and red.bytes_reduced == n * 4
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}
