#!/usr/bin/env python
"""Timeline of ONE data-parallel training step (bench.py's `detail.train` workload: Merlin-7B, 5 frames, S=4096):
where the gradient all-reduce kernels run relative to the backward kernels.

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
             --master-port 29511 scripts/train_timeline.py [--out gpurun_out/r2_train_timeline]
Rank 0 profiles one step with torch.profiler (CUPTI kernel records — the kernels of libmerlin_b200.so and NCCL's
alike) and writes  <out>.json (summary), <out>_kernels.csv (every kernel: stream, start µs, duration µs, name)
and <out>.trace.json.gz (chrome trace).  It also times the step with the exchange disabled in the same
process, so `exposed_comm_ms` = step − step-without-exchange on the same box.
Numbers measured under the profiler are never reported as bench values; the un-profiled step times are CUDA-event
times taken before the profiler is attached.
"""
from __future__ import annotations

import argparse
import csv
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def intervals_union(iv):
    iv = sorted(iv)
    out = []
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return out


def overlap(a, b):
    """Total length of the intersection of two unions of intervals."""
    i = j = 0
    tot = 0.0
    while i < len(a) and j < len(b):
        lo, hi = max(a[i][0], b[j][0]), min(a[i][1], b[j][1])
        if hi > lo:
            tot += hi - lo
        if a[i][1] < b[j][1]:
            i += 1
        else:
            j += 1
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/r2_train_timeline")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import bench
    from merlin_b200.mmgpt.engine.train.trainer import MMGPTTrainer, TrainingArguments
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    model = bench.build_model(device)
    for p in model.parameters():
        p.requires_grad_(True)
    for p in model.get_model().vision_tower.parameters():
        p.requires_grad_(False)
    ids_h, img_h = bench.make_inputs(rank)
    img_d = img_h.to(device=device, dtype=torch.bfloat16)
    trainer = MMGPTTrainer(model, TrainingArguments(learning_rate=1e-5, max_steps=1000))
    labels = ids_h.clone()
    labels[:, : bench.SEQ_LEN // 2] = -100
    batch = {"input_ids": ids_h, "labels": labels, "attention_mask": None, "images": [img_d]}

    def timed(n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            trainer.training_step(batch)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for _ in range(2):
        trainer.training_step(batch)
    step_ms = timed(args.steps)
    # the same step with the exchange switched off (every rank keeps its own gradients): the compute-only time
    exchange = trainer.reducer.exchange_enabled
    trainer.reducer.exchange_enabled = False
    trainer.training_step(batch)
    step_noex_ms = timed(args.steps)
    trainer.reducer.exchange_enabled = exchange

    from torch.profiler import ProfilerActivity, profile
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        trainer.training_step(batch)
        torch.cuda.synchronize()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        trace = args.out + ".trace.json"
        prof.export_chrome_trace(trace)
        ev = json.load(open(trace))["traceEvents"]
        ks = [e for e in ev if e.get("cat") == "kernel"]
        t0 = min(e["ts"] for e in ks)
        rows = sorted(((e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"]) for e in ks))
        with open(args.out + "_kernels.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["start_us", "dur_us", "stream", "name"])
            for r in rows:
                w.writerow([f"{r[0]:.1f}", f"{r[1]:.1f}", r[2], r[3][:120]])
        nccl = [(s, s + d) for s, d, _, n in rows if "nccl" in n.lower()]
        comp = [(s, s + d) for s, d, _, n in rows if "nccl" not in n.lower()]
        un, uc = intervals_union(nccl), intervals_union(comp)
        span = max(s + d for s, d, _, _ in rows)
        nccl_ms = sum(b - a for a, b in un) / 1e3
        ov_ms = overlap(un, uc) / 1e3
        by = {}
        for s, d, _, n in rows:
            key = n.split("<")[0].split("(")[0][:60]
            by[key] = by.get(key, 0.0) + d / 1e3
        top = sorted(by.items(), key=lambda kv: -kv[1])[:14]
        summary = {"world": world, "step_ms": step_ms, "step_without_exchange_ms": step_noex_ms,
                   "exposed_comm_ms": step_ms - step_noex_ms, "profiled_step_span_ms": span / 1e3,
                   "nccl_kernels": len(nccl), "nccl_busy_ms": nccl_ms, "nccl_overlapped_with_compute_ms": ov_ms,
                   "nccl_first_start_ms": (un[0][0] / 1e3 if un else None), "nccl_last_end_ms": (un[-1][1] / 1e3 if un else None),
                   "compute_busy_ms": sum(b - a for a, b in uc) / 1e3, "top_kernels_ms": top,
                   "env": {k: v for k, v in os.environ.items() if k.startswith(("NCCL_", "MERLIN_"))}}
        json.dump(summary, open(args.out + ".json", "w"), indent=1)
        print(json.dumps(summary), flush=True)
        with open(trace, "rb") as f, gzip.open(trace + ".gz", "wb") as g:
            g.write(f.read())
        os.remove(trace)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
