#!/usr/bin/env python
"""Beam-search decode timing at the bench shape (7B, S=4096, 5 frames): ms per decode POSITION for k beams through
merlin_decode_step_batched (graph replay) vs the greedy single-sequence decode; and generate(num_beams=k) end to end.
Writes gpurun_out/r2_beam_bench.json.   usage: python scripts/beam_bench.py [k ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch  # noqa: E402

from config_bench import build  # noqa: E402
from merlin_b200 import decode as D  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    S, n_img = 4096, 5
    model = build("7b", S, dev)
    g = torch.Generator().manual_seed(1234)
    ids = [1]
    for _ in range(n_img):
        ids += [32001] + [32000] * 144 + [32002, 13]
    ids += torch.randint(3, 32000, (S - len(ids) - 1,), generator=g).tolist() + [2]
    ids = torch.tensor([ids], device=dev)
    img = torch.empty(n_img, 3, 336, 336).normal_(0, 1, generator=g).to(device=dev, dtype=torch.bfloat16)
    out = {}

    def timed(fn, reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # greedy reference point
    for _ in range(2):
        model.generate(ids, images=[img], max_new_tokens=33, eos_token_id=None)
    full = timed(lambda: model.generate(ids, images=[img], max_new_tokens=33, eos_token_id=None), 2)
    pre = timed(lambda: model.generate(ids, images=[img], max_new_tokens=1, eos_token_id=None), 2)
    out["greedy_ms_per_token"] = (full - pre) / 32
    out["prefill_ms"] = pre
    print("greedy", out, flush=True)
    for k in [int(a) for a in sys.argv[1:]] or [1, 2, 5, 8]:
        runner = model._beam_runner(k, S, 64)
        model.generate(ids, images=[img], num_beams=k, max_new_tokens=4, eos_token_id=None)   # fills the prefix cache
        runner.begin(S)
        toks = list(range(100, 100 + k))
        for _ in range(3):
            runner.step(toks, None)
        # graph replays only (positions advance; 32 steps)
        kernel = timed(lambda: runner.graph.replay(), 32)
        runner.begin(S)
        full = timed(lambda: model.generate(ids, images=[img], num_beams=k, max_new_tokens=33, eos_token_id=None), 2)
        out[f"beams_{k}"] = {"step_ms": kernel, "generate_ms": full, "ms_per_position_e2e": (full - pre) / 32,
                             "vs_greedy_step": kernel / out["greedy_ms_per_token"]}
        print(k, out[f"beams_{k}"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_beam_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
