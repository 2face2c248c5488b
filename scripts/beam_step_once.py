#!/usr/bin/env python
"""Three eager batched beam-decode steps at the bench shape (7B, S=4096) for an ncu launch list:
MERLIN_BEAM_GRAPH=0 ncu -k regex:'gemv_beams|prefix_attention|suffix_combine' --launch-skip 386 --launch-count 193 ...
usage: python scripts/beam_step_once.py K"""
import os
import sys

os.environ["MERLIN_BEAM_GRAPH"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch  # noqa: E402

from config_bench import build  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
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
model.generate(ids, images=[img], num_beams=k, max_new_tokens=1, eos_token_id=None)
runner = model._beam_runner(k, S, 1)
runner.begin(S)
for _ in range(3):
    runner.step(list(range(100, 100 + k)), None)
torch.cuda.synchronize()
print("done", runner.t)
