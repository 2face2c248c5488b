"""Decode-only timing at the bench shape (ctx 4096): ms/token of the graph-replayed greedy loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from merlin_b200 import decode

dev = torch.device("cuda", 0)
model = bench.build_model(dev)
ids_h, img_h = bench.make_inputs(0)
gen = decode.GraphedGenerator(model, bench.SEQ_LEN, img_h.shape, bench.SEQ_LEN + 200)
gen.generate(ids_h, img_h, 8, None)
torch.cuda.synchronize()
r = gen.runner
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 96
gen.cache.seq_len = bench.SEQ_LEN
r.pos.fill_(bench.SEQ_LEN)
r.n_out.zero_()
e0.record()
for _ in range(n):
    r.graph.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
floor = (13.48e9 + 0.524e6 * (bench.SEQ_LEN + n / 2)) / 6545e9 * 1e3
print(f"decode {ms:.3f} ms/token  ({floor / ms * 100:.1f}% of HBM roofline)  env=" +
      " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("MERLIN_")))
