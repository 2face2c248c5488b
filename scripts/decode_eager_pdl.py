"""Eager (non-graph) decode-step timing with PDL on/off: isolates whether programmatic dependent launch helps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from merlin_b200 import decode

dev = torch.device("cuda", 0)
model = bench.build_model(dev)
ids_h, img_h = bench.make_inputs(0)
gen = decode.GraphedGenerator(model, bench.SEQ_LEN, img_h.shape, bench.SEQ_LEN + 200)
gen.generate(ids_h, img_h, 4, None)
r = gen.runner
torch.cuda.synchronize()
for mode in ("eager", "graph"):
    gen.cache.seq_len = bench.SEQ_LEN
    r.pos.fill_(bench.SEQ_LEN); r.n_out.zero_()
    n = 64
    for _ in range(4):
        r._greedy_body() if mode == "eager" else r.graph.replay()
    r.pos.fill_(bench.SEQ_LEN); r.n_out.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r._greedy_body() if mode == "eager" else r.graph.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{mode}: {e0.elapsed_time(e1) / n:.3f} ms/token  PDL={os.environ.get('MERLIN_PDL', '1')}", flush=True)
