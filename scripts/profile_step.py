"""Short run for ncu: one prefill + a few un-graphed decode steps of the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from merlin_b200 import ops

dev = torch.device("cuda", 0)
model = bench.build_model(dev)
ids_h, img_h = bench.make_inputs(0)
ids, img = ids_h.to(dev), img_h.to(device=dev, dtype=torch.bfloat16)
n_dec = int(os.environ.get("N_DEC", "3"))
out = model(input_ids=ids, images=[img], use_cache=True)
cache = out.past_key_values
first = ops.argmax(out.logits[:, -1, :]).view(1, 1)
runner = model._get_decoder(cache)
toks = runner.greedy(first, n_dec, None, use_graph=False)
torch.cuda.synchronize()
print("tokens", toks.tolist())
