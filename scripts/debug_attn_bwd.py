import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from merlin_b200 import ops
B, S, H, hd = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 128, 1, 128
torch.manual_seed(0)
qkv = torch.randn(B * S, 3 * H * hd, device="cuda").to(torch.bfloat16)
dout = torch.randn(B * S, H * hd, device="cuda").to(torch.bfloat16)
out, lse = ops.flash_attn(qkv, B, S, H, hd, causal=True, return_lse=True)
torch.cuda.synchronize(); print("fwd ok", flush=True)
dqkv = ops.flash_attn_bwd(qkv, out, dout, lse, B, S, H, hd, causal=True)
torch.cuda.synchronize(); print("bwd returned", flush=True)
ref_in = qkv.float().clone().requires_grad_(True)
q, k, v = ref_in.view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
s = (q @ k.transpose(-1, -2)) / hd ** 0.5
idx = torch.arange(S, device="cuda")
s = s.masked_fill((idx[None, :] > idx[:, None])[None, None], float("-inf"))
o = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, H * hd)
o.backward(dout.float())
g = ref_in.grad.view(B * S, 3, H * hd); mine = dqkv.float().view(B * S, 3, H * hd)
for blk, n in enumerate("qkv"):
    e = (mine[:, blk] - g[:, blk]).abs().max().item(); m = g[:, blk].abs().max().item()
    print(f"d{n}: max err {e:.4g} / ref max {m:.4g}", flush=True)
