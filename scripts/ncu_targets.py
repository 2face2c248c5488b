"""Launches each hot kernel a few times at bench shapes (for `ncu --set full -k regex:...`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from merlin_b200 import ops
from merlin_b200.lib import EPI_SWIGLU

dev = "cuda"
which = sys.argv[1]
torch.manual_seed(0)
if which == "attn":
    B, S, H, hd = 1, 4096, 32, 128
    qkv = torch.randn(B * S, 3 * H * hd, device=dev).to(torch.bfloat16)
    for _ in range(3):
        ops.flash_attn(qkv, B, S, H, hd, causal=True)
elif which == "gemm":
    a = torch.randn(4096, 4096, device=dev).to(torch.bfloat16)
    w = (torch.randn(22016, 4096, device=dev) * 0.02).to(torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, epilogue=EPI_SWIGLU)
torch.cuda.synchronize()
