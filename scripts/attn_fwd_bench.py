#!/usr/bin/env python
"""Microbenchmark of merlin_flash_attn_fwd (causal, head_dim 128, 32 heads): CUDA-event time per launch and
TFLOP/s on the algorithmic count 2 S^2 D (causal), plus max|Δ| against torch fp32 math on a slice."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from merlin_b200 import ops  # noqa: E402

H, hd = 32, 128
for S in [int(x) for x in os.environ.get("S_LIST", "2048,4096,8192").split(",")]:
    torch.manual_seed(0)
    qkv = (torch.randn(S, 3 * H * hd, device="cuda") * 0.5).to(torch.bfloat16)
    for _ in range(3):
        out = ops.flash_attn(qkv, 1, S, H, hd, causal=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        out = ops.flash_attn(qkv, 1, S, H, hd, causal=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * S * S * H * hd
    q, k, v = (qkv.view(S, 3, H, hd)[:, i, :2].float().transpose(0, 1) for i in range(3))   # 2 heads
    att = torch.softmax((q @ k.transpose(-1, -2)) / hd ** 0.5 + torch.triu(torch.full((S, S), float("-inf"), device="cuda"), 1), -1)
    ref = (att @ v).transpose(0, 1).reshape(S, 2 * hd)
    err = (out[:, :2 * hd].float() - ref).abs().max().item()
    print(f"S={S}: {ms * 1e3:.1f} us, {fl / ms / 1e9:.0f} TFLOP/s, max|d| vs fp32 {err:.2e} (ref max {ref.abs().max().item():.2f})")
