#!/usr/bin/env python
"""Microbenchmark of merlin_flash_attn_bwd (causal, head_dim 128, S=4096, 32 heads — one decoder layer of the C3
training step): CUDA-event time per launch and tensor TFLOP/s (2.5 x forward FLOPs).  MERLIN_ATTN_BWD_DBG masks
switch parts of the kernel off to localise the time (results are then wrong on purpose)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from merlin_b200 import ops  # noqa: E402

S, H, hd = int(os.environ.get("S", 4096)), 32, 128
torch.manual_seed(0)
qkv = (torch.randn(S, 3 * H * hd, device="cuda") * 0.5).to(torch.bfloat16)
out, lse = ops.flash_attn(qkv, 1, S, H, hd, causal=True, return_lse=True)
dout = torch.randn_like(out)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    ops.flash_attn_bwd(qkv, out, dout, lse, 1, S, H, hd, causal=True, dqkv=dqkv)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    ops.flash_attn_bwd(qkv, out, dout, lse, 1, S, H, hd, causal=True, dqkv=dqkv)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 2.5 * 2 * S * S * H * hd          # causal: half of 4 S^2 hd H, x2.5 for the backward
print(f"dbg={os.environ.get('MERLIN_ATTN_BWD_DBG', '0')} S={S}: {ms * 1e3:.1f} us per bwd call (incl. delta, memset, dq->bf16), "
      f"{fl / ms / 1e9:.0f} TFLOP/s")
if os.environ.get("MERLIN_B200_LIB", "").endswith("_dbg.so"):
    import ctypes
    from merlin_b200 import lib as _lib
    L = _lib.load()
    buf = (ctypes.c_ulonglong * (64 * 16))()
    L.merlin_debug_attn_bwd_trace.argtypes = [ctypes.c_void_p]
    assert L.merlin_debug_attn_bwd_trace(buf) == 0
    t = [[buf[s * 16 + e] for e in range(16)] for s in range(32)]
    t0 = min(v for row in t for v in row if v)
    names = {0: "mma:dP ready", 1: "mma:dP issued", 2: "mma:p_full", 3: "mma:dV issued", 4: "mma:S+1 issued", 5: "mma:ds_full", 6: "mma:dK,dQ issued",
             8: "sm:s_full", 14: "sm:stg_free", 15: "sm:P stored", 9: "sm:P pub", 10: "sm:dp_full", 11: "sm:dS pub", 12: "sm:dq_full", 13: "sm:drained"}
    for s in (0, 1, 2, 3, 10, 11):
        ev = sorted((t[s][e] - t0, names[e]) for e in names if t[s][e])
        print(f"step {s}: " + "  ".join(f"{n}@{v / 1e3:.2f}" for v, n in ev))
    print("step period (sm:s_full) us:", [round((t[s + 1][8] - t[s][8]) / 1e3, 2) for s in range(0, 12)])
