"""GPU bring-up harness: runs one named group of kernel checks against torch fp32 references and
prints diagnostics (never raises on a numeric mismatch, so one call surveys everything).
Usage: python scripts/bringup.py <group> ; groups: gemm1 gemm2 epi1 epi2 attn128 attn64 elt perf
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from merlin_b200 import ops
from merlin_b200.lib import EPI_QUICKGELU, EPI_RESIDUAL, EPI_STORE, EPI_SWIGLU

dev = "cuda"
torch.manual_seed(0)


def report(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-9
    rel = err.max().item() / denom
    bad = (err > tol * denom).sum().item()
    nan = torch.isnan(got).sum().item()
    status = "OK " if (rel <= tol and nan == 0) else "BAD"
    print(f"[{status}] {name}: max_abs={err.max().item():.4g} rel={rel:.3g} bad={bad}/{err.numel()} nan={nan} ref_max={denom:.4g}", flush=True)
    if status == "BAD" and err.dim() == 2:
        R, C = err.shape
        rb, cb = max(1, min(128, R)), max(1, min(64, C))
        pr = (R + rb - 1) // rb
        pc = (C + cb - 1) // cb
        lines = []
        for i in range(min(pr, 6)):
            row = []
            for j in range(min(pc, 12)):
                blk = err[i * rb:(i + 1) * rb, j * cb:(j + 1) * cb]
                row.append(f"{(blk > tol * denom).float().mean().item():.2f}")
            lines.append(" ".join(row))
        print("   bad-fraction per (128 rows x 64 cols) block:\n   " + "\n   ".join(lines), flush=True)
        i, j = divmod(err.argmax().item(), C)
        print(f"   worst at ({i},{j}): got {got[i, j].item():.5g} ref {ref[i, j].item():.5g}", flush=True)
    return status == "OK "


def ints(shape, lo=-2, hi=3):
    return torch.randint(lo, hi, shape, device=dev).to(torch.bfloat16)


def gemm_cases(cg):
    # exact integer-valued operands first (any layout bug shows up as a hard mismatch)
    for (M, N, K) in [(128 * cg, 256, 64), (128 * cg, 256, 256), (256 * cg, 512, 512), (4096, 4096, 4096),
                      (577 * 5, 1024, 1024), (300, 1000, 592), (2048, 32003, 512), (64, 256, 128)]:
        a = ints((M, K))
        w = ints((N, K))
        ref = a.float() @ w.float().t()
        got = ops.gemm(a, w, cta_group=cg)
        torch.cuda.synchronize()
        report(f"gemm cg{cg} int M={M} N={N} K={K}", got, ref, tol=1e-6 if K <= 512 else 8e-3)
    a = torch.randn(1024, 4096, device=dev).to(torch.bfloat16)
    w = (torch.randn(4096, 4096, device=dev) * 0.02).to(torch.bfloat16)
    report(f"gemm cg{cg} randn 1024x4096x4096", ops.gemm(a, w, cta_group=cg), a.float() @ w.float().t(), tol=1e-2)


def epi_cases(cg):
    M, N, K = 1024 + 37, 1024, 512
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=dev).to(torch.bfloat16)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16)
    base = a.float() @ w.float().t()
    report(f"epi cg{cg} bias", ops.gemm(a, w, bias=bias, cta_group=cg), base + bias.float())
    report(f"epi cg{cg} residual", ops.gemm(a, w, residual=res, epilogue=EPI_RESIDUAL, cta_group=cg), base + res.float())
    report(f"epi cg{cg} bias+residual", ops.gemm(a, w, bias=bias, residual=res, epilogue=EPI_RESIDUAL, cta_group=cg),
           base + bias.float() + res.float())
    x = res.clone()
    ops.gemm(a, w, residual=x, out=x, epilogue=EPI_RESIDUAL, cta_group=cg)
    report(f"epi cg{cg} residual in-place", x, base + res.float())
    v = base + bias.float()
    report(f"epi cg{cg} quick_gelu", ops.gemm(a, w, bias=bias, epilogue=EPI_QUICKGELU, cta_group=cg), v * torch.sigmoid(1.702 * v))
    FF = 1408
    gw = (torch.randn(FF, K, device=dev) * 0.05).to(torch.bfloat16)
    uw = (torch.randn(FF, K, device=dev) * 0.05).to(torch.bfloat16)
    packed = ops.swiglu_interleave(gw, uw)
    g = a.float() @ gw.float().t()
    u = a.float() @ uw.float().t()
    report(f"epi cg{cg} swiglu", ops.gemm(a, packed, epilogue=EPI_SWIGLU, cta_group=cg), torch.nn.functional.silu(g) * u)
    # rope epilogue
    H, S = 4, 300
    D = H * 128
    a2 = torch.randn(2 * S, 256, device=dev).to(torch.bfloat16)
    wq = (torch.randn(3 * D, 256, device=dev) * 0.05).to(torch.bfloat16)
    cos, sin = ops.rope_tables(512)
    got = ops.gemm_qkv_rope(a2, wq, cos, sin, S, cta_group=cg)
    raw = (a2.float() @ wq.float().t()).view(2, S, 3, H, 128)
    c = torch.cat([cos[:S], cos[:S]], -1)[None, :, None, :]
    s_ = torch.cat([sin[:S], sin[:S]], -1)[None, :, None, :]

    def rot(x):
        x1, x2 = x[..., :64], x[..., 64:]
        return x * c + torch.cat([-x2, x1], -1) * s_

    ref = torch.stack([rot(raw[:, :, 0]), rot(raw[:, :, 1]), raw[:, :, 2]], dim=2).reshape(2 * S, 3 * D)
    report(f"epi cg{cg} qkv+rope", got, ref)
    qkv_plain = ops.gemm(a2, wq, cta_group=cg)
    ops.rope_qk_inplace(qkv_plain, cos, sin, H, S)
    report(f"rope_qk_inplace (after cg{cg} gemm)", qkv_plain, ref)


def attn_ref(qkv, B, S, H, hd, causal, seqlens=None):
    q, k, v = qkv.float().view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / hd ** 0.5
    idx = torch.arange(S, device=qkv.device)
    mask = torch.zeros(B, 1, S, S, dtype=torch.bool, device=qkv.device)
    if causal:
        mask |= (idx[None, :] > idx[:, None])[None, None]
    if seqlens is not None:
        mask |= (idx[None, None, None, :] >= seqlens.view(B, 1, 1, 1))
    s = s.masked_fill(mask, float("-inf"))
    o = torch.softmax(s, -1) @ v
    lse = torch.logsumexp(s, -1)
    o = o.permute(0, 2, 1, 3).reshape(B * S, H * hd)
    if seqlens is not None:
        valid = (idx[None, :] < seqlens[:, None]).reshape(B * S, 1)
        o = o * valid
    return o, lse


def attn_cases(hd):
    causal_modes = [True, False]
    for causal in causal_modes:
        for (B, S, H) in [(1, 128, 1), (1, 256, 2), (2, 384, 2), (1, 577, 3), (1, 1000, 2), (1, 2048, 4)]:
            qkv = torch.randn(B * S, 3 * H * hd, device=dev).to(torch.bfloat16)
            out, lse = ops.flash_attn(qkv, B, S, H, hd, causal=causal, return_lse=True)
            torch.cuda.synchronize()
            ref, ref_lse = attn_ref(qkv, B, S, H, hd, causal)
            report(f"attn hd{hd} causal={causal} B={B} S={S} H={H}", out, ref)
            report("   lse", lse.reshape(B * H, S), ref_lse.reshape(B * H, S), tol=1e-3)
    # large-magnitude scores exercise the lazy max refresh / O rescale
    B, S, H = 1, 512, 2
    qkv = (torch.randn(B * S, 3 * H * hd, device=dev) * 4).to(torch.bfloat16)
    out = ops.flash_attn(qkv, B, S, H, hd, causal=True)
    report(f"attn hd{hd} causal large-scores", out, attn_ref(qkv, B, S, H, hd, True)[0])
    # right padding
    B, S, H = 3, 300, 2
    qkv = torch.randn(B * S, 3 * H * hd, device=dev).to(torch.bfloat16)
    sl = torch.tensor([300, 129, 7], dtype=torch.int32, device=dev)
    out = ops.flash_attn(qkv, B, S, H, hd, causal=True, seqlens=sl)
    report(f"attn hd{hd} causal seqlens", out, attn_ref(qkv, B, S, H, hd, True, sl)[0])


def elt_cases():
    x = torch.randn(300, 4096, device=dev).to(torch.bfloat16)
    w = torch.randn(4096, device=dev).to(torch.bfloat16)
    b = torch.randn(4096, device=dev).to(torch.bfloat16)
    xf = x.float()
    report("rmsnorm", ops.rmsnorm(x, w, 1e-5), w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)))
    x1 = x[:, :1024].contiguous()
    report("layernorm", ops.layernorm(x1, w[:1024].contiguous(), b[:1024].contiguous(), 1e-5),
           torch.nn.functional.layer_norm(x1.float(), (1024,), w[:1024].float(), b[:1024].float(), 1e-5))
    px = torch.randn(2, 3, 336, 336, device=dev).to(torch.bfloat16)
    col = ops.patch_im2col(px, 14, 592)
    ref = torch.nn.functional.unfold(px.float(), 14, stride=14).transpose(1, 2).reshape(2 * 576, 588)
    report("patch_im2col", col[:, :588], ref, tol=0)
    print("   pad zero:", bool((col[:, 588:] == 0).all().item()))
    # vit embed
    po = torch.randn(2 * 576, 1024, device=dev).to(torch.bfloat16)
    cls = torch.randn(1024, device=dev).to(torch.bfloat16)
    pos = torch.randn(577, 1024, device=dev).to(torch.bfloat16)
    lw, lb = w[:1024].contiguous(), b[:1024].contiguous()
    emb = torch.cat([cls.expand(2, 1, 1024), po.view(2, 576, 1024)], 1) + pos
    report("vit_embed_ln", ops.vit_embed_ln(po, cls, pos, lw, lb, 2, 1e-5),
           torch.nn.functional.layer_norm(emb.float(), (1024,), lw.float(), lb.float(), 1e-5).reshape(-1, 1024))
    # conv im2col
    hid = torch.randn(2 * 577, 1024, device=dev).to(torch.bfloat16)
    col = ops.conv3x3_im2col(hid, 2, 24, 1024, 2, 577, 1)
    feat = hid.view(2, 577, 1024)[:, 1:].permute(0, 2, 1).reshape(2, 1024, 24, 24).float()
    un = torch.nn.functional.unfold(feat, 3, padding=1, stride=2)  # [2, C*9, 144], k = c*9 + tap
    ref = un.view(2, 1024, 9, 144).permute(0, 3, 2, 1).reshape(2 * 144, 9 * 1024)
    report("conv3x3_im2col", col, ref, tol=0)
    # splice
    V, D = 1000, 512
    table = torch.randn(V, D, device=dev).to(torch.bfloat16)
    ids = torch.randint(0, V - 3, (2, 700), device=dev)
    ids[0, 10:154] = V - 3
    ids[1, 300:444] = V - 3
    ids[1, 500:644] = V - 3
    feats = torch.randn(3 * 144, D, device=dev).to(torch.bfloat16)
    out, ws = ops.embed_splice(ids, table, feats, V - 3, torch.tensor([0, 144], dtype=torch.int32, device=dev),
                               torch.tensor([144, 288], dtype=torch.int32, device=dev))
    ref = table[ids.view(-1)].clone()
    ref[ids.view(-1) == V - 3] = feats
    print("[%s] embed_splice bit-exact, counts=%s" % ("OK " if torch.equal(out, ref) else "BAD", ws.tolist()))
    # CE + argmax
    B, S, Vv = 2, 50, 32003
    logits = torch.randn(B * S, 32008, device=dev).to(torch.bfloat16)[:, :Vv]
    labels = torch.randint(0, Vv, (B, S), device=dev)
    labels[:, :20] = -100
    loss, row = ops.cross_entropy(logits, labels, B, S)
    ref = torch.nn.functional.cross_entropy(logits.float().view(B, S, Vv)[:, :-1].reshape(-1, Vv), labels[:, 1:].reshape(-1))
    print("[%s] cross_entropy: got %.6f ref %.6f" % ("OK " if abs(loss.item() - ref.item()) < 1e-4 else "BAD", loss.item(), ref.item()))
    am = ops.argmax(logits)
    print("[%s] argmax" % ("OK " if torch.equal(am, logits.float().argmax(-1)) else "BAD"))


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def perf_cases():
    for (M, N, K) in [(4096, 12288, 4096), (4096, 4096, 4096), (4096, 22016, 4096), (4096, 4096, 11008), (4096, 32003, 4096), (2885, 3072, 1024), (2885, 4096, 1024), (2885, 1024, 4096), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        line = f"gemm M={M} N={N} K={K}:"
        for cg in (1, 2):
            out = ops.gemm(a, w, cta_group=cg)
            ms = timeit(lambda: ops.gemm(a, w, out=out, cta_group=cg))
            line += f"  cg{cg} {ms * 1e3:.1f}us {2 * M * N * K / ms / 1e9:.0f} TF/s"
        wt = w.t()
        ms = timeit(lambda: torch.matmul(a, wt))
        line += f"  cublas {ms * 1e3:.1f}us {2 * M * N * K / ms / 1e9:.0f} TF/s"
        print(line, flush=True)
    for (B, S, H, hd, causal) in [(1, 4096, 32, 128, True), (1, 2048, 32, 128, True), (1, 8192, 32, 128, True), (5, 577, 16, 64, False)]:
        qkv = torch.randn(B * S, 3 * H * hd, device=dev).to(torch.bfloat16)
        out = ops.flash_attn(qkv, B, S, H, hd, causal=causal)
        ms = timeit(lambda: ops.flash_attn(qkv, B, S, H, hd, causal=causal, out=out))
        fl = 4 * B * H * S * S * hd * (0.5 if causal else 1.0)
        q, k, v = qkv.view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
        ms2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal))
        print(f"attn B={B} S={S} H={H} hd={hd} causal={causal}: {ms * 1e3:.1f}us {fl / ms / 1e9:.0f} TF/s | torch sdpa {ms2 * 1e3:.1f}us {fl / ms2 / 1e9:.0f} TF/s", flush=True)
    x = torch.randn(4096, 4096, device=dev).to(torch.bfloat16)
    w = torch.randn(4096, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)
    ms = timeit(lambda: ops.rmsnorm(x, w, 1e-5, out=y))
    print(f"rmsnorm 4096x4096: {ms * 1e3:.1f}us {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s")


if __name__ == "__main__":
    group = sys.argv[1]
    t0 = time.time()
    print(f"=== {group} (lib: {os.environ.get('MERLIN_B200_LIB', 'release')})", flush=True)
    {"gemm1": lambda: gemm_cases(1), "gemm2": lambda: gemm_cases(2), "epi1": lambda: epi_cases(1),
     "epi2": lambda: epi_cases(2), "attn128": lambda: attn_cases(128), "attn64": lambda: attn_cases(64),
     "elt": elt_cases, "perf": perf_cases}[group]()
    torch.cuda.synchronize()
    print(f"=== {group} done in {time.time() - t0:.1f}s", flush=True)
