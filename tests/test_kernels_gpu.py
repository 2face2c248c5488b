"""GPU: every kernel behind the C ABI against a torch fp32 restatement of the same op.
Integer / index work is bit-exact; floating work within the stated tolerance
(max|Δ| <= 2^-7 · max|ref| per op)."""
import pytest
import torch

from merlin_b200 import ops
from merlin_b200.lib import EPI_QUICKGELU, EPI_RESIDUAL, EPI_SWIGLU

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2.0 ** -7


def close(got, ref, tol=TOL):
    got, ref = got.float(), ref.float()
    assert not torch.isnan(got).any()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    assert err <= tol * scale, f"max|Δ|={err:.4g} > {tol:.3g}·{scale:.4g}"


def ints(shape):
    return torch.randint(-1, 2, shape, device=DEV).to(torch.bfloat16)


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("mnk", [(256, 256, 64), (512, 512, 512), (577 * 5, 1024, 1024), (300, 1000, 592),
                                 (1024, 32003, 256), (64, 256, 128), (1, 512, 256)])
def test_gemm_integer_operands_exact(cg, mnk):
    torch.manual_seed(0)
    M, N, K = mnk
    a, w = ints((M, K)), ints((N, K))
    got = ops.gemm(a, w, cta_group=cg)
    assert torch.equal(got.float(), a.float() @ w.float().t())   # operands in {-1,0,1}: |sums| < 2^8, exact in bf16


@pytest.mark.parametrize("cg", [1, 2])
def test_gemm_large_and_linearity(cg):
    torch.manual_seed(1)
    a = torch.randn(4096, 4096, device=DEV).to(torch.bfloat16)
    w = (torch.randn(4096, 4096, device=DEV) * 0.02).to(torch.bfloat16)
    y = ops.gemm(a, w, cta_group=cg)
    close(y, a.float() @ w.float().t())
    # size-independent property: A·(2W)ᵀ == 2·(A·Wᵀ) exactly (power-of-two scaling commutes with rounding)
    assert torch.equal(ops.gemm(a, (w * 2).contiguous(), cta_group=cg), y * 2)


@pytest.mark.parametrize("cg", [1, 2])
def test_gemm_epilogues(cg):
    torch.manual_seed(2)
    M, N, K = 1024 + 37, 1024, 512
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV).to(torch.bfloat16)
    res = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    base = a.float() @ w.float().t()
    close(ops.gemm(a, w, bias=bias, cta_group=cg), base + bias.float())
    close(ops.gemm(a, w, residual=res, epilogue=EPI_RESIDUAL, cta_group=cg), base + res.float())
    close(ops.gemm(a, w, bias=bias, residual=res, epilogue=EPI_RESIDUAL, cta_group=cg), base + bias.float() + res.float())
    x = res.clone()
    ops.gemm(a, w, residual=x, out=x, epilogue=EPI_RESIDUAL, cta_group=cg)   # in place (residual stream)
    close(x, base + res.float())
    v = base + bias.float()
    close(ops.gemm(a, w, bias=bias, epilogue=EPI_QUICKGELU, cta_group=cg), v * torch.sigmoid(1.702 * v))
    FF = 1408
    gw = (torch.randn(FF, K, device=DEV) * 0.05).to(torch.bfloat16)
    uw = (torch.randn(FF, K, device=DEV) * 0.05).to(torch.bfloat16)
    ref = torch.nn.functional.silu(a.float() @ gw.float().t()) * (a.float() @ uw.float().t())
    close(ops.gemm(a, ops.swiglu_interleave(gw, uw), epilogue=EPI_SWIGLU, cta_group=cg), ref)


def rope_ref(raw, cos, sin, S):
    c = torch.cat([cos[:S], cos[:S]], -1)[None, :, None, :]
    s = torch.cat([sin[:S], sin[:S]], -1)[None, :, None, :]

    def rot(x):
        x1, x2 = x[..., :64], x[..., 64:]
        return x * c + torch.cat([-x2, x1], -1) * s

    return torch.stack([rot(raw[:, :, 0]), rot(raw[:, :, 1]), raw[:, :, 2]], dim=2)


@pytest.mark.parametrize("cg", [1, 2])
def test_qkv_gemm_with_rope_epilogue(cg):
    torch.manual_seed(3)
    H, S, B = 4, 300, 2
    D = H * 128
    a = torch.randn(B * S, 256, device=DEV).to(torch.bfloat16)
    wq = (torch.randn(3 * D, 256, device=DEV) * 0.05).to(torch.bfloat16)
    cos, sin = ops.rope_tables(512)
    raw = (a.float() @ wq.float().t()).view(B, S, 3, H, 128)
    ref = rope_ref(raw, cos, sin, S).reshape(B * S, 3 * D)
    close(ops.gemm_qkv_rope(a, wq, cos, sin, S, cta_group=cg), ref)
    plain = ops.gemm(a, wq, cta_group=cg)
    close(ops.rope_qk_inplace(plain, cos, sin, H, S), ref)


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
    o = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, H * hd)
    if seqlens is not None:
        o = o * (idx[None, :] < seqlens[:, None]).reshape(B * S, 1)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("hd", [64, 128])
@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("bsh", [(1, 128, 1), (2, 384, 2), (1, 577, 3), (1, 1000, 2), (1, 2048, 4), (1, 1, 1), (1, 129, 1)])
def test_flash_attention(hd, causal, bsh):
    torch.manual_seed(4)
    B, S, H = bsh
    qkv = torch.randn(B * S, 3 * H * hd, device=DEV).to(torch.bfloat16)
    out, lse = ops.flash_attn(qkv, B, S, H, hd, causal=causal, return_lse=True)
    ref, ref_lse = attn_ref(qkv, B, S, H, hd, causal)
    close(out, ref)
    close(lse, ref_lse, tol=1e-4)


@pytest.mark.parametrize("hd", [64, 128])
def test_flash_attention_rescale_and_padding(hd):
    torch.manual_seed(5)
    B, S, H = 1, 512, 2
    qkv = (torch.randn(B * S, 3 * H * hd, device=DEV) * 4).to(torch.bfloat16)    # forces max refresh / O rescale
    close(ops.flash_attn(qkv, B, S, H, hd, causal=True), attn_ref(qkv, B, S, H, hd, True)[0])
    B, S = 3, 300
    qkv = torch.randn(B * S, 3 * H * hd, device=DEV).to(torch.bfloat16)
    sl = torch.tensor([300, 129, 7], dtype=torch.int32, device=DEV)
    out = ops.flash_attn(qkv, B, S, H, hd, causal=True, seqlens=sl)
    close(out, attn_ref(qkv, B, S, H, hd, True, sl)[0])
    assert bool((out.view(B, S, -1)[2, 7:] == 0).all())                           # pad rows are zeros (pad_input)


def test_norms():
    torch.manual_seed(6)
    x = torch.randn(300, 4096, device=DEV).to(torch.bfloat16)
    w = torch.randn(4096, device=DEV).to(torch.bfloat16)
    b = torch.randn(4096, device=DEV).to(torch.bfloat16)
    xf = x.float()
    close(ops.rmsnorm(x, w, 1e-5), w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)))
    x5 = torch.randn(7, 5120, device=DEV).to(torch.bfloat16)
    w5 = torch.randn(5120, device=DEV).to(torch.bfloat16)
    close(ops.rmsnorm(x5, w5, 1e-5), w5.float() * (x5.float() * torch.rsqrt(x5.float().pow(2).mean(-1, keepdim=True) + 1e-5)))
    x1, w1, b1 = x[:, :1024].contiguous(), w[:1024].contiguous(), b[:1024].contiguous()
    close(ops.layernorm(x1, w1, b1, 1e-5), torch.nn.functional.layer_norm(x1.float(), (1024,), w1.float(), b1.float(), 1e-5))


def test_im2col_gathers_are_bit_exact():
    torch.manual_seed(7)
    px = torch.randn(2, 3, 336, 336, device=DEV).to(torch.bfloat16)
    col = ops.patch_im2col(px, 14, 592)
    ref = torch.nn.functional.unfold(px.float(), 14, stride=14).transpose(1, 2).reshape(2 * 576, 588)
    assert torch.equal(col[:, :588].float(), ref) and bool((col[:, 588:] == 0).all())
    hid = torch.randn(2 * 577, 1024, device=DEV).to(torch.bfloat16)
    for stride in (1, 2):
        col = ops.conv3x3_im2col(hid, 2, 24, 1024, stride, 577, 1)
        feat = hid.view(2, 577, 1024)[:, 1:].permute(0, 2, 1).reshape(2, 1024, 24, 24).float()
        un = torch.nn.functional.unfold(feat, 3, padding=1, stride=stride)
        P = un.shape[-1]
        assert torch.equal(col.float(), un.view(2, 1024, 9, P).permute(0, 3, 2, 1).reshape(2 * P, 9 * 1024))


def test_vit_embed_ln():
    torch.manual_seed(8)
    po = torch.randn(2 * 576, 1024, device=DEV).to(torch.bfloat16)
    cls = torch.randn(1024, device=DEV).to(torch.bfloat16)
    pos = torch.randn(577, 1024, device=DEV).to(torch.bfloat16)
    lw = torch.randn(1024, device=DEV).to(torch.bfloat16)
    lb = torch.randn(1024, device=DEV).to(torch.bfloat16)
    emb = torch.cat([cls.expand(2, 1, 1024), po.view(2, 576, 1024)], 1) + pos
    close(ops.vit_embed_ln(po, cls, pos, lw, lb, 2, 1e-5),
          torch.nn.functional.layer_norm(emb.float(), (1024,), lw.float(), lb.float(), 1e-5).reshape(-1, 1024))


def test_embed_splice_bit_exact_with_ragged_samples():
    torch.manual_seed(9)
    V, D, P = 1000, 512, 144
    pid = V - 3
    table = torch.randn(V, D, device=DEV).to(torch.bfloat16)
    ids = torch.randint(0, pid, (3, 700), device=DEV)
    ids[0, 10:10 + P] = pid                       # sample 0: 1 image
    ids[1, 300:300 + P] = pid                     # sample 1: 2 images
    ids[1, 500:500 + P] = pid
    #                                               sample 2: text only, but owns one (unused) zero image
    feats = torch.randn(4 * P, D, device=DEV).to(torch.bfloat16)
    off = torch.tensor([0, P, 3 * P], dtype=torch.int32, device=DEV)
    lim = torch.tensor([P, 2 * P, P], dtype=torch.int32, device=DEV)
    out, counts = ops.embed_splice(ids, table, feats, pid, off, lim)
    ref = table[ids.view(-1)].clone().view(3, 700, D)
    ref[0][ids[0] == pid] = feats[:P]
    ref[1][ids[1] == pid] = feats[P:3 * P]
    assert torch.equal(out.view(3, 700, D), ref)
    assert counts.tolist() == [P, 2 * P, 0]
    # more placeholders than feature rows: the surplus keeps its token embedding (zip truncation, base_mmgpt.py:121)
    lim2 = torch.tensor([P, P, P], dtype=torch.int32, device=DEV)
    out2, _ = ops.embed_splice(ids, table, feats, pid, off, lim2)
    ref[1, 500:500 + P] = table[pid]
    assert torch.equal(out2.view(3, 700, D), ref)
    # empty batch
    e, _ = ops.embed_splice(torch.zeros((0, 5), dtype=torch.int64, device=DEV), table, None, pid)
    assert e.shape == (0, D)


def test_cross_entropy_and_argmax():
    torch.manual_seed(10)
    B, S, V = 2, 50, 32003
    logits = torch.randn(B * S, 32008, device=DEV).to(torch.bfloat16)[:, :V]
    labels = torch.randint(0, V, (B, S), device=DEV)
    labels[:, :20] = -100
    loss, row = ops.cross_entropy(logits, labels, B, S)
    ref = torch.nn.functional.cross_entropy(logits.float().view(B, S, V)[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1))
    assert abs(loss.item() - ref.item()) < 1e-4
    assert bool((row.view(B, S)[:, :19] == 0).all()) and bool((row.view(B, S)[:, -1] == 0).all())
    assert torch.equal(ops.argmax(logits), logits.float().argmax(-1))
    tie = torch.zeros(1, 64, device=DEV, dtype=torch.bfloat16)
    tie[0, 5] = tie[0, 40] = 3.0
    assert ops.argmax(tie).item() == 5                                           # lowest index on ties
    # unaligned rows (odd pitch -> scalar path), ties across the vector / tail boundary, maximum in the 3-element tail
    odd = torch.randn(3, 32003, device=DEV).to(torch.bfloat16)
    assert torch.equal(ops.argmax(odd), odd.float().argmax(-1))
    t2 = torch.zeros(1, 32008, device=DEV, dtype=torch.bfloat16)[:, :V]
    t2[0, 32001] = t2[0, 31999] = t2[0, 8191] = 7.0
    assert ops.argmax(t2).item() == 8191
    t2[0, 8191] = 0.0
    assert ops.argmax(t2).item() == 31999
    t2[0, 31999] = 0.0
    assert ops.argmax(t2).item() == 32001


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("mnk", [(256, 256, 64), (512, 384, 200), (4096, 1024, 2048), (304, 1000, 576), (1000, 264, 2888)])
def test_gemm_backward_layouts(cg, mnk):
    """dgrad (B transposed) and wgrad (both transposed) operand layouts, exact on {-1,0,1} operands."""
    torch.manual_seed(11)
    M, N, K = mnk
    a, bt, at = ints((M, K)), ints((K, N)), ints((K, M))
    assert torch.equal(ops.gemm_nt(a, bt, cta_group=cg).float(), a.float() @ bt.float())
    assert torch.equal(ops.gemm_tn(at, bt, cta_group=cg).float(), at.float().t() @ bt.float())
    acc = ints((M, N))
    want = acc.float() + at.float().t() @ bt.float()
    ops.gemm_tn(at, bt, out=acc, accumulate=True, cta_group=cg)          # gradient accumulation
    assert torch.equal(acc.float(), want)


def test_linear_backward_matches_autograd():
    torch.manual_seed(12)
    x = torch.randn(777, 512, device=DEV).to(torch.bfloat16)
    w = (torch.randn(1024, 512, device=DEV) * 0.05).to(torch.bfloat16)
    dy = torch.randn(777, 1024, device=DEV).to(torch.bfloat16)
    close(ops.gemm_nt(dy, w), dy.float() @ w.float())                      # dX = dY·W
    close(ops.gemm_tn(dy, x), dy.float().t() @ x.float())                  # dW = dYᵀ·X


@pytest.mark.parametrize("causal", [True, False])
@pytest.mark.parametrize("bsh", [(1, 128, 1), (1, 256, 2), (2, 384, 2), (1, 1000, 2), (1, 2048, 2)])
def test_flash_attention_backward(causal, bsh):
    torch.manual_seed(13)
    B, S, H, hd = *bsh, 128
    qkv = torch.randn(B * S, 3 * H * hd, device=DEV).to(torch.bfloat16)
    dout = torch.randn(B * S, H * hd, device=DEV).to(torch.bfloat16)
    out, lse = ops.flash_attn(qkv, B, S, H, hd, causal=causal, return_lse=True)
    dqkv = ops.flash_attn_bwd(qkv, out, dout, lse, B, S, H, hd, causal=causal)
    ref_in = qkv.float().clone().requires_grad_(True)
    q, k, v = ref_in.view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / hd ** 0.5
    if causal:
        idx = torch.arange(S, device=DEV)
        s = s.masked_fill((idx[None, :] > idx[:, None])[None, None], float("-inf"))
    o = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, H * hd)
    o.backward(dout.float())
    g = ref_in.grad.view(B * S, 3, H * hd)
    mine = dqkv.float().view(B * S, 3, H * hd)
    for blk in range(3):
        close(mine[:, blk], g[:, blk], tol=2.0 ** -6)


def test_flash_attention_backward_padding():
    torch.manual_seed(14)
    B, S, H, hd = 2, 300, 2, 128
    qkv = torch.randn(B * S, 3 * H * hd, device=DEV).to(torch.bfloat16)
    dout = torch.randn(B * S, H * hd, device=DEV).to(torch.bfloat16)
    sl = torch.tensor([300, 77], dtype=torch.int32, device=DEV)
    out, lse = ops.flash_attn(qkv, B, S, H, hd, causal=True, seqlens=sl, return_lse=True)
    dqkv = ops.flash_attn_bwd(qkv, out, dout, lse, B, S, H, hd, causal=True, seqlens=sl).view(B, S, -1)
    ref_in = qkv.float().clone().requires_grad_(True)
    ref, _ = attn_ref(ref_in, B, S, H, hd, True, sl)
    ref.backward(dout.float())
    g = ref_in.grad.view(B, S, -1)
    close(dqkv[0], g[0], tol=2.0 ** -6)
    close(dqkv[1, :77], g[1, :77], tol=2.0 ** -6)
    assert bool((dqkv[1, 77:] == 0).all())


def test_elementwise_backward_kernels():
    torch.manual_seed(15)
    M, D, FF = 333, 4096, 1408
    x = torch.randn(M, D, device=DEV).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(D, device=DEV)).to(torch.bfloat16)
    dy = torch.randn(M, D, device=DEV).to(torch.bfloat16)
    dres = torch.randn(M, D, device=DEV).to(torch.bfloat16)
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    y = wr * (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-5))
    y.backward(dy.float())
    dw = torch.zeros(D, dtype=torch.float32, device=DEV)
    dx = ops.rmsnorm_bwd(x, w, dy, dres, dw, 1e-5)
    close(dx, xr.grad + dres.float())
    close(dw, wr.grad, tol=2.0 ** -9)
    close(ops.rmsnorm_bwd(x, w, dy, None, None, 1e-5), xr.grad)
    # every width class of the warp-per-row kernel (D <= 256 / 1024 / 2048 / 4096 / 5120) and the CTA-per-row
    # fallback (6144), with a row count that is not a multiple of the 8 rows a CTA takes per sweep
    for D2, M2 in ((256, 77), (1024, 130), (1536, 9), (5120, 1001), (6144, 50)):
        x2 = torch.randn(M2, D2, device=DEV).to(torch.bfloat16)
        w2 = (1 + 0.1 * torch.randn(D2, device=DEV)).to(torch.bfloat16)
        dy2 = torch.randn(M2, D2, device=DEV).to(torch.bfloat16)
        dr2 = torch.randn(M2, D2, device=DEV).to(torch.bfloat16)
        xr2, wr2 = x2.float().requires_grad_(True), w2.float().requires_grad_(True)
        (wr2 * (xr2 * torch.rsqrt(xr2.pow(2).mean(-1, keepdim=True) + 1e-5))).backward(dy2.float())
        dw2 = torch.full((D2,), 0.5, dtype=torch.float32, device=DEV)          # accumulates onto what is there
        close(ops.rmsnorm_bwd(x2, w2, dy2, dr2, dw2, 1e-5), xr2.grad + dr2.float())
        close(dw2 - 0.5, wr2.grad, tol=2.0 ** -9)
    # swiglu
    g = torch.randn(M, FF, device=DEV).to(torch.bfloat16)
    u = torch.randn(M, FF, device=DEV).to(torch.bfloat16)
    gu = torch.stack((g.view(M, FF // 128, 128), u.view(M, FF // 128, 128)), dim=2).reshape(M, 2 * FF).contiguous()
    gr, ur = g.float().requires_grad_(True), u.float().requires_grad_(True)
    mid = torch.nn.functional.silu(gr) * ur
    close(ops.swiglu_fwd(gu), mid)
    dmid = torch.randn(M, FF, device=DEV).to(torch.bfloat16)
    mid.backward(dmid.float())
    dgu = ops.swiglu_bwd(gu, dmid).view(M, FF // 128, 2, 128)
    close(dgu[:, :, 0].reshape(M, FF), gr.grad)
    close(dgu[:, :, 1].reshape(M, FF), ur.grad)
    # rope inverse undoes rope
    H, S = 4, 200
    qkv = torch.randn(S, 3 * H * 128, device=DEV).to(torch.bfloat16)
    cos, sin = ops.rope_tables(256)
    rot = ops.rope_qk_inplace(qkv.clone(), cos, sin, H, S)
    close(ops.rope_qk_inverse_inplace(rot, cos, sin, H, S), qkv, tol=2.0 ** -6)
    # cross-entropy backward
    B, S2, V = 2, 40, 1003
    logits = torch.randn(B * S2, 1008, device=DEV).to(torch.bfloat16)[:, :V]
    labels = torch.randint(0, V, (B, S2), device=DEV)
    labels[:, :10] = -100
    lr_ = logits.float().clone().requires_grad_(True)
    loss_ref = torch.nn.functional.cross_entropy(lr_.view(B, S2, V)[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1))
    loss_ref.backward()
    loss, dlog = ops.cross_entropy_with_grad(logits, labels, B, S2)
    assert abs(loss.item() - loss_ref.item()) < 1e-4
    close(dlog, lr_.grad, tol=2.0 ** -7)
    # embedding / splice backward + column sums
    Vv, Dd = 500, 256
    ids = torch.randint(0, Vv - 1, (1, 300), device=DEV)
    ids[0, 50:60] = Vv - 1
    ids[0, 100] = ids[0, 101] = 7                          # duplicate token: gradients add
    table = torch.randn(Vv, Dd, device=DEV).to(torch.bfloat16)
    feats = torch.randn(10, Dd, device=DEV).to(torch.bfloat16)
    off = torch.zeros(1, dtype=torch.int32, device=DEV)
    lim = torch.full((1,), 10, dtype=torch.int32, device=DEV)
    _, _ = ops.embed_splice(ids, table, feats, Vv - 1, off, lim)
    rank = torch.full((300,), -1, dtype=torch.int32, device=DEV)
    rank[50:60] = torch.arange(10, dtype=torch.int32, device=DEV)
    dxx = torch.randn(300, Dd, device=DEV).to(torch.bfloat16)
    dtable = torch.zeros(Vv, Dd, device=DEV, dtype=torch.bfloat16)
    dfeats = torch.zeros(10, Dd, device=DEV, dtype=torch.bfloat16)
    ops.embed_splice_bwd(ids.view(-1), rank, dxx, dtable, dfeats, Vv)
    assert torch.equal(dfeats, dxx[50:60])
    ref_t = torch.zeros(Vv, Dd, device=DEV)
    text = torch.ones(300, dtype=torch.bool, device=DEV)
    text[50:60] = False
    ref_t.index_add_(0, ids.view(-1)[text], dxx.float()[text])
    close(dtable, ref_t, tol=2.0 ** -7)
    cs = torch.zeros(Dd, dtype=torch.float32, device=DEV)
    ops.colsum_accumulate(dxx, cs)
    close(cs, dxx.float().sum(0), tol=1e-5)


def test_adamw_matches_fp32_reference_rule():
    torch.manual_seed(16)
    n, chunk = 8192 * 3, 8192
    p = torch.randn(n, device=DEV).to(torch.bfloat16)
    g = (torch.randn(n, device=DEV) * 0.1).to(torch.bfloat16)
    m = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    v = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    lr_mult = torch.tensor([1.0, 0.5, 0.0], device=DEV)
    wd = torch.tensor([0.05, 0.0, 0.05], device=DEV)
    ss = torch.zeros(4, device=DEV)
    ops.grad_sumsq(g, ss)
    close(ss[:1], g.float().pow(2).sum().view(1), tol=1e-4)
    first = ss[0].item()
    for _ in range(5):                                   # order-independent accumulation: bitwise repeatable
        ops.grad_sumsq(g, ss)
        assert ss[0].item() == first
    pr, mr, vr = p.float().clone(), m.float().clone(), v.float().clone()
    lr, b1, b2, eps, max_norm = 1e-2, 0.9, 0.95, 1e-8, 1.0
    for step in (1, 2):
        ops.grad_sumsq(g, ss)
        ops.adamw_step(p, g, m, v, lr_mult, wd, chunk, lr, b1, b2, eps, step, 1.0, max_norm, ss)
        clip = min(1.0, max_norm / (g.float().norm().item() + 1e-6))
        gg = g.float() * clip
        for c in range(3):
            sl = slice(c * chunk, (c + 1) * chunk)
            lrc = lr * lr_mult[c].item()
            if lrc == 0:
                continue
            mr[sl] = (b1 * mr[sl] + (1 - b1) * gg[sl]).to(torch.bfloat16).float()
            vr[sl] = (b2 * vr[sl] + (1 - b2) * gg[sl] ** 2).to(torch.bfloat16).float()
        # (reference rule evaluated in fp32 on the bf16-rounded state, like the kernel)
        for c in range(3):
            sl = slice(c * chunk, (c + 1) * chunk)
            lrc = lr * lr_mult[c].item()
            if lrc == 0:
                continue
            m_pre = (mr[sl])
            denom = vr[sl].sqrt() / (1 - b2 ** step) ** 0.5 + eps
            pr[sl] = (pr[sl] * (1 - lrc * wd[c].item()) - lrc / (1 - b1 ** step) * m_pre / denom).to(torch.bfloat16).float()
    assert torch.equal(p[2 * chunk:], p[2 * chunk:])           # frozen chunk untouched below
    close(p[:2 * chunk], pr[:2 * chunk], tol=2.0 ** -6)
    close(m[:2 * chunk], mr[:2 * chunk], tol=2.0 ** -6)


def test_vision_backward_kernels():
    torch.manual_seed(17)
    M, D = 300, 1024
    x = torch.randn(M, D, device=DEV).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(D, device=DEV)).to(torch.bfloat16)
    b = (0.1 * torch.randn(D, device=DEV)).to(torch.bfloat16)
    dy = torch.randn(M, D, device=DEV).to(torch.bfloat16)
    dres = torch.randn(M, D, device=DEV).to(torch.bfloat16)
    xr, wr, br = x.float().requires_grad_(True), w.float().requires_grad_(True), b.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), wr, br, 1e-5).backward(dy.float())
    dw = torch.zeros(D, device=DEV)
    db = torch.zeros(D, device=DEV)
    close(ops.layernorm_bwd(x, w, dy, dres, dw, db, 1e-5), xr.grad + dres.float())
    close(dw, wr.grad, tol=2.0 ** -9)
    close(db, br.grad, tol=2.0 ** -9)
    # quick_gelu forward / backward
    xq = x.float().clone().requires_grad_(True)
    y = xq * torch.sigmoid(1.702 * xq)
    y.backward(dy.float())
    close(ops.quick_gelu(x), y)
    close(ops.quick_gelu(x, dy=dy), xq.grad)
    # col2im == autograd of the unfold-based im2col, strides 1 and 2
    n, G, C = 2, 24, 256
    for stride in (1, 2):
        feat = torch.randn(n, C, G, G, device=DEV, requires_grad=True)
        un = torch.nn.functional.unfold(feat, 3, padding=1, stride=stride)          # [n, C*9, P]
        P = un.shape[-1]
        col_ref = un.view(n, C, 9, P).permute(0, 3, 2, 1).reshape(n * P, 9 * C)
        dcol = torch.randn(n * P, 9 * C, device=DEV).to(torch.bfloat16)
        col_ref.backward(dcol.float())
        dfeat = torch.zeros(n * 577, C, device=DEV, dtype=torch.bfloat16)
        ops.conv3x3_col2im(dcol, dfeat, n, G, C, stride, 577, 1)
        want = feat.grad.permute(0, 2, 3, 1).reshape(n, G * G, C)
        close(dfeat.view(n, 577, C)[:, 1:], want, tol=2.0 ** -7)
        assert bool((dfeat.view(n, 577, C)[:, 0] == 0).all())
    # embedding assembly + backward
    po = torch.randn(2 * 576, 256, device=DEV).to(torch.bfloat16)
    cls = torch.randn(256, device=DEV).to(torch.bfloat16)
    pos = torch.randn(577, 256, device=DEV).to(torch.bfloat16)
    emb = ops.vit_embed(po, cls, pos, 2)
    ref = (torch.cat([cls.expand(2, 1, 256), po.view(2, 576, 256)], 1) + pos).reshape(-1, 256)
    assert torch.equal(emb, ref)
    d = torch.randn(2 * 577, 256, device=DEV).to(torch.bfloat16)
    dpos = torch.zeros(577, 256, device=DEV)
    dcls = torch.zeros(256, device=DEV)
    dpatch = ops.vit_embed_bwd(d, dpos, dcls, 2)
    assert torch.equal(dpatch.view(2, 576, 256), d.view(2, 577, 256)[:, 1:])
    close(dpos, d.float().view(2, 577, 256).sum(0), tol=1e-6)
    close(dcls, d.float().view(2, 577, 256)[:, 0].sum(0), tol=1e-6)


@pytest.mark.parametrize("causal", [True, False])
def test_flash_attention_backward_hd64(causal):
    torch.manual_seed(18)
    B, S, H, hd = 2, 577, 2, 64
    qkv = torch.randn(B * S, 3 * H * hd, device=DEV).to(torch.bfloat16)
    dout = torch.randn(B * S, H * hd, device=DEV).to(torch.bfloat16)
    out, lse = ops.flash_attn(qkv, B, S, H, hd, causal=causal, return_lse=True)
    dqkv = ops.flash_attn_bwd(qkv, out, dout, lse, B, S, H, hd, causal=causal)
    ref_in = qkv.float().clone().requires_grad_(True)
    ref, _ = attn_ref(ref_in, B, S, H, hd, causal)
    ref.backward(dout.float())
    g = ref_in.grad.view(B * S, 3, H * hd)
    mine = dqkv.float().view(B * S, 3, H * hd)
    for blk in range(3):
        close(mine[:, blk], g[:, blk], tol=2.0 ** -6)


@pytest.mark.parametrize("hw", [(480, 640), (336, 336), (1080, 1920), (100, 57), (37, 336)])
def test_image_preprocess_is_bit_exact_with_pillow_resize(hw):
    """merlin_image_preprocess: the uint8 pixels after the bicubic resize equal Pillow's `image.resize((S, S))`
    bit for bit (and the numpy oracle's), the normalised fp32 output equals the oracle's rescale / normalize
    exactly, bf16 output is its RNE rounding (base_dataset.py:190-192)."""
    import numpy as np
    from oracle import image_oracle as IO
    rng = np.random.default_rng(hw[0] + hw[1])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    img[: hw[0] // 3] = 255
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
    out, resized = ops.image_preprocess(torch.from_numpy(img).cuda(), 336, mean, std, return_resized=True)
    want_u8 = IO.resize_bicubic_u8(img, 336)
    assert torch.equal(resized.cpu(), torch.from_numpy(want_u8))
    try:
        from PIL import Image
        assert np.array_equal(resized.cpu().numpy(), np.asarray(Image.fromarray(img).resize((336, 336))))
    except ImportError:
        pass
    want = torch.from_numpy(IO.rescale_normalize(want_u8, mean, std))
    assert torch.equal(out.cpu(), want)
    out16 = ops.image_preprocess(torch.from_numpy(img).cuda(), 336, mean, std, dtype=torch.bfloat16)
    assert torch.equal(out16.cpu(), want.to(torch.bfloat16))


def test_interpair_dataset_image_path_runs_the_gpu_kernel(tmp_path):
    """InterPairWebDataset over a local tar shard with its default image loader: JPEG bytes -> PIL decode ->
    merlin_image_preprocess; equals the reference's PIL resize + CLIP processor arithmetic (oracle)."""
    import io
    import json
    import tarfile
    import types
    import numpy as np
    from PIL import Image
    from oracle import image_oracle as IO
    from merlin_b200.mmgpt.data.dataset.interpair_webdataset import InterPairWebDataset
    rng = np.random.default_rng(5)
    buf = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, (90, 120, 3), dtype=np.uint8)).save(buf, format="JPEG")
    jpg = buf.getvalue()
    shard = tmp_path / "s.tar"
    with tarfile.open(shard, "w") as tf:
        for name, data in (("r0.json", json.dumps(dict(prompt="Look <image>", txt="a thing", image_name_list=["v-f0"])).encode()),
                           ("r0.f0.jpg", jpg)):
            ti = tarfile.TarInfo(name)
            ti.size = len(data)
            tf.addfile(ti, io.BytesIO(data))

    class Tok:
        eos_token, pad_token_id, model_max_length = "</s>", 0, 128

        def convert_tokens_to_ids(self, toks):
            return [32000, 32001, 32002]

        def __call__(self, text, padding=None, max_length=None, truncation=False):
            return types.SimpleNamespace(input_ids=[1] + [5] * min(len(text.split()), 20))

    ds = InterPairWebDataset([str(shard)], Tok(), dict(image_token_len=4, use_im_start_end=True, image_size=56,
                                                       image_processor=None))
    sample = ds[0]
    want = IO.image_preprocess(np.asarray(Image.open(io.BytesIO(jpg)).convert("RGB")), 56,
                               (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))
    assert len(sample["image"]) == 1 and torch.equal(sample["image"][0], torch.from_numpy(want))


def test_image_preprocess_pad_mode_matches_expand2square_then_pillow():
    """mode 'pad' (base_dataset.py:186-188): expand2square on int(255 * mean), then the processor's bicubic resize."""
    import numpy as np
    from PIL import Image
    from oracle import image_oracle as IO
    from merlin_b200.mmgpt.utils.mm_utils import expand2square
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
    rng = np.random.default_rng(3)
    for hw in ((200, 321), (400, 150)):
        img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
        sq = expand2square(Image.fromarray(img), tuple(int(x * 255) for x in mean))
        want_u8 = np.asarray(sq.resize((224, 224), resample=Image.BICUBIC))
        out, resized = ops.image_preprocess(torch.from_numpy(img).cuda(), 224, mean, std, return_resized=True,
                                            pad_to_square=True)
        assert np.array_equal(resized.cpu().numpy(), want_u8)
        assert torch.equal(out.cpu(), torch.from_numpy(IO.rescale_normalize(want_u8, mean, std)))


@pytest.mark.parametrize("B,S", [(1, 333), (1, 1024), (2, 256), (1, 40)])
def test_kv_append_fused_into_the_qkv_gemm_epilogue_is_bit_exact(B, S):
    """merlin_gemm_qkv_rope_kv writes the rotated keys and the values into the block-of-32 KV cache from the GEMM's own
    epilogue (second TMA store of the staged chunk): same packed qkv, and the cache rows equal merlin_kv_append's
    bit for bit — incl. a ragged last tile (S = 333, 40) and two tile-aligned samples."""
    torch.manual_seed(B * 1000 + S)
    H, D = 2, 256
    x = torch.randn(B * S, D, device=DEV).to(torch.bfloat16)
    w = (torch.randn(3 * D, D, device=DEV) * 0.05).to(torch.bfloat16)
    cos, sin = ops.rope_tables(2048)
    cap = (S + 95) // 32 * 32
    plain = ops.gemm_qkv_rope(x, w, cos, sin, S)
    want = torch.zeros(B, cap // 32, H, 32, 256, dtype=torch.bfloat16, device=DEV)
    ops.kv_append(plain, want, B, S, D, H)
    got = torch.zeros_like(want)
    fused = ops.gemm_qkv_rope(x, w, cos, sin, S, kv_cache=got, batch=B, heads=H)
    assert torch.equal(fused, plain)
    t = torch.arange(S, device=DEV)
    assert torch.equal(got[:, t // 32, :, t % 32], want[:, t // 32, :, t % 32])      # every appended (token, head) row
    if S % 128 == 0:
        assert torch.equal(got, want)                                                 # and nothing else was touched
