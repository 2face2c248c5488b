"""Bring-up / parity / timing of the single-launch streaming decode step (csrc/decode_stream.cu) against the
per-kernel decode path (csrc/decode.cu) on the same weights and KV cache.

  python scripts/stream_check.py small     # odd shapes, several context lengths (tile / tail edge cases)
  python scripts/stream_check.py full      # Vicuna-7B dims, context 4096: parity + ms/token of both paths
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from merlin_b200 import decode as D
from merlin_b200.mmgpt.model.mmgpt.llama_mmgpt import MMGPTLlamaForCausalLM, MerlinKVCache, vicuna_7b_config

DEV = torch.device("cuda:0")


def runner(model, cache, mode):
    os.environ["MERLIN_DECODE"] = mode
    return D.DecodeRunner(model, cache)


def clone_cache(model, cache):
    c = MerlinKVCache(len(cache.kv), 1, cache.capacity, cache.kv[0].shape[2] * 128, DEV)
    for a, b in zip(c.kv, cache.kv):
        a.copy_(b)
    c.seq_len = cache.seq_len
    return c


def compare(model, ctx, steps=3, label="", full_ref=True, verbose=False):
    cfg = model.config
    torch.manual_seed(ctx)
    ids = torch.randint(3, cfg.vocab_size - 1, (1, ctx), device=DEV)
    model._cache_capacity = ctx + 64
    pre = model(input_ids=ids, use_cache=True)
    ca, cb = pre.past_key_values, None
    cb = clone_cache(model, ca)
    ra, rb = runner(model, ca, "kernels"), runner(model, cb, "stream")
    worst = 0.0
    toks = ids
    for s in range(steps):
        tok = torch.randint(3, cfg.vocab_size - 1, (1, 1), device=DEV)
        toks = torch.cat([toks, tok], dim=1)
        emb = model.model.embed_tokens[tok.view(1)].clone()
        la = ra.step_logits(emb).float().clone()
        torch.cuda.synchronize()
        if verbose:
            print(f"  kernels step {s} done", flush=True)
        lb = rb.step_logits(emb).float().clone()
        torch.cuda.synchronize()
        if verbose:
            print(f"  stream step {s} done", flush=True)
        ref = model(input_ids=toks).logits[0, -1].float() if full_ref else la[0]
        pos = ctx + s
        dkv = max((a[0, pos // 32, :, pos % 32].float() - b[0, pos // 32, :, pos % 32].float()).abs().max().item()
                  for a, b in zip(ca.kv, cb.kv))
        dl = (la - lb).abs().max().item()
        scale = ref.abs().max().item()
        ea, eb = (la[0] - ref).abs().max().item(), (lb[0] - ref).abs().max().item()
        worst = max(worst, eb / max(scale, 1e-6))
        assert eb <= max(1.5 * ea, 1e-2 * scale), "stream path is further from the full forward than the per-kernel path"
        print(f"{label} ctx={ctx} step={s}: stream-kernels max|d|={dl:.3e}; vs full forward: kernels {ea:.3e}, stream {eb:.3e} "
              f"(|ref| {scale:.3f}); new kv rows max|d|={dkv:.3e}; finite={bool(torch.isfinite(lb).all())}", flush=True)
    return worst


def small():
    cfg = vicuna_7b_config(hidden_size=512, num_attention_heads=4, num_key_value_heads=4, intermediate_size=1408,
                           num_hidden_layers=3, vocab_size=1003)
    torch.manual_seed(0)
    model = MMGPTLlamaForCausalLM(cfg).to(device=DEV, dtype=torch.bfloat16)
    worst = 0.0
    for ctx in (5, 31, 32, 33, 100, 700, 1500):
        worst = max(worst, compare(model, ctx, label="small"))
    # wide MLP: a down_proj row no longer fits one ring slot and is streamed in K-chunks
    cfg = vicuna_7b_config(hidden_size=512, num_attention_heads=4, num_key_value_heads=4, intermediate_size=11008,
                           num_hidden_layers=2, vocab_size=1003)
    model = MMGPTLlamaForCausalLM(cfg).to(device=DEV, dtype=torch.bfloat16)
    for ctx in (33, 700):
        worst = max(worst, compare(model, ctx, label="small-wide"))
    print("small worst rel diff", worst)


def timed_replays(r, n):
    r.ensure_graph()
    r.pos.fill_(r.cache.seq_len)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        r.graph.replay()
    r.pos.fill_(r.cache.seq_len)
    e0.record()
    for _ in range(n):
        r.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def copy_bandwidth():
    """Reference point for box-to-box variance: torch copy of 1 GiB (read + write), best of 5."""
    a = torch.empty(1 << 29, dtype=torch.bfloat16, device=DEV)
    b = torch.empty_like(a)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * a.numel() * 2 / (best / 1e3) / 1e9


def full(ctx=4096, layers=32, check=True):
    cfg = vicuna_7b_config(num_hidden_layers=layers)
    torch.manual_seed(0)
    t0 = time.time()
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(DEV):
        model = MMGPTLlamaForCausalLM(cfg)
    torch.set_default_dtype(prev)
    print(f"model built in {time.time() - t0:.1f}s", flush=True)
    print(f"copy bandwidth {copy_bandwidth():.0f} GB/s", flush=True)
    if check:
        compare(model, ctx, steps=3, label="7B", verbose=os.environ.get("VERBOSE") == "1")
    for mode in os.environ.get("MODES", "kernels,stream").split(","):
        cache = MerlinKVCache(layers, 1, ctx + 128, cfg.hidden_size, DEV)
        for kv in cache.kv:
            kv.normal_(0, 0.5)
        cache.seq_len = ctx
        r = runner(model, cache, mode)
        ms = timed_replays(r, 32)
        gb = 13.48 + 0.524e-3 * ctx
        print(f"{mode}: {ms:.3f} ms/token at ctx {ctx} -> {gb / (ms / 1e3):.0f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    with torch.no_grad():
        if what == "small":
            small()
        elif what == "full":
            full(int(sys.argv[2]) if len(sys.argv) > 2 else 4096)
        elif what == "time":
            full(int(sys.argv[2]) if len(sys.argv) > 2 else 4096, check=False)
        else:
            small()
            full()
