"""Decode-step driver: binds a model + KV cache to merlin_decode_step and (for greedy decoding)
captures embed -> decoder -> argmax -> advance into one CUDA graph that is replayed per token, so the
launch-bound loop of 32 layers x 5 kernels costs one graph launch (SURVEY.md K12 / §3.3 hot loop #4).
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p

import torch

from . import lib as _lib
from . import ops as _ops
from .lib import DecoderWeights, check

import os

N_SPLITS = int(os.environ.get("MERLIN_FD_SPLITS", "32"))  # flash-decode key splits (x 32 heads CTAs per layer)


def decode_mode() -> str:
    """"stream": the whole token in ONE persistent kernel (csrc/decode_stream.cu); "kernels": one launch
    per GEMV / attention (csrc/decode.cu).  MERLIN_DECODE overrides the default."""
    mode = os.environ.get("MERLIN_DECODE", "stream")
    if mode not in ("stream", "kernels"):
        raise ValueError(f"MERLIN_DECODE={mode!r}: expected 'stream' or 'kernels'")
    return mode


def _ptr_array(tensors):
    arr = (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


class DecodeRunner:
    def __init__(self, lm, cache):
        m, cfg = lm.model, lm.config
        self.lm, self.cache = lm, cache
        dev = lm.device
        if cache.kv[0].shape[0] != 1:
            raise NotImplementedError("decode kernels handle batch 1")
        L = cfg.num_hidden_layers
        cos, sin = m.rope_tables(cache.capacity)
        self._keep = [_ptr_array(list(m.qkv_w)), _ptr_array(list(m.o_w)), _ptr_array(list(m.gate_up_w)),
                      _ptr_array(list(m.down_w)), _ptr_array(list(m.ln1_w)), _ptr_array(list(m.ln2_w)),
                      _ptr_array(cache.kv), cos, sin]
        cast = lambda a: ctypes.cast(a, ctypes.POINTER(c_void_p))
        self.w = DecoderWeights(L, cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size,
                                cfg.vocab_size, float(cfg.rms_norm_eps), cast(self._keep[0]), cast(self._keep[1]),
                                cast(self._keep[2]), cast(self._keep[3]), cast(self._keep[4]), cast(self._keep[5]),
                                m.norm_w.data_ptr(), lm.lm_head_w.data_ptr(), cast(self._keep[6]), cache.capacity,
                                cos.data_ptr(), sin.data_ptr())
        L_ = _lib.load()
        nbytes = L_.merlin_decode_workspace_bytes(cfg.hidden_size, cfg.intermediate_size,
                                                  cfg.num_attention_heads, N_SPLITS)
        # zeroed: the per-head arrival tickets of the flash-decode merge must start at 0 (the allocator may hand
        # back a block that an earlier runner left dirty)
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.x = torch.empty(cfg.hidden_size, dtype=torch.bfloat16, device=dev)
        self.logits = torch.empty((1, (cfg.vocab_size + 7) // 8 * 8), dtype=torch.bfloat16, device=dev)[:, :cfg.vocab_size]
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.token = torch.zeros(1, dtype=torch.int64, device=dev)
        self.n_out = torch.zeros(1, dtype=torch.int32, device=dev)
        self.out_ids = torch.zeros(max(cache.capacity, 16), dtype=torch.int64, device=dev)
        self.graph = None
        self.mode = decode_mode()
        if self.mode == "stream":
            sbytes = L_.merlin_decode_stream_workspace_bytes(L, cfg.hidden_size, cfg.num_attention_heads,
                                                             cfg.intermediate_size)
            self.stream_ws = torch.empty(sbytes + 256, dtype=torch.uint8, device=dev)
            off = (-self.stream_ws.data_ptr()) % 256
            self.stream_ws_ptr = self.stream_ws.data_ptr() + off
            check(L_.merlin_decode_stream_prepare(torch.cuda.current_stream().cuda_stream, ctypes.byref(self.w),
                                                  self.stream_ws_ptr), "merlin_decode_stream_prepare")
            self.launches_per_step = 1
        else:
            self.launches_per_step = 5 * L + 1  # 4 GEMVs + 1 flash-decode per layer, + lm_head

    def _decode_step(self, st):
        """One decoder pass for the embedding in self.x at position *self.pos -> self.logits."""
        L_ = _lib.load()
        if self.mode == "stream":
            check(L_.merlin_decode_stream_step(st, ctypes.byref(self.w), self.x.data_ptr(), self.pos.data_ptr(),
                                               self.stream_ws_ptr, self.logits.data_ptr()),
                  "merlin_decode_stream_step")
        else:
            check(L_.merlin_decode_step(st, ctypes.byref(self.w), self.x.data_ptr(), self.pos.data_ptr(),
                                        self.ws.data_ptr(), self.logits.data_ptr(), N_SPLITS), "merlin_decode_step")

    # -- one decoder pass for an externally supplied embedding (used by forward(past_key_values=...))
    def step_logits(self, embeds: torch.Tensor) -> torch.Tensor:
        """embeds [1, D] -> logits [1, V]; appends to the cache at position cache.seq_len."""
        if self.cache.seq_len >= self.cache.capacity:
            raise RuntimeError("KV cache is full")
        self.x.copy_(embeds.reshape(-1))
        self.pos.fill_(self.cache.seq_len)
        self._decode_step(torch.cuda.current_stream().cuda_stream)
        self.cache.seq_len += 1
        _ops.PROFILE.launches += self.launches_per_step
        return self.logits

    # -- device-resident greedy loop
    def _greedy_body(self):
        L_ = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        cfg = self.lm.config
        check(L_.merlin_embed_token(st, self.lm.model.embed_tokens.data_ptr(), self.token.data_ptr(),
                                    self.x.data_ptr(), cfg.hidden_size, cfg.vocab_size), "merlin_embed_token")
        self._decode_step(st)
        check(L_.merlin_argmax(st, self.logits.data_ptr(), self.logits.stride(0), 1, cfg.vocab_size,
                               self.token.data_ptr()), "merlin_argmax")
        check(L_.merlin_record_advance(st, self.token.data_ptr(), self.out_ids.data_ptr(),
                                       self.n_out.data_ptr(), self.pos.data_ptr(), self.out_ids.numel()),
              "merlin_record_advance")

    def ensure_graph(self):
        """Capture one greedy step (embed -> decoder -> argmax -> record/advance).  The warm-up run
        outside capture scribbles on cache row `pos`; callers (re)fill the cache afterwards."""
        if self.graph is not None:
            return
        self.pos.fill_(min(self.cache.seq_len, self.cache.capacity - 1))
        self._greedy_body()  # warm-up outside capture (one-time kernel attributes)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._greedy_body()
        self.graph = g

    def greedy(self, first_token: torch.Tensor, n_tokens: int, eos_token_id=None, check_every: int = 16,
               use_graph: bool = True) -> torch.Tensor:
        """Feeds `first_token` (already chosen from the prefill logits) and decodes n_tokens more
        greedily without host round trips; returns the generated ids (truncated after eos)."""
        if use_graph:
            self.ensure_graph()
        self.token.copy_(first_token.reshape(1))
        self.pos.fill_(self.cache.seq_len)
        self.n_out.zero_()
        return self.run(n_tokens, eos_token_id, check_every, use_graph, base=0)

    def run(self, n_tokens: int, eos_token_id=None, check_every: int = 16, use_graph: bool = True,
            base: int = 0) -> torch.Tensor:
        """Decode n_tokens more from the device-resident (token, pos, n_out) state; `base` tokens were
        already recorded in out_ids by the caller.  Returns out_ids[base : base + done]."""
        n_tokens = min(n_tokens, self.cache.capacity - self.cache.seq_len)
        done = 0
        while done < n_tokens:
            chunk = min(check_every, n_tokens - done)
            for _ in range(chunk):
                if use_graph:
                    self.graph.replay()
                else:
                    self._greedy_body()
                _ops.PROFILE.launches += self.launches_per_step + 3
            done += chunk
            if eos_token_id is not None:
                ids = self.out_ids[base:base + done]
                hit = (ids == eos_token_id).nonzero()
                if hit.numel():
                    done = int(hit[0]) + 1
                    break
        self.cache.seq_len += done
        return self.out_ids[base:base + done].clone()


def beam_splits(heads: int) -> int:
    """Prefix key splits of the beam attention kernel: heads x splits prefix CTAs fit ONE wave of two CTAs per SM —
    one more CTA row would wait for a second wave and double the kernel's time.  (Three CTAs per SM with 12 splits
    measured slower: 3.62 vs 3.52 ms per position at 5 beams.)  The per-beam suffix CTAs are short and scheduled first."""
    env = os.environ.get("MERLIN_BEAM_SPLITS")
    if env:
        return int(env)
    sms = _lib.load().merlin_sm_count()
    return max(1, min(16, (2 * sms) // heads - 1))


MAX_BEAMS = 8


class BeamDecodeRunner:
    """All beams of a beam search advance one position per `step` (merlin_decode_step_batched): the weights are
    streamed once for the k beams, the prompt's KV cache (`prefix`, written by the prefill) is shared, and only each
    beam's own generated tokens live in the per-beam `suffix` cache [L, k, cap/32, H, 32, 256] that a re-ordering has
    to move.  embed -> decoder pass -> position advance is one CUDA graph; positions live in device memory, so the
    graph serves every prompt that fits the prefix cache."""

    def __init__(self, lm, prefix, num_beams: int, suffix_capacity: int):
        if not 1 <= num_beams <= MAX_BEAMS:
            raise NotImplementedError(f"the batched beam decode handles 1..{MAX_BEAMS} beams (got {num_beams}); "
                                      "use use_hf_loop=True")
        m, cfg = lm.model, lm.config
        dev = lm.device
        self.lm, self.prefix, self.nb = lm, prefix, num_beams
        L, H, D = cfg.num_hidden_layers, cfg.num_attention_heads, cfg.hidden_size
        self.cap = (suffix_capacity + 31) // 32 * 32
        self.suffix = torch.zeros((L, num_beams, self.cap // 32, H, 32, 256), dtype=torch.bfloat16, device=dev)
        cos, sin = m.rope_tables(prefix.capacity + self.cap)
        self._keep = [_ptr_array(list(m.qkv_w)), _ptr_array(list(m.o_w)), _ptr_array(list(m.gate_up_w)),
                      _ptr_array(list(m.down_w)), _ptr_array(list(m.ln1_w)), _ptr_array(list(m.ln2_w)),
                      _ptr_array(prefix.kv), _ptr_array([self.suffix[i] for i in range(L)]), cos, sin]
        cast = lambda a: ctypes.cast(a, ctypes.POINTER(c_void_p))
        self.w = DecoderWeights(L, D, H, cfg.intermediate_size, cfg.vocab_size, float(cfg.rms_norm_eps),
                                cast(self._keep[0]), cast(self._keep[1]), cast(self._keep[2]), cast(self._keep[3]),
                                cast(self._keep[4]), cast(self._keep[5]), m.norm_w.data_ptr(), lm.lm_head_w.data_ptr(),
                                cast(self._keep[6]), prefix.capacity, cos.data_ptr(), sin.data_ptr())
        self.pos = torch.zeros(2, dtype=torch.int32, device=dev)          # {prefix_len, t}
        self.bs = _lib.BeamState(num_beams, self.cap, cast(self._keep[6]), cast(self._keep[7]), self.pos.data_ptr())
        L_ = _lib.load()
        self.splits = beam_splits(H)
        self.ws = torch.zeros(L_.merlin_decode_batched_workspace_bytes(D, cfg.intermediate_size, H, self.splits,
                                                                       num_beams), dtype=torch.uint8, device=dev)
        self.ctl = torch.zeros((2, num_beams), dtype=torch.int64, device=dev)     # [tokens | parents] of this step
        self.ctl_host = torch.zeros((2, num_beams), dtype=torch.int64).pin_memory()
        self.tokens = self.ctl[0]
        self._ctl_ev = torch.cuda.Event()
        self.x = torch.empty((num_beams, D), dtype=torch.bfloat16, device=dev)
        vpad = (cfg.vocab_size + 7) // 8 * 8
        self.logits = torch.empty((num_beams, vpad), dtype=torch.bfloat16, device=dev)[:, :cfg.vocab_size]
        self.one = torch.tensor([0, 1], dtype=torch.int32, device=dev)
        self.t = 0
        self.graph = None
        self.launches_per_step = 7 * L + 2      # per layer: 2 norms, 4 projections, attention; final norm + lm_head
        self.fingerprint = weight_fingerprint(lm)
        if os.environ.get("MERLIN_BEAM_GRAPH", "1") == "1":
            self.ensure_graph()

    def _body(self):
        torch.index_select(self.lm.model.embed_tokens, 0, self.tokens, out=self.x)
        check(_lib.load().merlin_decode_step_batched(torch.cuda.current_stream().cuda_stream, ctypes.byref(self.w),
                                                     ctypes.byref(self.bs), self.x.data_ptr(), self.ws.data_ptr(),
                                                     self.logits.data_ptr(), self.logits.stride(0), self.splits),
              "merlin_decode_step_batched")
        self.pos.add_(self.one)

    def ensure_graph(self):
        if self.graph is not None:
            return
        self.pos.zero_()
        self._body()                      # warm-up outside capture; scribbles on suffix row 0 (no search is running yet)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._body()
        self.graph = g

    def begin(self, prefix_len: int):
        if prefix_len > self.prefix.capacity:
            raise RuntimeError("prompt longer than the prefix cache")
        self.pos.copy_(torch.tensor([prefix_len, 0], dtype=torch.int32), non_blocking=False)
        self.t = 0

    def step(self, tokens, parents=None, use_graph: bool = True) -> torch.Tensor:
        """tokens[k]: the token each beam appends; parents[k]: the beam (of the previous step) it descends from.
        Returns the [k, V] bf16 logits of the next position (a view of a static buffer)."""
        if self.t >= self.cap:
            raise RuntimeError("beam suffix cache is full")
        ident = list(range(self.nb))
        parents = ident if parents is None else list(parents)
        # one pinned staging buffer, one async copy (the event guards the buffer against a caller that does not read
        # a step's logits back before the next step)
        self._ctl_ev.synchronize()
        self.ctl_host[0] = torch.as_tensor(list(tokens), dtype=torch.int64)
        self.ctl_host[1] = torch.as_tensor(parents, dtype=torch.int64)
        self.ctl.copy_(self.ctl_host, non_blocking=True)
        self._ctl_ev.record()
        if self.t > 0 and parents != ident:
            nblk = (self.t + 31) // 32      # only the blocks that hold generated tokens move
            self.suffix[:, :, :nblk].copy_(self.suffix[:, :, :nblk].index_select(1, self.ctl[1]))
        if use_graph and self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        self.t += 1
        _ops.PROFILE.launches += self.launches_per_step + 2
        return self.logits


class DecodeState:
    """KV cache + DecodeRunner (+ its per-token CUDA graph) for one capacity bucket.  The decode graph reads the
    position from device memory, so ONE state serves every prompt length that fits the bucket."""

    def __init__(self, lm, capacity: int):
        from .mmgpt.model.mmgpt.llama_mmgpt import MerlinKVCache
        cfg = lm.config
        self.capacity = capacity
        self.cache = MerlinKVCache(cfg.num_hidden_layers, 1, capacity, cfg.hidden_size, lm.device)
        self.runner = DecodeRunner(lm, self.cache)
        self.runner.ensure_graph()


class GraphedGenerator:
    """Multimodal prefill for one (sequence length, image batch) shape: ViT -> projector -> splice -> decoder
    layers -> lm_head -> argmax of the last row, into the KV cache of a shared DecodeState.  The first call for a
    shape runs the launches eagerly; a shape that comes back is captured into a CUDA graph (static input buffers)
    and replayed from then on — evaluation prompts rarely repeat a length, benchmark / serving shapes do."""

    def __init__(self, lm, S: int, image_shape, state: DecodeState):
        dev = lm.device
        self.lm, self.S, self.state = lm, S, state
        self.cache, self.runner = state.cache, state.runner
        self.ids = torch.zeros((1, S), dtype=torch.int64, device=dev)
        self.img = torch.zeros(tuple(image_shape), dtype=torch.bfloat16, device=dev)
        tower = lm.get_model().vision_tower
        rows = image_shape[0] * tower.num_patches
        self.feat_offset = torch.zeros(1, dtype=torch.int32, device=dev)
        self.feat_limit = torch.full((1,), rows, dtype=torch.int32, device=dev)
        self.scratch_pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.per_image = tower.num_patches
        self.launches = 0
        self.graph = None
        self.calls = 0

    def capture(self):
        self._body()                       # warm-up outside capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            self._body()
        self.graph = g

    def _body(self):
        L_ = _lib.load()
        r = self.runner
        before = _ops.PROFILE.launches
        logits2d = self.lm._prefill_core(self.ids, self.img, self.feat_offset, self.feat_limit, self.cache)
        st = torch.cuda.current_stream().cuda_stream
        V = self.lm.config.vocab_size
        last = logits2d[self.S - 1:self.S]
        check(L_.merlin_argmax(st, last.data_ptr(), logits2d.stride(0), 1, V, r.token.data_ptr()), "merlin_argmax")
        r.n_out.zero_()
        r.pos.fill_(self.S)
        check(L_.merlin_record_advance(st, r.token.data_ptr(), r.out_ids.data_ptr(), r.n_out.data_ptr(),
                                       self.scratch_pos.data_ptr(), r.out_ids.numel()), "merlin_record_advance")
        self.launches = _ops.PROFILE.launches - before + 2
        self.logits2d = logits2d

    def generate(self, input_ids: torch.Tensor, images: torch.Tensor, max_new_tokens: int, eos_token_id=None,
                 capture_after: int = 1):
        ids_host = input_ids if not input_ids.is_cuda else input_ids.cpu()
        self.lm._validate_image_grammar(ids_host, [images], self.per_image)
        self.ids.copy_(input_ids, non_blocking=True)
        self.img.copy_(images, non_blocking=True)
        self.calls += 1
        if self.graph is None and self.calls > capture_after:
            self.capture()
        if self.graph is not None:
            self.graph.replay()
            _ops.PROFILE.launches += self.launches
        else:
            self._body()
        self.cache.seq_len = self.S
        new = None
        if max_new_tokens > 1:
            if eos_token_id is not None and int(self.runner.out_ids[0]) == eos_token_id:
                pass
            else:
                self.runner.run(max_new_tokens - 1, eos_token_id, base=1)
        n = int(self.cache.seq_len - self.S) + 1
        new = self.runner.out_ids[:n]
        return torch.cat([self.ids, new.view(1, -1)], dim=1)


def weight_fingerprint(lm) -> int:
    """Changes whenever a parameter the captured graphs / the streaming kernel's pointer tables refer to moves
    (e.g. the trainer re-homes the parameters into its flat arena)."""
    m = lm.model
    probes = (m.embed_tokens, lm.lm_head_w, m.norm_w, m.qkv_w[0], m.o_w[0], m.gate_up_w[0], m.down_w[0],
              m.qkv_w[-1], m.down_w[-1], m.ln1_w[0], m.ln2_w[-1])
    return hash(tuple(p.data_ptr() for p in probes))


CAPACITY_BUCKET = 1024      # KV-cache capacities are rounded up to a multiple of this
MAX_DECODE_STATES = 2       # capacity buckets kept alive (each holds a full KV cache)
MAX_PREFILL_SHAPES = 8      # (length, image shape) entries kept (static input buffers; a graph once a shape repeats)


def graphed_generate(lm, input_ids, images, max_new_tokens, eos_token_id, capture_after: int = 1):
    """Greedy generate() from device-resident state.  ONE KV cache + decode runner + per-token graph per capacity
    bucket serves every prompt length (evaluation prompts almost never repeat a length); the prefill of a shape
    is captured into a CUDA graph only once that shape has been seen `capture_after` times."""
    pool = lm.__dict__.setdefault("_graphed", {})
    states = lm.__dict__.setdefault("_decode_states", {})
    fp = weight_fingerprint(lm)
    if lm.__dict__.get("_graphed_fp") != fp:      # graphs captured for other weight addresses are stale
        pool.clear()
        states.clear()
        lm.__dict__["_graphed_fp"] = fp
    need = input_ids.shape[1] + max_new_tokens + 8
    cap = (need + CAPACITY_BUCKET - 1) // CAPACITY_BUCKET * CAPACITY_BUCKET
    state = states.get(cap)
    if state is None:
        bigger = [c for c in states if c >= cap]
        if bigger:                                 # a larger bucket already exists: use it
            cap = min(bigger)
            state = states[cap]
        else:
            while len(states) >= MAX_DECODE_STATES:
                old = next(iter(states))
                states.pop(old)
                for k in [k for k in pool if k[2] == old]:
                    pool.pop(k)
            state = states[cap] = DecodeState(lm, cap)
    key = (int(input_ids.shape[1]), tuple(images.shape), cap)
    gen = pool.get(key)
    if gen is None:
        while len(pool) >= MAX_PREFILL_SHAPES:
            pool.pop(next(iter(pool)))
        gen = pool[key] = GraphedGenerator(lm, int(input_ids.shape[1]), images.shape, state)
    return gen.generate(input_ids, images, max_new_tokens, eos_token_id, capture_after)
