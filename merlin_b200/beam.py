"""Beam search bookkeeping for generate(num_beams > 1) — host-side control flow, batch size 1.

The reference reaches beam search through HF `generate(num_beams=5)` (mmgpt/engine/eval/eval_box.py:246-254,
with the `repeat_interleave(5)` contract of base_mmgpt.py:162-163); the algorithm itself lives in the
third-party library (`transformers`, pyproject.toml:22).  This module restates `GenerationMixin._beam_search`
of the installed transformers 5.5.0 (generation/utils.py:2876-3370 — the executable spec SURVEY.md §8c names)
for one sequence, and tests/test_generate_cpu.py pins it against HF `generate` itself:

  every step   cand[b, v] = log_softmax(logits[b])[v] + running_score[b]; the 2·num_beams best (b, v) pairs are
               the candidates; a candidate "hits" when its token is EOS, when it reaches max_length or when a
               stopping criterion fires on it;
  running      the num_beams best candidates that did not hit continue;
  finished     candidates that hit AND rank inside the first num_beams are scored
               sum_logprobs / generated_length**length_penalty and merged into the num_beams best finished
               hypotheses (EOS included in the hypothesis);
  stop         when the best running beam can no longer beat the worst finished one (heuristic of
               `_check_early_stop_heuristic`, sticky), when early_stopping=True and all slots are filled, or when
               every candidate hit;
  result       the best finished hypothesis.

transformers 4.31 (the version the reference pins) normalises by the TOTAL length (prompt included) instead of
the generated length; `length_norm="total"` switches the two denominators, everything else is shared.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

NEG = -1.0e9


class BeamSearch:
    """Call `step(logits)` once per decoded position with the [num_beams, V] logits of the running beams (all
    rows equal at the first step).  It returns (tokens, parents): the token each new running beam appends and
    the running beam it descends from (for the KV-cache reordering), or None when the search is over.

    Host cost per step is kept small because the GPU waits for it: hypotheses are Python lists of the GENERATED
    tokens (the shared prompt is attached only for stopping criteria and at the end), log-softmax + top-2k run where
    the logits live and come back in one transfer; the score arithmetic stays in fp32 tensors (HF's dtype)."""

    def __init__(self, prompt: torch.Tensor, num_beams: int, eos_token_id: Optional[int], max_new_tokens: int,
                 length_penalty: float = 1.0, early_stopping=False, length_norm: str = "generated",
                 stopping_criteria: Optional[Sequence[Callable]] = None):
        if length_norm not in ("total", "generated"):
            raise ValueError("length_norm must be 'total' or 'generated'")
        self.nb, self.eos = num_beams, eos_token_id
        self.prompt = prompt.reshape(1, -1).cpu().to(torch.int64)
        self.prompt_len = int(self.prompt.shape[-1])
        self.max_len = self.prompt_len + max_new_tokens
        self.lp, self.early_stopping, self.norm = length_penalty, early_stopping, length_norm
        self.criteria = list(stopping_criteria or [])
        self.keep = 2 * num_beams                                  # max(2, 1 + n_eos) * num_beams
        self.gen: List[List[int]] = [[] for _ in range(num_beams)]  # generated tokens of the running beams
        self.running_scores = torch.zeros(num_beams, dtype=torch.float32)
        self.running_scores[1:] = NEG                              # all beams start identical
        self.fin_gen: List[List[int]] = [[] for _ in range(num_beams)]
        self.fin_scores = torch.full((num_beams,), NEG, dtype=torch.float32)
        self.fin_flags = torch.zeros(num_beams, dtype=torch.bool)
        self.unsat = True                                          # early-stop heuristic not yet satisfied
        self.done = False

    def _full(self, gen: List[int]) -> torch.Tensor:
        return torch.cat([self.prompt, torch.tensor([gen], dtype=torch.int64)], dim=1)

    @property
    def running(self) -> torch.Tensor:
        """[num_beams, prompt + generated] ids of the running beams."""
        return torch.cat([self._full(g) for g in self.gen], dim=0)

    def _len(self, total_len: int) -> int:
        return total_len if self.norm == "total" else total_len - self.prompt_len

    def step(self, logits: torch.Tensor):
        if self.done:
            return None
        nb, keep = self.nb, self.keep
        cur_len = self.prompt_len + len(self.gen[0])
        # log-softmax and the top-2k selection run where the logits live (the GPU during generate): only the 2k
        # candidates cross to the host, in one transfer (an index < 2^24 is exact in fp32)
        logp = torch.log_softmax(logits.detach().to(dtype=torch.float32), dim=-1)
        V = logp.shape[-1]
        cand = (logp + self.running_scores.to(logp.device, non_blocking=True)[:, None]).reshape(-1)
        top_scores, top_idx = torch.topk(cand, k=keep)
        if nb * V < (1 << 24):
            both = torch.stack([top_scores, top_idx.to(torch.float32)]).cpu()
            top_scores, top_idx = both[0], both[1].to(torch.int64)
        else:
            top_scores, top_idx = top_scores.cpu(), top_idx.cpu()
        parents, tokens = (top_idx // V).tolist(), (top_idx % V).tolist()
        cand_gen = [self.gen[parents[i]] + [tokens[i]] for i in range(keep)]
        hits = torch.zeros(keep, dtype=torch.bool)
        if self.eos is not None:
            hits |= torch.tensor([t == self.eos for t in tokens])
        if cur_len + 1 >= self.max_len:
            hits[:] = True
        for crit in self.criteria:                                  # evaluated per candidate sequence
            for i in range(keep):
                r = crit(self._full(cand_gen[i]), None)
                hits[i] |= bool(r.all()) if torch.is_tensor(r) else bool(r)

        # running beams of the next iteration
        run_scores = top_scores + hits.float() * NEG
        nxt = torch.topk(run_scores, k=nb)[1].tolist()
        new_scores = run_scores[nxt]
        out_tokens, out_parents = [tokens[i] for i in nxt], [parents[i] for i in nxt]

        # finished hypotheses
        fin = top_scores / (self._len(cur_len + 1) ** self.lp)
        if bool(self.fin_flags.all()) and self.early_stopping is True:
            fin = fin + NEG
        if not self.unsat:
            fin = fin + NEG
        just_finished = hits & (torch.arange(keep) < nb)
        fin = fin + (~just_finished).float() * NEG
        merged_scores = torch.cat([self.fin_scores, fin])
        merged_flags = torch.cat([self.fin_flags, just_finished])
        merged_gen = self.fin_gen + cand_gen
        sel = torch.topk(merged_scores, k=nb)[1].tolist()
        self.fin_scores, self.fin_flags = merged_scores[sel], merged_flags[sel]
        self.fin_gen = [merged_gen[i] for i in sel]

        self.gen, self.running_scores = [cand_gen[i] for i in nxt], new_scores
        cur_len += 1
        # early-stop heuristic (generation/utils.py:2876-2921)
        if self.early_stopping == "never" and self.lp > 0.0:
            hyp_len = self._len(self.max_len)
        else:
            hyp_len = self._len(cur_len)
        best_possible = float(self.running_scores[0]) / (hyp_len ** self.lp)
        worst_finished = torch.where(self.fin_flags, self.fin_scores.min(), torch.tensor(NEG))
        self.unsat = self.unsat and bool((best_possible > worst_finished).any())
        open_beam = not (bool(self.fin_flags.all()) and self.early_stopping is True)
        valid = not bool(hits.all())
        if not (self.unsat and open_beam and valid):
            self.done = True
            return None
        return out_tokens, out_parents

    def finalize(self) -> torch.Tensor:
        """Best finished hypothesis as a [1, len] id tensor (prompt included)."""
        return self._full(self.fin_gen[0])
