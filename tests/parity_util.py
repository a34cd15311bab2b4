"""Helpers of the GPU parity tests: greedy token streams are compared with the oracle position by position, teacher-forced.

Why teacher-forced: the CUDA path computes with fp16 operands (fp32 accumulate), the oracle in fp32. Over thousands of
arg-max decisions on seeded random weights some top-2 margins fall below the fp16 logit error, and a flip there changes
the rest of that sequence. So besides the plain comparison with `oracle.greedy` (identical until the first such flip)
the oracle is run on the tokens the GPU chose, and every choice must be the oracle's arg-max or lie within TOL_TIE of
it — a tie the reference's own fp16 CoreML graph could resolve either way. TOL_TIE sits below the stated logit tolerance.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch
import torch.nn.functional as F

TOL_TIE = 2e-2        # |oracle max logit - oracle logit of the GPU's token| that still counts as a tie (logit tolerance: 5e-2)


@dataclass
class GreedyReport:
    decisions: int
    exact: int                 # GPU token == oracle arg-max given the GPU's own prefix
    ties: int                  # not the arg-max, but within TOL_TIE of it
    bad: int                   # neither
    worst_gap: float           # largest (oracle max - oracle logit of the GPU token)
    min_top2_margin: float     # smallest oracle top-1/top-2 margin over all decisions
    sum_logprob: np.ndarray    # oracle log-prob of the GPU's tokens, summed per sequence (EOT rule applied)
    bad_at: List[tuple]

    def line(self) -> str:
        return (f"{self.decisions} decisions: {self.exact} exact, {self.ties} ties (gap <= {TOL_TIE}), {self.bad} bad; worst gap "
                f"{self.worst_gap:.3e}, min oracle top-2 margin {self.min_top2_margin:.3e}")


def teacher_forced_check(oracle, xa_ref: torch.Tensor, tokens: np.ndarray, n_init: int, suppress, suppress_begin, eot: int,
                         batch: int = 4) -> GreedyReport:
    """tokens [B, L] from the GPU (prompt included). Runs the oracle decoder on tokens[:, :-1] in slices of `batch`
    sequences, applies the static logit filters, and grades every sampled position."""
    tok = torch.from_numpy(np.ascontiguousarray(tokens).astype(np.int64))
    B, L = tok.shape
    sup = torch.tensor(sorted(set(int(s) for s in suppress)), dtype=torch.long)
    supb = torch.tensor(sorted(set(int(s) for s in suppress_begin)), dtype=torch.long)
    exact = ties = bad = 0
    worst, min_margin = 0.0, float("inf")
    slp = np.zeros(B, dtype=np.float64)
    bad_at = []
    for b0 in range(0, B, batch):
        sl = slice(b0, min(B, b0 + batch))
        lg = oracle.decoder_logits(tok[sl, :-1], xa_ref[sl])[:, n_init - 1:]          # [b, L - n_init, V]: position p predicts token p+1
        if len(sup):
            lg[:, :, sup] = float("-inf")
        if len(supb):
            lg[:, 0, supb] = float("-inf")
        chosen = tok[sl, n_init:]                                                      # [b, L - n_init]
        top2 = lg.topk(2, dim=-1).values
        got = lg.gather(-1, chosen[..., None])[..., 0]
        gap = top2[..., 0] - got
        lp = got - torch.logsumexp(lg.float(), dim=-1)
        prev = tok[sl, n_init - 1:-1]
        live = prev != eot                                                             # after an EOT the row is forced to EOT, no log-prob
        for i in range(chosen.shape[0]):
            for p in range(chosen.shape[1]):
                if not bool(live[i, p]):
                    assert int(chosen[i, p]) == eot, f"sequence {b0 + i}: token after EOT at position {n_init + p} is not EOT"
                    continue
                g = float(gap[i, p])
                min_margin = min(min_margin, float(top2[i, p, 0] - top2[i, p, 1]))
                worst = max(worst, g)
                if g == 0.0:
                    exact += 1
                elif g <= TOL_TIE:
                    ties += 1
                else:
                    bad += 1
                    bad_at.append((b0 + i, n_init + p, g))
                slp[b0 + i] += float(lp[i, p])
    return GreedyReport(exact + ties + bad, exact, ties, bad, worst, min_margin, slp, bad_at)


def first_divergence(tokens: np.ndarray, tok_ref: torch.Tensor):
    """Per sequence: index of the first position where the GPU stream and oracle.greedy differ (-1: identical)."""
    n = min(tokens.shape[1], tok_ref.shape[1])
    diff = torch.from_numpy(tokens[:, :n].astype(np.int64)) != tok_ref[:, :n]
    out = []
    for b in range(diff.shape[0]):
        nz = diff[b].nonzero()
        out.append(int(nz[0]) if nz.numel() else -1)
    return out
