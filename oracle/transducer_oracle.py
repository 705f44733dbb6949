"""CPU oracle for the transducer greedy search that follows the encoder on `chunkformer-rnnt-*` models.

TEST INFRASTRUCTURE ONLY (same rules as chunkformer_oracle.py: imported by tests/ and bench tools as the checker, never
by the product package).  From-scratch restatement, in plain fp32 torch-CPU algebra and Python loops, of

  * RNNPredictor.forward_step      /root/reference/chunkformer/transducer/predictor.py:190-207  (embedding -> LSTM stack
    -> projection; torch.nn.LSTM cell: gate order i, f, g, o)
  * TransducerJoint.forward        /root/reference/chunkformer/transducer/joint.py:69-101      (prejoin linears, add, tanh,
    ffn_out; the non-HAT branch every shipped config uses)
  * optimized_search / batch_greedy_search  /root/reference/chunkformer/transducer/search/greedy_search.py:6-92

Pinning: tests/golden/make_golden_transducer.py runs the UNMODIFIED reference classes on synthetic weights and
tests/test_transducer_oracle.py checks this file against tests/golden/transducer.npz (token grids bit-exact).

Because every utterance of `optimized_search` only ever touches its own row of the state, the batched search is restated
per utterance: with (input token u, LSTM state s) the predictor output p = P(u, s) and the candidate state s' are fixed
until the next non-blank symbol; frame t, step n (1..n_steps) emits k = argmax joint(enc_t, p); k != blank -> record it
at column n-1 of frame t, (u, s) <- (k, s'), stay on the frame unless n == n_steps; k == blank -> next frame, step 1.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

State = List[Tuple[torch.Tensor, torch.Tensor]]     # per layer (h, c), each (hidden,)


def num_layers(sd: Dict[str, torch.Tensor]) -> int:
    n = 0
    while f"predictor.rnn.weight_ih_l{n}" in sd:
        n += 1
    return n


def init_state(sd) -> State:
    """predictor.py:176-188: zeros (n_layers, hidden)."""
    hidden = sd["predictor.rnn.weight_hh_l0"].shape[1]
    return [(torch.zeros(hidden), torch.zeros(hidden)) for _ in range(num_layers(sd))]


def predictor_step(sd, token: int, state: State):
    """predictor.py:190-207 for one utterance: returns (projection output (P,), new state)."""
    x = sd["predictor.embed.weight"][int(token)]
    new_state: State = []
    for layer, (h, c) in enumerate(state):
        gates = (sd[f"predictor.rnn.weight_ih_l{layer}"] @ x + sd[f"predictor.rnn.bias_ih_l{layer}"]
                 + sd[f"predictor.rnn.weight_hh_l{layer}"] @ h + sd[f"predictor.rnn.bias_hh_l{layer}"])
        i, f, g, o = gates.chunk(4)
        c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h2 = torch.sigmoid(o) * torch.tanh(c2)
        new_state.append((h2, c2))
        x = h2
    out = sd["predictor.projection.weight"] @ x + sd["predictor.projection.bias"]
    return out, new_state


def joint_logits(sd, enc_t: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """joint.py:80-101 (prejoin_linear, joint_mode add, tanh, ffn_out). enc_t (E,) or (n, E); pred (P,)."""
    e = enc_t @ sd["joint.enc_ffn.weight"].T + sd["joint.enc_ffn.bias"]
    p = sd["joint.pred_ffn.weight"] @ pred + sd["joint.pred_ffn.bias"]
    return torch.tanh(e + p) @ sd["joint.ffn_out.weight"].T + sd["joint.ffn_out.bias"]


def greedy_search_one(sd, enc: torch.Tensor, n_frames: int, n_steps: int = 64, blank: int = 0, want_margin: bool = False):
    """One utterance of optimized_search (greedy_search.py:6-80): (T, n_steps) int64 grid, blank where nothing was emitted
    [, fp32 top-2 logit margin of every decision as a list of (frame, step, margin)]."""
    T = enc.shape[0]
    grid = torch.full((T, n_steps), blank, dtype=torch.int64)
    margins = []
    token, state = blank, init_state(sd)
    pred, cand = predictor_step(sd, token, state)
    for t in range(min(int(n_frames), T)):
        for step in range(1, n_steps + 1):
            logits = joint_logits(sd, enc[t], pred)
            k = int(torch.argmax(torch.log_softmax(logits, dim=-1)))
            if want_margin:
                top = torch.topk(logits, 2).values
                margins.append((t, step, float(top[0] - top[1])))
            if k == blank:
                break
            grid[t, step - 1] = k
            token, state = k, cand
            pred, cand = predictor_step(sd, token, state)
    return (grid, margins) if want_margin else grid


def optimized_search(sd, encoder_out: torch.Tensor, encoder_out_lens: Sequence[int], n_steps: int = 64, blank: int = 0):
    """greedy_search.py:6-80: (B, T * n_steps) int64."""
    B, T, _ = encoder_out.shape
    return torch.stack([greedy_search_one(sd, encoder_out[b], int(encoder_out_lens[b]), n_steps, blank).reshape(-1)
                        for b in range(B)])


def batch_greedy_search(sd, encoder_out, encoder_out_lens, n_steps: int = 64, blank: int = 0) -> List[List[int]]:
    """greedy_search.py:84-99: non-blank symbols of every utterance in emission order."""
    out = optimized_search(sd, encoder_out, encoder_out_lens, n_steps, blank)
    return [row[row != blank].tolist() for row in out]
