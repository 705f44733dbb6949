"""Deterministic synthetic checkpoints and fbank workloads.

There is no network for real checkpoints, so parity and benchmarks run on random weights in the
reference's checkpoint layout (flat state_dict of the inner model):
the same (geometry, seed) gives the same tensors here, in the fixture generator that loads
them into the unmodified reference (tests/golden/make_golden.py) and on the GPU box.
Unlike the HF wrapper's own init (N(0, 0.02) weights, zero biases), every bias and norm
parameter is non-trivial so that no term of the path is silently skipped by a test.
"""
import math
from typing import Dict, List

import torch

from .geometry import EncoderGeometry


def _u(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _n(gen, shape, std, mean=0.0):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std + mean


def synth_state_dict(geo: EncoderGeometry, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random weights keyed exactly like the reference checkpoint (encoder.* and ctc.ctc_lo.*)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    d, H, F, K, V = geo.d_model, geo.heads, geo.ffn, geo.kernel, geo.vocab
    sd: Dict[str, torch.Tensor] = {}

    def linear(prefix, out_f, in_f, bias=True, shape=None):
        b = 1.0 / math.sqrt(in_f)
        sd[prefix + ".weight"] = _u(g, shape or (out_f, in_f), b)
        if bias:
            sd[prefix + ".bias"] = _u(g, (out_f,), b)

    def norm(prefix):
        sd[prefix + ".weight"] = _n(g, (d,), 0.1, 1.0)
        sd[prefix + ".bias"] = _n(g, (d,), 0.1)

    if geo.has_cmvn:
        sd["encoder.global_cmvn.mean"] = _n(g, (geo.feat_dim,), 0.5)
        sd["encoder.global_cmvn.istd"] = 1.0 / (1.0 + 0.2 * torch.rand((geo.feat_dim,), generator=g))
    e = "encoder.embed."
    linear(e + "conv.0", d, 9, shape=(d, 1, 3, 3))
    linear(e + "conv.2", d, 9, shape=(d, 1, 3, 3))
    linear(e + "conv.3", d, d, shape=(d, d, 1, 1))
    linear(e + "conv.5", d, 9, shape=(d, 1, 3, 3))
    linear(e + "conv.6", d, d, shape=(d, d, 1, 1))
    freq = geo.feat_dim
    for _ in range(3):
        freq = (freq - 3) // 2 + 1
    linear(e + "out", d, d * freq)
    for i in range(geo.layers):
        p = f"encoder.encoders.{i}."
        for name in ("linear_q", "linear_k", "linear_v", "linear_out"):
            linear(p + "self_attn." + name, d, d)
        linear(p + "self_attn.linear_pos", d, d, bias=False)
        sd[p + "self_attn.pos_bias_u"] = _n(g, (H, d // H), 0.2)
        sd[p + "self_attn.pos_bias_v"] = _n(g, (H, d // H), 0.2)
        for ff in ("feed_forward", "feed_forward_macaron"):
            linear(p + ff + ".w_1", F, d)
            linear(p + ff + ".w_2", d, F)
        linear(p + "conv_module.pointwise_conv1", 2 * d, d, shape=(2 * d, d, 1))
        linear(p + "conv_module.depthwise_conv", d, K, shape=(d, 1, K))
        norm(p + "conv_module.norm")
        if geo.conv_norm == "batch_norm":          # BatchNorm1d buffers (inference uses the running statistics)
            sd[p + "conv_module.norm.running_mean"] = _n(g, (d,), 0.3)
            sd[p + "conv_module.norm.running_var"] = 0.5 + torch.rand((d,), generator=g, dtype=torch.float32)
            sd[p + "conv_module.norm.num_batches_tracked"] = torch.tensor(100, dtype=torch.long)
        linear(p + "conv_module.pointwise_conv2", d, d, shape=(d, d, 1))
        for name in ("norm_ff", "norm_mha", "norm_ff_macaron", "norm_conv", "norm_final"):
            norm(p + name)
    norm("encoder.after_norm")
    if V > 0:
        linear("ctc.ctc_lo", V, d)
    return sd


def synth_fbank(num_frames: int, seed: int = 1, feat_dim: int = 80) -> torch.Tensor:
    """N(0,1) 'fbank' of `num_frames` 10-ms frames (post-CMVN statistics)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn((num_frames, feat_dim), generator=g, dtype=torch.float32)


def frames_of_seconds(seconds: float) -> int:
    """kaldi fbank, 25 ms window / 10 ms shift, snip_edges: T = 100*s - 2."""
    return int(round(seconds * 100)) - 2


# BASELINE.json configs[1] / SURVEY.md 8(d) workload 2: 19 utterances, 14 400 s in total.
MASKED_BATCH_SECONDS: List[float] = [1, 30, 60, 900, 1800, 3600] * 2 + [1000, 500, 100, 10, 5, 2, 1]


def masked_batch_lengths(scale: float = 1.0) -> List[int]:
    return [max(frames_of_seconds(s * scale), 16) for s in MASKED_BATCH_SECONDS]


def synth_transducer_state_dict(vocab: int, embed: int, hidden: int, layers: int, pred_out: int, enc_dim: int, join_dim: int,
                                blank_bias: float = 3.0, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random predictor + joint weights keyed like the reference transducer checkpoint (`predictor.*`, `joint.*`;
    transducer/predictor.py:69-96, joint.py:36-52).  `blank_bias` is added to the blank logit so that, as with a trained
    model, most frames emit blank (a random joint would emit a symbol on nearly every step)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {"predictor.embed.weight": _n(g, (vocab, embed), 1.0)}
    b = 1.0 / math.sqrt(hidden)
    for layer in range(layers):
        inp = embed if layer == 0 else hidden
        sd[f"predictor.rnn.weight_ih_l{layer}"] = _u(g, (4 * hidden, inp), b)
        sd[f"predictor.rnn.weight_hh_l{layer}"] = _u(g, (4 * hidden, hidden), b)
        sd[f"predictor.rnn.bias_ih_l{layer}"] = _u(g, (4 * hidden,), b)
        sd[f"predictor.rnn.bias_hh_l{layer}"] = _u(g, (4 * hidden,), b)

    def linear(prefix, out_f, in_f, scale=1.0):
        bound = scale / math.sqrt(in_f)
        sd[prefix + ".weight"] = _u(g, (out_f, in_f), bound)
        sd[prefix + ".bias"] = _u(g, (out_f,), bound)

    linear("predictor.projection", pred_out, hidden, 2.0)
    linear("joint.enc_ffn", join_dim, enc_dim, 2.0)
    linear("joint.pred_ffn", join_dim, pred_out, 0.7)
    linear("joint.ffn_out", vocab, join_dim, 4.0)
    sd["joint.ffn_out.bias"][0] += blank_bias
    return sd


def check_bench_batch(golden, out: torch.Tensor, tokens: torch.Tensor, n_chunks, enc_lens, margin_tol: float) -> dict:
    """Compare one encode of the benchmark batch (BASELINE.json configs[1]) with the golden of the UNMODIFIED reference for
    that batch (tests/golden/bench_batch.npz, written by tests/golden/make_golden_bench.py; `golden` = the loaded npz).
    out (n, c, d) / tokens (n, c) as forward_parallel_chunk / ctc_greedy return them.  Used by the -m gpu parity test and by
    bench.py after its timed steps (`parity_checked`).  Returns the measured errors; the caller applies the bar."""
    import numpy as np
    c, d = out.shape[1], out.shape[2]
    stride = int(golden["cfg"][3])
    if [int(v) for v in n_chunks] != [int(v) for v in golden["n_chunks"]] or \
            [int(v) for v in enc_lens] != [int(v) for v in golden["enc_lens"]]:
        raise AssertionError("chunk counts / encoder lengths differ from the reference golden")
    idx_all, idx_sample, row = [], [], 0
    for nck, m in zip(n_chunks, enc_lens):
        m = max(int(m), 0)
        base = row * c
        idx_all.append(torch.arange(base, base + m))
        idx_sample.append(torch.arange(base, base + m, stride))
        row += int(nck)
    idx_all = torch.cat(idx_all).to(out.device)
    idx_sample = torch.cat(idx_sample).to(out.device)
    flat = out.reshape(-1, d)
    rows = flat[idx_sample].float().cpu()
    want_rows = torch.from_numpy(golden["rows"].astype(np.float32))
    diff = rows - want_rows
    rowsum = flat[idx_all].float().sum(1).cpu()
    tok = tokens.reshape(-1)[idx_all].cpu()
    want_tok = torch.from_numpy(golden["tokens"].astype(np.int64))
    margin = torch.from_numpy(golden["margin"].astype(np.float32))
    bad = (tok != want_tok)
    return {"rows_compared": int(rows.shape[0]), "max_abs": float(diff.abs().max()),
            "rel_rms": float(diff.pow(2).mean().sqrt() / want_rows.pow(2).mean().sqrt()),
            "rowsum_max_abs": float((rowsum - torch.from_numpy(golden["rowsum"])).abs().max()),
            "tokens_compared": int(tok.numel()), "token_mismatches": int(bad.sum()),
            "token_mismatches_above_tol": int((bad & (margin >= margin_tol)).sum()), "margin_tol": margin_tol}
