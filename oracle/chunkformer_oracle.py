"""CPU oracle for the ChunkFormer masked-chunk encoder forward + greedy CTC.

TEST INFRASTRUCTURE ONLY.  This file is a from-scratch restatement (closed-form index math +
plain fp32 tensor algebra on the CPU) of the algorithm that the reference implements in
/root/reference/chunkformer/modules/{encoder,encoder_layer,attention,convolution,subsampling,
embedding,ctc}.py.  It exists to check the CUDA path; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product package
(chunkformer_b200/) never imports it and has no CPU fallback.

Pinning: the reference holds no golden vectors for this path (its tests need Hugging Face
downloads), so the oracle is pinned against outputs of the UNMODIFIED
reference run in the build container on synthetic checkpoints: tests/golden/make_golden.py
generates tests/golden/*.npz and tests/test_oracle_golden.py checks this file against them
(integer tables bit-exact, floats to 2e-4 abs).

Arithmetic is numpy int64 for the packer / tables and torch CPU fp32 for the float path (the
reference's own arithmetic is ATen fp32; pinned as torch>=2.5.1 in its pyproject.toml:27).

Notation (SURVEY.md 8a'): utterance with T input frames, M = 1 + floor((T-15)/8) valid encoder
frames, stream offset o, chunk j of size c, left/right context l/r, W = l+c+r, R = 2c+l+r-1,
absolute encoder frame f = c*j + i.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SUB = 8            # subsampling rate            (subsampling.py:43-45)
CTX = 15           # right_context + 1           (encoder.py:540)


# --------------------------------------------------------------------------------------------
# integer part: packer and bound tables (must match the reference bit-exactly)
# --------------------------------------------------------------------------------------------
def calc_length(T: int) -> int:
    """subsampling.py:270-288 — three times floor((L-3)/2)+1, evaluated in float then cast."""
    L = float(T)
    for _ in range(3):
        L = math.floor((L - 3.0) / 2.0) + 1.0
    return int(L)


@dataclass
class Plan:
    c: int
    l: int
    r: int
    lorder: int
    lens: np.ndarray        # (B,) input frames
    offsets: np.ndarray     # (B,) stream offsets (encoder frames already emitted)
    n_chunks: np.ndarray    # (B,) chunks per utterance               encoder.py:562
    pad: np.ndarray         # (B,) zero frames appended               encoder.py:557-560
    valid: np.ndarray       # (B,) M = 1 + floor((T-15)/8)            encoder.py:567
    enc_lens: np.ndarray    # (B,) calc_length(T)                     encoder.py:673
    chunk_utt: np.ndarray   # (n,) utterance of each chunk
    chunk_idx: np.ndarray   # (n,) index j of the chunk inside its utterance
    chunk_in_len: np.ndarray  # (n,) input frames present in the chunk  encoder.py:598
    att_mask: np.ndarray    # (n, W) bool                              encoder.py:637-645
    conv_mask: np.ndarray   # (n, c + 2*lorder) bool                   encoder.py:627-633

    @property
    def n(self) -> int:
        return int(self.chunk_utt.shape[0])


def make_plan(lens: Sequence[int], offsets: Optional[Sequence[int]], c: int, l: int, r: int,
              kernel: int = 15) -> Plan:
    """Masked-batch packer of forward_parallel_chunk (encoder.py:538-645) in closed form.

    Chunk j of an utterance covers input frames [8c*j, 8c*j + 8(c-1)+15); the bound tables reduce to
      attention key slot q of chunk j (frame f = c*j - l + q)  valid  <=>  -o <= f < M
      conv window slot q of chunk j (frame f = c*j - lorder + q) valid <=>  -o <= f < min(M, c*(j+1) + r)
    """
    lens = np.asarray(lens, dtype=np.int64)
    B = lens.shape[0]
    offs = np.zeros(B, dtype=np.int64) if offsets is None else np.asarray(offsets, dtype=np.int64)
    lorder = kernel // 2
    size = (c - 1) * SUB + CTX
    step = SUB * c
    pad = np.where(lens >= size, (step - ((lens - size) % step)) % step, size - lens)
    n_chunks = (lens + pad - size) // step + 1
    valid = 1 + np.floor_divide(lens - CTX, SUB)
    enc_lens = np.array([calc_length(int(t)) for t in lens], dtype=np.int64)
    chunk_utt = np.repeat(np.arange(B), n_chunks)
    first = np.concatenate([[0], np.cumsum(n_chunks)[:-1]])
    chunk_idx = np.arange(int(n_chunks.sum())) - np.repeat(first, n_chunks)
    chunk_in_len = np.full(chunk_utt.shape, size, dtype=np.int64)
    chunk_in_len[np.cumsum(n_chunks) - 1] = size - pad
    W = l + c + r
    Mj = valid[chunk_utt][:, None]
    oj = offs[chunk_utt][:, None]
    f_att = (c * chunk_idx)[:, None] - l + np.arange(W)[None, :]
    att_mask = (f_att >= -oj) & (f_att < Mj)
    f_cv = (c * chunk_idx)[:, None] - lorder + np.arange(c + 2 * lorder)[None, :]
    hi = np.minimum(Mj, (c * (chunk_idx + 1))[:, None] + r)
    conv_mask = (f_cv >= -oj) & (f_cv < hi)
    return Plan(c, l, r, lorder, lens, offs, n_chunks, pad, valid, enc_lens, chunk_utt, chunk_idx,
                chunk_in_len, att_mask, conv_mask)


def endless_segments(xs_len: int, c: int, r: int, layers: int, total_batch_duration: float,
                     lorder: int = 7) -> Tuple[int, int, List[Tuple[int, int, bool]]]:
    """Segment arithmetic of endless_decode (chunkformer_model.py:344-371, 391-434).

    Returns (truncated_context_size, rel_right_context (input frames), [(start, end, is_last)])
    where the encoder sees input frames [start, end) for each sequential segment."""
    max_len = int(total_batch_duration // 0.01) // 2
    multiply_n = max_len // c // SUB
    trunc = c * multiply_n
    rr = max(r, lorder)
    rel_right = (rr + max(c, rr) * (layers - 1)) * SUB
    segs = []
    idx = 0
    for _ in range(0, xs_len, max(trunc * SUB, 1)):
        start = trunc * SUB * idx
        end = min(trunc * SUB * (idx + 1) + 7, xs_len)
        last = not (trunc * SUB * idx + rel_right < xs_len)
        segs.append((start, min(end + rel_right, xs_len), last))
        if last:
            break
        idx += 1
    return trunc, rel_right, segs


def batch_groups(lens: Sequence[int], total_batch_duration: float) -> List[List[int]]:
    """Greedy arrival-order admission of batch_decode (chunkformer_model.py:481-504)."""
    budget0 = int(total_batch_duration // 0.01) // 2
    groups, cur, budget = [], [], budget0
    for i, t in enumerate(lens):
        cur.append(i)
        budget -= int(t)
        if budget <= 0 or i == len(lens) - 1:
            groups.append(cur)
            cur, budget = [], budget0
    return groups


def ctc_collapse(tokens: Sequence[int], blank: int = 0) -> List[int]:
    """remove_duplicates_and_blank (utils/model_utils.py:23-32)."""
    out, prev = [], None
    for t in tokens:
        t = int(t)
        if t != prev and t != blank:
            out.append(t)
        prev = t
    return out


# --------------------------------------------------------------------------------------------
# float part
# --------------------------------------------------------------------------------------------
def rel_pos_table(d: int, c: int, l: int, r: int) -> torch.Tensor:
    """embedding.py:119-174: row p of the (R, d) table encodes relative distance rho = c+l-1-p
    (positive = key to the left of the query): PE[2m] = sin(rho w_m), PE[2m+1] = cos(rho w_m)."""
    R = 2 * c + l + r - 1
    rho = (c + l - 1 - torch.arange(R, dtype=torch.float32)).unsqueeze(1)
    w = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    # the reference builds the table from non-negative positions and mirrors the sine
    ang = rho.abs() * w
    pe = torch.zeros(R, d, dtype=torch.float32)
    pe[:, 0::2] = torch.sin(ang) * torch.sign(rho)
    pe[:, 1::2] = torch.cos(ang)
    return pe


def _conv_norm(sd, cm, z):
    """The conv module's norm over the channel dimension (convolution.py:83-89): LayerNorm, or eval-mode BatchNorm1d (running
    statistics) when the checkpoint carries them (cnn_module_norm: batch_norm, the constructor default)."""
    if cm + "norm.running_mean" in sd:
        return (z - sd[cm + "norm.running_mean"]) / torch.sqrt(sd[cm + "norm.running_var"] + 1e-5) * sd[cm + "norm.weight"] \
            + sd[cm + "norm.bias"]
    return _ln(z, sd[cm + "norm.weight"], sd[cm + "norm.bias"])


def _ln(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _ffn(sd, p, x):
    """positionwise_feed_forward.py:51-60: W2 SiLU(W1 x + b1) + b2."""
    h = x @ sd[p + ".w_1.weight"].T + sd[p + ".w_1.bias"]
    h = h * torch.sigmoid(h)
    return h @ sd[p + ".w_2.weight"].T + sd[p + ".w_2.bias"]


def subsample_chunks(sd: Dict[str, torch.Tensor], chunks: torch.Tensor) -> torch.Tensor:
    """subsampling.py:70-112,128-175 + embedding.py:198 on (n, size, 80) chunk inputs -> (n, c, d).
    Flatten order of the last Linear is channel-major then frequency (subsampling.py:163-164)."""
    e = "encoder.embed."
    d = sd[e + "conv.0.weight"].shape[0]
    x = chunks.unsqueeze(1)
    x = F.relu(F.conv2d(x, sd[e + "conv.0.weight"], sd[e + "conv.0.bias"], stride=2))
    for dw, pw in ((2, 3), (5, 6)):
        x = F.conv2d(x, sd[e + f"conv.{dw}.weight"], sd[e + f"conv.{dw}.bias"], stride=2, groups=d)
        x = F.relu(F.conv2d(x, sd[e + f"conv.{pw}.weight"], sd[e + f"conv.{pw}.bias"]))
    n, ch, t, fq = x.shape
    x = x.permute(0, 2, 1, 3).reshape(n, t, ch * fq)
    x = x @ sd[e + "out.weight"].T + sd[e + "out.bias"]
    return x * math.sqrt(d)


def _attention(sd, p, H, y, plan: Plan, pe, att_cache, trunc):
    """attention.py:420-505 + 104-150 on flat frames.  y: (n*c, d) normalised input.
    Returns (out (n*c, d), new_cache (l, H, 2*d_k) or None)."""
    c, l, r = plan.c, plan.l, plan.r
    n = plan.n
    d = y.shape[1]
    dk = d // H
    W = l + c + r
    q = (y @ sd[p + "linear_q.weight"].T + sd[p + "linear_q.bias"]).view(n, c, H, dk)
    k = (y @ sd[p + "linear_k.weight"].T + sd[p + "linear_k.bias"]).view(n * c, H, dk)
    v = (y @ sd[p + "linear_v.weight"].T + sd[p + "linear_v.bias"]).view(n * c, H, dk)
    if att_cache is not None:
        kc, vc = att_cache[..., :dk], att_cache[..., dk:]
    else:
        kc = vc = torch.zeros(l, H, dk)
    # flat buffers [cache (l) | all frames | r zeros]; chunk j's window is rows [c*j, c*j + W)
    kf = torch.cat([kc, k, torch.zeros(r, H, dk)], 0)
    vf = torch.cat([vc, v, torch.zeros(r, H, dk)], 0)
    new_cache = None
    if att_cache is not None:
        t0 = min(trunc, n * c)              # kv[: trunc + l][-l:] clamps at the end of the buffer (attention.py:466-467)
        new_cache = torch.cat([kf, vf], -1)[t0:t0 + l].clone()
    win = (c * torch.arange(n)).unsqueeze(1) + torch.arange(W).unsqueeze(0)      # (n, W)
    kw = kf[win].permute(0, 2, 1, 3)                                                # (n, H, W, dk)
    vw = vf[win].permute(0, 2, 1, 3)
    pos = (pe @ sd[p + "linear_pos.weight"].T).view(-1, H, dk).permute(1, 0, 2)     # (H, R, dk)
    qu = (q + sd[p + "pos_bias_u"]).permute(0, 2, 1, 3)                             # (n, H, c, dk)
    qv = (q + sd[p + "pos_bias_v"]).permute(0, 2, 1, 3)
    ac = qu @ kw.transpose(-1, -2)                                                  # (n, H, c, W)
    bd_full = qv @ pos.transpose(-1, -2).unsqueeze(0)                               # (n, H, c, R)
    # rel_shift (attention.py:242-266): bd[i, q] = bd_full[i, (c-1) - i + q]
    idx = (c - 1) - torch.arange(c).unsqueeze(1) + torch.arange(W).unsqueeze(0)     # (c, W)
    bd = bd_full.gather(-1, idx.expand(n, H, c, W))
    s = (ac + bd) / math.sqrt(dk)
    m = torch.from_numpy(plan.att_mask).view(n, 1, 1, W)
    s = s.masked_fill(~m, float("-inf"))
    smax = s.amax(-1, keepdim=True)
    smax = torch.where(torch.isinf(smax), torch.zeros_like(smax), smax)
    e = torch.exp(s - smax)
    den = e.sum(-1, keepdim=True)
    a = torch.where(den > 0, e / den.clamp_min(1e-38), torch.zeros_like(e))         # all-masked row -> 0
    ctx = (a @ vw).permute(0, 2, 1, 3).reshape(n * c, d)
    return ctx @ sd[p + "linear_out.weight"].T + sd[p + "linear_out.bias"], new_cache


def _conv_module(sd, p, y, plan: Plan, cnn_cache, trunc, sub_chunk: int = 0):
    """convolution.py:194-255 on flat frames.  y: (n*c, d).  Returns (out, new_cache (d, lorder) or None).
    sub_chunk > 0 (frame-synchronous streaming with right context, encoder.py:310-385 -> convolution.py:101-192 with
    chunk_size = sub_chunk): the plan's chunk holds chunk + right-context frames, and the dynamic conv cuts it into sub-chunks
    of `sub_chunk` frames, each seeing real left context but ZEROS to its right (convolution.py:150-167)."""
    c, lo = plan.c, plan.lorder
    n = plan.n
    d = y.shape[1]
    w1 = sd[p + "pointwise_conv1.weight"].view(2 * d, d)
    h = y @ w1.T + sd[p + "pointwise_conv1.bias"]
    g = h[:, :d] * torch.sigmoid(h[:, d:])                                          # GLU: value * sigmoid(gate)
    left = cnn_cache.T if cnn_cache is not None else torch.zeros(lo, d)
    gf = torch.cat([left, g, torch.zeros(lo, d)], 0)                                # row f + lo  <->  frame f
    new_cache = None
    if cnn_cache is not None:
        t0 = min(trunc, n * c)              # x[:, : trunc + lo][:, -lo:] (convolution.py:228-230)
        new_cache = gf[t0:t0 + lo].T.clone()
    win = (c * torch.arange(n)).unsqueeze(1) + torch.arange(c + 2 * lo).unsqueeze(0)
    gw = gf[win] * torch.from_numpy(plan.conv_mask).unsqueeze(-1)                   # (n, c+2lo, d)
    wd = sd[p + "depthwise_conv.weight"].view(d, -1)                                # (d, K)
    z = torch.zeros(n, c, d)
    fr = torch.arange(c)
    for tau in range(wd.shape[1]):
        term = gw[:, tau:tau + c, :] * wd[:, tau]
        if sub_chunk > 0:                   # window slot f + tau <-> frame f - lo + tau must lie left of the sub-chunk's end
            term = term * ((fr + tau) < ((fr // sub_chunk + 1) * sub_chunk + lo)).view(1, c, 1)
        z = z + term
    z = z + sd[p + "depthwise_conv.bias"]
    z = _conv_norm(sd, p, z)
    z = z * torch.sigmoid(z)
    w2 = sd[p + "pointwise_conv2.weight"].view(d, d)
    out = z.view(n * c, d) @ w2.T + sd[p + "pointwise_conv2.bias"]
    centre = torch.from_numpy(plan.conv_mask[:, lo:lo + c]).reshape(n * c, 1)
    return out * centre, new_cache


def _layer(sd, i, H, x, plan, pe, att_cache, cnn_cache, trunc, sub_chunk: int = 0):
    """encoder_layer.py:155-248 (pre-norm, macaron, dropout = identity)."""
    p = f"encoder.encoders.{i}."
    ln = lambda name, t: _ln(t, sd[p + name + ".weight"], sd[p + name + ".bias"])  # noqa: E731
    x = x + 0.5 * _ffn(sd, p + "feed_forward_macaron", ln("norm_ff_macaron", x))
    a, new_att = _attention(sd, p + "self_attn.", H, ln("norm_mha", x), plan, pe, att_cache, trunc)
    x = x + a
    cv, new_cnn = _conv_module(sd, p + "conv_module.", ln("norm_conv", x), plan, cnn_cache, trunc, sub_chunk)
    x = x + cv
    x = x + 0.5 * _ffn(sd, p + "feed_forward", ln("norm_ff", x))
    return ln("norm_final", x), new_att, new_cnn


def pack_chunks(xs: Sequence[torch.Tensor], plan: Plan) -> torch.Tensor:
    """encoder.py:552-606: zero-pad each utterance and cut it into overlapping (size, 80) chunks."""
    size = (plan.c - 1) * SUB + CTX
    step = SUB * plan.c
    out = []
    for u, x in enumerate(xs):
        x = F.pad(x, (0, 0, 0, int(plan.pad[u])))
        for j in range(int(plan.n_chunks[u])):
            out.append(x[step * j: step * j + size])
    return torch.stack(out, 0)


@torch.no_grad()
def forward_parallel_chunk(sd: Dict[str, torch.Tensor], heads: int, xs: Sequence[torch.Tensor],
                           lens: Sequence[int], c: int, l: int, r: int,
                           att_cache: Optional[torch.Tensor] = None,
                           cnn_cache: Optional[torch.Tensor] = None,
                           truncated_context_size: int = 0,
                           offsets: Optional[Sequence[int]] = None, num_layers: Optional[int] = None, conv_sub_chunk: int = 0):
    """ChunkFormerEncoder.forward_parallel_chunk (encoder.py:503-681).

    att_cache (L, l, H, 2*d_k) / cnn_cache (L, d, lorder) or None (= the reference's empty caches).
    Returns (xs (n, c, d), enc_lens (B,), n_chunks list, new_att_cache|None, new_cnn_cache|None,
    new_offsets (B,))."""
    L = num_layers if num_layers is not None else \
        1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.encoders."))
    kernel = sd["encoder.encoders.0.conv_module.depthwise_conv.weight"].shape[-1]
    d = sd["encoder.after_norm.weight"].shape[0]
    plan = make_plan(lens, offsets, c, l, r, kernel)
    chunks = pack_chunks([x.float() for x in xs], plan)
    if "encoder.global_cmvn.mean" in sd:                       # cmvn.py:32-43, after zero padding
        chunks = (chunks - sd["encoder.global_cmvn.mean"]) * sd["encoder.global_cmvn.istd"]
    x = subsample_chunks(sd, chunks).reshape(plan.n * c, d)
    pe = rel_pos_table(d, c, l, r)
    new_att, new_cnn = [], []
    for i in range(L):
        x, a, cv = _layer(sd, i, heads, x, plan, pe,
                          None if att_cache is None else att_cache[i],
                          None if cnn_cache is None else cnn_cache[i], truncated_context_size, conv_sub_chunk)
        new_att.append(a)
        new_cnn.append(cv)
    x = _ln(x, sd["encoder.after_norm.weight"], sd["encoder.after_norm.bias"])
    enc_lens = torch.from_numpy(plan.enc_lens).to(torch.int32)
    new_off = torch.from_numpy(plan.offsets + plan.enc_lens)
    return (x.view(plan.n, c, d), enc_lens, [int(v) for v in plan.n_chunks],
            None if att_cache is None else torch.stack(new_att, 0),
            None if cnn_cache is None else torch.stack(new_cnn, 0), new_off)


@torch.no_grad()
def ctc_greedy(sd: Dict[str, torch.Tensor], enc: torch.Tensor):
    """ctc.py:73-91 + chunkformer_model.py:437-438: argmax of log_softmax(W x + b).
    Returns (tokens int64 (...,), top-2 logit margin float32 (...,))."""
    logits = enc @ sd["ctc.ctc_lo.weight"].T + sd["ctc.ctc_lo.bias"]
    top2 = logits.topk(2, dim=-1).values
    return logits.argmax(-1), (top2[..., 0] - top2[..., 1])


@torch.no_grad()
def endless_decode_tokens(sd: Dict[str, torch.Tensor], heads: int, layers: int, x: torch.Tensor, c: int, l: int, r: int,
                          total_batch_duration: float):
    """The segment loop of endless_decode (chunkformer_model.py:391-438) on this oracle: sequential segments with the K/V and
    conv caches carried over, rows computed only as right context dropped, greedy CTC over the concatenation.
    Returns (tokens (T',), top-2 margins (T',), [dict(len, trunc, offset) per encoder call])."""
    kernel = sd["encoder.encoders.0.conv_module.depthwise_conv.weight"].shape[-1]
    d = sd["encoder.after_norm.weight"].shape[0]
    trunc, _, segs = endless_segments(int(x.shape[0]), c, r, layers, total_batch_duration, kernel // 2)
    att = torch.zeros(layers, l, heads, 2 * d // heads)
    cnn = torch.zeros(layers, d, kernel // 2)
    off, outs, calls = [0], [], []
    for (s, e, last) in segs:
        calls.append(dict(len=e - s, trunc=trunc, offset=off[0]))
        o, ol, _, att, cnn, noff = forward_parallel_chunk(sd, heads, [x[s:e]], [e - s], c, l, r, att, cnn, trunc, off, layers)
        o = o.reshape(-1, d)[: max(int(ol[0]), 0)]
        if not last:
            o = o[:trunc]
        off = [int(noff[0]) - int(ol[0]) + o.shape[0]]
        outs.append(o)
    tok, margin = ctc_greedy(sd, torch.cat(outs, 0))
    return tok, margin, calls


@torch.no_grad()
def batch_decode_tokens(sd: Dict[str, torch.Tensor], heads: int, xs: Sequence[torch.Tensor], c: int, l: int, r: int,
                        total_batch_duration: float):
    """batch_decode (chunkformer_model.py:481-529) on this oracle: arrival-order admission, one masked batch per group.
    Returns ([tokens per utterance], [margins per utterance], group sizes)."""
    lens = [int(t.shape[0]) for t in xs]
    toks, margins, sizes = [], [], []
    for grp in batch_groups(lens, total_batch_duration):
        out, enc_lens, n_chunks, _, _, _ = forward_parallel_chunk(sd, heads, [xs[i] for i in grp], [lens[i] for i in grp], c, l, r)
        tok, mg = ctc_greedy(sd, out)
        row = 0
        for u, nck in enumerate(n_chunks):
            m = max(int(enc_lens[u]), 0)
            toks.append(tok[row:row + nck].reshape(-1)[:m])
            margins.append(mg[row:row + nck].reshape(-1)[:m])
            row += nck
        sizes.append(len(grp))
    return toks, margins, sizes


@torch.no_grad()
def ctc_log_softmax(sd: Dict[str, torch.Tensor], enc: torch.Tensor) -> torch.Tensor:
    logits = enc @ sd["ctc.ctc_lo.weight"].T + sd["ctc.ctc_lo.bias"]
    return torch.log_softmax(logits, dim=-1)


@torch.no_grad()
def forward_encoder(sd: Dict[str, torch.Tensor], heads: int, xs: torch.Tensor, xs_lens: Sequence[int],
                    c: int, l: int, r: int, num_layers: Optional[int] = None):
    """ChunkFormerEncoder.forward_encoder (encoder.py:220-308) = the path behind ChunkFormerModel.encode.

    Differences from the masked-batch path, all restated from the reference's non-parallel modules:
      * padded batch (B, T, 80) subsampled whole (subsampling.py:120-175), T' = 1 + (T-15)//8 rows each;
      * key j of utterance b valid <=> 0 <= f < len_b inside the window (attention.py:349-383);
      * conv module input rows >= len_b are zeroed BEFORE pointwise_conv1 (convolution.py:125-127), so those
        frames carry GLU(bias); each chunk sees lorder real frames on its left (zeros before frame 0) and
        ZEROS on its right and beyond row T' (convolution.py:150-167); output rows >= len_b are zeroed.
    Returns (out (B, T', d), mask (B, 1, T') bool)."""
    L = num_layers if num_layers is not None else \
        1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.encoders."))
    kernel = sd["encoder.encoders.0.conv_module.depthwise_conv.weight"].shape[-1]
    lo = kernel // 2
    d = sd["encoder.after_norm.weight"].shape[0]
    H = heads
    dk = d // H
    B, T, _ = xs.shape
    x = xs.float()
    if "encoder.global_cmvn.mean" in sd:
        x = (x - sd["encoder.global_cmvn.mean"]) * sd["encoder.global_cmvn.istd"]
    x = subsample_chunks(sd, x)                                            # (B, T', d), whole utterance
    Tp = x.shape[1]
    lens = torch.tensor([calc_length(int(t)) for t in xs_lens])
    valid = torch.arange(Tp).unsqueeze(0) < lens.unsqueeze(1)              # (B, T')
    nck = (Tp + c - 1) // c
    Tpad = nck * c
    W = l + c + r
    pe = rel_pos_table(d, c, l, r)
    # window frame index per (chunk, slot) and validity per (b, chunk, slot)
    fwin = (c * torch.arange(nck)).unsqueeze(1) - l + torch.arange(W).unsqueeze(0)           # (nck, W)
    kvalid = (fwin >= 0).unsqueeze(0) & (fwin.unsqueeze(0) < lens.view(B, 1, 1))             # (B, nck, W)
    fidx = fwin.clamp(0, Tpad - 1)
    sh = (c - 1) - torch.arange(c).unsqueeze(1) + torch.arange(W).unsqueeze(0)               # (c, W)
    cwin = (c * torch.arange(nck)).unsqueeze(1) - lo + torch.arange(c + 2 * lo).unsqueeze(0)  # (nck, c+2lo)
    cvalid = (cwin >= 0) & (cwin < Tp) & (cwin < (c * (torch.arange(nck) + 1)).unsqueeze(1))
    cidx = cwin.clamp(0, Tpad - 1)

    def pad_t(t):
        return F.pad(t, (0, 0, 0, Tpad - Tp))

    for i in range(L):
        p = f"encoder.encoders.{i}."
        ln = lambda name, t: _ln(t, sd[p + name + ".weight"], sd[p + name + ".bias"])  # noqa: E731
        x = x + 0.5 * _ffn(sd, p + "feed_forward_macaron", ln("norm_ff_macaron", x))
        # ---- attention (attention.py:268-418)
        y = ln("norm_mha", x)
        a = p + "self_attn."
        q = pad_t(y @ sd[a + "linear_q.weight"].T + sd[a + "linear_q.bias"]).view(B, nck, c, H, dk)
        k = pad_t(y @ sd[a + "linear_k.weight"].T + sd[a + "linear_k.bias"]).view(B, Tpad, H, dk)
        v = pad_t(y @ sd[a + "linear_v.weight"].T + sd[a + "linear_v.bias"]).view(B, Tpad, H, dk)
        kw = k[:, fidx].permute(0, 1, 3, 2, 4)                               # (B, nck, H, W, dk)
        vw = v[:, fidx].permute(0, 1, 3, 2, 4)
        pos = (pe @ sd[a + "linear_pos.weight"].T).view(-1, H, dk).permute(1, 0, 2)
        qu = (q + sd[a + "pos_bias_u"]).permute(0, 1, 3, 2, 4)               # (B, nck, H, c, dk)
        qv = (q + sd[a + "pos_bias_v"]).permute(0, 1, 3, 2, 4)
        ac = qu @ kw.transpose(-1, -2)
        bd = (qv @ pos.transpose(-1, -2)).gather(-1, sh.expand(B, nck, H, c, W))
        s = ((ac + bd) / math.sqrt(dk)).masked_fill(~kvalid.view(B, nck, 1, 1, W), float("-inf"))
        smax = s.amax(-1, keepdim=True)
        smax = torch.where(torch.isinf(smax), torch.zeros_like(smax), smax)
        e = torch.exp(s - smax)
        den = e.sum(-1, keepdim=True)
        att = torch.where(den > 0, e / den.clamp_min(1e-38), torch.zeros_like(e))
        ctx = (att @ vw).permute(0, 1, 3, 2, 4).reshape(B, Tpad, d)[:, :Tp]
        ctx = ctx * valid.unsqueeze(-1)                                      # query-side mask
        x = x + ctx @ sd[a + "linear_out.weight"].T + sd[a + "linear_out.bias"]
        # ---- conv module (convolution.py:101-192)
        cm = p + "conv_module."
        y = ln("norm_conv", x) * valid.unsqueeze(-1)
        h = y @ sd[cm + "pointwise_conv1.weight"].view(2 * d, d).T + sd[cm + "pointwise_conv1.bias"]
        g = pad_t(h[..., :d] * torch.sigmoid(h[..., d:]))                    # (B, Tpad, d), rows >= T' zero
        gw = g[:, cidx] * cvalid.view(1, nck, c + 2 * lo, 1)                 # (B, nck, c+2lo, d)
        wd = sd[cm + "depthwise_conv.weight"].view(d, -1)
        z = torch.zeros(B, nck, c, d)
        for tau in range(wd.shape[1]):
            z = z + gw[:, :, tau:tau + c, :] * wd[:, tau]
        z = z + sd[cm + "depthwise_conv.bias"]
        z = _conv_norm(sd, cm, z)
        z = (z * torch.sigmoid(z)).view(B, Tpad, d)[:, :Tp]
        cv = z @ sd[cm + "pointwise_conv2.weight"].view(d, d).T + sd[cm + "pointwise_conv2.bias"]
        x = x + cv * valid.unsqueeze(-1)
        x = x + 0.5 * _ffn(sd, p + "feed_forward", ln("norm_ff", x))
        x = ln("norm_final", x)
    x = _ln(x, sd["encoder.after_norm.weight"], sd["encoder.after_norm.bias"])
    return x, valid.unsqueeze(1)


# --------------------------------------------------------------------------------------------
# frame-synchronous streaming: forward_chunk / forward_chunk_by_chunk with right context 0
# --------------------------------------------------------------------------------------------
def forward_chunk(sd, heads: int, xs: torch.Tensor, att_cache: torch.Tensor, cnn_cache: torch.Tensor, c: int, l: int, r: int = 0,
                  offset: int = 0):
    """ChunkFormerEncoder.forward_chunk (encoder.py:310-390).  xs (B, 8(c + r - 1) + 15, feat); att_cache (L, B, H, l, 2 d_k);
    cnn_cache (L, B, d, lorder).

    One step is, per stream, the masked-chunk call on a single chunk of c + r frames with the stream's caches: the embedding
    and the attention treat chunk + right context as ONE chunk of c + r frames with left context l and no right context
    (encoder.py:341-347, attention.py:323-330; keys = [cache (l rows, valid where >= l - offset, the flipped mask of
    encoder.py:349-355) | c + r frames]); the conv module runs with chunk_size = c, i.e. it cuts the c + r frames into
    sub-chunks of c frames with zeros to the right of each (convolution.py:150-167); the returned caches end where the
    CHUNK ends, not the right context: rows [c, c + l) / [c - 7, c) of cache + frames (encoder.py:376-385).
    Pinned against the reference in tests/test_oracle_golden.py::test_streaming_*.
    Returns (out (B, c + r, d), new att_cache, new cnn_cache)."""
    B = xs.shape[0]
    cc = c + r
    outs, atts, cnns = [], [], []
    for b in range(B):
        a_in = att_cache[:, b].transpose(1, 2)                      # (L, l, H, 2 d_k)
        o, _, _, a, cv, _ = forward_parallel_chunk(sd, heads, [xs[b]], [int(xs.shape[1])], cc, l, 0, a_in, cnn_cache[:, b], c,
                                                   [int(offset)], conv_sub_chunk=c if r > 0 else 0)
        outs.append(o.reshape(-1, o.shape[-1])[:cc])
        atts.append(a.transpose(1, 2))                              # back to (L, H, l, 2 d_k)
        cnns.append(cv)
    return torch.stack(outs), torch.stack(atts, 1), torch.stack(cnns, 1)


def forward_chunk_by_chunk(sd, heads: int, xs: torch.Tensor, xs_lens: Sequence[int], c: int, l: int, r: int = 0):
    """ChunkFormerEncoder.forward_chunk_by_chunk (encoder.py:392-459): (out (B, steps * c [+ r for the last step], d),
    mask (B, 1, T'))."""
    B, T, _ = xs.shape
    L = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.encoders."))
    d = sd["encoder.after_norm.weight"].shape[0]
    lo = sd["encoder.encoders.0.conv_module.depthwise_conv.weight"].shape[-1] // 2
    size, stride = SUB * (c - 1) + CTX + SUB * r, SUB * c
    pad = stride - ((T - size) % stride)                            # encoder.py:417-419 (always pads, a full stride if aligned)
    xp = F.pad(xs.float(), (0, 0, 0, pad))
    att = torch.zeros((L, B, heads, l, 2 * d // heads))
    cnn = torch.zeros((L, B, d, lo))
    outs, offset = [], 0
    for i in range(0, xp.shape[1] - size + stride, stride):
        o, att, cnn = forward_chunk(sd, heads, xp[:, i:i + size], att, cnn, c, l, r, offset)
        outs.append(o[:, :c] if i + size < xp.shape[1] else o)      # encoder.py:449: the right-context rows only of the last step
        offset += c
    out = torch.cat(outs, 1)
    enc_lens = [calc_length(int(t) + pad) for t in xs_lens]
    mask = torch.arange(max(enc_lens)).unsqueeze(0) < torch.tensor(enc_lens).unsqueeze(1)
    return out, mask.unsqueeze(1)
