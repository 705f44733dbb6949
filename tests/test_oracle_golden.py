"""Pin oracle/chunkformer_oracle.py against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from chunkformer_b200.geometry import EncoderGeometry
from chunkformer_b200.synth import synth_fbank, synth_state_dict
from oracle import chunkformer_oracle as O

TINY = EncoderGeometry(d_model=64, heads=2, ffn=128, layers=2, kernel=15, vocab=50)
TINY_CMVN = EncoderGeometry(d_model=64, heads=2, ffn=128, layers=2, kernel=15, vocab=50, has_cmvn=True)
LARGE = EncoderGeometry(d_model=512, heads=8, ffn=2048, layers=17, kernel=15, vocab=5000)
FTOL = 2e-4   # fp32 reference vs fp32 restatement (different summation order)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=True)


def test_plan_tables_bit_exact(golden_dir):
    g = _load(golden_dir, "plan_cases.npz")
    for case in g["plan_cases"]:
        plan = O.make_plan(case["lens"], case["offsets"], case["c"], case["l"], case["r"], 15)
        assert plan.n == case["n"]
        W = case["l"] + case["c"] + case["r"]
        att = np.unpackbits(case["att"])[: plan.n * W].reshape(plan.n, W).astype(bool)
        cv = np.unpackbits(case["conv"])[: plan.n * (case["c"] + 14)].reshape(plan.n, case["c"] + 14).astype(bool)
        assert np.array_equal(plan.att_mask, att), case
        assert np.array_equal(plan.conv_mask, cv), case


def test_calc_length(golden_dir):
    g = _load(golden_dir, "plan_cases.npz")
    got = np.array([O.calc_length(int(t)) for t in g["calc_len_T"]])
    assert np.array_equal(got, g["calc_len"])


@pytest.mark.parametrize("tag,geo,seed", [("tiny", TINY, 3), ("tiny_cmvn", TINY_CMVN, 4)])
def test_forward_parallel_chunk_masked_batch(golden_dir, tag, geo, seed):
    g = _load(golden_dir, f"{tag}.npz")
    sd = synth_state_dict(geo, seed)
    for ci in range(4):
        c, l, r = (int(v) for v in g[f"c{ci}_cfg"])
        lens = [int(v) for v in g[f"c{ci}_lens"]]
        xs = [synth_fbank(t, seed=100 + k) for k, t in enumerate(lens)]
        out, enc_lens, n_chunks, _, _, off = O.forward_parallel_chunk(sd, geo.heads, xs, lens, c, l, r)
        assert n_chunks == [int(v) for v in g[f"c{ci}_n_chunks"]]
        assert np.array_equal(enc_lens.numpy(), g[f"c{ci}_enc_lens"])
        assert np.array_equal(off.numpy(), g[f"c{ci}_offset"])
        ref = torch.from_numpy(g[f"c{ci}_out"])
        # only rows < enc_len of each utterance are defined by the reference
        row = 0
        for u, nck in enumerate(n_chunks):
            m = max(int(enc_lens[u]), 0)
            a = out[row:row + nck].reshape(-1, geo.d_model)[:m]
            b = ref[row:row + nck].reshape(-1, geo.d_model)[:m]
            if m:
                assert (a - b).abs().max().item() < FTOL, (tag, ci, u)
            row += nck
        tok, margin = O.ctc_greedy(sd, out)
        ref_tok = torch.from_numpy(g[f"c{ci}_tokens"])
        row = 0
        for u, nck in enumerate(n_chunks):
            m = max(int(enc_lens[u]), 0)
            a = tok[row:row + nck].reshape(-1)[:m]
            b = ref_tok[row:row + nck].reshape(-1)[:m]
            mg = margin[row:row + nck].reshape(-1)[:m]
            assert bool(((a == b) | (mg < 10 * FTOL)).all())
            row += nck


@pytest.mark.parametrize("tag,geo,seed", [("tiny", TINY, 3), ("tiny_cmvn", TINY_CMVN, 4)])
def test_streaming_segments_with_caches(golden_dir, tag, geo, seed):
    g = _load(golden_dir, f"{tag}.npz")
    sd = synth_state_dict(geo, seed)
    c, l, r, T, trunc, rel_right = (int(v) for v in g["stream_cfg"])
    x = synth_fbank(T, seed=200)
    L, H, d = geo.layers, geo.heads, geo.d_model
    att = torch.zeros(L, l, H, 2 * d // H)
    cnn = torch.zeros(L, d, 7)
    offset = [0]
    outs = []
    for idx in range(3):
        start = trunc * 8 * idx
        end = min(trunc * 8 * (idx + 1) + 7, T)
        seg = x[start:end + rel_right]
        o, ol, _, att, cnn, off = O.forward_parallel_chunk(sd, H, [seg], [seg.shape[0]], c, l, r, att, cnn,
                                                           trunc, offset)
        o = o.reshape(-1, d)[: int(ol[0])][:trunc]
        offset = [int(off[0]) - int(ol[0]) + o.shape[0]]
        outs.append(o)
        assert offset[0] == int(g[f"stream_offset_{idx}"][0])
        assert np.abs(att.numpy() - g[f"stream_att_cache_{idx}"]).max() < FTOL
        assert np.abs(cnn.numpy() - g[f"stream_cnn_cache_{idx}"]).max() < FTOL
    assert np.abs(torch.cat(outs, 0).numpy() - g["stream_out"]).max() < FTOL


@pytest.mark.parametrize("tag,geo,seed", [("tiny", TINY, 3), ("tiny_cmvn", TINY_CMVN, 4)])
def test_forward_encoder_padded_batch(golden_dir, tag, geo, seed):
    g = _load(golden_dir, f"{tag}.npz")
    sd = synth_state_dict(geo, seed)
    for ei in range(3):
        c, l, r = (int(v) for v in g[f"e{ei}_cfg"])
        lens = [int(v) for v in g[f"e{ei}_lens"]]
        xb = torch.zeros(len(lens), max(lens), 80)
        for k, t in enumerate(lens):
            xb[k, :t] = synth_fbank(t, seed=300 + k)
        out, mask = O.forward_encoder(sd, geo.heads, xb, lens, c, l, r)
        out_lens = mask.squeeze(1).sum(-1)
        assert np.array_equal(out_lens.numpy(), g[f"e{ei}_out_lens"])
        ref = torch.from_numpy(g[f"e{ei}_out"])
        assert out.shape == ref.shape
        for b, m in enumerate(out_lens.tolist()):
            assert (out[b, :m] - ref[b, :m]).abs().max().item() < FTOL, (tag, ei, b)


def test_ctc_large_60s(golden_dir):
    """BASELINE.json configs[0]: CTC-large geometry, one 60 s utterance, 64/128/128, greedy CTC."""
    g = _load(golden_dir, "ctc_large_60s.npz")
    sd = synth_state_dict(LARGE, 0)
    T = int(g["cfg"][3])
    x = synth_fbank(T, seed=1)
    out, enc_lens, n_chunks, _, _, _ = O.forward_parallel_chunk(sd, LARGE.heads, [x], [T], 64, 128, 128)
    m = int(enc_lens[0])
    assert m == int(g["enc_len"][0]) and n_chunks == [int(v) for v in g["n_chunks"]]
    flat = out.reshape(-1, 512)[:m]
    assert (flat[::8] - torch.from_numpy(g["out_rows"])).abs().max().item() < 5e-4
    assert np.abs(flat.double().sum(0).numpy() - g["out_colsum"]).max() < 2e-2
    tok, margin = O.ctc_greedy(sd, flat)
    same = tok.numpy() == g["tokens"]
    assert bool((same | (margin.numpy() < 1e-3)).all())


def _stream_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "stream.npz"))
    for k, (c, l, B, T) in enumerate(g["cases"].tolist()):
        xs = torch.stack([synth_fbank(T, seed=20 + 7 * k + b) for b in range(B)])
        yield k, c, l, B, T, xs, g


def test_streaming_forward_chunk_by_chunk(golden_dir):
    """Frame-synchronous streaming (right context 0) == the reference's forward_chunk_by_chunk."""
    sd = synth_state_dict(TINY, 3)
    for k, c, l, B, T, xs, g in _stream_cases(golden_dir):
        out, mask = O.forward_chunk_by_chunk(sd, TINY.heads, xs, [T] * B, c, l, 0)
        want = torch.from_numpy(g[f"c{k}_out"])
        assert out.shape == want.shape and float((out - want).abs().max()) < FTOL
        assert torch.equal(mask, torch.from_numpy(g[f"c{k}_mask"]))


def test_streaming_forward_chunk_caches(golden_dir):
    """Three explicit forward_chunk steps: outputs and the returned (L, B, H, l, 2 d_k) / (L, B, d, 7) caches."""
    sd = synth_state_dict(TINY, 3)
    L, H, d = TINY.layers, TINY.heads, TINY.d_model
    for k, c, l, B, T, xs, g in _stream_cases(golden_dir):
        size, stride = 8 * (c - 1) + 15, 8 * c
        att, cnn = torch.zeros((L, B, H, l, 2 * d // H)), torch.zeros((L, B, d, 7))
        for step in range(3):
            o, att, cnn = O.forward_chunk(sd, H, xs[:, step * stride: step * stride + size], att, cnn, c, l, 0, offset=step * c)
        for got, key in ((o, "out"), (att, "att"), (cnn, "cnn")):
            want = torch.from_numpy(g[f"c{k}_step3_{key}"])
            assert got.shape == want.shape and float((got - want).abs().max()) < FTOL, key


def test_streaming_with_right_context(golden_dir):
    """forward_chunk / forward_chunk_by_chunk with right_context_size > 0 (encoder.py:310-385): chunk + right context embedded
    and attended as one chunk, conv cut at the chunk grid, caches ending at the chunk: outputs of every step, the outputs of the
    third explicit step (c + r rows) and both caches against the unmodified reference (tests/golden/stream_right.npz)."""
    sd = synth_state_dict(TINY, 3)
    L, H, d = TINY.layers, TINY.heads, TINY.d_model
    g = _load(golden_dir, "stream_right.npz")
    for k, (c, l, r, B, T) in enumerate(g["cases"].tolist()):
        xs = torch.stack([synth_fbank(T, seed=40 + 7 * k + b) for b in range(B)])
        out, mask = O.forward_chunk_by_chunk(sd, H, xs, [T] * B, c, l, r)
        want = torch.from_numpy(g[f"c{k}_out"])
        assert out.shape == want.shape and float((out - want).abs().max()) < FTOL, (c, l, r)
        assert torch.equal(mask, torch.from_numpy(g[f"c{k}_mask"]))
        size, stride = 8 * (c - 1) + 15 + 8 * r, 8 * c
        att, cnn = torch.zeros((L, B, H, l, 2 * d // H)), torch.zeros((L, B, d, 7))
        for step in range(3):
            o, att, cnn = O.forward_chunk(sd, H, xs[:, step * stride: step * stride + size], att, cnn, c, l, r, offset=step * c)
        for got, key in ((o, "out"), (att, "att"), (cnn, "cnn")):
            want = torch.from_numpy(g[f"c{k}_step3_{key}"])
            assert got.shape == want.shape and float((got - want).abs().max()) < FTOL, (key, c, l, r)


def test_batch_norm_conv_module_matches_reference(golden_dir):
    """cnn_module_norm: batch_norm (the reference constructor's default): eval-mode BatchNorm1d with running statistics in the
    conv module, masked batch and padded-batch encode() against the unmodified reference (tests/golden/make_golden_bn.py)."""
    geo = EncoderGeometry(d_model=64, heads=2, ffn=128, layers=2, kernel=15, vocab=50, conv_norm="batch_norm")
    g = _load(golden_dir, "tiny_bn.npz")
    sd = synth_state_dict(geo, 6)
    for ci in range(2):
        c, l, r = (int(v) for v in g[f"c{ci}_cfg"])
        lens = [int(v) for v in g[f"c{ci}_lens"]]
        xs = [synth_fbank(t, seed=100 + k) for k, t in enumerate(lens)]
        out, enc_lens, n_chunks, _, _, _ = O.forward_parallel_chunk(sd, geo.heads, xs, lens, c, l, r)
        assert n_chunks == [int(v) for v in g[f"c{ci}_n_chunks"]] and np.array_equal(enc_lens.numpy(), g[f"c{ci}_enc_lens"])
        ref = torch.from_numpy(g[f"c{ci}_out"])
        row = 0
        for u, nck in enumerate(n_chunks):
            m = max(int(enc_lens[u]), 0)
            if m:
                a = out[row:row + nck].reshape(-1, geo.d_model)[:m]
                b = ref[row:row + nck].reshape(-1, geo.d_model)[:m]
                assert float((a - b).abs().max()) < FTOL, (ci, u)
            row += nck
    c, l, r = (int(v) for v in g["e0_cfg"])
    lens = [int(v) for v in g["e0_lens"]]
    xb = torch.zeros(len(lens), max(lens), 80)
    for k, t in enumerate(lens):
        xb[k, :t] = synth_fbank(t, seed=300 + k)
    out, mask = O.forward_encoder(sd, geo.heads, xb, lens, c, l, r)
    ref = torch.from_numpy(g["e0_out"])
    for b, m in enumerate(g["e0_out_lens"].tolist()):
        assert float((out[b, :m] - ref[b, :m]).abs().max()) < FTOL, b
