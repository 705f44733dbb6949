"""Drop-in facade on a B200: from_pretrained on a reference-layout checkpoint directory, encode, endless_decode,
batch_decode, classify_audio, and chunk-range sharding of one long recording — each against the CPU oracle driven the
way the reference's facade drives its encoder (chunkformer_model.py:256-640)."""
import json
import os

import pytest
import torch
import yaml

from chunkformer_b200.geometry import EncoderGeometry
from chunkformer_b200.model import ChunkFormerModel
from chunkformer_b200.shard import split_recording
from chunkformer_b200.synth import synth_fbank, synth_state_dict
from oracle import chunkformer_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GEO = EncoderGeometry(d_model=256, heads=4, ffn=512, layers=3, kernel=15, vocab=120, has_cmvn=True)
MARGIN_TOL = 0.08


def _encoder_conf(geo):
    return dict(output_size=geo.d_model, attention_heads=geo.heads, linear_units=geo.ffn, num_blocks=geo.layers,
                input_layer="dw_striding", cnn_module_kernel=geo.kernel, cnn_module_norm="layer_norm", dynamic_conv=True,
                activation_type="swish", pos_enc_layer_type="chunk_rel_pos", selfattention_layer_type="chunk_rel_seflattn")


@pytest.fixture(scope="module")
def asr_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("asr_model")
    sd = synth_state_dict(GEO, 21)
    mean, istd = sd.pop("encoder.global_cmvn.mean"), sd.pop("encoder.global_cmvn.istd")
    sd["decoder.embed.0.weight"] = torch.zeros(4, 4)            # ignored like load_state_dict(strict=False)
    torch.save(sd, d / "pytorch_model.bin")
    yaml.safe_dump(dict(input_dim=80, output_dim=GEO.vocab, model="asr_model", encoder="chunkformer",
                        encoder_conf=_encoder_conf(GEO), ctc_conf=dict(ctc_blank_id=0)), open(d / "config.yaml", "w"))
    n = 1000.0   # global_cmvn json: mean_stat / var_stat / frame_num (utils/cmvn.py:23-46)
    var = (1.0 / istd) ** 2
    json.dump(dict(mean_stat=(mean * n).tolist(), var_stat=((var + mean * mean) * n).tolist(), frame_num=n),
              open(d / "global_cmvn", "w"))
    with open(d / "vocab.txt", "w", encoding="utf8") as f:
        f.write("<blank> 0\n<unk> 1\n")
        for i in range(2, GEO.vocab):
            f.write(("▁" if i % 4 == 0 else "") + f"t{i} {i}\n")
    return str(d)


@pytest.fixture(scope="module")
def asr(asr_dir):
    return ChunkFormerModel.from_pretrained(asr_dir, device=DEV)


def _oracle_sd():
    sd = synth_state_dict(GEO, 21)
    # the json round trip of the CMVN statistics is exact to fp32 rounding; use the same values the model loaded
    return sd


def test_from_pretrained_errors(tmp_path):
    with pytest.raises(ValueError):
        ChunkFormerModel.from_pretrained(str(tmp_path), device=DEV)          # no config
    yaml.safe_dump(dict(input_dim=80, output_dim=10, encoder_conf=_encoder_conf(GEO)), open(tmp_path / "config.yaml", "w"))
    with pytest.raises(ValueError):
        ChunkFormerModel.from_pretrained(str(tmp_path), device=DEV)          # no checkpoint


def test_encode_matches_oracle(asr):
    sd = _oracle_sd()
    lens = [280, 150]
    xb = torch.zeros(2, 280, 80)
    for k, t in enumerate(lens):
        xb[k, :t] = synth_fbank(t, seed=30 + k)
    out, out_lens = asr.encode(xb, torch.tensor(lens), chunk_size=16, left_context_size=32, right_context_size=16)
    ref, mask = O.forward_encoder(sd, GEO.heads, xb, lens, 16, 32, 16)
    assert out_lens.tolist() == mask.squeeze(1).sum(-1).tolist()
    for b, m in enumerate(out_lens.tolist()):
        assert (out[b, :m].cpu() - ref[b, :m]).abs().max().item() < 0.12
    with pytest.raises(NotImplementedError):
        asr.forward()


def test_endless_decode_matches_reference_driver(asr):
    """Segmented long-form decoding with carried caches == the oracle driven by the reference's segment arithmetic."""
    sd = _oracle_sd()
    c, l, r, tbd = 16, 32, 16, 20
    T = 3300
    x = synth_fbank(T, seed=40)
    res = asr.endless_decode(x, c, l, r, total_batch_duration=tbd, return_timestamps=True, max_silence_duration=0.16)
    text = asr.endless_decode(x, c, l, r, total_batch_duration=tbd, return_timestamps=False, max_silence_duration=0.16)
    assert isinstance(res, list) and isinstance(text, str) and all(set(i) == {"decode", "start", "end"} for i in res)
    # oracle driven exactly like chunkformer_model.py:391-434
    trunc, rel_right, segs = O.endless_segments(T, c, r, GEO.layers, tbd)
    assert len(segs) >= 3
    L, H, d = GEO.layers, GEO.heads, GEO.d_model
    att, cnn, off, outs = torch.zeros(L, l, H, 2 * d // H), torch.zeros(L, d, 7), [0], []
    for (s, e, last) in segs:
        o, ol, _, att, cnn, noff = O.forward_parallel_chunk(sd, H, [x[s:e]], [e - s], c, l, r, att, cnn, trunc, off)
        o = o.reshape(-1, d)[: int(ol[0])]
        if not last:
            o = o[:trunc]
        off = [int(noff[0]) - int(ol[0]) + o.shape[0]]
        outs.append(o)
    enc = torch.cat(outs, 0)
    tok, margin = O.ctc_greedy(sd, enc)
    asr.char_dict, keep = None, asr.char_dict
    got = asr.endless_decode(x, c, l, r, total_batch_duration=tbd).reshape(-1).cpu()
    asr.char_dict = keep
    assert got.shape == tok.shape
    assert bool(((got == tok) | (margin < MARGIN_TOL)).all())


def test_batch_decode_matches_oracle(asr):
    sd = _oracle_sd()
    lens = [900, 77, 1500, 300, 2200]
    xs = [synth_fbank(t, seed=60 + k) for k, t in enumerate(lens)]
    texts = asr.batch_decode(xs, 16, 32, 16, total_batch_duration=30)      # budget 1500 frames -> several groups
    assert len(texts) == len(lens) and all(isinstance(t, str) for t in texts)
    groups = O.batch_groups(lens, 30)
    assert len(groups) > 1
    asr.char_dict, keep = None, asr.char_dict
    hyps = asr.batch_decode(xs, 16, 32, 16, total_batch_duration=30)
    asr.char_dict = keep
    k = 0
    for grp in groups:
        out, enc_lens, n_chunks, _, _, _ = O.forward_parallel_chunk(sd, GEO.heads, [xs[i] for i in grp], [lens[i] for i in grp], 16, 32, 16)
        tok, margin = O.ctc_greedy(sd, out)
        row = 0
        for u, nck in enumerate(n_chunks):
            m = max(int(enc_lens[u]), 0)
            a = hyps[k].cpu()
            b = tok[row:row + nck].reshape(-1)[:m]
            mg = margin[row:row + nck].reshape(-1)[:m]
            assert a.shape == b.shape and bool(((a == b) | (mg < MARGIN_TOL)).all())
            row += nck
            k += 1


def _record_encoder_calls(enc, calls):
    orig = enc.forward_parallel_chunk

    def wrapped(*a, **kw):
        calls.append(dict(lens=[int(v) for v in kw["xs_origin_lens"].tolist()], trunc=int(kw.get("truncated_context_size", 0)),
                          offset=[int(v) for v in kw["offset"].tolist()]))
        return orig(*a, **kw)
    enc.forward_parallel_chunk = wrapped
    return orig


def test_endless_decode_reference_golden(asr, golden_dir):
    """endless_decode of THIS facade against goldens of the unmodified reference's endless_decode (fbank loader patched,
    tests/golden/make_golden_decode.py): every encoder call gets the reference's segment length / truncated_context_size /
    offset (exact), and the returned token ids are the reference's wherever its fp32 top-2 margin exceeds the tolerance."""
    import numpy as np
    g = np.load(os.path.join(golden_dir, "decode.npz"))
    texts = json.load(open(os.path.join(golden_dir, "decode_texts.json"), encoding="utf8"))
    n = len([k for k in g.files if k.startswith("endless") and k.endswith("_cfg")])
    assert n >= 5
    calls = []
    orig = _record_encoder_calls(asr.encoder, calls)
    keep = asr.char_dict
    try:
        for k in range(n):
            T, seed, c, l, r = (int(v) for v in g[f"endless{k}_cfg"])
            tbd, ms = (float(v) for v in g[f"endless{k}_tbd_ms"])
            x = synth_fbank(T, seed=seed)
            calls.clear()
            asr.char_dict = None
            got = asr.endless_decode(x, c, l, r, total_batch_duration=tbd).reshape(-1).cpu()
            assert [cl["lens"][0] for cl in calls] == g[f"endless{k}_seg_lens"].tolist(), k
            assert [cl["trunc"] for cl in calls] == g[f"endless{k}_seg_trunc"].tolist(), k
            assert [cl["offset"][0] for cl in calls] == g[f"endless{k}_seg_offset"].tolist(), k
            want = torch.from_numpy(g[f"endless{k}_tokens"].astype(np.int64))
            margin = torch.from_numpy(g[f"endless{k}_margin"])
            assert got.shape == want.shape, k
            ok = (got == want) | (margin < MARGIN_TOL)
            assert bool(ok.all()), (k, int((~ok).sum()))
            asr.char_dict = keep
            if bool((got == want).all()):                 # identical ids -> identical strings and stamps
                assert asr.endless_decode(x, c, l, r, total_batch_duration=tbd, return_timestamps=True,
                                          max_silence_duration=ms) == texts[f"endless{k}_stamps"]
    finally:
        asr.char_dict = keep
        asr.encoder.forward_parallel_chunk = orig


def test_batch_decode_reference_golden(asr, golden_dir):
    """batch_decode against goldens of the unmodified reference's batch_decode: same admission groups, same token ids above
    the margin tolerance."""
    import numpy as np
    g = np.load(os.path.join(golden_dir, "decode.npz"))
    n = len([k for k in g.files if k.startswith("batch") and k.endswith("_cfg")])
    calls = []
    orig = _record_encoder_calls(asr.encoder, calls)
    keep = asr.char_dict
    try:
        asr.char_dict = None
        for k in range(n):
            seed0, c, l, r = (int(v) for v in g[f"batch{k}_cfg"])
            tbd = float(g[f"batch{k}_tbd"][0])
            lens = g[f"batch{k}_lens"].tolist()
            xs = [synth_fbank(t, seed=seed0 + j) for j, t in enumerate(lens)]
            calls.clear()
            hyps = asr.batch_decode(xs, c, l, r, total_batch_duration=tbd)
            assert [len(cl["lens"]) for cl in calls] == g[f"batch{k}_group_sizes"].tolist(), k
            assert [h.numel() for h in hyps] == g[f"batch{k}_hyp_lens"].tolist(), k
            got = torch.cat([h.reshape(-1).cpu() for h in hyps])
            want = torch.from_numpy(g[f"batch{k}_tokens"].astype(np.int64))
            margin = torch.from_numpy(g[f"batch{k}_margin"])
            ok = (got == want) | (margin < MARGIN_TOL)
            assert bool(ok.all()), (k, int((~ok).sum()))
    finally:
        asr.char_dict = keep
        asr.encoder.forward_parallel_chunk = orig


@pytest.mark.parametrize("two_gpus", [False, True])
def test_batch_decode_devices_balanced_equals_oracle(asr, two_gpus):
    """batch_decode(devices=[...]): duration-balanced scheduling over several encoders.  Results come back in
    input order and every utterance's ids are the oracle's for that utterance alone (utterances are independent in a masked
    batch) wherever the fp32 margin exceeds the tolerance.  With one GPU the two 'devices' are the same one (the scheduling
    logic is what is tested); with two, weights are replicated to the second GPU."""
    if two_gpus and torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    devices = ["cuda:0", "cuda:1"] if two_gpus else ["cuda:0", "cuda:0"]
    sd = _oracle_sd()
    lens = [900, 77, 1500, 300, 2200, 15, 640, 3100]
    xs = [synth_fbank(t, seed=90 + k) for k, t in enumerate(lens)]
    texts = asr.batch_decode(xs, 16, 32, 16, total_batch_duration=30, devices=devices)
    assert len(texts) == len(lens) and all(isinstance(t, str) for t in texts)
    asr.char_dict, keep = None, asr.char_dict
    try:
        hyps = asr.batch_decode(xs, 16, 32, 16, total_batch_duration=30, devices=devices)
    finally:
        asr.char_dict = keep
    toks, margins, _ = O.batch_decode_tokens(sd, GEO.heads, xs, 16, 32, 16, 1e9)
    for k in range(len(lens)):
        a = hyps[k].cpu()
        assert a.shape == toks[k].shape, k
        assert bool(((a == toks[k]) | (margins[k] < MARGIN_TOL)).all()), k


def test_classify_audio_full_and_chunked(tmp_path):
    geo = EncoderGeometry(d_model=256, heads=4, ffn=512, layers=2, kernel=15, vocab=0)
    sd = synth_state_dict(geo, 31)
    g = torch.Generator().manual_seed(9)
    tasks = {"gender": 2, "emotion": 7}
    for name, n in tasks.items():
        sd[f"classification_heads.{name}.linear.weight"] = torch.randn(n, 256, generator=g) * 0.2
        sd[f"classification_heads.{name}.linear.bias"] = torch.randn(n, generator=g) * 0.1
    torch.save(sd, tmp_path / "pytorch_model.pt")
    yaml.safe_dump(dict(input_dim=80, model="classification", encoder="chunkformer", encoder_conf=_encoder_conf(geo),
                        model_conf=dict(tasks=tasks)), open(tmp_path / "config.yaml", "w"))
    json.dump({"gender": {"0": "female", "1": "male"}}, open(tmp_path / "label_mapping.json", "w"))
    m = ChunkFormerModel.from_pretrained(str(tmp_path), device=DEV)
    assert m.is_classification and m.get_tasks() == tasks
    x = synth_fbank(500, seed=70)
    for cfg in ((-1, -1, -1), (16, 32, 16)):
        res = m.classify_audio(x, *cfg)
        c, l, r = (61, 0, 0) if cfg[0] < 0 else cfg          # full attention = one chunk of T' = 61 frames
        ref, mask = O.forward_encoder(sd, geo.heads, x.unsqueeze(0), [500], c, l, r)
        pooled = (ref * mask.transpose(1, 2).float()).sum(1) / mask.sum()
        for name in tasks:
            logits = pooled @ sd[f"classification_heads.{name}.linear.weight"].T + sd[f"classification_heads.{name}.linear.bias"]
            p = torch.softmax(logits, -1)[0]
            assert set(res[name]) == {"label", "label_id", "prob"}
            top2 = p.topk(2).values
            if float(top2[0] - top2[1]) > 0.05:
                assert res[name]["label_id"] == int(p.argmax())
                assert abs(res[name]["prob"] - float(p.max())) < 0.03
        assert res["gender"]["label"] in ("female", "male")
    res2 = m.classify_audio(x, -1, -1, -1)
    assert abs(res2["emotion"]["prob"] - m.classify_audio(x, -1, -1, -1)["emotion"]["prob"]) < 1e-6   # deterministic
    with pytest.raises(ValueError):
        m.endless_decode(x)


def test_long_recording_chunk_range_shards_equal_unsharded(asr):
    """contiguous chunk ranges with recomputed halos, each run as a stand-alone utterance, reproduce the
    one-shot result (ranks emulated one after another on one GPU)."""
    c, l, r = 16, 32, 16
    T = 9000
    x = synth_fbank(T, seed=80)
    enc = asr.encoder

    def encode(frames):
        out, el, *_ = enc.forward_parallel_chunk([frames], torch.tensor([frames.shape[0]], dtype=torch.int32), c, l, r,
                                                 offset=torch.zeros(1, dtype=torch.int32))
        return out.reshape(-1, GEO.d_model)[: int(el[0])]
    full = encode(x)
    parts = []
    for sh in split_recording(T, c, l, r, GEO.layers, 4, "exact"):
        parts.append(encode(x[sh.in_start:sh.in_end])[sh.keep_lo:sh.keep_hi])
    sharded = torch.cat(parts, 0)
    assert sharded.shape == full.shape
    assert (sharded - full).abs().max().item() < 0.03


# ------------------------------------------------------------------------------------------------ device CTC compaction
def _host_compact(tokens, starts, lens, mode):
    from chunkformer_b200.postprocess import ctc_collapse
    res = []
    for s, n in zip(starts, lens):
        seg = tokens[s:s + n].tolist()
        if mode == 1:
            keep = [(t, i) for i, t in enumerate(seg) if t != 0]
        else:
            keep = [(t, i) for i, t in enumerate(seg) if t != 0 and (i == 0 or seg[i - 1] != t)]
            assert [t for t, _ in keep] == ctc_collapse(seg)
        res.append(keep)
    return res


@pytest.mark.parametrize("rows,blank_p", [(1, 0.5), (37, 0.0), (2048, 0.9), (2049, 0.5), (50000, 0.97), (70001, 0.3)])
def test_ctc_compact_bit_exact(asr, rows, blank_p):
    """cf_ctc_compact == remove_duplicates_and_blank / blank filter per utterance (bit-exact ids, frame indices, offsets):
    ragged utterances with chunk padding rows between them, empty utterances (also sharing a start row, also at the very
    end of the buffer), runs of repeats crossing the 2048-row tile boundary."""
    gen = torch.Generator().manual_seed(rows)
    tok = torch.randint(1, 6, (rows,), generator=gen)
    tok = tok.repeat_interleave(torch.randint(1, 4, (rows,), generator=gen))[:rows]       # runs of repeated ids
    tok[torch.rand(rows, generator=gen) < blank_p] = 0
    starts, lens, pos = [], [], 0
    while pos < rows:
        n = int(torch.randint(0, max(2, rows // 3), (1,), generator=gen))
        n = min(n, rows - pos)
        starts.append(pos); lens.append(n)
        if torch.rand(1, generator=gen) < 0.3:
            starts.append(pos); lens.append(0)            # empty utterance in front of the next one's rows
            starts[-2], starts[-1] = starts[-1], starts[-2]; lens[-2], lens[-1] = lens[-1], lens[-2]
        pos += n + int(torch.randint(0, 70, (1,), generator=gen))                           # padding rows owned by nobody
    starts.append(rows); lens.append(0)                   # empty utterance at the end of the buffer
    dtok = tok.to(DEV)
    for mode in (0, 1):
        got = asr.encoder.ctc_compact(dtok, starts, lens, mode=mode)
        want = _host_compact(tok, starts, lens, mode)
        assert len(got) == len(want)
        for (gt, gf), w in zip(got, want):
            assert gt.tolist() == [t for t, _ in w] and gf.tolist() == [i for _, i in w]


def test_decode_text_equals_host_postprocessing_of_raw_tokens(asr):
    """endless_decode / batch_decode texts (device compaction + short-list segmentation) == the reference's host
    post-processing (utils/model_utils.py:164-222, pinned in tests/test_postprocess.py) applied to the raw token ids."""
    from chunkformer_b200.postprocess import get_output, get_output_with_timestamps
    x = synth_fbank(3300, seed=40)
    for ms in (0.0, 0.16, 0.5):
        res = asr.endless_decode(x, 16, 32, 16, total_batch_duration=20, return_timestamps=True, max_silence_duration=ms)
        cd, asr.char_dict = asr.char_dict, None
        raw = asr.endless_decode(x, 16, 32, 16, total_batch_duration=20)
        asr.char_dict = cd
        assert res == get_output_with_timestamps(raw.cpu(), cd, "asr_model", ms)[0]
    lens = [900, 77, 1500, 10, 300]
    xs = [synth_fbank(t, seed=60 + k) for k, t in enumerate(lens)]
    texts = asr.batch_decode(xs, 16, 32, 16, total_batch_duration=30)
    cd, asr.char_dict = asr.char_dict, None
    hyps = asr.batch_decode(xs, 16, 32, 16, total_batch_duration=30)
    asr.char_dict = cd
    assert texts == get_output([h.cpu() for h in hyps], cd, "asr_model")


# ------------------------------------------------------------------------------------------------ transducer facade
@pytest.fixture(scope="module")
def rnnt(tmp_path_factory):
    from chunkformer_b200.synth import synth_transducer_state_dict
    d = tmp_path_factory.mktemp("rnnt_model")
    geo = EncoderGeometry(d_model=256, heads=4, ffn=512, layers=3, kernel=15, vocab=0, has_cmvn=False)
    sd = synth_state_dict(geo, 33)
    sd.update(synth_transducer_state_dict(60, 32, 64, 2, 48, geo.d_model, 64, blank_bias=4.0, seed=34))
    torch.save(sd, d / "pytorch_model.bin")
    yaml.safe_dump(dict(input_dim=80, output_dim=60, model="transducer", encoder="chunkformer", encoder_conf=_encoder_conf(geo),
                        predictor="rnn", predictor_conf=dict(embed_size=32, output_size=48, hidden_size=64, num_layers=2,
                                                             rnn_type="lstm"),
                        joint="transducer_joint", joint_conf=dict(enc_output_size=256, pred_output_size=48, join_dim=64,
                                                                 prejoin_linear=True, postjoin_linear=False, joint_mode="add",
                                                                 activation="tanh"),
                        ctc_conf=dict(ctc_blank_id=0)), open(d / "config.yaml", "w"))
    with open(d / "vocab.txt", "w", encoding="utf8") as f:
        f.write("<blank> 0\n<unk> 1\n")
        for i in range(2, 60):
            f.write(("▁" if i % 4 == 0 else "") + f"t{i} {i}\n")
    return ChunkFormerModel.from_pretrained(str(d), device=DEV), sd


def test_transducer_decode_paths(rnnt):
    """endless_decode / batch_decode on a `model: transducer` checkpoint: the facade's symbols are the CPU oracle's greedy
    search on the facade's own encoder output (search parity with identical inputs is tests/test_gpu_transducer.py), and
    texts / stamps equal the reference's host post-processing of the symbol grid (chunkformer_model.py:440-456, 532-543)."""
    from chunkformer_b200.postprocess import get_output, get_output_with_timestamps
    from oracle import transducer_oracle as TO
    model, sd = rnnt
    assert model.transducer is not None and model.ctc is None
    x = synth_fbank(2500, seed=70)
    cd, model.char_dict = model.char_dict, None
    grid = model.endless_decode(x, 16, 32, 16, total_batch_duration=20)          # (1, T', 64) like optimized_search
    model.char_dict = cd
    assert grid.dim() == 3 and grid.shape[0] == 1 and grid.shape[2] == 64 and int((grid != 0).sum()) > 5
    for ms in (0.0, 0.24):
        res = model.endless_decode(x, 16, 32, 16, total_batch_duration=20, return_timestamps=True, max_silence_duration=ms)
        assert res == get_output_with_timestamps([grid[0].cpu()], cd, "transducer", ms)[0]
    # the same symbols as the oracle search on the encoder rows the facade computed
    out, lens, n_chunks, _, _, _ = model.encoder.forward_parallel_chunk([x], torch.tensor([2500], dtype=torch.int), 16, 32, 16,
                                                                        offset=torch.zeros(1, dtype=torch.int))
    rows = out.reshape(-1, out.shape[-1])[: int(lens[0])].cpu()
    want, margins = TO.greedy_search_one(sd, rows, rows.shape[0], 64, 0, want_margin=True)
    lens5 = [600, 90, 1400]
    xs = [synth_fbank(t, seed=80 + k) for k, t in enumerate(lens5)]
    texts = model.batch_decode(xs, 16, 32, 16, total_batch_duration=30)
    cd, model.char_dict = model.char_dict, None
    hyps = model.batch_decode(xs, 16, 32, 16, total_batch_duration=30)
    model.char_dict = cd
    assert len(texts) == 3 and texts == get_output(hyps, cd, "transducer") and all(isinstance(h, list) for h in hyps)
    got = model.batch_decode([x], 16, 32, 16, total_batch_duration=60)
    cd, model.char_dict = model.char_dict, None
    hyp = model.batch_decode([x], 16, 32, 16, total_batch_duration=60)[0]
    model.char_dict = cd
    want_list = want[want != 0].tolist()
    if hyp != want_list:
        k = next((i for i, (a, b) in enumerate(zip(hyp, want_list)) if a != b), min(len(hyp), len(want_list)))
        assert min(m for _, _, m in margins) < 1e-4, f"diverges at symbol {k} without a near-tie"
    assert isinstance(got[0], str)
