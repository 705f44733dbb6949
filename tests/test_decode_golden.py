"""The facade drivers against goldens of the UNMODIFIED reference (tests/golden/make_golden_decode.py): endless_decode's
segment arithmetic / cache carry-over (chunkformer_model.py:320-459) and batch_decode's admission (:461-552).

CPU part: the oracle's drivers (oracle.endless_decode_tokens / batch_decode_tokens) reproduce, call for call, the arguments the
reference's drivers hand to the encoder (segment lengths, truncated_context_size, offsets, group sizes: exact), the token ids
(identical except at fp32 near-ties) and, through chunkformer_b200.postprocess, the texts and time stamps.  The GPU part of the
same goldens is tests/test_gpu_model.py::test_*_reference_golden."""
import json
import os

import numpy as np
import pytest
import torch

from chunkformer_b200.geometry import EncoderGeometry
from chunkformer_b200.postprocess import get_output, get_output_with_timestamps
from chunkformer_b200.synth import synth_fbank, synth_state_dict
from oracle import chunkformer_oracle as O

GEO = EncoderGeometry(d_model=256, heads=4, ffn=512, layers=3, kernel=15, vocab=120, has_cmvn=True)
SEED = 21
TIE = 1e-3          # fp32 oracle vs fp32 reference: summation order only


def char_dict(vocab):
    cd = {0: "<blank>", 1: "<unk>"}
    for i in range(2, vocab):
        cd[i] = ("▁" if i % 4 == 0 else "") + f"t{i}"
    return cd


@pytest.fixture(scope="module")
def gold(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode.npz"))
    texts = json.load(open(os.path.join(golden_dir, "decode_texts.json"), encoding="utf8"))
    return g, texts


def n_cases(g, prefix):
    return len([k for k in g.files if k.startswith(prefix) and k.endswith("_cfg")])


def test_endless_decode_oracle_matches_reference_golden(gold):
    g, texts = gold
    sd = synth_state_dict(GEO, SEED)
    cd = char_dict(GEO.vocab)
    assert n_cases(g, "endless") >= 5
    for k in range(n_cases(g, "endless")):
        T, seed, c, l, r = (int(v) for v in g[f"endless{k}_cfg"])
        tbd, ms = (float(v) for v in g[f"endless{k}_tbd_ms"])
        tok, margin, calls = O.endless_decode_tokens(sd, GEO.heads, GEO.layers, synth_fbank(T, seed=seed), c, l, r, tbd)
        assert [cl["len"] for cl in calls] == g[f"endless{k}_seg_lens"].tolist(), k
        assert [cl["trunc"] for cl in calls] == g[f"endless{k}_seg_trunc"].tolist(), k
        assert [cl["offset"] for cl in calls] == g[f"endless{k}_seg_offset"].tolist(), k
        want = torch.from_numpy(g[f"endless{k}_tokens"].astype(np.int64))
        wm = torch.from_numpy(g[f"endless{k}_margin"])
        assert tok.shape == want.shape, k
        assert bool(((tok == want) | (wm < TIE)).all()), k
        assert float((margin - wm).abs().max()) < 1e-3
        # texts / stamps through this repo's post-processing of the REFERENCE's tokens == the reference's own strings
        res = get_output_with_timestamps([want.reshape(-1, 1)], cd, "asr_model", ms)[0]
        assert res == texts[f"endless{k}_stamps"], k
        assert " ".join(i["decode"] for i in res).strip() == texts[f"endless{k}_text"], k


def test_batch_decode_oracle_matches_reference_golden(gold):
    g, texts = gold
    sd = synth_state_dict(GEO, SEED)
    cd = char_dict(GEO.vocab)
    for k in range(n_cases(g, "batch")):
        seed0, c, l, r = (int(v) for v in g[f"batch{k}_cfg"])
        tbd = float(g[f"batch{k}_tbd"][0])
        lens = g[f"batch{k}_lens"].tolist()
        xs = [synth_fbank(t, seed=seed0 + j) for j, t in enumerate(lens)]
        toks, margins, sizes = O.batch_decode_tokens(sd, GEO.heads, xs, c, l, r, tbd)
        assert sizes == g[f"batch{k}_group_sizes"].tolist(), k
        assert [t.numel() for t in toks] == g[f"batch{k}_hyp_lens"].tolist(), k
        want = torch.from_numpy(g[f"batch{k}_tokens"].astype(np.int64))
        wm = torch.from_numpy(g[f"batch{k}_margin"])
        got = torch.cat(toks)
        assert bool(((got == want) | (wm < TIE)).all()), k
        hyps = list(want.split(g[f"batch{k}_hyp_lens"].tolist()))
        assert get_output(hyps, cd, "asr_model") == texts[f"batch{k}_texts"], k
