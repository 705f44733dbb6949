"""CTC prefix beam search: oracle and product host search against goldens of the reference's function
(tests/golden/make_golden_beam.py: the unmodified modules/search.py:131-249 with its log_add helper repaired to take the
two-argument call of PrefixScore.score(), without which it raises on the first frame)."""
import json
import os

import numpy as np
import pytest
import torch

from chunkformer_b200 import postprocess as P
from oracle import search_oracle as S

HERE = os.path.dirname(os.path.abspath(__file__))


def _cases():
    z = np.load(os.path.join(HERE, "golden", "beam.npz"))
    cases = json.loads(str(z["cases"]))
    return [(torch.from_numpy(z[f"logp_{k}"]), c) for k, c in enumerate(cases)]


def test_oracle_matches_reference_golden():
    n = 0
    for logp, c in _cases():
        for b, (num_t, want) in enumerate(zip(c["lens"], c["results"])):
            rows = [[float(x) for x in r] for r in logp[b]]
            nbest, scores, times = S.ctc_prefix_beam_search(rows, num_t, c["beam"])
            assert nbest == want["nbest"] and times == want["nbest_times"]
            assert max(abs(a - w) for a, w in zip(scores, want["nbest_scores"])) < 1e-9
            n += 1
    assert n == 15


def test_host_search_matches_reference_golden():
    """The product's host search on top-k lists: same n-best lists, frame times and scores as the reference on full rows."""
    for logp, c in _cases():
        res = P.ctc_prefix_beam_search(logp, torch.tensor(c["lens"]), c["beam"])
        for r, want in zip(res, c["results"]):
            assert r.tokens == want["tokens"] and r.times == want["times"]
            assert r.nbest == want["nbest"] and r.nbest_times == want["nbest_times"]
            assert abs(r.score - want["score"]) < 1e-9
            assert max(abs(a - w) for a, w in zip(r.nbest_scores, want["nbest_scores"])) < 1e-9


def test_host_search_edge_cases():
    """Empty utterance, beam wider than the vocabulary, all-blank frames, a context graph is rejected."""
    logp = torch.log_softmax(torch.randn(2, 7, 5, generator=torch.Generator().manual_seed(3)), -1)
    res = P.ctc_prefix_beam_search(logp, torch.tensor([0, 7]), 50)
    assert res[0].tokens == [] and res[0].score == 0.0 and len(res[1].nbest) <= 50
    blank = torch.full((1, 6, 4), -20.0)
    blank[..., 0] = 0.0
    r = P.ctc_prefix_beam_search(torch.log_softmax(blank, -1), torch.tensor([6]), 3)[0]
    assert r.tokens == [] and r.times == []
    with pytest.raises(NotImplementedError):
        P.ctc_prefix_beam_search(logp, torch.tensor([7, 7]), 3, context_graph=object())
    # beam 1 on a peaky distribution is the greedy collapse
    peaky = torch.full((1, 9, 6), -30.0)
    for t, v in enumerate([0, 2, 2, 0, 2, 3, 3, 0, 1]):
        peaky[0, t, v] = 0.0
    r = P.ctc_prefix_beam_search(torch.log_softmax(peaky, -1), torch.tensor([9]), 1)[0]
    assert r.tokens == P.ctc_collapse([0, 2, 2, 0, 2, 3, 3, 0, 1]) == [2, 2, 3, 1]


@pytest.mark.gpu
def test_encoder_beam_search_on_device_log_probs():
    """ChunkFormerEncoderB200.ctc_prefix_beam_search: the library's log-probabilities and the device top-k feed the same host search
    as the oracle run on those log-probabilities; the best hypothesis scores at least the greedy path's collapse."""
    from chunkformer_b200.encoder import ChunkFormerEncoderB200
    from chunkformer_b200.geometry import EncoderGeometry
    from chunkformer_b200.synth import synth_fbank, synth_state_dict
    geo = EncoderGeometry(d_model=256, heads=4, ffn=512, layers=2, kernel=15, vocab=200)
    enc = ChunkFormerEncoderB200(geo, synth_state_dict(geo, 5), "cuda:0")
    xb = torch.zeros(2, 500, 80)
    xb[0], xb[1, :260] = synth_fbank(500, seed=21), synth_fbank(260, seed=22)
    out, masks = enc.forward_encoder(xb, torch.tensor([500, 260]), 16, 32, 16)
    lens = masks.squeeze(1).sum(1)
    res = enc.ctc_prefix_beam_search(out, lens, beam_size=6)
    for b in range(2):
        n = int(lens[b])
        tok, logp = enc.ctc_greedy(out[b, :n], want_logp=True)
        rows = [[float(x) for x in r] for r in logp.cpu()]
        nbest, scores, times = S.ctc_prefix_beam_search(rows, n, 6)
        assert res[b].nbest == nbest and res[b].nbest_times == times
        assert max(abs(a - w) for a, w in zip(res[b].nbest_scores, scores)) < 1e-6
        greedy_path = float(logp.max(dim=1).values.sum())
        assert res[b].score >= greedy_path - 1e-4            # the beam sums over alignments of its best prefix
        assert len(res[b].times) == len(res[b].tokens)
