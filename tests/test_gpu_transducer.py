"""Device transducer greedy search (cf_rnnt_greedy) against the reference's golden token grids and the CPU oracle."""
import pytest
import torch

from chunkformer_b200.synth import synth_transducer_state_dict
from chunkformer_b200.transducer import TransducerGreedyB200
from oracle import transducer_oracle as T
from test_transducer_oracle import load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# fp32 on both sides; only the summation order differs (~1e-6 on logits of spread ~1.5).  A decision whose fp32 top-2 logit
# margin is below this may legitimately differ; everything after such a decision is a different (equally valid) search path.
MARGIN_TOL = 1e-4


def _assert_same_until_near_tie(sd, enc, n_frames, n_steps, got_tok, got_fr, label):
    grid, margins = T.greedy_search_one(sd, enc, n_frames, n_steps, 0, want_margin=True)
    want = [(t, int(grid[t, k])) for t in range(grid.shape[0]) for k in range(n_steps) if int(grid[t, k]) != 0]
    got = list(zip(got_fr.tolist(), got_tok.tolist()))
    if got == want:
        return
    tight = [(t, s) for (t, s, m) in margins if m < MARGIN_TOL]
    k = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), min(len(got), len(want)))
    t_div = min(got[k][0] if k < len(got) else 10 ** 9, want[k][0] if k < len(want) else 10 ** 9)
    assert any(t <= t_div for t, _ in tight), f"{label}: hypotheses diverge at symbol {k} (frame {t_div}) without a near-tie"


@pytest.mark.parametrize("name", ["tiny", "tiny_cap", "mid"])
def test_search_matches_reference_golden(golden_dir, name):
    """optimized_search grid and batch_greedy_search hypotheses == the unmodified reference's (bit-exact token ids)."""
    sd, enc, lens, n_steps, grid, hyps = load_case(golden_dir, name)
    srch = TransducerGreedyB200(sd, blank=0, device=DEV)
    got = srch.optimized_search(enc.to(DEV), lens, n_steps)
    if not torch.equal(got, grid):
        for b in range(enc.shape[0]):
            tok, fr = srch.search_flat(enc[b].to(DEV), [0], [int(lens[b])], n_steps)[0]
            _assert_same_until_near_tie(sd, enc[b], int(lens[b]), n_steps, tok, fr, f"{name}[{b}]")
    else:
        assert srch.batch_greedy_search(enc.to(DEV), lens, n_steps) == hyps
    assert srch.last_iterations > 0


def test_flat_ragged_batch_matches_oracle():
    """Utterances packed in one flat row buffer with unowned rows between them, lengths 0 .. 300, the rnnt-large head sizes."""
    V, emb, hid, nl, po, E, J = 1024, 256, 512, 2, 512, 512, 512
    sd = synth_transducer_state_dict(V, emb, hid, nl, po, E, J, blank_bias=7.2, seed=5)
    gen = torch.Generator().manual_seed(9)
    lens = [300, 0, 17, 1, 120, 64, 9]
    starts, pos = [], 0
    for n in lens:
        starts.append(pos)
        pos += n + 13
    enc = torch.randn((pos, E), generator=gen)
    srch = TransducerGreedyB200(sd, device=DEV)
    res = srch.search_flat(enc.to(DEV), starts, lens, n_steps=6)
    assert len(res) == len(lens)
    total = 0
    for b, (tok, fr) in enumerate(res):
        _assert_same_until_near_tie(sd, enc[starts[b]:starts[b] + lens[b]], lens[b], 6, tok, fr, f"utt{b}")
        total += tok.numel()
    assert total > 50 and res[1][0].numel() == 0
    # the speculative blank-run batching: far fewer iterations than (frames + symbols) of the longest utterance
    assert srch.last_iterations <= 64 * ((300 + total) // 64 + 1)


def test_capacity_overflow_is_reported_and_retried():
    sd = synth_transducer_state_dict(12, 8, 16, 1, 16, 16, 16, blank_bias=-3.0, seed=12)    # emits on nearly every step
    enc = torch.randn((40, 16), generator=torch.Generator().manual_seed(3))
    srch = TransducerGreedyB200(sd, device=DEV)
    with pytest.raises(RuntimeError, match="capacity"):
        srch.search_flat(enc.to(DEV), [0], [40], n_steps=8, capacity=10)
    tok, fr = srch.search_flat(enc.to(DEV), [0], [40], n_steps=8)[0]                           # automatic capacity grows
    _assert_same_until_near_tie(sd, enc, 40, 8, tok, fr, "cap")
    assert tok.numel() > 4 * 40


def test_unsupported_heads_are_rejected():
    sd = synth_transducer_state_dict(12, 8, 16, 1, 16, 16, 16, seed=1)
    bad = dict(sd); bad["joint.post_ffn.weight"] = torch.zeros(16, 16)
    with pytest.raises(ValueError):
        TransducerGreedyB200(bad, device=DEV)
    with pytest.raises(ValueError):
        TransducerGreedyB200({k: v for k, v in sd.items() if not k.startswith("predictor.rnn")}, device=DEV)


@pytest.mark.parametrize("V", [300, 512, 256])
def test_many_utterances_use_several_predictor_tiles(V):
    """21 utterances = three tiles of 8 in the predictor kernels, several joint CTAs per utterance, lengths 0 .. 40; every
    hypothesis against the oracle.  V = 300 -> 3 vocabulary tiles (five-launch path), 512 / 256 -> clusters of 4 / 2 CTAs
    (projection + joint + greedy control in one cluster launch; the golden `mid` case covers clusters of 8)."""
    emb, hid, nl, po, E, J = 64, 128, 2, 96, 80, 112
    sd = synth_transducer_state_dict(V, emb, hid, nl, po, E, J, blank_bias=4.5, seed=21)
    gen = torch.Generator().manual_seed(22)
    lens = [int(v) for v in torch.randint(0, 41, (21,), generator=gen)]
    lens[3], lens[17] = 40, 1
    starts, pos = [], 0
    for n in lens:
        starts.append(pos)
        pos += n + 3
    enc = torch.randn((pos, E), generator=gen)
    srch = TransducerGreedyB200(sd, device=DEV)
    res = srch.search_flat(enc.to(DEV), starts, lens, n_steps=5)
    total = 0
    for b, (tok, fr) in enumerate(res):
        _assert_same_until_near_tie(sd, enc[starts[b]:starts[b] + lens[b]], lens[b], 5, tok, fr, f"utt{b}")
        total += tok.numel()
    assert total > 100


def test_persistent_kernel_equals_launch_per_phase():
    """The persistent cooperative search kernel (weights resident in shared memory, grid barriers) against the launch-per-phase
    path on the same handle: same symbols and frames, or the first difference sits at a decision whose fp32 top-2 margin is
    below 1e-4 (the two sum the joint in different orders); rnnt-large head sizes, ragged batch with empty utterances."""
    V, emb, hid, nl, po, E, J = 1024, 256, 512, 2, 512, 512, 512
    sd = synth_transducer_state_dict(V, emb, hid, nl, po, E, J, blank_bias=7.2, seed=5)
    gen = torch.Generator().manual_seed(19)
    lens = [700, 0, 17, 1, 350, 64, 9, 1200, 33]
    starts, pos = [], 0
    for n in lens:
        starts.append(pos)
        pos += n + 5
    enc = torch.randn((pos, E), generator=gen)
    srch = TransducerGreedyB200(sd, device=DEV)
    out = {}
    for persistent in (1, 0):
        srch.set_option("persistent", persistent)
        out[persistent] = srch.search_flat(enc.to(DEV), starts, lens, n_steps=6)
        assert srch.last_iterations > 0
    srch.set_option("persistent", 1)
    for b, ((t1, f1), (t0, f0)) in enumerate(zip(out[1], out[0])):
        if t1.shape == t0.shape and bool((t1 == t0).all()) and bool((f1 == f0).all()):
            continue
        _assert_same_until_near_tie(sd, enc[starts[b]:starts[b] + lens[b]], lens[b], 6, t1, f1, f"persistent utt{b}")
        _assert_same_until_near_tie(sd, enc[starts[b]:starts[b] + lens[b]], lens[b], 6, t0, f0, f"per-phase utt{b}")
