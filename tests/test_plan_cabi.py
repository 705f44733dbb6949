"""C-ABI library on CPU: loads, exports every declared symbol, and its packer matches the reference bit-exactly
(golden masks captured from the reference's own packer) and the oracle on random cases."""
import os
import re

import numpy as np

from chunkformer_b200 import lib as cflib
from chunkformer_b200.plan import Plan
from oracle import chunkformer_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = cflib.load()
    header = open(os.path.join(ROOT, "include", "chunkformer_b200.h")).read()
    declared = set(re.findall(r"^CF_API [^;(]*?\b(cf_[a-z_]+)\(", header, flags=re.M))
    assert len(declared) >= 20
    assert declared == set(cflib.SIGNATURES), declared ^ set(cflib.SIGNATURES)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.cf_version()


def test_packer_matches_reference_masks(golden_dir):
    g = np.load(os.path.join(golden_dir, "plan_cases.npz"), allow_pickle=True)
    for case in g["plan_cases"]:
        p = Plan(case["c"], case["l"], case["r"], case["lens"], case["offsets"])
        assert p.n == case["n"]
        W = case["l"] + case["c"] + case["r"]
        att = np.unpackbits(case["att"])[: p.n * W].reshape(p.n, W).astype(bool)
        cv = np.unpackbits(case["conv"])[: p.n * (case["c"] + 14)].reshape(p.n, case["c"] + 14).astype(bool)
        a, c = p.masks()
        assert np.array_equal(a, att), case
        assert np.array_equal(c, cv), case


def test_packer_matches_oracle_random():
    rng = np.random.RandomState(7)
    for _ in range(300):
        c = int(rng.choice([4, 6, 8, 16, 64, 128]))
        l = int(rng.choice([0, 40, 64, 128, 256]))
        r = int(rng.choice([0, 5, 64, 128]))
        B = int(rng.randint(1, 8))
        lens = [int(rng.choice([rng.randint(1, 20), rng.randint(15, 6000), rng.randint(15, 400000)])) for _ in range(B)]
        offs = [int(rng.choice([0, rng.randint(0, 2000)])) for _ in range(B)]
        p = Plan(c, l, r, lens, offs)
        o = O.make_plan(lens, offs, c, l, r, 15)
        assert p.n == o.n and p.n_chunks == [int(v) for v in o.n_chunks]
        assert np.array_equal(p.enc_lens, o.enc_lens)
        if p.n * (l + c + r) < 2_000_000:
            a, cv = p.masks()
            assert np.array_equal(a, o.att_mask) and np.array_equal(cv, o.conv_mask)
        t = p.chunk_table()
        assert np.array_equal(t[:, 0], o.chunk_utt) and np.array_equal(t[:, 1], o.chunk_idx)


def test_calc_length_and_edge_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "plan_cases.npz"), allow_pickle=True)
    ts = [int(t) for t in g["calc_len_T"]]
    p = Plan(64, 128, 128, ts)
    assert np.array_equal(p.enc_lens, g["calc_len"])
    # utterances shorter than one frame of context still give one (fully masked) chunk
    p = Plan(64, 128, 128, [3])
    assert p.n == 1 and not p.masks()[0].any() and int(p.enc_lens[0]) == -1


def test_padded_plan_geometry():
    p = Plan(8, 16, 16, [333, 180, 95], padded_T=333)
    Tp = (333 - 15) // 8 + 1
    nck = (Tp + 7) // 8
    assert p.n == 3 * nck and p.n_chunks == [nck] * 3
    assert [int(v) for v in p.enc_lens] == [O.calc_length(t) for t in (333, 180, 95)]
    t = p.chunk_table()
    # key slots: frame f = 8 j - 16 + q valid iff 0 <= f < len'
    for row in t:
        u, j, alo, ahi = int(row[0]), int(row[1]), int(row[2]), int(row[3])
        m = O.calc_length([333, 180, 95][u])
        f = 8 * j - 16 + np.arange(40)
        ok = (f >= 0) & (f < m)
        exp = np.zeros(40, bool)
        exp[alo:ahi] = True
        assert np.array_equal(ok, exp)
