"""The transducer greedy-search oracle (oracle/transducer_oracle.py) against golden outputs of the unmodified reference
(tests/golden/make_golden_transducer.py -> transducer.npz): token grids of optimized_search and hypotheses of
batch_greedy_search must match bit-exactly."""
import os

import numpy as np
import pytest
import torch

from chunkformer_b200.synth import synth_transducer_state_dict
from oracle import transducer_oracle as T


def load_case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "transducer.npz"))
    V, emb, hid, nl, po, E, J, B, Tn, n_steps, seed = [int(v) for v in g[name + "_cfg"]]
    sd = synth_transducer_state_dict(V, emb, hid, nl, po, E, J, blank_bias=float(g[name + "_blank_bias"][0]), seed=seed)
    gen = torch.Generator().manual_seed(seed + 100)
    enc = torch.randn((B, Tn, E), generator=gen)
    lens = torch.tensor([Tn] + [int(v) for v in torch.randint(1, Tn + 1, (B - 1,), generator=gen)])
    assert lens.tolist() == g[name + "_lens"].tolist()
    hyps, pos = [], 0
    for n in g[name + "_hyp_len"].tolist():
        hyps.append(g[name + "_hyp_flat"][pos:pos + n].tolist())
        pos += n
    return sd, enc, lens, n_steps, torch.from_numpy(g[name + "_grid"]), hyps


@pytest.mark.parametrize("name", ["tiny", "tiny_cap", "mid"])
def test_oracle_matches_reference_search(golden_dir, name):
    sd, enc, lens, n_steps, grid, hyps = load_case(golden_dir, name)
    got = T.optimized_search(sd, enc, lens.tolist(), n_steps)
    assert got.shape == grid.shape and torch.equal(got, grid)
    assert T.batch_greedy_search(sd, enc, lens.tolist(), n_steps) == hyps
    assert any(len(h) > 0 for h in hyps) and any((grid[b].reshape(-1, n_steps)[:, 0] == 0).any() for b in range(grid.shape[0]))
