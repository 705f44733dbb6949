"""Golden outputs of the reference's transducer greedy search (predictor.py, joint.py, search/greedy_search.py) on synthetic
weights: token grids of optimized_search and hypotheses of batch_greedy_search, for a small and a mid-size geometry.

    python tests/golden/make_golden_transducer.py        # needs /root/reference; writes tests/golden/transducer.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ref_import import import_reference  # noqa: E402

import_reference()
from chunkformer.transducer.joint import TransducerJoint  # noqa: E402
from chunkformer.transducer.predictor import RNNPredictor  # noqa: E402
from chunkformer.transducer.search.greedy_search import batch_greedy_search, optimized_search  # noqa: E402

from chunkformer_b200.synth import synth_transducer_state_dict  # noqa: E402

CASES = {
    # name: (vocab, embed, hidden, layers, pred_out, enc_dim, join_dim, blank_bias, B, T, n_steps, seed)
    "tiny": (40, 16, 32, 2, 24, 20, 28, 5.0, 3, 30, 4, 11),
    "tiny_cap": (12, 8, 16, 1, 16, 16, 16, -3.0, 2, 9, 3, 12),       # blank is rare: the per-frame symbol cap binds
    "mid": (1024, 256, 512, 2, 512, 512, 512, 7.2, 4, 60, 64, 13),   # rnnt-large-vie head sizes (yaml:25-45)
}
out = {}
for name, (V, emb, hid, nl, po, E, J, bb, B, T, n_steps, seed) in CASES.items():
    sd = synth_transducer_state_dict(V, emb, hid, nl, po, E, J, blank_bias=bb, seed=seed)
    pred = RNNPredictor(V, emb, po, 0.1, hid, nl, rnn_type="lstm", dropout=0.1).eval()
    joint = TransducerJoint(V, E, po, J, prejoin_linear=True, postjoin_linear=False, joint_mode="add", activation="tanh").eval()
    pred.load_state_dict({k[len("predictor."):]: v for k, v in sd.items() if k.startswith("predictor.")}, strict=True)
    joint.load_state_dict({k[len("joint."):]: v for k, v in sd.items() if k.startswith("joint.")}, strict=True)
    model = types.SimpleNamespace(predictor=pred, joint=joint, blank=0)
    g = torch.Generator().manual_seed(seed + 100)
    enc = torch.randn((B, T, E), generator=g)
    lens = torch.tensor([T] + [int(v) for v in torch.randint(1, T + 1, (B - 1,), generator=g)])
    with torch.no_grad():
        grid = optimized_search(model, enc, lens, n_steps)
        hyps = batch_greedy_search(model, enc, lens, n_steps)
    out[name + "_cfg"] = np.array([V, emb, hid, nl, po, E, J, B, T, n_steps, seed], dtype=np.int64)
    out[name + "_blank_bias"] = np.array([bb], dtype=np.float64)
    out[name + "_lens"] = lens.numpy()
    out[name + "_grid"] = grid.numpy()
    out[name + "_hyp_flat"] = np.array([t for h in hyps for t in h], dtype=np.int64)
    out[name + "_hyp_len"] = np.array([len(h) for h in hyps], dtype=np.int64)
    print(name, "tokens per utterance", [len(h) for h in hyps], "frames", lens.tolist())
np.savez_compressed(os.path.join(HERE, "transducer.npz"), **out)
