"""Golden for `cnn_module_norm: batch_norm` (the reference constructor's default, modules/encoder.py:59, convolution.py:83-89):
the UNMODIFIED reference in eval mode on a tiny model with non-trivial BatchNorm running statistics.
    python tests/golden/make_golden_bn.py          # -> tests/golden/tiny_bn.npz
forward_parallel_chunk on a ragged masked batch (two window shapes) and encode() on a padded batch."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_golden import build_reference  # noqa: E402
from chunkformer_b200.geometry import EncoderGeometry  # noqa: E402
from chunkformer_b200.synth import synth_fbank  # noqa: E402

TINY_BN = EncoderGeometry(d_model=64, heads=2, ffn=128, layers=2, kernel=15, vocab=50, conv_norm="batch_norm")


@torch.no_grad()
def main():
    torch.set_num_threads(4)
    model, sd = build_reference(TINY_BN, seed=6)
    enc = model.model.encoder
    assert not enc.training and isinstance(enc.encoders[0].conv_module.norm, torch.nn.BatchNorm1d)
    res = {}
    lens = [700, 9, 131, 600, 15]
    xs = [synth_fbank(t, seed=100 + k) for k, t in enumerate(lens)]
    for ci, (c, l, r) in enumerate([(8, 16, 16), (16, 32, 0)]):
        o, ol, nck, _, _, _ = enc.forward_parallel_chunk(
            xs=xs, xs_origin_lens=torch.tensor(lens, dtype=torch.int), chunk_size=c, left_context_size=l, right_context_size=r,
            offset=torch.zeros(len(lens), dtype=torch.int))
        res[f"c{ci}_cfg"] = np.array([c, l, r])
        res[f"c{ci}_lens"] = np.array(lens)
        res[f"c{ci}_out"] = o.numpy()
        res[f"c{ci}_enc_lens"] = ol.numpy()
        res[f"c{ci}_n_chunks"] = np.array(nck)
    lens = [333, 180]
    xb = torch.zeros(len(lens), max(lens), 80)
    for k, t in enumerate(lens):
        xb[k, :t] = synth_fbank(t, seed=300 + k)
    o, lens_o = model.encode(xb, torch.tensor(lens), chunk_size=8, left_context_size=16, right_context_size=16)
    res["e0_cfg"], res["e0_lens"], res["e0_out"], res["e0_out_lens"] = np.array([8, 16, 16]), np.array(lens), o.numpy(), lens_o.numpy()
    np.savez_compressed(os.path.join(HERE, "tiny_bn.npz"), **res)
    print("tiny_bn.npz")


if __name__ == "__main__":
    main()
