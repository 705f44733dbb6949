"""Golden outputs of the reference's frame-synchronous streaming entry points (modules/encoder.py:310-459) on the tiny
synthetic model: `forward_chunk_by_chunk` outputs and the caches `forward_chunk` returns, at the shipped streaming presets'
shape (right context 0; apps/realtime-asr/config.py:86-110 uses chunk 4/6/8, left 40/50/60).

    python tests/golden/make_golden_stream.py        # needs /root/reference; writes tests/golden/stream.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import TINY, build_reference  # noqa: E402
from chunkformer_b200.synth import synth_fbank  # noqa: E402

CASES = [(8, 40, 2, 555), (4, 12, 1, 300), (6, 20, 3, 411)]     # (chunk, left, batch, T)
CASES_R = [(8, 16, 4, 2, 555), (4, 12, 4, 1, 300), (8, 24, 12, 2, 411), (16, 32, 7, 1, 700)]   # (chunk, left, right, batch, T)


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    model, sd = build_reference(TINY, seed=3)
    enc = model.model.encoder
    L, H, d = TINY.layers, TINY.heads, TINY.d_model
    out = {"cases": np.array(CASES, dtype=np.int64)}
    for k, (c, l, B, T) in enumerate(CASES):
        xs = torch.stack([synth_fbank(T, seed=20 + 7 * k + b) for b in range(B)])
        y, mask = enc.forward_chunk_by_chunk(xs, torch.full((B,), T, dtype=torch.long), c, l, 0)
        out[f"c{k}_out"] = y.numpy()
        out[f"c{k}_mask"] = mask.numpy()
        # three explicit forward_chunk steps: pins the cache layouts (L, B, H, l, 2 d_k) / (L, B, d, 7) and the offset mask
        size, stride = 8 * (c - 1) + 15, 8 * c
        att = torch.zeros((L, B, H, l, 2 * d // H))
        cnn = torch.zeros((L, B, d, 7))
        for step in range(3):
            o, _, att, cnn = enc.forward_chunk(xs[:, step * stride: step * stride + size], att, cnn, c, l, 0, offset=step * c)
        out[f"c{k}_step3_out"] = o.numpy()
        out[f"c{k}_step3_att"] = att.numpy()
        out[f"c{k}_step3_cnn"] = cnn.numpy()
        print((c, l, B, T), tuple(y.shape), tuple(att.shape), tuple(cnn.shape))
    np.savez_compressed(os.path.join(HERE, "stream.npz"), **out)
    # ---- right context > 0 (encoder.py:310-385: chunk + right context embedded and attended as one chunk, conv cut at the
    # chunk grid, caches end at the chunk): a separate file so that stream.npz stays as round 1 generated it
    out = {"cases": np.array(CASES_R, dtype=np.int64)}
    for k, (c, l, r, B, T) in enumerate(CASES_R):
        xs = torch.stack([synth_fbank(T, seed=40 + 7 * k + b) for b in range(B)])
        y, mask = enc.forward_chunk_by_chunk(xs, torch.full((B,), T, dtype=torch.long), c, l, r)
        out[f"c{k}_out"] = y.numpy()
        out[f"c{k}_mask"] = mask.numpy()
        size, stride = 8 * (c - 1) + 15 + 8 * r, 8 * c
        att = torch.zeros((L, B, H, l, 2 * d // H))
        cnn = torch.zeros((L, B, d, 7))
        for step in range(3):
            o, _, att, cnn = enc.forward_chunk(xs[:, step * stride: step * stride + size], att, cnn, c, l, r, offset=step * c)
        out[f"c{k}_step3_out"] = o.numpy()
        out[f"c{k}_step3_att"] = att.numpy()
        out[f"c{k}_step3_cnn"] = cnn.numpy()
        print((c, l, r, B, T), tuple(y.shape), tuple(o.shape), tuple(att.shape), tuple(cnn.shape))
    np.savez_compressed(os.path.join(HERE, "stream_right.npz"), **out)


if __name__ == "__main__":
    main()
