"""Goldens for the facade drivers: `endless_decode` (chunkformer_model.py:320-459) and `batch_decode` (:461-552) of the
UNMODIFIED reference, run on CPU with `_load_audio_and_extract_features` replaced by a function that returns synthetic
fbank (the reference's loader needs pydub + an audio file; everything after it is the reference's own code).

    python tests/golden/make_golden_decode.py          # -> tests/golden/decode.npz (a few seconds)

Recorded per case: the arguments of every `encoder.forward_parallel_chunk` call the driver made (segment input lengths,
truncated_context_size, incoming offset: pins oracle.endless_segments / oracle.batch_groups and the facade's own
arithmetic), the raw greedy token ids the driver returns with `char_dict = None`, the fp32 top-2 logit margin of every
frame, and the text / timestamp output with a synthetic vocabulary.
Model: d 256, H 4, F 512, L 3, V 120 with CMVN, synth_state_dict seed 21 (the geometry of tests/test_gpu_model.py)."""
import json
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

GEO = EncoderGeometry(d_model=256, heads=4, ffn=512, layers=3, kernel=15, vocab=120, has_cmvn=True)
SEED = 21

# (T input frames, fbank seed, c, l, r, total_batch_duration, max_silence_duration)
ENDLESS_CASES = [
    (3300, 40, 16, 32, 16, 20, 0.16),      # several segments, the last one short
    (5000, 41, 8, 16, 8, 10, 0.5),         # stops on the rel-right-context test
    (2055, 42, 16, 32, 0, 20.5, 0.0),      # right context 0 (the conv lorder sets the look-ahead), fractional duration
    (700, 43, 16, 32, 16, 1800, 0.5),      # one segment (default budget)
    (4111, 44, 4, 40, 0, 6, 0.24),         # streaming-preset sizes, many segments
    (2576, 45, 16, 32, 16, 20.48, 0.5),    # xs_len a multiple of the segment stride + 8: range() ends on its own
]
# (lens, first fbank seed, c, l, r, total_batch_duration)
BATCH_CASES = [
    ([900, 77, 1500, 300, 2200], 60, 16, 32, 16, 30),
    ([120, 2400, 15, 9, 640, 640, 3000, 50], 70, 8, 16, 8, 25),
    ([500], 80, 16, 32, 16, 1800),
]


def char_dict(vocab):
    cd = {0: "<blank>", 1: "<unk>"}
    for i in range(2, vocab):
        cd[i] = ("▁" if i % 4 == 0 else "") + f"t{i}"
    return cd


@torch.no_grad()
def main():
    torch.set_num_threads(4)
    model, _ = build_reference(GEO, SEED)
    enc, ctc = model.model.encoder, model.model.ctc
    feats = {}
    model._load_audio_and_extract_features = lambda key: (feats[key], int(feats[key].shape[0]))
    calls, enc_outs = [], []
    orig_fpc, orig_ls = enc.forward_parallel_chunk, ctc.log_softmax

    def fpc(**kw):
        calls.append(dict(lens=[int(v) for v in kw["xs_origin_lens"].tolist()], trunc=int(kw.get("truncated_context_size", 0)),
                          offset=[int(v) for v in kw["offset"].tolist()]))
        return orig_fpc(**kw)

    def ls(x):
        enc_outs.append(x.detach().clone())
        return orig_ls(x)
    enc.forward_parallel_chunk, ctc.log_softmax = fpc, ls

    def margins(x):
        top2 = ctc.ctc_lo(x).topk(2, -1).values
        return (top2[..., 0] - top2[..., 1])

    res, texts = {}, {}
    for k, (T, seed, c, l, r, tbd, ms) in enumerate(ENDLESS_CASES):
        feats["a"] = synth_fbank(T, seed=seed)
        calls.clear(); enc_outs.clear()
        model.char_dict = None
        tok = model.endless_decode("a", c, l, r, total_batch_duration=tbd)
        res[f"endless{k}_cfg"] = np.array([T, seed, c, l, r], dtype=np.int64)
        res[f"endless{k}_tbd_ms"] = np.array([tbd, ms], dtype=np.float64)
        res[f"endless{k}_seg_lens"] = np.array([cl["lens"][0] for cl in calls])
        res[f"endless{k}_seg_trunc"] = np.array([cl["trunc"] for cl in calls])
        res[f"endless{k}_seg_offset"] = np.array([cl["offset"][0] for cl in calls])
        res[f"endless{k}_tokens"] = tok.reshape(-1).numpy().astype(np.int16)
        res[f"endless{k}_margin"] = margins(enc_outs[0][0]).numpy().astype(np.float32)
        model.char_dict = char_dict(GEO.vocab)
        texts[f"endless{k}_stamps"] = model.endless_decode("a", c, l, r, total_batch_duration=tbd, return_timestamps=True,
                                                           max_silence_duration=ms)
        texts[f"endless{k}_text"] = model.endless_decode("a", c, l, r, total_batch_duration=tbd, return_timestamps=False,
                                                         max_silence_duration=ms)
        print(f"endless {k}: {len(res[f'endless{k}_seg_lens'])} segments, {tok.numel()} frames")
    for k, (lens, seed0, c, l, r, tbd) in enumerate(BATCH_CASES):
        keys = []
        for j, t in enumerate(lens):
            feats[f"u{j}"] = synth_fbank(t, seed=seed0 + j)
            keys.append(f"u{j}")
        calls.clear(); enc_outs.clear()
        model.char_dict = None
        hyps = model.batch_decode(keys, c, l, r, total_batch_duration=tbd)
        res[f"batch{k}_cfg"] = np.array([seed0, c, l, r], dtype=np.int64)
        res[f"batch{k}_tbd"] = np.array([tbd], dtype=np.float64)
        res[f"batch{k}_lens"] = np.array(lens)
        res[f"batch{k}_group_sizes"] = np.array([len(cl["lens"]) for cl in calls])
        res[f"batch{k}_hyp_lens"] = np.array([h.numel() for h in hyps])
        res[f"batch{k}_tokens"] = np.concatenate([h.numpy().astype(np.int16) for h in hyps])
        # margins of every valid frame, utterance by utterance in arrival order
        mg, u = [], 0
        for call, out in zip(calls, enc_outs):    # one (n, c, d) tensor per admitted group
            m, row = margins(out), 0
            for t in call["lens"]:
                nck = max(1, -(-(t - 7) // (8 * c)))                 # chunks of an utterance of t frames (encoder.py:557-562)
                mg.append(m[row:row + nck].reshape(-1)[: hyps[u].numel()].numpy().astype(np.float32))
                row += nck
                u += 1
            assert row == out.shape[0]
        res[f"batch{k}_margin"] = np.concatenate(mg)
        model.char_dict = char_dict(GEO.vocab)
        texts[f"batch{k}_texts"] = model.batch_decode(keys, c, l, r, total_batch_duration=tbd)
        print(f"batch {k}: groups {res[f'batch{k}_group_sizes'].tolist()}")
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **res)
    with open(os.path.join(HERE, "decode_texts.json"), "w", encoding="utf8") as f:
        json.dump(texts, f, ensure_ascii=False, indent=0)
    print("decode.npz, decode_texts.json")


if __name__ == "__main__":
    main()
