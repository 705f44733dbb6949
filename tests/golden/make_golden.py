"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run once in the build container:  python tests/golden/make_golden.py
The fixtures pin oracle/chunkformer_oracle.py (tests/test_oracle_golden.py) and, through the
oracle, the CUDA path.  Weights come from chunkformer_b200.synth.synth_state_dict(geometry, seed)
loaded into the reference model with load_state_dict, inputs from synth_fbank(T, seed); both are
re-generated bit-identically by the tests, so only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from ref_import import import_reference, reference_config_dict  # noqa: E402
from chunkformer_b200.geometry import EncoderGeometry  # noqa: E402
from chunkformer_b200.synth import synth_fbank, synth_state_dict  # noqa: E402

TINY = EncoderGeometry(d_model=64, heads=2, ffn=128, layers=2, kernel=15, vocab=50)
TINY_CMVN = EncoderGeometry(d_model=64, heads=2, ffn=128, layers=2, kernel=15, vocab=50, has_cmvn=True)
LARGE = EncoderGeometry(d_model=512, heads=8, ffn=2048, layers=17, kernel=15, vocab=5000)


def build_reference(geo: EncoderGeometry, seed: int):
    Model, Config = import_reference()
    cfg = reference_config_dict(geo.d_model, geo.heads, geo.ffn, geo.layers, geo.vocab, geo.kernel, geo.conv_norm)
    model = Model(Config.from_dict(cfg)).eval()
    sd = synth_state_dict(geo, seed)
    enc = model.model.encoder
    if geo.has_cmvn:
        from chunkformer.modules.cmvn import GlobalCMVN
        enc.global_cmvn = GlobalCMVN(sd["encoder.global_cmvn.mean"].clone(),
                                     sd["encoder.global_cmvn.istd"].clone())
    missing, unexpected = model.model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("decoder.") for k in missing), [k for k in missing if not k.startswith("decoder.")]
    return model, sd


def capture_masks(enc, xs, lens, c, l, r, offsets):
    """Run the reference packer + embed and grab the two masks handed to layer 0."""
    rec = {}
    layer0 = enc.encoders[0]
    orig = layer0.forward_parallel_chunk

    class Stop(Exception):
        pass

    def hook(x, att_mask, pos_emb, mask_pad=None, **kw):
        rec["att"] = att_mask.clone()
        rec["conv"] = mask_pad.clone()
        rec["pos"] = pos_emb.clone()
        rec["n"] = x.shape[0]
        raise Stop()

    layer0.forward_parallel_chunk = hook
    try:
        enc.forward_parallel_chunk(xs=xs, xs_origin_lens=torch.tensor(lens, dtype=torch.int),
                                   chunk_size=c, left_context_size=l, right_context_size=r,
                                   offset=torch.tensor(offsets, dtype=torch.int))
    except Stop:
        pass
    finally:
        layer0.forward_parallel_chunk = orig
    return rec


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    out = {}

    # ---- 1. packer / masks: random ragged cases through the reference's own packer -----------------
    model, sd = build_reference(TINY, seed=3)
    enc = model.model.encoder
    rng = np.random.RandomState(1234)
    cases = []
    for case in range(120):
        c = int(rng.choice([4, 8, 16, 64]))
        l = int(rng.choice([0, 8, 16, 40, 128]))
        r = int(rng.choice([0, 3, 8, 16, 128]))
        B = int(rng.randint(1, 5))
        lens = [int(rng.choice([rng.randint(1, 15), rng.randint(15, 8 * c + 40), rng.randint(15, 2500)]))
                for _ in range(B)]
        offsets = [int(rng.choice([0, 0, rng.randint(1, 300)])) for _ in range(B)]
        xs = [torch.zeros(t, 80) for t in lens]
        rec = capture_masks(enc, xs, lens, c, l, r, offsets)
        cases.append(dict(c=c, l=l, r=r, lens=lens, offsets=offsets, n=rec["n"],
                          att=np.packbits(rec["att"].numpy().reshape(-1)),
                          conv=np.packbits(rec["conv"].numpy().reshape(-1))))
    out["plan_cases"] = np.array(cases, dtype=object)
    # calc_length / xs_lens over a range of T
    ts = list(range(1, 200)) + [519, 520, 1030, 5998, 6000, 359998]
    out["calc_len_T"] = np.array(ts)
    out["calc_len"] = enc.embed.calc_length(torch.tensor(ts, dtype=torch.int)).numpy()
    np.savez_compressed(os.path.join(HERE, "plan_cases.npz"), **out)
    print("plan_cases.npz", len(cases))

    # ---- 2. tiny model: full forward_parallel_chunk, masked batch, with and without CMVN ----------
    for tag, geo, seed in (("tiny", TINY, 3), ("tiny_cmvn", TINY_CMVN, 4)):
        model, sd = build_reference(geo, seed)
        enc = model.model.encoder
        res = {}
        for ci, (c, l, r) in enumerate([(8, 16, 16), (16, 32, 0), (8, 16, 3), (64, 128, 128)]):
            lens = [700, 9, 131, 64 * c + 77, 15]
            xs = [synth_fbank(t, seed=100 + k) for k, t in enumerate(lens)]
            o, ol, nck, _, _, off = enc.forward_parallel_chunk(
                xs=xs, xs_origin_lens=torch.tensor(lens, dtype=torch.int), chunk_size=c,
                left_context_size=l, right_context_size=r, offset=torch.zeros(len(lens), dtype=torch.int))
            logp = model.model.ctc.log_softmax(o)
            res[f"c{ci}_cfg"] = np.array([c, l, r])
            res[f"c{ci}_lens"] = np.array(lens)
            res[f"c{ci}_out"] = o.numpy()
            res[f"c{ci}_enc_lens"] = ol.numpy()
            res[f"c{ci}_n_chunks"] = np.array(nck)
            res[f"c{ci}_offset"] = off.numpy()
            res[f"c{ci}_tokens"] = logp.argmax(-1).numpy()
            res[f"c{ci}_logp_sample"] = logp[:, ::7, ::5].numpy()
        # streaming: three sequential segments with caches, as endless_decode drives it
        c, l, r = 8, 16, 16
        T = 1500
        x = synth_fbank(T, seed=200)
        trunc = c * 6
        L, H, d = geo.layers, geo.heads, geo.d_model
        att_cache = torch.zeros(L, l, H, 2 * d // H)
        cnn_cache = torch.zeros(L, d, 7)
        offset = torch.zeros(1, dtype=torch.int)
        rel_right = (max(r, 7) + max(c, max(r, 7)) * (L - 1)) * 8
        outs = []
        for idx in range(3):
            start = trunc * 8 * idx
            end = min(trunc * 8 * (idx + 1) + 7, T)
            seg = x[start:end + rel_right]
            o, ol, _, att_cache, cnn_cache, offset = enc.forward_parallel_chunk(
                xs=[seg], xs_origin_lens=torch.tensor([seg.shape[0]], dtype=torch.int), chunk_size=c,
                left_context_size=l, right_context_size=r, att_cache=att_cache, cnn_cache=cnn_cache,
                truncated_context_size=trunc, offset=offset)
            o = o.reshape(1, -1, d)[:, :ol]
            o = o[:, :trunc]
            offset = offset - ol + o.shape[1]
            outs.append(o[0].numpy())
            res[f"stream_att_cache_{idx}"] = att_cache.numpy()
            res[f"stream_cnn_cache_{idx}"] = cnn_cache.numpy()
            res[f"stream_offset_{idx}"] = offset.numpy().copy()  # the next call mutates `offset` in place (encoder.py:674)
        res["stream_out"] = np.concatenate(outs, 0)
        res["stream_cfg"] = np.array([c, l, r, T, trunc, rel_right])
        # ---- forward_encoder (the path behind ChunkFormerModel.encode), padded batch
        lens = [333, 180, 95]
        Tm = max(lens)
        xb = torch.zeros(len(lens), Tm, 80)
        for k, t in enumerate(lens):
            xb[k, :t] = synth_fbank(t, seed=300 + k)
        for ei, (c, l, r) in enumerate([(8, 16, 16), (16, 32, 0), (4, 40, 0)]):
            o, lens_o = model.encode(xb, torch.tensor(lens), chunk_size=c, left_context_size=l,
                                     right_context_size=r)
            res[f"e{ei}_cfg"] = np.array([c, l, r])
            res[f"e{ei}_lens"] = np.array(lens)
            res[f"e{ei}_out"] = o.numpy()
            res[f"e{ei}_out_lens"] = lens_o.numpy()
        np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **res)
        print(f"{tag}.npz")

    # ---- 3. CTC-large geometry (BASELINE.json configs[0]): one 60 s utterance, 64/128/128 -----------
    model, sd = build_reference(LARGE, seed=0)
    enc = model.model.encoder
    T = 5998
    x = synth_fbank(T, seed=1)
    o, ol, nck, _, _, _ = enc.forward_parallel_chunk(
        xs=[x], xs_origin_lens=torch.tensor([T], dtype=torch.int), chunk_size=64, left_context_size=128,
        right_context_size=128, offset=torch.zeros(1, dtype=torch.int))
    n_valid = int(ol[0])
    flat = o.reshape(-1, 512)[:n_valid]
    logits = model.model.ctc.ctc_lo(flat)
    top2 = logits.topk(2, -1).values
    res = dict(cfg=np.array([64, 128, 128, T]), enc_len=np.array([n_valid]), n_chunks=np.array(nck),
               out_rows=flat[::8].numpy(), out_rms=np.array([float(flat.pow(2).mean().sqrt())]),
               out_absmax=np.array([float(flat.abs().max())]),
               out_colsum=flat.double().sum(0).numpy(), out_rowsum=flat.double().sum(1).numpy(),
               tokens=model.model.ctc.log_softmax(flat.unsqueeze(0))[0].argmax(-1).numpy().astype(np.int32),
               margin=(top2[:, 0] - top2[:, 1]).numpy())
    np.savez_compressed(os.path.join(HERE, "ctc_large_60s.npz"), **res)
    print("ctc_large_60s.npz", n_valid, nck)


if __name__ == "__main__":
    main()
