"""Host-side multi-GPU logic on CPU: LPT partition, halo arithmetic, and world_size-2 gloo runs in which the oracle
stands in for the device encoder (sharded result == unsharded result)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chunkformer_b200.geometry import EncoderGeometry
from chunkformer_b200.plan import Plan
from chunkformer_b200.shard import (chunks_of, decode_batch_sharded, encode_recording_sharded, halo_chunks,
                                    partition_by_chunks, split_recording)
from chunkformer_b200.synth import masked_batch_lengths, synth_fbank, synth_state_dict
from oracle import chunkformer_oracle as O

TINY = EncoderGeometry(d_model=64, heads=2, ffn=128, layers=2, kernel=15, vocab=50)


def test_chunks_of_matches_packer():
    rng = np.random.RandomState(0)
    for _ in range(200):
        c = int(rng.choice([4, 8, 16, 64]))
        T = int(rng.randint(1, 20000))
        assert chunks_of(T, c) == Plan(c, 0, 0, [T]).n


def test_lpt_partition_balanced_and_complete():
    lens = masked_batch_lengths()
    for world in (1, 2, 4, 8):
        bins = partition_by_chunks(lens, 64, world)
        assert sorted(i for b in bins for i in b) == list(range(len(lens)))
        loads = [sum(chunks_of(lens[i], 64) for i in b) for b in bins]
        biggest = max(chunks_of(t, 64) for t in lens)
        assert max(loads) <= sum(loads) / world + biggest


def test_split_recording_covers_every_row_once():
    for T, world in ((5759998, 8), (100000, 4), (3000, 2), (700, 4)):
        c, l, r, L = 64, 128, 128, 17
        shards = split_recording(T, c, l, r, L, world)
        M = 1 + (T - 15) // 8
        rows = 0
        for s in shards:
            assert 0 <= s.in_start <= s.in_end <= T
            rows += s.keep_hi - s.keep_lo
            if s.chunk_hi > s.chunk_lo:
                assert s.in_start % (8 * c) == 0          # cut on the global chunk grid
        assert rows == M
    assert halo_chunks(64, 128, 128, 17, "exact") == (51, 51)
    assert halo_chunks(64, 128, 128, 17, "reference")[1] == 34


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd = synth_state_dict(TINY, 3)
    c, l, r = 8, 16, 8
    # ---- batch sharding: tokens of every utterance, gathered in the original order
    lens = [300, 45, 700, 120, 15, 510]
    xs = [synth_fbank(t, seed=50 + k) for k, t in enumerate(lens)]

    def tokens_fn(sub_xs, sub_lens):
        out, enc_lens, n_chunks, _, _, _ = O.forward_parallel_chunk(sd, TINY.heads, sub_xs, sub_lens, c, l, r)
        tok, _ = O.ctc_greedy(sd, out)
        res, row = [], 0
        for u, nck in enumerate(n_chunks):
            res.append(tok[row:row + nck].reshape(-1)[: max(int(enc_lens[u]), 0)])
            row += nck
        return res
    got = decode_batch_sharded(tokens_fn, xs, lens, c)
    ref = tokens_fn(xs, lens)
    ok_batch = all(torch.equal(a, b) for a, b in zip(got, ref))
    # ---- one long recording split by chunk range with exact halos
    T = 2600
    x = synth_fbank(T, seed=77)

    def encode_fn(frames):
        out, enc_lens, _, _, _, _ = O.forward_parallel_chunk(sd, TINY.heads, [frames], [frames.shape[0]], c, l, r)
        return out.reshape(-1, TINY.d_model)[: int(enc_lens[0])]
    full = encode_fn(x)
    sharded = encode_recording_sharded(encode_fn, x, c, l, r, TINY.layers, "exact")
    err = float((full - sharded).abs().max()) if full.shape == sharded.shape else float("inf")
    if rank == 0:
        q.put((ok_batch, err, tuple(sharded.shape), tuple(full.shape)))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    ok_batch, err, shp, full_shp = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_batch
    assert shp == full_shp and err < 1e-4, (shp, full_shp, err)


def _worker_idle_ranks(rank, world, port, q):
    """More ranks than utterances / chunks: idle ranks take part in the gathers with empty tensors on the same device."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd = synth_state_dict(TINY, 3)
    c, l, r = 8, 16, 8
    lens = [300, 45]
    xs = [synth_fbank(t, seed=50 + k) for k, t in enumerate(lens)]

    def tokens_fn(sub_xs, sub_lens):
        out, enc_lens, n_chunks, _, _, _ = O.forward_parallel_chunk(sd, TINY.heads, sub_xs, sub_lens, c, l, r)
        tok, _ = O.ctc_greedy(sd, out)
        res, row = [], 0
        for u, nck in enumerate(n_chunks):
            res.append(tok[row:row + nck].reshape(-1)[: max(int(enc_lens[u]), 0)])
            row += nck
        return res
    got = decode_batch_sharded(tokens_fn, xs, lens, c)
    ref = tokens_fn(xs, lens)
    ok_batch = all(torch.equal(a, b) for a, b in zip(got, ref))
    x = synth_fbank(100, seed=78)                        # 2 chunks over 3 ranks: rank 0 owns none

    def encode_fn(frames):
        out, enc_lens, _, _, _, _ = O.forward_parallel_chunk(sd, TINY.heads, [frames], [frames.shape[0]], c, l, r)
        return out.reshape(-1, TINY.d_model)[: int(enc_lens[0])]
    full = encode_fn(x)
    sharded = encode_recording_sharded(encode_fn, x, c, l, r, TINY.layers, "exact")
    err = float((full - sharded).abs().max()) if full.shape == sharded.shape else float("inf")
    if rank == 0:
        q.put((ok_batch, err, tuple(sharded.shape), tuple(full.shape)))
    dist.destroy_process_group()


def test_world_size_3_gloo_with_idle_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_idle_ranks, args=(rk, 3, port, q)) for rk in range(3)]
    for p in procs:
        p.start()
    ok_batch, err, shp, full_shp = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_batch
    assert shp == full_shp and err < 1e-4, (shp, full_shp, err)
