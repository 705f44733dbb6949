"""Multi-GPU partitioning of the path: only what shards naturally, no per-layer collective.

* many utterances: masked batches are independent (the bound tables isolate utterances, encoder.py:567-645), so they are
  split across ranks by chunk count with longest-processing-time-first bin packing; every rank holds a weight replica.
* one long recording: contiguous chunk ranges per rank, cut on the global chunk grid (as the reference's own segments are,
  chunkformer_model.py:364-365), each extended by recomputed context halos; results are gathered once at the end.

Halo sizes (chunks): exact math needs ceil(L*(l+c)/c) on the left and ceil(L*(r+c)/c) on the right because the 15-tap
conv leaks +-7 frames across chunk edges in every layer; the reference's own streaming uses the smaller right halo
r + max(c, r)*(L-1) frames (chunkformer_model.py:344-345) and exact left state through caches.
"""
import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import torch


def chunks_of(T: int, c: int) -> int:
    """Chunks the packer makes for an utterance of T input frames (encoder.py:557-562)."""
    size, step = 8 * (c - 1) + 15, 8 * c
    pad = (step - ((T - size) % step)) % step if T >= size else size - T
    return (T + pad - size) // step + 1


def partition_by_chunks(lens: Sequence[int], c: int, world: int) -> List[List[int]]:
    """LPT greedy bin packing of utterances over `world` ranks by chunk count; returns utterance indices per rank
    (each list in original order)."""
    order = sorted(range(len(lens)), key=lambda i: -chunks_of(int(lens[i]), c))
    load = [0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        g = min(range(world), key=lambda k: (load[k], k))
        bins[g].append(i)
        load[g] += chunks_of(int(lens[i]), c)
    return [sorted(b) for b in bins]


@dataclass
class RecordingShard:
    rank: int
    chunk_lo: int        # first chunk owned (global chunk grid)
    chunk_hi: int        # one past the last chunk owned
    in_start: int        # input frame range handed to the encoder (own chunks + halos)
    in_end: int
    keep_lo: int         # encoder rows of the shard's output to keep: [keep_lo, keep_hi)
    keep_hi: int


def halo_chunks(c: int, l: int, r: int, layers: int, mode: str = "exact") -> Tuple[int, int]:
    if mode == "exact":
        return math.ceil(layers * (l + c) / c), math.ceil(layers * (r + c) / c)
    if mode == "reference":
        rr = max(r, 7)
        return math.ceil(layers * l / c), math.ceil((rr + max(c, rr) * (layers - 1)) / c)
    raise ValueError("mode must be 'exact' or 'reference'")


def split_recording(T: int, c: int, l: int, r: int, layers: int, world: int, mode: str = "exact") -> List[RecordingShard]:
    """Contiguous chunk ranges with recomputed halos for one recording of T input frames."""
    n = chunks_of(T, c)
    M = max(0, 1 + (T - 15) // 8)           # valid encoder frames
    hL, hR = halo_chunks(c, l, r, layers, mode)
    shards = []
    for g in range(world):
        a, b = (n * g) // world, (n * (g + 1)) // world
        if a == b:
            shards.append(RecordingShard(g, a, b, 0, 0, 0, 0))
            continue
        a2, b2 = max(0, a - hL), min(n, b + hR)
        in_start = 8 * c * a2
        in_end = T if b2 == n else min(T, 8 * c * b2 + 7)
        keep_lo = (a - a2) * c
        keep_hi = min((b - a2) * c, M - a2 * c)
        shards.append(RecordingShard(g, a, b, in_start, in_end, keep_lo, max(keep_lo, keep_hi)))
    return shards


def gather_variable(t: torch.Tensor, group=None) -> List[torch.Tensor]:
    """All-gather 1-D or 2-D tensors whose first dimension differs per rank (one size exchange + one padded gather)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[: int(s)] for o, s in zip(outs, sizes)]


def comm_device(group=None, device=None) -> torch.device:
    """Device the collectives of `group` run on: the caller's choice, else the current CUDA device under NCCL and the CPU
    otherwise.  EVERY rank must use the same kind of device, also a rank that got no work (its tensors are empty, not absent)."""
    import torch.distributed as dist
    if device is not None:
        return torch.device(device)
    if "nccl" in str(dist.get_backend(group)):
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def decode_batch_sharded(tokens_fn: Callable[[List[torch.Tensor], List[int]], List[torch.Tensor]],
                         xs: Sequence[torch.Tensor], lens: Sequence[int], c: int, group=None, device=None) -> List[torch.Tensor]:
    """Every rank encodes its LPT share of the batch with `tokens_fn(xs_subset, lens_subset) -> [tokens per utterance]`
    and the per-utterance greedy token ids are gathered once, back in the original order, on every rank (on `device`, see
    comm_device; ranks whose bin is empty, i.e. more ranks than utterances, take part with empty tensors)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bins = partition_by_chunks(lens, c, world)
    mine = bins[rank]
    toks = tokens_fn([xs[i] for i in mine], [int(lens[i]) for i in mine]) if mine else []
    dev = comm_device(group, device)
    flat = torch.cat([t.reshape(-1).to(dev, torch.int64) for t in toks]) if toks else torch.zeros(0, dtype=torch.int64, device=dev)
    counts = torch.tensor([t.numel() for t in toks], dtype=torch.int64, device=dev)
    all_flat = gather_variable(flat, group)
    all_counts = gather_variable(counts, group)
    out: List[Optional[torch.Tensor]] = [None] * len(lens)
    for g in range(world):
        pos = 0
        for k, i in enumerate(bins[g]):
            n = int(all_counts[g][k])
            out[i] = all_flat[g][pos:pos + n]
            pos += n
    return out  # type: ignore[return-value]


def encode_recording_sharded(encode_fn: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor, c: int, l: int, r: int,
                             layers: int, mode: str = "exact", group=None, device=None) -> torch.Tensor:
    """Each rank encodes its chunk range (+ halos) of one long recording with `encode_fn(frames) -> (rows, d)` (all valid
    rows of that slice encoded as a stand-alone utterance) and keeps its own rows; one gather at the end returns the
    full (M, d) output on every rank (on `device`, see comm_device)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sh = split_recording(int(x.shape[0]), c, l, r, layers, world, mode)[rank]
    if sh.chunk_hi > sh.chunk_lo and sh.keep_hi > sh.keep_lo:
        out = encode_fn(x[sh.in_start:sh.in_end])[sh.keep_lo:sh.keep_hi]
    else:
        out = None
    dev = comm_device(group, device)
    # feature width: ranks without work learn it from the others (they still take part in the gather, with zero rows)
    d = torch.tensor([0 if out is None else out.shape[1]], dtype=torch.int64, device=dev)
    ds = [torch.zeros_like(d) for _ in range(world)]
    dist.all_gather(ds, d, group=group)
    width = int(max(int(v) for v in ds))
    out = torch.zeros((0, width), dtype=torch.float32, device=dev) if out is None else out.to(dev, torch.float32)
    parts = gather_variable(out.contiguous(), group)
    return torch.cat(parts, dim=0)
