"""Timings of BASELINE.json configs[2..4] (the default bench.py line is configs[1]); one JSON line per workload on rank 0.

    python tools/bench_configs.py long16h                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_configs.py long16h rnnt stream16            # N GPUs

long16h  configs[2]: ONE 16 h recording (T = 5 759 998 frames, 11 250 chunks), CTC-large 64/128/128, split into N contiguous
         chunk ranges with recomputed context halos (shard.split_recording, exact halos of 51 chunks per side), greedy CTC
         per rank, token ids gathered once.  Strong scaling: total work fixed.
rnnt     configs[3]: rnnt-large encoder (d512 H4 L12, no CTC head), N duration-balanced shards of N masked batches
         (LPT over 19 N utterances), encoder only.  Weak scaling.
stream16 configs[4]: CTC-large at the low-latency preset chunk 16 / left 64 / right 0 on the masked batch, one batch per GPU.

Timing as in bench.py: W warm-up steps, K timed steps between barrier + synchronize, CUDA events, max over ranks."""
import argparse, json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE, RNNT_LARGE
from chunkformer_b200.plan import Plan
from chunkformer_b200 import shard
from chunkformer_b200.synth import masked_batch_lengths, synth_fbank, synth_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("which", nargs="*", default=["long16h"])
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def timed(step):
    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def emit(name, audio_s, ms, scaling, extra):
    if rank == 0:
        print(json.dumps(dict({"workload": name, "n_gpus": world, "ms_per_step": ms, "audio_hours_per_s": audio_s / ms / 3.6,
                               "scaling": scaling, "steps": args.steps, "warmup": args.warmup}, **extra)), flush=True)


for which in args.which:
    if which == "long16h":
        c, l, r = 64, 128, 128
        T = 16 * 3600 * 100 - 2
        enc = ChunkFormerEncoderB200(CTC_LARGE, synth_state_dict(CTC_LARGE, 0), dev)
        sh = shard.split_recording(T, c, l, r, CTC_LARGE.layers, world, "exact")[rank]
        # the rank's input slice (with halos) of the synthetic recording; generated per rank from a slice-specific seed
        x = torch.randn((sh.in_end - sh.in_start, 80), device=dev, generator=torch.Generator(dev).manual_seed(100 + rank))
        lens = [int(x.shape[0])]

        def step():
            plan = Plan(c, l, r, lens)
            out, _ = enc.encode_plan(plan, x, out_dtype=torch.bfloat16)
            tok = enc.ctc_greedy(out[sh.keep_lo:sh.keep_hi])
            if world > 1:
                shard.gather_variable(tok)
            return tok
        ms = timed(step)
        hL, hR = shard.halo_chunks(c, l, r, CTC_LARGE.layers, "exact")
        emit("configs[2]: one 16 h recording, CTC-large 64/128/128, chunk-range shards with halos, encoder + greedy CTC",
             (T + 2) / 100.0, ms, "strong", {"chunks_total": shard.chunks_of(T, c), "halo_chunks": [hL, hR],
                                             "chunks_rank0": shard.chunks_of(lens[0], c)})
        del enc, x
    elif which == "rnnt":
        c, l, r = 64, 128, 128
        enc = ChunkFormerEncoderB200(RNNT_LARGE, synth_state_dict(RNNT_LARGE, 0), dev)
        base = masked_batch_lengths(1.0)
        all_lens = base * world
        mine = shard.partition_by_chunks(all_lens, c, world)[rank]
        lens = [all_lens[i] for i in mine]
        feats = torch.randn((sum(lens), 80), device=dev, generator=torch.Generator(dev).manual_seed(200 + rank))

        def step():
            plan = Plan(c, l, r, lens)
            out, _ = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
            return out
        ms = timed(step)
        emit("configs[3]: rnnt-large encoder (d512 H4 L12), masked batches of 14 400 s per GPU, LPT duration-balanced shards, "
             "encoder only", sum((t + 2) / 100.0 for t in all_lens), ms, "weak",
             {"utterances_rank0": len(lens), "chunks_rank0": sum(shard.chunks_of(t, c) for t in lens)})
        del enc, feats
    elif which == "stream16":
        c, l, r = 16, 64, 0
        enc = ChunkFormerEncoderB200(CTC_LARGE, synth_state_dict(CTC_LARGE, 0), dev)
        lens = masked_batch_lengths(1.0)
        feats = torch.randn((sum(lens), 80), device=dev, generator=torch.Generator(dev).manual_seed(300 + rank))

        def step():
            plan = Plan(c, l, r, lens)
            out, _ = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
            return enc.ctc_greedy(out)
        ms = timed(step)
        emit("configs[4]: CTC-large at chunk 16 / left 64 / right 0 on the 14 400 s masked batch per GPU, encoder + greedy CTC",
             world * sum((t + 2) / 100.0 for t in lens), ms, "weak", {})
        del enc, feats
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
