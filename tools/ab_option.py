"""A/B of a per-handle library option on the benchmark step (same process, same box, alternating):
    python tools/ab_option.py fused_layernorm            # ms per step with the option 0 / 1, three rounds each"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE
from chunkformer_b200.plan import Plan
from chunkformer_b200.synth import masked_batch_lengths, synth_state_dict

name = sys.argv[1] if len(sys.argv) > 1 else "fused_layernorm"
values = [int(v) for v in sys.argv[2:]] or [0, 1]
geo = CTC_LARGE
enc = ChunkFormerEncoderB200(geo, synth_state_dict(geo, 0), "cuda:0")
lens = masked_batch_lengths()
feats = torch.randn((sum(lens), 80), device="cuda")


def step():
    plan = Plan(64, 128, 128, lens, None, geo.kernel)
    _, o16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
    return enc.ctc_greedy(o16)


for rnd in range(3):
    for v in values:
        enc.set_option(name, v)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            tok = step()
        e1.record()
        torch.cuda.synchronize()
        print(f"round {rnd} {name}={v}: {e0.elapsed_time(e1) / 5:.2f} ms/step  (token checksum {int(tok.sum())})", flush=True)
