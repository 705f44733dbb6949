"""One encoder + CTC step of the benchmark workload for profiling (ncu launch list); --scale shrinks the batch."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE
from chunkformer_b200.plan import Plan
from chunkformer_b200.synth import masked_batch_lengths, synth_fbank, synth_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
a = ap.parse_args()
enc = ChunkFormerEncoderB200(CTC_LARGE, synth_state_dict(CTC_LARGE, 0), "cuda:0")
lens = masked_batch_lengths(a.scale)
feats = torch.cat([synth_fbank(t, seed=1 + k) for k, t in enumerate(lens)], 0).cuda()
def step():
    plan = Plan(64, 128, 128, lens)
    out, out16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
    return enc.ctc_greedy(out16)
for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(a.steps):
    step()
ev1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ms per step", ev0.elapsed_time(ev1) / a.steps)
