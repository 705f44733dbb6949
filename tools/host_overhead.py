"""Host-side cost of one step: how long the Python/C++ driver takes to enqueue everything vs the GPU time."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE
from chunkformer_b200.plan import Plan
from chunkformer_b200.synth import masked_batch_lengths, synth_fbank, synth_state_dict
enc = ChunkFormerEncoderB200(CTC_LARGE, synth_state_dict(CTC_LARGE, 0), "cuda:0")
lens = masked_batch_lengths(1.0)
feats = torch.cat([synth_fbank(t, seed=1 + k) for k, t in enumerate(lens)], 0).cuda()
for _ in range(3):
    plan = Plan(64, 128, 128, lens); out, o16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16); tok = enc.ctc_greedy(o16)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    plan = Plan(64, 128, 128, lens)
    t1 = time.perf_counter()
    out, o16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
    t2 = time.perf_counter()
    tok = enc.ctc_greedy(o16)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"plan {1e3*(t1-t0):.2f} ms | encode enqueue {1e3*(t2-t1):.2f} ms | ctc enqueue {1e3*(t3-t2):.2f} ms | drain {1e3*(t4-t3):.2f} ms | total {1e3*(t4-t0):.2f} ms")
