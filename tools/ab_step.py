"""A/B timing of one benchmark step in a single process: option toggles of the library (same box, same clocks)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200 import lib as cflib
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE
from chunkformer_b200.plan import Plan
from chunkformer_b200.synth import masked_batch_lengths, synth_fbank, synth_state_dict
L = cflib.load()
enc = ChunkFormerEncoderB200(CTC_LARGE, synth_state_dict(CTC_LARGE, 0), "cuda:0")
lens = masked_batch_lengths(1.0)
feats = torch.cat([synth_fbank(t, seed=1 + k) for k, t in enumerate(lens)], 0).cuda()
def step():
    plan = Plan(64, 128, 128, lens)
    out, out16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
    return out, enc.ctc_greedy(out16)
def timeit(n=5):
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): o = step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, o
for rnd in range(2):
    L.cf_set_fused_layernorm(0); t0, (o0, k0) = timeit()
    L.cf_set_fused_layernorm(1); t1, (o1, k1) = timeit()
    print(f"round {rnd}: separate LN {t0:.2f} ms | fused LN {t1:.2f} ms | max |diff| {float((o0.float()-o1.float()).abs().max()):.4f} tokens differ {int((k0!=k1).sum())}")
