"""Per-kernel time table (CUPTI) of one step for the other geometries / window shapes of BASELINE.json's configs:
    python tools/variants_table.py rnnt_large | ctc_small | stream16"""
import os, sys, re, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE, RNNT_LARGE, CTC_SMALL
from chunkformer_b200.plan import Plan
from chunkformer_b200.synth import masked_batch_lengths, synth_fbank, synth_state_dict
which = sys.argv[1] if len(sys.argv) > 1 else "rnnt_large"
geo, (c, l, r) = {"rnnt_large": (RNNT_LARGE, (64, 128, 128)), "ctc_small": (CTC_SMALL, (64, 128, 128)),
                  "stream16": (CTC_LARGE, (16, 64, 0)), "ctc_large": (CTC_LARGE, (64, 128, 128))}[which]
enc = ChunkFormerEncoderB200(geo, synth_state_dict(geo, 0), "cuda:0")
lens = masked_batch_lengths(1.0)
feats = torch.cat([synth_fbank(t, seed=1 + k) for k, t in enumerate(lens)], 0).cuda()
def step():
    plan = Plan(c, l, r, lens)
    out, out16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
    return out16
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"{which}: {geo} window {c}/{l}/{r}: {e0.elapsed_time(e1):.1f} ms per 14400 s batch (encoder only) = {14400 / e0.elapsed_time(e1) / 3.6:.1f} audio-h/s")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA: continue
    name = re.sub(r"\(.*$", "", re.sub(r"^void ", "", ev.name)).replace("cf::", "")
    tot[name][0] += 1; tot[name][1] += ev.device_time
total = sum(v[1] for v in tot.values())
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"  {name[:70]:70s} {n:5d} {us / 1e3:8.3f} ms {100 * us / total:5.1f}%")

if which == "rnnt_large":
    # full-attention padded batch (classification-style call: forward_encoder with chunk_size = -1), 64 x 30 s
    xs = torch.randn((64, 2998, 80), device="cuda")
    xl = torch.full((64,), 2998, dtype=torch.int32)
    for _ in range(2): enc.forward_encoder(xs, xl, -1, -1, -1)
    torch.cuda.synchronize()
    e0.record(); o, m = enc.forward_encoder(xs, xl, -1, -1, -1); e1.record(); torch.cuda.synchronize()
    print(f"full attention, padded batch 64 x 30 s (T' = {o.shape[1]}): {e0.elapsed_time(e1):.1f} ms = {64 * 30 / e0.elapsed_time(e1) / 3.6:.1f} audio-h/s")
