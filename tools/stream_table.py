"""Per-kernel time table (CUPTI) of one multi-stream streaming step: python tools/stream_table.py [B] [c] [l] [small]"""
import os, sys, re, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE, CTC_SMALL
from chunkformer_b200.synth import synth_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
c = int(sys.argv[2]) if len(sys.argv) > 2 else 16
l = int(sys.argv[3]) if len(sys.argv) > 3 else 64
geo = CTC_SMALL if "small" in sys.argv else CTC_LARGE
enc = ChunkFormerEncoderB200(geo, synth_state_dict(geo, 0), "cuda:0")
x = torch.randn((B, 8 * (c - 1) + 15, 80), device="cuda")
att, cnn = torch.zeros((0, 0, 0, 0, 0)), torch.zeros((0, 0, 0, 0))
for s in range(l // c + 3):
    o, _, att, cnn = enc.forward_chunk(x, att, cnn, c, l, 0, offset=s * c, donate_caches=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
o, _, att, cnn = enc.forward_chunk(x, att, cnn, c, l, 0, offset=1000, donate_caches=True); tok = enc.ctc_greedy(o)
e1.record(); torch.cuda.synchronize()
print(f"B={B} chunk {c} left {l}: {e0.elapsed_time(e1):.2f} ms per step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    o, _, att, cnn = enc.forward_chunk(x, att, cnn, c, l, 0, offset=1016, donate_caches=True); tok = enc.ctc_greedy(o)
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA: continue
    name = re.sub(r"\(.*$", "", re.sub(r"^void ", "", ev.name)).replace("cf::", "")
    tot[name][0] += 1; tot[name][1] += ev.device_time
total = sum(v[1] for v in tot.values())
print(f"kernel time sum {total / 1e3:.2f} ms, {sum(v[0] for v in tot.values())} launches")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {name[:72]:72s} {n:5d} {us / 1e3:8.3f} ms {100 * us / total:5.1f}%  {us / n:7.1f} us/launch")
