import os, sys, re, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chunkformer_b200.synth import synth_transducer_state_dict
from chunkformer_b200.transducer import TransducerGreedyB200
sd = synth_transducer_state_dict(1024, 256, 512, 2, 512, 512, 512, blank_bias=9.0, seed=13)
srch = TransducerGreedyB200(sd, device="cuda:0")
B, T = 19, 600
enc = torch.randn((B * T, 512), device="cuda", generator=torch.Generator("cuda").manual_seed(1))
st, ln = [b * T for b in range(B)], [T] * B
srch.search_flat(enc, st, ln)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    srch.search_flat(enc, st, ln); torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA: continue
    name = re.sub(r"\(.*$", "", re.sub(r"^void ", "", ev.name)).replace("cf::", "")
    tot[name][0] += 1; tot[name][1] += ev.device_time
print("iterations", srch.last_iterations)
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {name[:60]:60s} {n:6d} {us / n:8.2f} us avg {us / 1e3:9.2f} ms")
