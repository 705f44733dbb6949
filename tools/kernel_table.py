"""Per-kernel time table of one benchmark step via the CUPTI activity trace (torch.profiler): no serialisation, no
replays; shares and per-launch averages for choosing what to optimise.  Bench values never come from this run."""
import argparse, os, sys, re, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE
from chunkformer_b200.plan import Plan
from chunkformer_b200.synth import masked_batch_lengths, synth_fbank, synth_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--out", default="")
ap.add_argument("--opt", action="append", default=[], help="library option name=value (cf_set_option), repeatable")
ap.add_argument("--host-input", action="store_true", help="step = forward_parallel_chunk(pinned host fbank) + ctc (the end-to-end path)")
a = ap.parse_args()
enc = ChunkFormerEncoderB200(CTC_LARGE, synth_state_dict(CTC_LARGE, 0), "cuda:0")
for kv in a.opt:
    k, v = kv.split("=")
    enc.set_option(k, int(v))
lens = masked_batch_lengths(a.scale)
feats = torch.cat([synth_fbank(t, seed=1 + k) for k, t in enumerate(lens)], 0).cuda()
xs_host = [synth_fbank(t, seed=1 + k).pin_memory() for k, t in enumerate(lens)] if a.host_input else None
lens_t = torch.tensor(lens, dtype=torch.int32)
def step():
    if a.host_input:
        out, *_ = enc.forward_parallel_chunk(xs_host, lens_t, 64, 128, 128, offset=torch.zeros(len(lens), dtype=torch.int32))
        return enc.ctc_greedy(out)
    plan = Plan(64, 128, 128, lens)
    out, out16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
    return enc.ctc_greedy(out16)
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
plain = e0.elapsed_time(e1)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
t_min, t_max = None, None
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = ev.name
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("cf::", "")
    tot[name][0] += 1
    tot[name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    s, e = ev.time_range.start, ev.time_range.end
    t_min = s if t_min is None else min(t_min, s)
    t_max = e if t_max is None else max(t_max, e)
total = sum(v[1] for v in tot.values())
# idle time of the device between kernels (all streams merged): where a step loses time without any kernel getting slower
spans = sorted((ev.time_range.start, ev.time_range.end) for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA)
busy_end, idle, gaps = spans[0][0], 0.0, []
for s0, e0 in spans:
    if s0 > busy_end:
        idle += s0 - busy_end
        gaps.append((s0 - busy_end, s0 - t_min))
    busy_end = max(busy_end, e0)
gaps.sort(reverse=True)
lines = [f"step without profiler: {plain:.2f} ms; kernel time sum {total / 1e3:.2f} ms; trace span {(t_max - t_min) / 1e3:.2f} ms; "
         f"device idle inside the span {idle / 1e3:.2f} ms; largest gaps (us @ offset ms): " + ", ".join(f"{g:.0f}@{o / 1e3:.1f}" for g, o in gaps[:6]),
         f"{'kernel':100s} {'n':>5s} {'ms':>8s} {'share':>6s} {'us/launch':>9s}"]
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{name[:100]:100s} {n:5d} {us / 1e3:8.3f} {100 * us / total:5.1f}% {us / n:9.1f}")
print("\n".join(lines))
if a.out:
    open(a.out, "w").write("\n".join(lines) + "\n")
