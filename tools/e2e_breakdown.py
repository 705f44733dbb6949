"""Where the end-to-end step differs from the device-resident one (bench.py `e2e` vs `value`): per-step wall times of variants.
    python tools/e2e_breakdown.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE
from chunkformer_b200.plan import Plan
from chunkformer_b200.synth import masked_batch_lengths, synth_fbank, synth_state_dict

geo = CTC_LARGE
enc = ChunkFormerEncoderB200(geo, synth_state_dict(geo, 0), "cuda:0")
lens = masked_batch_lengths()
xs_host = [synth_fbank(t, seed=1 + k).pin_memory() for k, t in enumerate(lens)]
xs_dev = [x.cuda() for x in xs_host]
feats = torch.cat(xs_dev, 0)
lens_t = torch.tensor(lens, dtype=torch.int32)


def timed(name, fn, n=6):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:70s} " + " ".join(f"{t:6.2f}" for t in ts), flush=True)


def resident_bf16():
    plan = Plan(64, 128, 128, lens, None, geo.kernel)
    _, o16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
    return enc.ctc_greedy(o16)


def resident_f32_twin():
    plan = Plan(64, 128, 128, lens, None, geo.kernel)
    o, o16 = enc.encode_plan(plan, feats, out_dtype=torch.float32, want_bf16=True)
    return enc.ctc_greedy(o16)


def fpc(xs):
    out, *_ = enc.forward_parallel_chunk(xs, lens_t, 64, 128, 128, offset=torch.zeros(len(lens), dtype=torch.int32))
    return out


timed("resident: encode_plan(bf16) + ctc", resident_bf16)
timed("resident: encode_plan(fp32 + bf16 twin) + ctc", resident_f32_twin)
timed("forward_parallel_chunk(device list) only", lambda: fpc(xs_dev))
timed("forward_parallel_chunk(device list) + ctc", lambda: enc.ctc_greedy(fpc(xs_dev)))
timed("forward_parallel_chunk(host pinned list) + ctc", lambda: enc.ctc_greedy(fpc(xs_host)))
timed("forward_parallel_chunk(host pinned list) + ctc + tokens.cpu()", lambda: enc.ctc_greedy(fpc(xs_host)).cpu())
timed("upload only (host pinned -> device)", lambda: enc.upload_async(xs_host, lens))
