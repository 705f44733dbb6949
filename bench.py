#!/usr/bin/env python
"""Benchmark of the hot path: masked-chunk Conformer encoder forward + greedy CTC, CTC-large, chunk 64 / left 128 /
right 128 (BASELINE.json).  One "step" = one pass of the path over one masked batch of synthetic fbank.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun (one rank per GPU); every rank encodes its own batch (weak scaling, no data-path
collective; token ids are gathered once per step).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from chunkformer_b200.geometry import CTC_LARGE  # noqa: E402
from chunkformer_b200.synth import MASKED_BATCH_SECONDS, masked_batch_lengths, synth_fbank, synth_state_dict  # noqa: E402

METRIC = "encoder audio-hours/sec (RTFx), CTC-large chunk64/L128/R128"
UNIT = "audio-hours/s"
C, L, R = 64, 128, 128


def audio_seconds(frames):
    return (frames + 2) / 100.0


def layer_flops_per_chunk(geo, c, l, r):
    d, F = geo.d_model, geo.ffn
    W, Rr = l + c + r, 2 * c + l + r - 1
    return 2 * (4 * c * d * F + 4 * c * d * d + c * d * (2 * W + Rr) + 3 * c * d * d + 15 * c * d)


def path_flops_per_chunk(geo, c, l, r):
    """sub + L * layer + ctc, 2*MAC."""
    d, V = geo.d_model, geo.vocab
    t1, t2 = 4 * c + 3, 2 * c + 1
    f1, f2, f3 = 39, 19, 9
    sub = 2 * (9 * t1 * f1 * d + 9 * t2 * f2 * d + t2 * f2 * d * d + 9 * c * f3 * d + c * f3 * d * d + c * f3 * d * d)
    return sub + geo.layers * layer_flops_per_chunk(geo, c, l, r) + 2 * c * d * V


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/r02_ncu_gemm_ffn1_full.csv, else the round-1 capture of the same, unchanged kernel; first profiled launch);
    None if no summary is there."""
    try:
        tot, seen = 0.0, 0
        path = os.path.join(ROOT, "profiles", "r02_ncu_gemm_ffn1_full.csv")
        if not os.path.exists(path):
            path = os.path.join(ROOT, "profiles", "r01_ncu_gemm_ffn1_full.csv")
        for line in open(path):
            if line.startswith("#"):
                seen += 1
                if seen > 1:
                    break
                continue
            parts = line.strip().split(",", 2)
            if len(parts) != 3:
                continue
            name, unit, val = parts
            if name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(val) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return tot or None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  The process is started before the warm-up steps (nvidia-smi
    needs a few hundred ms to come up) and every line is stamped on arrival; `stop(t0, t1)` keeps the samples that fall inside
    the timed region [t0, t1] (host clock around the barrier + synchronize pair)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "50", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True, bufsize=1)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)                     # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [ln for (t, ln) in self.lines if t0 is None or (t0 <= t <= t1 + 0.06)]
        window = "timed region"
        if not inside and self.lines:        # region shorter than one sampling period: take the samples closest to it
            mid = 0.5 * (t0 + t1)
            inside = [ln for (_, ln) in sorted(self.lines, key=lambda tl: abs(tl[0] - mid))[:2]]
            window = "nearest samples (timed region shorter than the 50 ms sampling period)"
        sm, mx, reasons, power = [], [], set(), []
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


SAMPLE_SECONDS = [1, 30, 60]      # bounded CPU sample: the {1 s, 30 s, 60 s} utterances of the masked batch as one masked batch


def cpu_reference_baseline(threads, steps=2, warmup=1):
    """The reference's own CPU implementation of the path on the host cores, fp32, on a bounded sample of the workload.

    kind "reference": the UNMODIFIED reference installed under baseline/_ref (baseline/reference_arm.py drives its
    forward_parallel_chunk + ctc.log_softmax().argmax()); kind "port": the oracle restatement, only if baseline/_ref is absent."""
    from baseline import reference_arm as RA
    torch.set_num_threads(threads)
    sd = synth_state_dict(CTC_LARGE, 0)
    lens = [int(round(s * 100)) - 2 for s in SAMPLE_SECONDS]
    xs = [synth_fbank(t, seed=1 + k) for k, t in enumerate(lens)]
    audio = sum(audio_seconds(t) for t in lens)
    if RA.available():
        model = RA.build_reference(CTC_LARGE, sd)
        times = RA.time_reference_cpu(model, xs, lens, C, L, R, steps, warmup)
        kind, what = "reference", "unmodified reference (baseline/_ref) forward_parallel_chunk + ctc.log_softmax().argmax(), fp32 torch CPU"
    else:
        from oracle import chunkformer_oracle as O
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out, enc_lens, n_chunks, _, _, _ = O.forward_parallel_chunk(sd, CTC_LARGE.heads, xs, lens, C, L, R)
            O.ctc_greedy(sd, out)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kind, what = "port", "oracle port of the reference path (baseline/_ref not installed), fp32 torch CPU"
    best = min(times)
    return {"value": audio / best / 3600.0, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"masked batch of the {SAMPLE_SECONDS} s utterances ({audio:.0f} s audio), {what}, best of {steps}",
            "seconds_per_sample": best}, times, audio


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    base, times, audio = cpu_reference_baseline(threads, steps=args.steps, warmup=max(args.warmup, 1))
    total = sum(times)
    value = audio * len(times) / total / 3600.0
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": max(args.warmup, 1),
            "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "bounded sample of the CTC-large 64/128/128 masked batch: the 1 s, 30 s and 60 s utterances "
                                   "(91 s audio) per step, " + ("the unmodified reference on CPU" if base["kind"] == "reference"
                                                                 else "CPU oracle port of the reference path"),
                       "chunk": C, "left": L, "right": R},
            "cpu_baseline": dict(base, value=value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


PARITY_MAX_ABS, PARITY_REL_RMS, PARITY_MARGIN_TOL = 0.05, 0.01, 0.08     # the bar of tests/test_gpu_encoder.py


def strong_scaling_legs(enc, geo, dev, rank, world, barrier, steps=2):
    """BASELINE.json north_star splits under the driver's own command, device-resident inputs, CUDA events,
    max over ranks:
      * batch: ONE global batch of world x 14 400 s (world copies of the 19-utterance list) split over the ranks by
        chunk count (LPT, shard.partition_by_chunks); every rank encodes + greedy-decodes its share, token ids gathered once;
      * recording: ONE 16 h recording (BASELINE configs[2]) split into contiguous chunk ranges with recomputed context halos
        (shard.split_recording), "exact" halos (51 + 51 chunks) and the reference's own look-ahead ("reference": 34 right)."""
    import torch.distributed as dist
    from chunkformer_b200.plan import Plan
    from chunkformer_b200.shard import chunks_of, gather_variable, halo_chunks, partition_by_chunks, split_recording

    def timed(fn):
        fn()                                             # warm-up (workspace growth, position tables)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def loads(mine):
        t = torch.tensor([float(mine)], device=dev)
        if world > 1:
            all_t = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(all_t, t)
            vals = [float(v.item()) for v in all_t]
        else:
            vals = [float(mine)]
        return {"max": max(vals), "min": min(vals)}

    res = {}
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    # ---- (a) one global masked batch, LPT by chunk count
    lens_all = masked_batch_lengths() * world
    mine = partition_by_chunks(lens_all, C, world)[rank]
    my_lens = [lens_all[i] for i in mine]
    feats = torch.randn((sum(my_lens), geo.feat_dim), device=dev, generator=gen)

    def step_batch():
        plan = Plan(C, L, R, my_lens, None, geo.kernel)
        _, o16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
        tok = enc.ctc_greedy(o16)
        if world > 1:
            gather_variable(tok)
    ms = timed(step_batch)
    audio = sum(audio_seconds(t) for t in lens_all)
    res["batch_lpt"] = {"what": "one global batch of %d utterances (%d s) split over %d GPU(s) by chunk count (LPT)" % (len(lens_all), round(audio), world),
                        "ms_per_step": ms, "value": audio / (ms / 1e3) / 3600.0, "unit": UNIT,
                        "rank_load_chunks": loads(sum(chunks_of(t, C) for t in my_lens))}
    del feats
    # ---- (b) one 16 h recording split by chunk range with halos
    T = 16 * 3600 * 100 - 2
    for mode in ("exact", "reference"):
        sh = split_recording(T, C, L, R, geo.layers, world, mode)[rank]
        n_in = sh.in_end - sh.in_start
        x = torch.randn((n_in, geo.feat_dim), device=dev, generator=gen)

        def step_rec():
            plan = Plan(C, L, R, [n_in], None, geo.kernel)
            _, o16 = enc.encode_plan(plan, x, out_dtype=torch.bfloat16)
            tok = enc.ctc_greedy(o16)[sh.keep_lo:sh.keep_hi]
            if world > 1:
                gather_variable(tok)
        ms = timed(step_rec)
        res["recording_16h_" + mode] = {"what": "one 16 h recording (%d chunks) as %d contiguous chunk range(s), halos %s chunks (%s)"
                                                % (chunks_of(T, C), world, halo_chunks(C, L, R, geo.layers, mode), mode),
                                        "ms_per_step": ms, "value": audio_seconds(T) / (ms / 1e3) / 3600.0, "unit": UNIT,
                                        "rank_load_chunks": loads(chunks_of(n_in, C))}
        del x
    enc._ws = None
    torch.cuda.empty_cache()
    return res


def reference_gpu_legs(dev, xs_host, lens, audio):
    """BASELINE.md section 3, second reference point: the UNMODIFIED reference's own eager path on this GPU (the existing
    library path: cuBLAS / cuDNN / ATen kernels), the same masked batch, fp32 and bf16 autocast.  If the whole batch does not fit
    (the reference materialises unfolded K/V windows and a (n, d, 259, 39) fp32 conv output: about 110 GiB), it is run as the
    four masked sub-batches of tests/golden/make_golden_bench.py and the times are added."""
    from baseline import reference_arm as RA
    if not RA.available():
        return {"unavailable": "baseline/_ref not installed"}
    res = {}
    try:
        model = RA.build_reference(CTC_LARGE, synth_state_dict(CTC_LARGE, 0)).to(dev)
        xs_dev = [x.to(dev) for x in xs_host]
        groups_full = [list(range(len(lens)))]
        groups_split = [[5], [11], [4, 10], [0, 1, 2, 3, 6, 7, 8, 9, 12, 13, 14, 15, 16, 17, 18]]
        for name, dt in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
            for groups in (groups_full, groups_split):
                try:
                    ms, peak = 0.0, 0
                    for g in groups:
                        m, pk, _ = RA.time_reference_gpu(model, [xs_dev[i] for i in g], [lens[i] for i in g], C, L, R, dev, dt)
                        ms, peak = ms + m, max(peak, pk)
                    res[name] = {"ms_per_step": ms, "value": audio / (ms / 1e3) / 3600.0, "unit": UNIT, "peak_memory_bytes": peak,
                                 "sub_batches": len(groups)}
                    break
                except torch.cuda.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    res[name] = {"unavailable": "out of memory"}
        res["what"] = ("unmodified reference (baseline/_ref) on this GPU, torch eager: forward_parallel_chunk + "
                       "ctc.log_softmax().argmax() on the same 14 400 s masked batch, 1 warm-up + 2 timed steps, CUDA events")
        del model, xs_dev
        torch.cuda.empty_cache()
    except Exception as e:                                  # a reported baseline, never a reason to lose the bench line
        res["unavailable"] = f"{type(e).__name__}: {e}"
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="scale every utterance duration (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling legs (LPT batch split, 16 h recording)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the unmodified reference's eager path on the GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from chunkformer_b200 import lib as cflib
    from chunkformer_b200.encoder import ChunkFormerEncoderB200
    from chunkformer_b200.plan import Plan

    geo = CTC_LARGE
    enc = ChunkFormerEncoderB200(geo, synth_state_dict(geo, 0), dev)
    lens = masked_batch_lengths(args.scale)
    xs_host = [synth_fbank(t, seed=1 + 100 * rank + k).pin_memory() for k, t in enumerate(lens)]
    lens_t = torch.tensor(lens, dtype=torch.int32)
    audio = sum(audio_seconds(t) for t in lens)
    feats_dev = torch.cat(xs_host, 0).to(dev)
    Llib = cflib.load()

    def step_resident():
        plan = Plan(C, L, R, lens, None, geo.kernel)                       # host packer is part of the path
        out, out16 = enc.encode_plan(plan, feats_dev, out_dtype=torch.bfloat16)
        tokens = enc.ctc_greedy(out16)
        return plan, tokens, out16

    def step_e2e():
        """One step through the public API with HOST inputs: features cross PCIe (pinned host -> device), encoder, greedy CTC,
        token ids back to the host."""
        out, enc_lens, n_chunks, _, _, _ = enc.forward_parallel_chunk(xs_host, lens_t, C, L, R,
                                                                      offset=torch.zeros(len(lens), dtype=torch.int32))
        tokens = enc.ctc_greedy(out)
        return tokens.to("cpu", non_blocking=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        plan, tokens, out16 = step_resident()
    barrier()
    t_region0 = time.perf_counter()
    launches0 = Llib.cf_launch_count()
    if rank == 0:
        # CUDA events around every FFN w_1 / w_2 launch of the timed steps
        cflib.check(Llib.cf_kernel_timing_begin(enc._h, (1 << cflib.FAMILY_FFN_W1) | (1 << cflib.FAMILY_FFN_W2)), enc._h, "timing")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        plan, tokens, out16 = step_resident()
        if world > 1:
            gathered = [torch.empty_like(tokens) for _ in range(world)]
            dist.all_gather(gathered, tokens)
    ev1.record()
    barrier()
    t_region1 = time.perf_counter()
    launches = Llib.cf_launch_count() - launches0
    dom_ms, dom_n = 0.0, 0
    if rank == 0:
        import ctypes
        fam = {}
        for name, f in (("w1", cflib.FAMILY_FFN_W1), ("w2", cflib.FAMILY_FFN_W2)):
            tot_ms, n_l = ctypes.c_double(0.0), ctypes.c_int(0)
            cflib.check(Llib.cf_kernel_timing_end(enc._h, f, ctypes.byref(tot_ms), ctypes.byref(n_l)), enc._h, "cf_kernel_timing_end")
            fam[name] = (float(tot_ms.value), int(n_l.value))
        dom_ms, dom_n = fam["w1"]
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    value = world * audio * args.steps / (ms_total / 1000.0) / 3600.0

    # ---- end to end through the public API with host buffers: every step uploads its features (pinned host -> device, inside
    # forward_parallel_chunk: the copy runs on a side stream and the front-end starts on the first rows while the rest is still
    # crossing PCIe), encodes, decodes greedily and reads the token ids back.  Wall clock around K steps.
    # (Also tried: starting the upload of step k + 1 before step k is encoded (upload_async); the copy then runs under the
    # attention / FFN kernels instead of under the front-end and the step got SLOWER: 73.1 vs 71.6 ms on one GPU, 121 vs 71 ms
    # with two ranks sharing the host; profiles/README.md.)
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tok_host = step_e2e()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * audio * args.steps / float(e2e_s.item()) / 3600.0
    h2d = int(sum(x.numel() * 4 for x in xs_host))
    d2h = int(tok_host.numel() * tok_host.element_size())
    # the link the end-to-end number depends on: one pinned-host -> device copy of the largest utterance, timed alone
    big = max(xs_host, key=lambda t: t.numel())
    dst = torch.empty_like(big, device=dev)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(); dst.copy_(big, non_blocking=True); c1.record()
    torch.cuda.synchronize()
    h2d_gbs = big.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del dst

    # ---- what was timed is what the reference computes: the last timed step of rank 0 (whose inputs are the golden's: weights
    # seed 0, fbank seeds 1 + k) against the golden the UNMODIFIED reference produced for this batch (tests/golden/bench_batch.npz)
    parity = None
    if rank == 0 and args.scale == 1.0:
        gpath = os.path.join(ROOT, "tests", "golden", "bench_batch.npz")
        if os.path.exists(gpath):
            import numpy as np
            from chunkformer_b200.synth import check_bench_batch
            rep = check_bench_batch(np.load(gpath), out16.view(plan.n, C, geo.d_model), tokens.view(plan.n, C), plan.n_chunks,
                                    [int(v) for v in plan.enc_lens], PARITY_MARGIN_TOL)
            rep["bar"] = {"max_abs": PARITY_MAX_ABS, "rel_rms": PARITY_REL_RMS, "token_mismatches_above_tol": 0}
            rep["ok"] = bool(rep["max_abs"] <= PARITY_MAX_ABS and rep["rel_rms"] <= PARITY_REL_RMS and
                             rep["token_mismatches_above_tol"] == 0)
            rep["against"] = "tests/golden/bench_batch.npz (unmodified reference, fp32 CPU); bf16 output of the last timed step"
            parity = rep
            if not rep["ok"]:
                raise SystemExit("bench.py: the timed output does not match the reference golden: " + json.dumps(rep))

    # ---- the north-star splits, measured in the same run (strong scaling: fixed total work split over the ranks)
    strong = None
    if not args.no_strong:
        strong = strong_scaling_legs(enc, geo, dev, rank, world, barrier)

    # ---- roofline of the dominant kernel family: the FFN up-projection GEMM (tcgen05, bf16 -> fp32 accumulate, SiLU epilogue;
    # 34 launches per step).  `achieved` = algorithmic flops per launch / average launch duration measured with CUDA events on
    # the launching stream inside the timed steps above; the denominator is therefore the SUSTAINED measured bf16 peak (the
    # kernel runs inside a long, power-capped step).  The same kernel timed alone (best single launch) is given against the
    # burst peak for reference.
    rows = plan.rows
    roofline = None
    if rank == 0:
        from ctypes import c_void_p
        d, F = geo.d_model, geo.ffn
        flops = 2.0 * rows * d * F
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
        burst = float(peaks.get("bf16_tflops", 1590.0))
        k_ms = dom_ms / max(dom_n, 1)
        achieved = flops / (k_ms * 1e-3) / 1e12 if dom_n else None
        # the same kernel alone: best of 10 individually timed launches (burst methodology of MEASURED_PEAKS.json)
        A = torch.randn((rows, d), device=dev).bfloat16()
        Wt = (torch.randn((F, d), device=dev) / d ** 0.5).bfloat16()
        bias = torch.zeros(F, device=dev)
        Hbuf = torch.empty((rows, F), device=dev, dtype=torch.bfloat16)
        st = c_void_p(torch.cuda.current_stream().cuda_stream)

        def ffn1():
            rc = Llib.cf_op_gemm(c_void_p(A.data_ptr()), d, c_void_p(Wt.data_ptr()), d, rows, F, d, cflib.EPI_BF16,
                                 cflib.ACT_SILU, c_void_p(bias.data_ptr()), None, 0, 1.0, None, 1,
                                 c_void_p(Hbuf.data_ptr()), F, None, None, None, -1, st)
            cflib.check(rc, None, "cf_op_gemm")
        for _ in range(3):
            ffn1()
        torch.cuda.synchronize()
        alone = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ffn1(); e1.record()
            torch.cuda.synchronize()
            alone.append(e0.elapsed_time(e1))
            time.sleep(0.02)
        best_alone = min(alone)
        roofline = {"bound": "tensor", "kernel": "gemm2_tcgen05_kernel<EPI_BF16, ACT_SILU> (FFN w_1 + SiLU on CTA pairs, cta_group::2, M=%d N=%d K=%d)" % (rows, F, d),
                    "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": (achieved / sustained) if achieved else None,
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the timed steps, %d launches)" % dom_n)
                    if peaks else "fallback 1.4 PFLOP/s sustained",
                    "traffic": ncu_traffic_bytes(), "ms_per_launch": k_ms, "launches_timed": dom_n,
                    "flops_per_launch": flops,
                    "alone": {"ms_best_of_10": best_alone, "achieved": flops / (best_alone * 1e-3) / 1e12, "peak": burst,
                              "frac": flops / (best_alone * 1e-3) / 1e12 / burst, "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)"},
                    "also": {"kernel": "gemm_ln_split_kernel (FFN w_2 + 0.5 * residual + the LayerNorm(s) behind it, CTA pair, M=%d N=%d K=%d; "
                                       "flops of the GEMM only)" % (rows, d, F),
                             "ms_per_launch": fam["w2"][0] / max(fam["w2"][1], 1), "launches_timed": fam["w2"][1],
                             "achieved": flops / (fam["w2"][0] / max(fam["w2"][1], 1) * 1e-3) / 1e12 if fam["w2"][1] else None,
                             "frac": flops / (fam["w2"][0] / max(fam["w2"][1], 1) * 1e-3) / 1e12 / sustained if fam["w2"][1] else None},
                    "path_tflops": path_flops_per_chunk(geo, C, L, R) * plan.n * args.steps * world / (ms_total * 1e-3) / 1e12,
                    "path_frac_of_sustained": None}
        roofline["path_frac_of_sustained"] = roofline["path_tflops"] / (sustained * world)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _, _ = cpu_reference_baseline(os.cpu_count() or 1)

    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_ref_gpu:
        del out16, tokens
        enc._ws = None
        torch.cuda.empty_cache()
        ref_gpu = reference_gpu_legs(dev, xs_host, lens, audio)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "masked batch of 19 utterances %s s (14400 s audio, %d chunks, %d encoder rows) per GPU, "
                                       "CTC-large d512 H8 F2048 L17 V5000, chunk 64 / left 128 / right 128, encoder + greedy CTC"
                                       % (MASKED_BATCH_SECONDS, plan.n, plan.rows),
                           "chunk": C, "left": L, "right": R, "global_batch_audio_s": world * audio,
                           "l2": "working set per step (>4 GB) exceeds the 126 MB L2; no flush needed",
                           "parallelism": f"dp{world} (one batch per GPU, token ids gathered once per step)"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1000.0 * float(e2e_s.item()) / args.steps,
                        "h2d_link_gbs": h2d_gbs,
                        "api": "ChunkFormerEncoderB200.forward_parallel_chunk(pinned host fbank) + ctc_greedy + tokens.cpu(), one step "
                               "after the other; every step's H2D and D2H inside the timed region"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_base,
                "parity_checked": bool(parity and parity["ok"]), "parity": parity, "strong": strong, "reference_gpu": ref_gpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
