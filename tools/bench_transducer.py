"""Transducer greedy search (cf_rnnt_greedy) on the rnnt-large head sizes (examples/asr/rnnt/conf/chunkformer-rnnt-large-vie.yaml:
vocab 1024, embed 256, LSTM 2 x 512, join 512) over the encoder rows of the benchmark batch (19 utterances, 14 400 s).  The CPU side of
the comparison (the oracle, i.e. the reference's per-frame loop restated) lives in tests/perf_transducer_cpu_oracle.py: only
tests/, smoke() and bench.py's CPU legs execute oracle/.
    python tools/bench_transducer.py [blank_bias ...]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.synth import masked_batch_lengths, synth_transducer_state_dict
from chunkformer_b200.transducer import TransducerGreedyB200
from chunkformer_b200 import lib as cflib

biases = [float(v) for v in sys.argv[1:]] or [7.2, 9.0]
c = 64
lens_in = masked_batch_lengths(1.0)
enc_lens = [max((t - 15) // 8 + 1, 0) for t in lens_in]
chunks = [max(1, -(-(t - 7) // (8 * c))) for t in lens_in]
starts, row = [], 0
for n in chunks:
    starts.append(row * c); row += n
rows = row * c
enc = torch.randn((rows, 512), device="cuda", generator=torch.Generator("cuda").manual_seed(1))
L = cflib.load()
for bb in biases:
    sd = synth_transducer_state_dict(1024, 256, 512, 2, 512, 512, 512, blank_bias=bb, seed=13)
    srch = TransducerGreedyB200(sd, device="cuda:0")
    ref_tokens = None
    for persistent in (1, 0):
        srch.set_option("persistent", persistent)
        srch.search_flat(enc[:4096], [0], [4096])          # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = srch.search_flat(enc, starts, enc_lens)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        syms = sum(t.numel() for t, _ in res)
        same = ""
        if ref_tokens is None:
            ref_tokens = [t for t, _ in res]
        else:
            same = "; same symbols as the persistent kernel: " + str(all(a.shape == b.shape and bool((a == b).all()) for a, (b, _) in zip(ref_tokens, res)))
        print(f"blank_bias {bb} persistent={persistent}: {len(enc_lens)} utterances, {sum(enc_lens)} frames (longest {max(enc_lens)}), "
              f"{syms} symbols (longest utterance {max(t.numel() for t, _ in res)}); {srch.last_iterations} iterations, {dt * 1e3:.1f} ms "
              f"= {dt * 1e6 / srch.last_iterations:.1f} us per iteration; {14400 / dt / 3600:.1f} audio-h/s{same}", flush=True)
