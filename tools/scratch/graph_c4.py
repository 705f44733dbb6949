import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200, StreamingGraph
from chunkformer_b200.geometry import CTC_LARGE
from chunkformer_b200.synth import synth_state_dict
enc = ChunkFormerEncoderB200(CTC_LARGE, synth_state_dict(CTC_LARGE, 0), "cuda:0")
for (c, l, B) in ((4, 40, 256), (16, 64, 256)):
    x = torch.randn((B, 8 * (c - 1) + 15, 80), device="cuda")
    att, cnn = torch.zeros((0, 0, 0, 0, 0)), torch.zeros((0, 0, 0, 0))
    for s in range(l // c + 3):
        o, _, att, cnn = enc.forward_chunk(x, att, cnn, c, l, 0, offset=s * c, donate_caches=True)
    def timed(fn, n=10):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
    def eager():
        global att, cnn
        o, _, att, cnn = enc.forward_chunk(x, att, cnn, c, l, 0, offset=4000, donate_caches=True)
        return enc.ctc_greedy(o)
    te = timed(eager)
    rows_e = int(enc._L.cf_encode_output_rows(enc._h))
    sg = StreamingGraph(enc, B, c, l)
    for s in range(l // c + 3): sg.step(x)
    rows_g = int(enc._L.cf_encode_output_rows(enc._h))
    tg = timed(lambda: sg.step(x))
    ts = timed(lambda: sg._steady_step())          # the same step eagerly on the pinned plan / private workspace
    print(f"c={c} l={l} B={B}: eager {te:.2f} ms (rows {rows_e}), graph replay {tg:.2f} ms (rows at capture {rows_g}), steady step eager {ts:.2f} ms")
