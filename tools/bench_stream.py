"""Frame-synchronous streaming step (forward_chunk, all streams in one encoder pass) at the shipped presets: ms per step for B
concurrent streams and the number of real-time streams one GPU sustains (a step consumes chunk x 80 ms of audio per stream).
    python tools/bench_stream.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200, StreamingGraph
from chunkformer_b200.geometry import CTC_LARGE, CTC_SMALL
from chunkformer_b200.synth import synth_state_dict

for name, geo in (("ctc-small (d256 H4 L12)", CTC_SMALL), ("ctc-large (d512 H8 L17)", CTC_LARGE)):
    enc = ChunkFormerEncoderB200(geo, synth_state_dict(geo, 0), "cuda:0")
    for (c, l) in ((8, 60), (4, 40), (16, 64)):
        for B in (1, 32, 256, 1024):
            x = torch.randn((B, 8 * (c - 1) + 15, 80), device="cuda")
            att, cnn = torch.zeros((0, 0, 0, 0, 0)), torch.zeros((0, 0, 0, 0))
            for s in range(3):
                o, _, att, cnn = enc.forward_chunk(x, att, cnn, c, l, 0, offset=s * c)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            n = 10
            for s in range(n):
                o, _, att, cnn = enc.forward_chunk(x, att, cnn, c, l, 0, offset=(3 + s) * c, donate_caches=True)
                tok = enc.ctc_greedy(o).cpu()
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
            # the same steady-state step as one captured CUDA graph (StreamingGraph)
            sg = StreamingGraph(enc, B, c, l)
            for s in range(l // c + 3):
                sg.step(x)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for s in range(n):
                o, tok = sg.step(x)
                tok = tok.cpu()
            torch.cuda.synchronize(); dg = (time.perf_counter() - t0) / n
            print(f"{name} chunk {c} left {l}: B={B:3d} {dt * 1e3:7.2f} ms per step ({dt * 1e3 / B:5.2f} per stream) "
                  f"-> {int(c * 0.08 / (dt / B))} real-time streams per GPU;  CUDA graph {dg * 1e3:7.2f} ms per step "
                  f"-> {int(c * 0.08 / (dg / B))} streams", flush=True)
            del sg
