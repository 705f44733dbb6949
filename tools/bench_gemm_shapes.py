"""1-CTA vs 2-CTA GEMM on the shapes of the path (M = 180544 rows; front-end pw1 M = 627456)."""
import os, sys
from ctypes import c_void_p
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200 import lib as cflib
L = cflib.load()
st = c_void_p(torch.cuda.current_stream().cuda_stream)
def p(t): return None if t is None else c_void_p(t.data_ptr())
shapes = [("ffn1 silu bf16", 180544, 2048, 512, 0, 2), ("qkv bf16", 180544, 2048, 512, 0, 0), ("ffn2 f32+res", 180544, 512, 2048, 2, 0),
          ("out/pw2 f32+res", 180544, 512, 512, 2, 0), ("pw1 glu", 180544, 1024, 512, 1, 0), ("fe pw1 relu", 627456, 512, 512, 0, 1),
          ("fe out f32", 16384, 512, 4608, 2, 0), ("ctc argmax", 180544, 5000, 512, 4, 0)]
for name, M, N, K, epi, act in shapes:
    A = torch.randn((M, K), device="cuda").bfloat16()
    W = (torch.randn((N, K), device="cuda") / K ** 0.5).bfloat16()
    b = torch.zeros(N, device="cuda")
    ocols = N // 2 if epi == 1 else N
    out = torch.empty((M, ocols), device="cuda", dtype=torch.float32 if epi == 2 else torch.bfloat16) if epi != 4 else None
    res = out if epi == 2 and "res" in name else None
    nt = 2 * ((N + 255) // 256)
    parts = (torch.empty((M, nt), device="cuda"), torch.empty((M, nt), device="cuda"), torch.empty((M, nt), device="cuda", dtype=torch.int32)) if epi == 4 else (None, None, None)
    line = f"{name:18s} M={M} N={N} K={K}: "
    for v in (0, 1):
        def run():
            cflib.check(L.cf_op_gemm(p(A), K, p(W), K, M, N, K, epi, act, p(b), p(res), ocols if res is not None else 0, 0.5, None, 1,
                                     p(out), ocols, p(parts[0]), p(parts[1]), p(parts[2]), v, st))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        line += f" {'2cta' if v else '1cta'} {ms:.3f} ms ({2.0 * M * N * K / ms / 1e9:.0f} TF)"
    print(line)
    del A, W, out, res
