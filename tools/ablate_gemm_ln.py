"""Phase ablation of gemm_ln_split_kernel (tools build; results are wrong with a bit set, only the time is of interest):
    python -m chunkformer_b200.build --ablation
    CHUNKFORMER_B200_LIB=chunkformer_b200/csrc/libchunkformer_b200_ablation.so python tools/ablate_gemm_ln.py
bits: 1 no residual loads, 2 no x TMA store, 4 no shared-memory staging accesses, 8 P2 without global stores, 16 P2 without its
last pass, 32 P1 without statistics"""
import os, sys
from ctypes import c_void_p
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200 import lib as cflib
L = cflib.load()
rows, d = 180544, 512
X = torch.randn((rows, d), device="cuda")
Y = torch.empty((rows, d), device="cuda", dtype=torch.bfloat16)
b = torch.zeros(d, device="cuda")
w1, b1 = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = c_void_p(torch.cuda.current_stream().cuda_stream)
def p(t): return c_void_p(t.data_ptr())
var = int(os.environ.get('CF_LN_VARIANT', '0'))     # 0 pair kernel, 32 cluster of four
combos = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 16, 32, 3, 7, 24, 31, 63]
for K, mode in ((512, 1), (2048, 1)):
    A = torch.randn((rows + 256, K), device="cuda").bfloat16()      # + 256 rows: CF_LN_ABLOCKED reads A as [K / 64][mpad][64]
    W = (torch.randn((d, K), device="cuda") / K ** 0.5).bfloat16()
    for dbg in combos:
        os.environ["CF_LN_DEBUG"] = str(dbg)
        def run():
            cflib.check(L.cf_op_gemm_ln(p(A), K, p(W), K, rows, d, K, p(b), p(X), d, 0.5, None, 1, mode + var, p(w1), p(b1), p(w1), p(b1),
                                        p(X), d, p(Y), d, None, 1, st))
        for _ in range(2): run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            junk.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"K={K} mode={mode} debug={dbg:3d}: {min(ts):.3f} ms", flush=True)
    del A, W
