"""Stand-alone launches of the hot GEMM shapes (FFN w_1 with SiLU, FFN w_2 with residual) for ncu --set full."""
import os, sys
from ctypes import c_void_p
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200 import lib as cflib
L = cflib.load()
rows, d, F = 180544, 512, 2048
A = torch.randn((rows, d), device="cuda").bfloat16()
W1 = (torch.randn((F, d), device="cuda") / d ** 0.5).bfloat16()
W2 = (torch.randn((d, F), device="cuda") / F ** 0.5).bfloat16()
b1, b2 = torch.zeros(F, device="cuda"), torch.zeros(d, device="cuda")
H = torch.empty((rows, F), device="cuda", dtype=torch.bfloat16)
X = torch.randn((rows, d), device="cuda")
st = c_void_p(torch.cuda.current_stream().cuda_stream)
def p(t): return c_void_p(t.data_ptr())
def ffn1(): cflib.check(L.cf_op_gemm(p(A), d, p(W1), d, rows, F, d, 0, 2, p(b1), None, 0, 1.0, None, 1, p(H), F, None, None, None, -1, st))
def ffn2(): cflib.check(L.cf_op_gemm(p(H), F, p(W2), F, rows, d, F, 2, 0, p(b2), p(X), d, 0.5, None, 1, p(X), d, None, None, None, -1, st))
for f in (ffn1, ffn2):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    print(f.__name__, "ms", e0.elapsed_time(e1) / 5)
