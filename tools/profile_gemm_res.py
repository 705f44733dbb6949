"""Stand-alone launches of the K = 512, N = 512 residual GEMM (attention out-projection / pointwise_conv2 shape) for ncu."""
import os, sys
from ctypes import c_void_p
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200 import lib as cflib
L = cflib.load()
rows, d = 180544, 512
A = torch.randn((rows, d), device="cuda").bfloat16()
W = (torch.randn((d, d), device="cuda") / d ** 0.5).bfloat16()
b = torch.zeros(d, device="cuda")
X = torch.randn((rows, d), device="cuda")
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = c_void_p(torch.cuda.current_stream().cuda_stream)
def p(t): return c_void_p(t.data_ptr())
def run(): cflib.check(L.cf_op_gemm(p(A), d, p(W), d, rows, d, d, 2, 0, p(b), p(X), d, 1.0, None, 1, p(X), d, None, None, None, -1, st))
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    junk.zero_()                       # flush L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("out/pw2 residual gemm ms (L2 flushed):", ["%.3f" % t for t in ts])
