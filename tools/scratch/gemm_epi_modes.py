"""FFN w_1 / QKV GEMM under the epilogue ablation modes of CF_GEMM_DEBUG (read once per process: run once per mode)."""
import os, sys
from ctypes import c_void_p
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from chunkformer_b200 import lib as cflib
L = cflib.load()
st = c_void_p(torch.cuda.current_stream().cuda_stream)
M, N, K = 180544, 2048, 512
A = torch.randn((M, K), device="cuda").bfloat16()
W = (torch.randn((N, K), device="cuda") / K ** 0.5).bfloat16()
b = torch.zeros(N, device="cuda")
out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
for name, act in (("ffn1 silu", 2), ("qkv", 0)):
    def run():
        cflib.check(L.cf_op_gemm(c_void_p(A.data_ptr()), K, c_void_p(W.data_ptr()), K, M, N, K, 0, act, c_void_p(b.data_ptr()), None, 0, 1.0,
                                 None, 1, c_void_p(out.data_ptr()), N, None, None, None, -1, st))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(f"CF_GEMM_DEBUG={os.environ.get('CF_GEMM_DEBUG', '0')} {name}: {e0.elapsed_time(e1) / 10:.3f} ms")
