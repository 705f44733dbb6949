"""Per-role cycle accounting of gemm_ln_split_kernel (tools build only):
    python -m chunkformer_b200.build --ablation
    CF_LN_PROF=1 CHUNKFORMER_B200_LIB=chunkformer_b200/csrc/libchunkformer_b200_ablation.so python tools/profile_gemm_ln_roles.py"""
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
st = c_void_p(torch.cuda.current_stream().cuda_stream)
def p(t): return c_void_p(t.data_ptr())
var = int(sys.argv[1]) if len(sys.argv) > 1 else 0      # 0 pair kernel, 32 cluster of four
for K, mode in ((512, 1 + var), (2048, 1 + var), (2048, 2 + var)):
    A = torch.randn((rows, K), device="cuda").bfloat16()
    W = (torch.randn((d, K), device="cuda") / K ** 0.5).bfloat16()
    os.environ.pop("CF_LN_PROF", None)
    for _ in range(2):
        cflib.check(L.cf_op_gemm_ln(p(A), K, p(W), K, rows, d, K, p(b), p(X), d, 0.5, None, 1, mode, p(w1), p(b1), p(w1), p(b1),
                                    p(X), d, p(Y), d, None, 1, st))
    torch.cuda.synchronize()
    os.environ["CF_LN_PROF"] = "1"
    cflib.check(L.cf_op_gemm_ln(p(A), K, p(W), K, rows, d, K, p(b), p(X), d, 0.5, None, 1, mode, p(w1), p(b1), p(w1), p(b1),
                                p(X), d, p(Y), d, None, 1, st))
    torch.cuda.synchronize()
    del A, W
