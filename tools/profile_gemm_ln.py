"""Stand-alone launches of the residual GEMM + fused LayerNorm kernel (gemm_ln_kernel) at the benchmark shapes, for timing and ncu:
    python tools/profile_gemm_ln.py            # K = 512 mode 1 (out-projection / pointwise_conv2), K = 2048 modes 1 and 2 (FFN w_2)"""
import os, sys
from ctypes import c_void_p
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200 import lib as cflib
L = cflib.load()
rows, d, F = 180544, 512, 2048
X = torch.randn((rows, d), device="cuda")
Y = torch.empty((rows, d), device="cuda", dtype=torch.bfloat16)
b = torch.zeros(d, device="cuda")
w1, b1 = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = c_void_p(torch.cuda.current_stream().cuda_stream)
def p(t): return c_void_p(t.data_ptr())
variants = [int(a) for a in sys.argv[1:]] or [0, 64, 32, 16]      # 0: by shape, 64: gemm_ln_split_kernel, 32: gemm_ln_quad_kernel, 16: gemm_ln_kernel
for K, mode, var in [(K, m, v) for (K, m) in ((512, 1), (2048, 1), (2048, 2)) for v in variants]:
    A = torch.randn((rows, K), device="cuda").bfloat16()
    W = (torch.randn((d, K), device="cuda") / K ** 0.5).bfloat16()
    def run():
        cflib.check(L.cf_op_gemm_ln(p(A), K, p(W), K, rows, d, K, p(b), p(X), d, 0.5, None, 1, mode + var, p(w1), p(b1), p(w1), p(b1),
                                    p(X), d, p(Y), d, None, 1, st))
    for _ in range(3): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        junk.zero_()                       # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    hbm = (rows * K * 2 + rows * d * (4 + 4 + 2)) / 1e9
    print(f"gemm_ln K={K} mode={mode} variant={var}: ms (L2 flushed) {['%.3f' % t for t in ts]}  algorithmic HBM {hbm:.2f} GB -> {hbm / min(ts):.2f} TB/s; "
          f"{2.0 * rows * d * K / min(ts) / 1e9:.0f} TFLOP/s")
    del A, W
