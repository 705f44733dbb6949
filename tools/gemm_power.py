"""Clock / power while one GEMM shape runs back to back for a few seconds (is the kernel power-capped?).
    CF_GEMM_DEBUG=0|1|2 python tools/gemm_power.py [variant]"""
import os, sys, subprocess, time
from ctypes import c_void_p
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200 import lib as cflib
L = cflib.load()
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
M, N, K = 180544, 2048, 512
A = torch.randn((M, K), device="cuda").bfloat16()
W = (torch.randn((N, K), device="cuda") / K ** 0.5).bfloat16()
b = torch.zeros(N, device="cuda")
out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
st = c_void_p(torch.cuda.current_stream().cuda_stream)
def p(t): return c_void_p(t.data_ptr())
def run(): cflib.check(L.cf_op_gemm(p(A), K, p(W), K, M, N, K, 0, 2, p(b), None, 0, 1.0, None, 1, p(out), N, None, None, None, variant, st))
for _ in range(5): run()
torch.cuda.synchronize()
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(50): run()
    n += 50
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
proc.terminate()
lines = [l.strip().split(",") for l in proc.stdout.read().strip().splitlines()]
clk = sorted(float(l[0]) for l in lines[3:]); pw = sorted(float(l[1]) for l in lines[3:])
print(f"debug={os.environ.get('CF_GEMM_DEBUG', '0')} variant={variant}: {ms:.3f} ms/launch ({2.0*M*N*K/ms/1e9:.0f} TF)  clock median {clk[len(clk)//2]:.0f} MHz  power median {pw[len(pw)//2]:.0f} W  ({len(clk)} samples)")
