"""Stand-alone launches of the non-GEMM kernels at benchmark size for ncu: attention (tcgen05), dwconv, layernorm."""
import os, sys
from ctypes import c_void_p
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200 import lib as cflib
L = cflib.load()
c, l, r, d, H, n = 64, 128, 128, 512, 8, 2821
rows = l + n * c + r + 2 * c + 128
def p(t): return c_void_p(t.data_ptr())
st = c_void_p(torch.cuda.current_stream().cuda_stream)
qkv = torch.zeros((rows, 4 * d), device="cuda", dtype=torch.bfloat16)
qkv[: l + n * c] = torch.randn((l + n * c, 4 * d), device="cuda").bfloat16()
R = 2 * c + l + r - 1
pos = torch.zeros(((R + 127) // 128 * 128, d), device="cuda", dtype=torch.bfloat16)
pos[:R] = torch.randn((R, d), device="cuda").bfloat16()
rng = torch.zeros((n + 2, 2), dtype=torch.int32); rng[:n, 1] = l + c + r; rng[0, 0] = l; rng = rng.cuda()
ctx = torch.zeros((n * c, d), device="cuda", dtype=torch.bfloat16)
def attn1(): cflib.check(L.cf_op_attention(1, p(qkv), p(pos), p(rng), p(ctx), n, c, l, r, d, H, 1, st))
def attn3(): cflib.check(L.cf_op_attention(3, p(qkv), p(pos), p(rng), p(ctx), n, c, l, r, d, H, 1, st))     # ring kernel (64-key blocks)
g = torch.randn((n * c + 14 + 64, d), device="cuda").bfloat16()
z = torch.empty((n * c, d), device="cuda", dtype=torch.bfloat16)
w, b = torch.randn((d, 15), device="cuda") * 0.3, torch.randn(d, device="cuda") * 0.1
lw, lb = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
crng = torch.zeros((n, 2), dtype=torch.int32); crng[:, 1] = c + 14; crng = crng.cuda()
def dwconv(): cflib.check(L.cf_op_dwconv(d, 15, p(g), p(z), p(w), p(b), p(lw), p(lb), p(crng), c, n, st))
x = torch.randn((n * c, d), device="cuda")
y = torch.empty((n * c, d), device="cuda", dtype=torch.bfloat16)
def ln0(): cflib.check(L.cf_op_layernorm(0, d, p(x), None, p(y), p(lw), p(lb), None, None, n * c, st))
only = sys.argv[1:]
for f in (attn1, attn3, dwconv, ln0):
    if only and f.__name__ not in only: continue
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    print(f.__name__, "ms", e0.elapsed_time(e1) / 5)
