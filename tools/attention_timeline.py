"""Timeline of CTA 0 of the version-2 attention kernel (SM clocks): where the producer, the MMA issuer and one softmax
thread wait.  Benchmark-size launch (2821 chunks, 64/128/128)."""
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
def attn(): cflib.check(L.cf_op_attention(2, p(qkv), p(pos), p(rng), p(ctx), n, c, l, r, d, H, 1, st))
for _ in range(3): attn()
torch.cuda.synchronize()
buf = torch.zeros(3 * 512, dtype=torch.int64, device="cuda")
L.cf_debug_attention_trace(p(buf))
attn(); torch.cuda.synchronize()
L.cf_debug_attention_trace(None)
t = buf.cpu().view(3, 64, 8)
t0 = int(t[t > 0].min())
rel = lambda v: (int(v) - t0) if int(v) > 0 else -1
print("blk | softmax thread (warp 2 lane 0): iter_start  pull_done  exp_done  pv_done_ok  p_arrive  skew_done(next blk)  barrier_done(next blk)  | iteration length")
prev = None
for b in range(3, 24):
    sm = t[2, b]
    nxt = t[2, b + 1]
    print(f"{b:3d} | " + " ".join(f"{rel(v):7d}" for v in sm[:5]) + f" {rel(nxt[5]):7d} {rel(nxt[6]):7d} | {int(t[2, b + 1, 0]) - int(sm[0])}")
