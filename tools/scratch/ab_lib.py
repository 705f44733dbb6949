"""Same-box A/B of two builds of the library (separate processes, alternating): python tools/scratch/ab_lib.py <lib_a.so> <lib_b.so>"""
import os, subprocess, sys
child = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from chunkformer_b200 import lib as L
old = L.SIGNATURES
import ctypes
_l = ctypes.CDLL(os.environ["CHUNKFORMER_B200_LIB"])
for k in list(old):
    if not hasattr(_l, k): del old[k]
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import CTC_LARGE
from chunkformer_b200.plan import Plan
from chunkformer_b200.synth import masked_batch_lengths, synth_state_dict
geo = CTC_LARGE
enc = ChunkFormerEncoderB200(geo, synth_state_dict(geo, 0), "cuda:0")
lens = masked_batch_lengths()
feats = torch.randn((sum(lens), 80), device="cuda")
def step():
    plan = Plan(64, 128, 128, lens, None, geo.kernel)
    _, o16 = enc.encode_plan(plan, feats, out_dtype=torch.bfloat16)
    return enc.ctc_greedy(o16)
for _ in range(3): step()
torch.cuda.synchronize()
res = []
for rnd in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): tok = step()
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 5)
print(os.environ["CHUNKFORMER_B200_LIB"], " ".join("%.2f" % r for r in res), "ms/step", int(tok.sum()))
'''
for rnd in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, CHUNKFORMER_B200_LIB=lib)
        subprocess.run([sys.executable, "-c", child], env=env)
