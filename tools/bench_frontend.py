"""Time the conv0+ReLU+dw1 front-end kernels (impl 0/1/2) on one slab of 256 chunks (c=64) with CUDA events."""
import sys, os, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200 import lib as cflib
from ctypes import POINTER, c_int32, c_int64

L = cflib.load()
dev = "cuda:0"
d, c, n = 512, 64, 256
size = 8 * (c - 1) + 15
feats = torch.randn((n * 512 + size, 80), device=dev)
rows = (np.arange(n) * 512).astype(np.int64)
lens = np.full(n, size, dtype=np.int32)
wpack = (torch.randn((d, 20), device=dev) * 0.3).contiguous()
T2, F2 = 2 * c + 1, 19
outs = []
for impl in [int(a) for a in sys.argv[1:]] or [1, 2]:
    out = torch.zeros((n * T2 * F2, d), device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    def run():
        cflib.check(L.cf_op_frontend_conv(impl, d, feats.data_ptr(), rows.ctypes.data_as(POINTER(c_int64)),
                                          lens.ctypes.data_as(POINTER(c_int32)), n, c, 80, wpack.data_ptr(), None, None,
                                          out.data_ptr(), st), None, "frontend")
    run(); torch.cuda.synchronize()
    # cf_op_frontend_conv allocates + syncs internally, so time with events around several calls and take the min
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    outs.append(out)
    gb = out.numel() * 2 / 1e9
    print(f"impl {impl}: {best:.3f} ms per slab of {n} chunks  ({gb / best * 1e3:.0f} GB/s written; x{2821 / n:.1f} slabs = {best * 2821 / n:.2f} ms/step)")
if len(outs) == 2:
    print("max abs diff between impls:", (outs[0].float() - outs[1].float()).abs().max().item())
