"""The "existing library path" on the same B200 (last bullet; not a pytest module): the oracle's plain-PyTorch
restatement of the reference's masked-chunk forward (same op sequence: gather-unfolded K/V windows, fp32 softmax, ATen convs)
executed by torch's own CUDA kernels in eager mode, fp32 and bf16 autocast, on the benchmark batch (or a part of it).
The unmodified reference cannot travel to the GPU box (/root/reference is absent there); the oracle is pinned to it.
    python tests/perf_torch_eager_gpu.py [n_utterances]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.geometry import CTC_LARGE  # noqa: E402
from chunkformer_b200.synth import masked_batch_lengths, synth_fbank, synth_state_dict  # noqa: E402
from oracle import chunkformer_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
_from_numpy = torch.from_numpy
torch.from_numpy = lambda a: _from_numpy(np.ascontiguousarray(a)).to(dev)      # the oracle's index tables follow the data
sd = {k: v.to(dev) for k, v in synth_state_dict(CTC_LARGE, 0).items()}
lens = masked_batch_lengths(1.0)
n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else len(lens)
lens = sorted(lens)[:n_utt]
xs = [synth_fbank(t, seed=1 + k).to(dev) for k, t in enumerate(lens)]
audio = sum((t + 2) / 100.0 for t in lens)
with torch.device(dev), torch.no_grad():
    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16 autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        with ctx:
            for it in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                out, enc_lens, n_chunks, _, _, _ = O.forward_parallel_chunk(sd, CTC_LARGE.heads, xs, lens, 64, 128, 128)
                tok = O.ctc_greedy(sd, out)
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"torch eager on {torch.cuda.get_device_name(0)}, {name}: {len(lens)} utterances, {audio:.0f} s audio, "
              f"{sum(n_chunks)} chunks: {dt * 1e3:.1f} ms = {audio / dt / 3600:.2f} audio-h/s "
              f"(peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)")
