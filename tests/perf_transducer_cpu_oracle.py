"""CPU side of the transducer search comparison (not a pytest module): the oracle (oracle/transducer_oracle.py, the reference's
per-frame greedy loop restated) on a bounded sample of the workload tools/bench_transducer.py runs on the GPU: the first 1500
encoder frames of one utterance, rnnt-large head sizes, all host threads.
    python tests/perf_transducer_cpu_oracle.py [blank_bias] [cuda]
With `cuda` the same per-frame loop runs on the GPU with torch's eager kernels (one predictor step, one joint call and a host
read-back per decision: the way the reference's search behaves on a GPU)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.synth import synth_transducer_state_dict  # noqa: E402
from oracle import transducer_oracle as T  # noqa: E402

bb = float(sys.argv[1]) if len(sys.argv) > 1 else 9.0
sd = synth_transducer_state_dict(1024, 256, 512, 2, 512, 512, 512, blank_bias=bb, seed=13)
sample = torch.randn((1500, 512), generator=torch.Generator().manual_seed(1))
torch.set_num_threads(os.cpu_count() or 1)
where = f"CPU oracle ({os.cpu_count()} threads)"
if "cuda" in sys.argv[1:]:
    dev = torch.device("cuda:0")
    sd = {k: v.to(dev) for k, v in sd.items()}
    sample = sample.to(dev)
    where = f"oracle loop in torch eager on {torch.cuda.get_device_name(0)}"
    with torch.device(dev):
        T.greedy_search_one(sd, sample[:100], 100, 64)          # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        grid = T.greedy_search_one(sd, sample, 1500, 64)
        torch.cuda.synchronize()
else:
    t0 = time.perf_counter()
    grid = T.greedy_search_one(sd, sample, 1500, 64)
dt = time.perf_counter() - t0
print(f"{where}, blank_bias {bb}: 1500 frames, {int((grid != 0).sum())} symbols in {dt:.2f} s "
      f"= {1500 * 0.08 / dt / 3600:.4f} audio-h/s")
