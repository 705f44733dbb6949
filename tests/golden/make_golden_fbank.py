"""Golden vectors for the fbank step from torchaudio itself (the third-party implementation the reference calls,
chunkformer_model.py:307-315).  Run in the build container: python tests/golden/make_golden_fbank.py"""
import os
import numpy as np
import torch
import torchaudio.compliance.kaldi as kaldi

out = {}
g = torch.Generator().manual_seed(1234)
for name, n in (("noise_1p3s", 20800), ("short_401", 401), ("mix_2s", 32000), ("exact_400", 400), ("quiet_0p5s", 8000)):
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    if name.startswith("mix"):
        w = 6000.0 * torch.sin(2 * torch.pi * 440.0 * t) + 2500.0 * torch.sin(2 * torch.pi * 3111.0 * t + 0.3) + 300.0 * torch.randn(n, generator=g) + 120.0
    elif name.startswith("quiet"):
        w = 3.0 * torch.randn(n, generator=g)
    else:
        w = 4000.0 * torch.randn(n, generator=g) - 57.0
    w = torch.round(w).clamp(-32768, 32767)                      # 16-bit samples as floats, like the reference's pydub path
    f = kaldi.fbank(w.unsqueeze(0), num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0, energy_floor=0.0,
                    sample_frequency=16000)
    out[name + "_wav"] = w.numpy().astype(np.float32)
    out[name + "_fbank"] = f.numpy().astype(np.float32)
    print(name, tuple(f.shape), float(f.mean()))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fbank_golden.npz"), **out)
