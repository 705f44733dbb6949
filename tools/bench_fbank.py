"""fbank of one hour of 16 kHz audio on the device: cf_fbank vs torchaudio.compliance.kaldi.fbank on the same GPU."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chunkformer_b200.encoder import ChunkFormerEncoderB200
from chunkformer_b200.geometry import EncoderGeometry
from chunkformer_b200.synth import synth_state_dict
geo = EncoderGeometry(d_model=256, heads=4, ffn=256, layers=1, kernel=15, vocab=16)
enc = ChunkFormerEncoderB200(geo, synth_state_dict(geo, 3), "cuda:0")
secs = 3600
w = torch.round(torch.randn(1, 16000 * secs, device="cuda") * 3000.0)
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, o
t_cf, a = timeit(lambda: enc.fbank(w))
print(f"cf_fbank: {t_cf:.3f} ms per {secs} s of audio ({a.shape[0]} frames; {(w.numel() * 4 + a.numel() * 4) / t_cf / 1e6:.0f} GB/s algorithmic)")
try:
    import torchaudio.compliance.kaldi as kaldi
    t_ta, b = timeit(lambda: kaldi.fbank(w, num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0, energy_floor=0.0, sample_frequency=16000), n=3)
    print(f"torchaudio kaldi.fbank on the same GPU: {t_ta:.3f} ms; max |diff| {float((a - b).abs().max()):.2e}")
except Exception as e:
    print("torchaudio comparison unavailable:", e)
