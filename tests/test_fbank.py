"""fbank step: the CPU oracle against torchaudio golden vectors (no GPU), the CUDA kernel against
both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import fbank_oracle as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fbank_golden.npz")
CASES = ["noise_1p3s", "short_401", "mix_2s", "exact_400", "quiet_0p5s"]
TOL = 2e-3     # log-mel, absolute: fp32 FFT implementations differ by < 1e-3 on these signals (oracle vs torchaudio 7.4e-4)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_torchaudio_golden(name):
    g = np.load(GOLD)
    a = F.fbank(g[name + "_wav"])
    b = g[name + "_fbank"]
    assert a.shape == b.shape
    assert np.abs(a - b).max() < TOL


def test_oracle_frame_count_and_edge_cases():
    assert F.num_frames(399) == 0 and F.num_frames(400) == 1 and F.num_frames(559) == 1 and F.num_frames(560) == 2
    assert F.fbank(np.zeros(100, dtype=np.float32)).shape == (0, 80)
    z = F.fbank(np.zeros(800, dtype=np.float32))                 # digital silence: every bin at log(eps)
    assert np.allclose(z, np.log(F.EPS))
    m = F.mel_banks()
    assert m.shape == (80, 257) and (m[:, 256] == 0).all() and (m.sum(1) > 0).all()


def _encoder():
    from chunkformer_b200.encoder import ChunkFormerEncoderB200
    from chunkformer_b200.geometry import EncoderGeometry
    from chunkformer_b200.synth import synth_state_dict
    geo = EncoderGeometry(d_model=256, heads=4, ffn=256, layers=1, kernel=15, vocab=16)
    return ChunkFormerEncoderB200(geo, synth_state_dict(geo, 3), "cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_torchaudio_golden(name):
    g = np.load(GOLD)
    enc = _encoder()
    out = enc.fbank(torch.from_numpy(g[name + "_wav"]).unsqueeze(0)).cpu().numpy()
    assert out.shape == g[name + "_fbank"].shape
    assert np.abs(out - g[name + "_fbank"]).max() < TOL


@pytest.mark.gpu
def test_kernel_matches_oracle_long_and_empty():
    enc = _encoder()
    rs = np.random.RandomState(5)
    w = np.round(rs.randn(16000 * 37 + 123) * 3000.0 + 40.0 * np.sin(np.arange(16000 * 37 + 123) * 0.01)).astype(np.float32)
    out = enc.fbank(torch.from_numpy(w)).cpu().numpy()
    ref = F.fbank(w)
    assert out.shape == ref.shape == (F.num_frames(w.shape[0]), 80)
    assert np.abs(out - ref).max() < TOL
    assert enc.fbank(torch.zeros(399)).shape == (0, 80)
    z = enc.fbank(torch.zeros(4000)).cpu().numpy()
    assert np.allclose(z, np.log(F.EPS))
    out40 = enc.fbank(torch.from_numpy(w[:16000]), num_mel_bins=40).cpu().numpy()
    assert np.abs(out40 - F.fbank(w[:16000], num_mel_bins=40)).max() < TOL
