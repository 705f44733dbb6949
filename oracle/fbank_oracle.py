"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy, float32/float64) of the Kaldi-compatible log-mel filterbank the
reference computes right before the hot path (chunkformer/chunkformer_model.py:307-315 calls
torchaudio.compliance.kaldi.fbank(waveform, num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0, energy_floor=0.0,
sample_frequency=16000); every other argument keeps torchaudio's default: snip_edges, remove_dc_offset, preemphasis 0.97,
povey window, round_to_power_of_two, low_freq 20, high_freq 0 (= Nyquist), use_power, use_log_fbank, no energy column).
torchaudio (pinned as torchaudio>=2.5.1 in the reference's pyproject.toml) is a third-party dependency that is not part
of /root/reference; this file restates its published algorithm and is pinned against outputs of torchaudio 2.11 itself
(tests/golden/make_golden_fbank.py -> tests/golden/fbank_golden.npz).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline may import it; the product path is the CUDA kernel in chunkformer_b200/csrc/fbank.cuh."""
import math

import numpy as np

EPS = np.float32(1.1920928955078125e-07)     # torch.finfo(torch.float).eps


def num_frames(n_samples: int, frame_len: int = 400, frame_shift: int = 160) -> int:
    """snip_edges=True (kaldi.py _get_strided)."""
    return 0 if n_samples < frame_len else 1 + (n_samples - frame_len) // frame_shift


def povey_window(n: int) -> np.ndarray:
    """hann(n, periodic=False) ** 0.85 (kaldi.py _feature_window_function)."""
    k = np.arange(n, dtype=np.float64)
    return ((0.5 - 0.5 * np.cos(2.0 * math.pi * k / (n - 1))) ** 0.85).astype(np.float32)


def mel_scale(f):
    return 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_banks(num_bins: int = 80, padded: int = 512, sample_rate: float = 16000.0, low_freq: float = 20.0,
              high_freq: float = 0.0) -> np.ndarray:
    """(num_bins, padded // 2 + 1) triangular filters in the mel domain (kaldi.py get_mel_banks); the Nyquist column is 0."""
    nfb = padded // 2
    nyq = 0.5 * sample_rate
    if high_freq <= 0.0:
        high_freq += nyq
    width = sample_rate / padded
    lo, hi = mel_scale(low_freq), mel_scale(high_freq)
    delta = (hi - lo) / (num_bins + 1)
    b = np.arange(num_bins, dtype=np.float64)[:, None]
    left, center, right = lo + b * delta, lo + (b + 1.0) * delta, lo + (b + 2.0) * delta
    mel = mel_scale(width * np.arange(nfb, dtype=np.float64))[None, :]
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    w = np.maximum(0.0, np.minimum(up, down))
    return np.concatenate([w, np.zeros((num_bins, 1))], axis=1).astype(np.float32)


def fbank(waveform: np.ndarray, num_mel_bins: int = 80, frame_length_ms: float = 25.0, frame_shift_ms: float = 10.0,
          sample_rate: float = 16000.0, preemph: float = 0.97) -> np.ndarray:
    """waveform (n,) float32 in int16 range -> (T, num_mel_bins) float32 log-mel energies."""
    x = np.asarray(waveform, dtype=np.float32).reshape(-1)
    flen = int(sample_rate * frame_length_ms * 0.001)
    fshift = int(sample_rate * frame_shift_ms * 0.001)
    padded = 1 << (flen - 1).bit_length()
    T = num_frames(x.shape[0], flen, fshift)
    if T == 0:
        return np.zeros((0, num_mel_bins), dtype=np.float32)
    idx = np.arange(T)[:, None] * fshift + np.arange(flen)[None, :]
    fr = x[idx].astype(np.float32)
    fr = fr - fr.mean(axis=1, keepdims=True, dtype=np.float32)                       # remove_dc_offset
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)                            # replicate-padded shift
    fr = fr - np.float32(preemph) * prev
    fr = fr * povey_window(flen)[None, :]
    fr = np.concatenate([fr, np.zeros((T, padded - flen), dtype=np.float32)], axis=1)
    spec = np.abs(np.fft.rfft(fr.astype(np.float64), axis=1)) ** 2                    # power spectrum, 257 bins
    mel = spec.astype(np.float32) @ mel_banks(num_mel_bins, padded, sample_rate).T
    return np.log(np.maximum(mel, EPS)).astype(np.float32)
