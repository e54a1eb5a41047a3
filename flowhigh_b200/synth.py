"""Deterministic speech-like synthetic audio: no dataset is available offline.

Harmonic stack sum_k sin(k*phi)/k with f0(t) = 120 + 30 sin(2 pi 0.7 t) Hz, harmonics up to
0.95 * Nyquist of the INPUT rate (band-limited by construction), 3 Hz syllabic AM, plus
0.01 * N(0,1) noise; peak 0.9; float32.
"""
import numpy as np


def synth_speech(n: int, sr: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr
    f0 = 120.0 + 30.0 * np.sin(2 * np.pi * 0.7 * t + 0.37 * seed)
    phi = 2 * np.pi * np.cumsum(f0) / sr
    x = np.zeros(n, dtype=np.float64)
    kmax = int(0.95 * (sr / 2) / 150.0)
    for k in range(1, kmax + 1):
        x += np.sin(k * phi) / k
    am = 0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t) ** 2
    x = x * am + 0.01 * rng.standard_normal(n)
    x = 0.9 * x / np.max(np.abs(x))
    return x.astype(np.float32)
