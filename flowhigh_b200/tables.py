"""Host-side constant tables of the DSP stages (computed once per process, uploaded once).

* resampler taps: scipy.signal.firwin, exactly what scipy.signal.resample_poly designs at
  flowhighsr.py:68 (Kaiser beta=5, 20*max(up,down)+1 taps), cast to fp32 and scaled by `up`
  in fp32 like scipy does for an fp32 input.
* mel filterbank: Slaney-scale / Slaney-norm triangles of librosa.filters.mel(sr=48000,
  n_fft=2048, n_mels=256, fmin=20, fmax=24000) (call site melvoco.py:64-70), stored sparse:
  every band touches a contiguous run of <= 33 FFT bins (2030 non-zeros of 262 400).
* FFT twiddles exp(-2 pi i k / 2048), k < 1024, rounded from fp64.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np


@lru_cache(maxsize=None)
def resample_plan(sr_in: int, sr_out: int):
    """(taps fp32, up, down, n_pre_pad, n_pre_remove) of scipy's polyphase design."""
    from scipy.signal import firwin
    g = math.gcd(sr_out, sr_in)
    up, down = sr_out // g, sr_in // g
    if up == 1 and down == 1:
        return None
    half = 10 * max(up, down)
    h = firwin(2 * half + 1, 1.0 / max(up, down), window=("kaiser", 5.0)).astype(np.float32)
    h = (h * np.float32(up)).astype(np.float32)
    n_pre_pad = down - half % down
    n_pre_remove = (half + n_pre_pad) // down
    return h, up, down, n_pre_pad, n_pre_remove


def soxr_hq_spec():
    """Band edges and rejection of libsoxr's HQ recipe (soxr.c `soxr_quality_spec`, quality 4 = SOXR_HQ as
    `librosa.resample(..., res_type='soxr_hq')` selects it at flowhighsr.py:76): 20-bit precision => 120.4 dB rejection,
    linear phase, stop band from 1.0 x the lower Nyquist, pass band to 1 - 0.05 / TO_3dB(rej) = 0.9136 x that."""
    rej = 20 * 20.0 * math.log10(2.0)
    to_3db = (1.6e-6 * rej - 7.5e-4) * rej + 0.646
    return rej, 1.0 - 0.05 / to_3db, 1.0


@lru_cache(maxsize=None)
def resample_plan_soxr_hq(sr_in: int, sr_out: int):
    """Same tuple as `resample_plan` for `upsampling_method='librosa'`: ONE linear-phase Kaiser-windowed-sinc polyphase
    filter designed to soxr_hq's published band edges and rejection.  libsoxr itself (not installable offline, cannot
    be pinned) reaches the same specification with a multi-stage half-band / polyphase cascade, so the two agree inside
    the pass band and the stop band to the 20-bit precision the recipe promises and differ only in the shape of the
    transition band (0.9136 .. 1.0 of the input Nyquist frequency) and in how the first / last filter-length of samples
    is extrapolated (here: zero extension, as in the scipy path)."""
    from scipy.signal import firwin
    g = math.gcd(sr_out, sr_in)
    up, down = sr_out // g, sr_in // g
    if up == 1 and down == 1:
        return None
    rej, fp, fs = soxr_hq_spec()
    q = max(up, down)
    beta = 0.1102 * (rej - 8.7)
    width = (fs - fp) / q                      # transition width as a fraction of the Nyquist rate of the up-rate grid
    numtaps = int(math.ceil((rej - 7.95) / (2.285 * math.pi * width))) + 1
    half = numtaps // 2 + 1
    h = firwin(2 * half + 1, 0.5 * (fp + fs) / q, window=("kaiser", beta))
    h = (h * up).astype(np.float32)
    n_pre_pad = down - half % down
    n_pre_remove = (half + n_pre_pad) // down
    return h, up, down, n_pre_pad, n_pre_remove


def resample_out_len(n_in: int, up: int, down: int) -> int:
    return -(-n_in * up // down)


def _slaney_hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    log_part = 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) / (np.log(6.4) / 27.0)
    return np.where(f >= 1000.0, log_part, lin)


def _slaney_mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * (200.0 / 3)
    log_part = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0))
    return np.where(m >= 15.0, log_part, lin)


@lru_cache(maxsize=None)
def mel_filterbank_dense(sr: int = 48000, n_fft: int = 2048, n_mels: int = 256, fmin: float = 20.0,
                         fmax: float = 24000.0) -> np.ndarray:
    n_freq = n_fft // 2 + 1
    freqs = np.linspace(0.0, sr / 2.0, n_freq)
    edges = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(fmin), _slaney_hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    # rising / falling slopes of every triangle, vectorised over bands
    rise = (freqs[None, :] - edges[:-2, None]) / width[:-1, None]
    fall = (edges[2:, None] - freqs[None, :]) / width[1:, None]
    tri = np.maximum(0.0, np.minimum(rise, fall)).astype(np.float32)
    norm = (2.0 / (edges[2:] - edges[:-2])).astype(np.float32)
    return tri * norm[:, None]


@lru_cache(maxsize=None)
def mel_filterbank_sparse():
    """(start[256] int32, length[256] int32, weights[256, stride] fp32, stride)."""
    dense = mel_filterbank_dense()
    n_mels = dense.shape[0]
    start = np.zeros(n_mels, dtype=np.int32)
    length = np.zeros(n_mels, dtype=np.int32)
    runs = []
    for m in range(n_mels):
        nz = np.nonzero(dense[m])[0]
        if nz.size == 0:
            runs.append(np.zeros(0, dtype=np.float32))
            continue
        start[m] = nz[0]
        length[m] = nz[-1] - nz[0] + 1
        runs.append(dense[m, nz[0]: nz[-1] + 1])
    stride = int(max(8, max(len(r) for r in runs)))
    w = np.zeros((n_mels, stride), dtype=np.float32)
    for m, r in enumerate(runs):
        w[m, : len(r)] = r
    return start, length, w, stride


@lru_cache(maxsize=None)
def fft_twiddles(n: int = 2048) -> np.ndarray:
    k = np.arange(n // 2, dtype=np.float64)
    ang = -2.0 * np.pi * k / n
    return np.stack([np.cos(ang), np.sin(ang)], axis=1).astype(np.float32)
