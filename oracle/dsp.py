"""ORACLE (test infrastructure, not product code) -- DSP stages of `FlowHighSR.generate`.

CPU restatement, in numpy / plain torch-CPU fp32, of the signal-processing stages of the
reference hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product (flowhigh_b200/) never does.

Third-party arithmetic restated here because the dependency is not under /root/reference:
  * scipy.signal.resample_poly (pyproject.toml:6 `scipy>=1.10.1`; installed 1.18.1) -- call
    site flowhighsr.py:68.  Restated from the published upfirdn algorithm
    (scipy/signal/_signaltools.py `resample_poly`, firwin + Kaiser(beta=5) design);
    PINNED: tests/test_oracle_cpu.py checks it against the installed scipy bit-for-bit-ish
    (<=1e-6) for every rate pair the configs use.
  * librosa.filters.mel (pyproject.toml:8 `librosa>=0.9.2`, NOT installed here) -- call site
    melvoco.py:64-70.  Restated from the published Slaney-scale algorithm (htk=False,
    norm='slaney').  PARITY UNPINNED for this one table: no librosa is available to check
    against; structural properties (shape, 2030 non-zeros, row sums) are tested instead.
  * torchdiffeq.odeint fixed-grid euler / midpoint (pyproject.toml:14) -- see oracle/cfm.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# resampler  (flowhighsr.py:62-69)
# --------------------------------------------------------------------------------------
def _kaiser_window(n: int, beta: float) -> np.ndarray:
    # scipy.signal.windows.kaiser(n, beta, sym=True)
    k = np.arange(n, dtype=np.float64)
    alpha = (n - 1) / 2.0
    return np.i0(beta * np.sqrt(np.clip(1 - ((k - alpha) / alpha) ** 2, 0, 1))) / np.i0(beta)


def firwin_kaiser_lowpass(numtaps: int, cutoff: float, beta: float = 5.0) -> np.ndarray:
    """scipy.signal.firwin(numtaps, cutoff, window=('kaiser', beta)) for one low-pass band.

    h[n] = cutoff*sinc(cutoff*(n-alpha)) * w[n], scaled so the DC gain is exactly 1.
    """
    alpha = 0.5 * (numtaps - 1)
    m = np.arange(numtaps, dtype=np.float64) - alpha
    h = cutoff * np.sinc(cutoff * m)
    h *= _kaiser_window(numtaps, beta)
    h /= h.sum()  # scale_frequency = 0 for a low-pass starting at DC
    return h


def resample_poly_params(sr_in: int, sr_out: int):
    g = math.gcd(sr_out, sr_in)
    up, down = sr_out // g, sr_in // g
    half = 10 * max(up, down)
    n_pre_pad = down - half % down
    n_pre_remove = (half + n_pre_pad) // down
    return up, down, half, n_pre_pad, n_pre_remove


def resample_poly(x: np.ndarray, sr_out: int, sr_in: int) -> np.ndarray:
    """y = scipy.signal.resample_poly(x, sr_out, sr_in) for 1-D x (zero edge extension).

    y[m] = sum_i x[i] * h[(m + n_pre_remove)*down - n_pre_pad - i*up],  h = up * firwin(...)
    cast to x.dtype (scipy casts the taps to the input dtype; flowhighsr.py:68 passes the
    caller's dtype through).
    """
    x = np.asarray(x)
    up, down, half, n_pre_pad, n_pre_remove = resample_poly_params(sr_in, sr_out)
    if up == 1 and down == 1:
        return x.copy()
    dt = x.dtype if x.dtype in (np.float32, np.float64) else np.float64
    h = (firwin_kaiser_lowpass(2 * half + 1, 1.0 / max(up, down)).astype(dt) * dt.type(up)).astype(dt)
    n_in = x.shape[0]
    n_out = -(-n_in * up // down)
    # zero-stuffed convolution, evaluated phase by phase in the input dtype
    xu = np.zeros(n_in * up, dtype=dt)
    xu[::up] = x.astype(dt)
    full = np.convolve(xu, h)  # full[n] = sum_k h[k] xu[n-k]
    idx = (np.arange(n_out) + n_pre_remove) * down - n_pre_pad
    y = np.zeros(n_out, dtype=dt)
    ok = (idx >= 0) & (idx < full.shape[0])
    y[ok] = full[idx[ok]]
    return y


def soxr_hq_filter(sr_in: int, sr_out: int):
    """FIR taps (fp64, unit DC gain) meeting the specification of libsoxr's HQ recipe, the resampler behind
    `librosa.resample(audio, sr, target, res_type='soxr_hq')` (flowhighsr.py:74-80).  libsoxr / librosa are third-party
    dependencies (pyproject.toml:8 `librosa>=0.9.2`) that are NOT installed offline: PARITY UNPINNED.  Restated from
    soxr.c `soxr_quality_spec`: quality HQ = 20-bit precision => rejection 20*6.02 = 120.4 dB, linear phase, stop band
    from 1.0 x the lower Nyquist frequency, pass band to 1 - 0.05/TO_3dB(rej) of it, TO_3dB(a) = (1.6e-6 a - 7.5e-4) a + 0.646.
    One Kaiser-windowed sinc (beta = 0.1102 (A - 8.7), length from Kaiser's formula) meets that specification; libsoxr
    meets it with a multi-stage cascade, so the two differ only inside the transition band and at the clip edges."""
    g = math.gcd(sr_out, sr_in)
    up, down = sr_out // g, sr_in // g
    rej = 20 * 20.0 * math.log10(2.0)
    fp = 1.0 - 0.05 / ((1.6e-6 * rej - 7.5e-4) * rej + 0.646)
    q = max(up, down)
    width = (1.0 - fp) / q
    numtaps = int(math.ceil((rej - 7.95) / (2.285 * math.pi * width))) + 1
    half = numtaps // 2 + 1
    return firwin_kaiser_lowpass(2 * half + 1, 0.5 * (fp + 1.0) / q, beta=0.1102 * (rej - 8.7)), up, down


def resample_soxr_hq(x: np.ndarray, sr_out: int, sr_in: int) -> np.ndarray:
    """Polyphase resampling with `soxr_hq_filter`, zero edge extension, ceil(n * ratio) output samples (librosa fixes the
    length to that).  scipy's upfirdn (through resample_poly with an explicit FIR) does the arithmetic in x's dtype."""
    from scipy.signal import resample_poly as _rp
    x = np.asarray(x)
    h, up, down = soxr_hq_filter(sr_in, sr_out)
    if up == 1 and down == 1:
        return x.copy()
    dt = x.dtype if x.dtype in (np.float32, np.float64) else np.float64
    return _rp(x.astype(dt), up, down, window=h.astype(dt))


def preprocess_audio(audio: np.ndarray, sr: int, target_sr: int = 48000, method: str = "scipy") -> np.ndarray:
    """flowhighsr.py:59-80: squeeze, int16 heuristic, resample (scipy branch :66-72 or librosa/soxr_hq branch :74-80),
    peak normalise."""
    audio = np.asarray(audio)
    if audio.ndim == 2:
        audio = audio.squeeze(0)
    if audio.max() > 1:
        audio = audio / 32768.0
    cond = resample_poly(audio, target_sr, sr) if method == "scipy" else resample_soxr_hq(audio, target_sr, sr)
    cond = cond / np.max(np.abs(cond))
    return cond


# --------------------------------------------------------------------------------------
# mel filterbank  (librosa.filters.mel restated; melvoco.py:64-70)
# --------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-12) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr=48000, n_fft=2048, n_mels=256, fmin=20.0, fmax=24000.0) -> np.ndarray:
    """Slaney-scale, Slaney-normalised triangular filterbank, float32 [n_mels, 1+n_fft/2]."""
    n_freq = 1 + n_fft // 2
    fftfreqs = np.linspace(0, sr / 2.0, n_freq, dtype=np.float64)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    mel_f = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, n_freq), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None].astype(np.float32)
    return w


_MEL_CACHE = {}


def mel_basis_48k() -> torch.Tensor:
    if "b" not in _MEL_CACHE:
        _MEL_CACHE["b"] = torch.from_numpy(mel_filterbank())
    return _MEL_CACHE["b"]


# --------------------------------------------------------------------------------------
# log-mel front end  (melvoco.py:56-86, modules.py:31-36)
# --------------------------------------------------------------------------------------
def encode_logmel(audio: torch.Tensor, n_fft=2048, hop=480, win=2048) -> torch.Tensor:
    """audio [B,T] fp32 -> log-mel [B,N,256] fp32."""
    pad = (n_fft - hop) // 2
    x = F.pad(audio.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    window = torch.hann_window(win, dtype=audio.dtype)
    spec = torch.stft(x, n_fft, hop_length=hop, win_length=win, window=window, center=False,
                      normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(torch.view_as_real(spec).pow(2).sum(-1) + 1e-9)
    mel = torch.matmul(mel_basis_48k().to(audio.dtype), mag)
    mel = torch.log(torch.clamp(mel, min=1e-5))
    return mel.transpose(1, 2).contiguous()


# --------------------------------------------------------------------------------------
# audio-domain post-processing  (postprocessing.py:6-41)
# --------------------------------------------------------------------------------------
def _stft_center_zero(x: torch.Tensor, n_fft=2048, hop=480) -> torch.Tensor:
    # torchaudio Spectrogram(power=None, pad_mode='constant', center=True, normalized=False)
    window = torch.hann_window(n_fft, dtype=x.dtype)
    return torch.stft(x, n_fft, hop_length=hop, win_length=n_fft, window=window, center=True,
                      pad_mode="constant", normalized=False, onesided=True, return_complex=True)


def cutoff_index(spec_src: torch.Tensor, threshold: float = 0.99) -> int:
    """postprocessing.py:10-16, vectorised: the loop scans from the top bin down and never
    tests bin 0; it returns F - i for the first i>=1 with energy[F-i] < thr, else 0."""
    energy = torch.cumsum(torch.sum(spec_src.squeeze().abs(), dim=-1), dim=0)
    thr = energy[-1] * threshold
    Fb = energy.shape[0]
    for i in range(1, Fb):
        if energy[-i] < thr:
            return Fb - i
    return 0


def postprocess(pred: torch.Tensor, src: torch.Tensor, length: int) -> torch.Tensor:
    """pred [1,T'], src [1,T] -> [1,length]; low band from src, high band from pred."""
    sp, ss = _stft_center_zero(pred), _stft_center_zero(src)
    cr = cutoff_index(ss)
    nt = min(sp.shape[-1], ss.shape[-1])
    res = torch.empty_like(sp[:, :, :nt])
    res[:, cr:] = sp[:, cr:, :nt]
    res[:, :cr] = ss[:, :cr, :nt]
    window = torch.hann_window(2048, dtype=pred.dtype)
    audio = torch.istft(res, 2048, hop_length=480, win_length=2048, window=window, center=True,
                        normalized=False, onesided=True, length=length)
    return audio / audio.abs().max() * 0.99
