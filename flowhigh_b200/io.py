"""WAV file I/O for the entry points.

The reference reads and writes audio with `torchaudio.load` / `torchaudio.save` (example.py:10-12): float32 in
[-1, 1], shape [channels, T].  These helpers give the same contract without a codec dependency: RIFF/WAVE PCM
(8/16/24/32-bit) and IEEE float (32/64-bit), plain or WAVE_FORMAT_EXTENSIBLE, little endian.
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Tuple, Union

import numpy as np
import torch

_PCM, _FLOAT, _EXTENSIBLE = 1, 3, 0xFFFE


def load_wav(path: Union[str, Path]) -> Tuple[torch.Tensor, int]:
    """-> (float32 tensor [channels, T] in [-1, 1], sample_rate), like `torchaudio.load`."""
    data = Path(path).read_bytes()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos: pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8: pos + 8 + size]
        if cid == b"fmt ":
            if size < 16:
                raise ValueError(f"{path}: truncated fmt chunk")
            tag, ch, sr, _, _, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == _EXTENSIBLE and size >= 26:
                tag = struct.unpack_from("<H", body, 24)[0]  # first two bytes of the sub-format GUID
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)  # chunks are word aligned
    if fmt is None or payload is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, sr, bits = fmt
    if ch < 1:
        raise ValueError(f"{path}: no channels")
    frame = ch * (bits // 8)
    payload = payload[: len(payload) // frame * frame]
    if tag == _PCM:
        if bits == 8:
            x = (np.frombuffer(payload, np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(payload, "<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(payload, np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(payload, "<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise ValueError(f"{path}: unsupported PCM width {bits}")
    elif tag == _FLOAT and bits in (32, 64):
        x = np.frombuffer(payload, "<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE format tag {tag} / {bits} bits")
    return torch.from_numpy(np.ascontiguousarray(x.reshape(-1, ch).T)), int(sr)


def save_wav(path: Union[str, Path], wav, sample_rate: int, bits_per_sample: int = 16) -> None:
    """Writes [channels, T] or [T] float audio in [-1, 1] as PCM16 (default) or float32 (bits_per_sample=32),
    like `torchaudio.save(path, wav.cpu(), sr)` in example.py:12."""
    x = wav.detach().cpu().float().numpy() if isinstance(wav, torch.Tensor) else np.asarray(wav, dtype=np.float32)
    if x.ndim == 1:
        x = x[None]
    if x.ndim != 2:
        raise ValueError("save_wav expects [channels, T] or [T]")
    ch, _ = x.shape
    inter = np.ascontiguousarray(x.T)
    if bits_per_sample == 16:
        body = np.clip(np.rint(inter * 32768.0), -32768, 32767).astype("<i2").tobytes()
        tag = _PCM
    elif bits_per_sample == 32:
        body = inter.astype("<f4").tobytes()
        tag = _FLOAT
    else:
        raise ValueError("bits_per_sample must be 16 (PCM) or 32 (float)")
    width = bits_per_sample // 8
    fmt = struct.pack("<HHIIHH", tag, ch, int(sample_rate), int(sample_rate) * ch * width, ch * width, bits_per_sample)
    out = b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + len(body) + (len(body) & 1)) + b"WAVE"
    out += b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(body)) + body
    if len(body) & 1:
        out += b"\0"
    Path(path).write_bytes(out)
