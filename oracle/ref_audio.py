"""TEST INFRASTRUCTURE (oracle side) -- CPU restatement of the reference's audio preparation steps.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

  get_audio_fragment   ref:src/dataset/vocaset.py:408-430 (with normalize_audio, :64-69)
  resample             torchaudio.functional.resample (third-party, version unpinned by ref:requirements.txt; restated from
                       torchaudio 2.11.0 functional.py _get_sinc_resample_kernel / _apply_sinc_resample_kernel), called at
                       ref:src/dataset/vocaset.py:279-283 (22 kHz -> 16 kHz) and ref:src/model/extractor.py:88.

Pinned by tests/golden/audio_prep.npz (tests/golden/make_golden_audio_prep.py: the reference's own get_audio_fragment,
extracted from its source file because the module itself needs `lightning`, and the live torchaudio resample).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def normalize_audio(audio: np.ndarray) -> np.ndarray:
    assert audio.dtype == np.int16
    return (audio / 32768).astype(np.float32)


def get_audio_fragment(audio: np.ndarray, idx: int, fps: int, sample_rate: int, length: float, shift: int):
    l_pad = int(sample_rate * length / 2) + shift
    n_pad = int(sample_rate * length / 2)
    pad_audio = np.concatenate([np.zeros(l_pad, dtype=audio.dtype), audio, np.zeros(2 * n_pad, dtype=audio.dtype)])
    start = idx * sample_rate // fps
    end = start + 2 * n_pad
    if end > len(pad_audio):
        return None
    return pad_audio[start:end]


def fragments(audio: np.ndarray, n_frames: int, fps: int = 60, sample_rate: int = 22000, length: float = 0.52,
              shift: int = 0, first_frame: int = 0) -> torch.Tensor:
    rows = [get_audio_fragment(audio, first_frame + f, fps, sample_rate, length, shift) for f in range(n_frames)]
    rows = [normalize_audio(r) if r.dtype == np.int16 else r.astype(np.float32) for r in rows]
    return torch.from_numpy(np.stack(rows))


def resample(x: torch.Tensor, orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99) -> torch.Tensor:
    if orig_freq == new_freq:                    # torchaudio returns the waveform untouched (no 0.99-Nyquist low-pass)
        return x
    gcd = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // gcd, int(new_freq) // gcd
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = torch.arange(-width, width + orig, dtype=x.dtype)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=x.dtype)[:, None, None] / new + idx
    t *= base
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t) * window * (base / orig)
    shape = x.shape
    w = x.reshape(-1, shape[-1])
    n = w.shape[1]
    w = F.pad(w, (width, width + orig))
    y = F.conv1d(w[:, None], kernels, stride=orig).transpose(1, 2).reshape(w.shape[0], -1)
    target = int(math.ceil(new * n / orig))
    return y[..., :target].reshape(shape[:-1] + (target,))
