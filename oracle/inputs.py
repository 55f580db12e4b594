"""TEST INFRASTRUCTURE -- deterministic synthetic inputs shared by the golden-fixture generator, the parity tests
and bench.py (numpy PCG64: identical on the build container and on the GPU box)."""
from __future__ import annotations

import os

import numpy as np
import torch

V = 5023
V3 = V * 3
_HERE = os.path.dirname(os.path.abspath(__file__))


def _rng(seed: int, tag: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, tag]))


def flame_like_template(seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    """[5023,3] head-sized point cloud (metres * scale).  The real FLAME sample mesh of the reference
    (ref:assets/FLAME_sample.obj) is not redistributed here; parity does not depend on the template values."""
    r = _rng(seed, 101)
    pts = r.standard_normal((V, 3)).astype(np.float32)
    pts /= np.linalg.norm(pts, axis=1, keepdims=True) + 1e-6
    pts *= np.array([0.08, 0.11, 0.09], dtype=np.float32)
    return torch.from_numpy(pts * np.float32(scale))


def one_hot(batch: int, n: int = 12, seed: int = 0) -> torch.Tensor:
    idx = _rng(seed, 102).integers(0, 8, size=batch)    # < 8 so VOCA's one_hot[:, :8] stays one-hot
    return torch.eye(n, dtype=torch.float32)[torch.from_numpy(idx)]


def voca_features(batch: int, seed: int = 0) -> torch.Tensor:
    """[B,29,16] DeepSpeech-like window features (ref:src/model/voca.py:38)."""
    return torch.from_numpy(_rng(seed, 103).standard_normal((batch, 29, 16)).astype(np.float32))


def a2m_features(batch: int, seed: int = 0) -> torch.Tensor:
    """[B,52,32] MFCC windows (ref:config.yaml:8-12, ref:src/model/audio2face.py:57)."""
    return torch.from_numpy(_rng(seed, 104).standard_normal((batch, 52, 32)).astype(np.float32))


def audio(batch: int, n_samples: int, seed: int = 0) -> torch.Tensor:
    """[B,N] raw 16 kHz audio, 0.1 * N(0,1) plus a per-utterance DC offset and gain so that the processor's
    zero-mean / unit-variance step is exercised."""
    r = _rng(seed, 105)
    a = 0.1 * r.standard_normal((batch, n_samples)).astype(np.float32)
    gain = r.uniform(0.5, 1.5, size=(batch, 1)).astype(np.float32)
    dc = 0.02 * r.standard_normal((batch, 1)).astype(np.float32)
    return torch.from_numpy(a * gain + dc)


def batch_templates(batch: int, seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    """[B,5023,3]: one base mesh plus a small per-sample identity offset."""
    base = flame_like_template(seed, scale)
    r = _rng(seed, 106)
    off = (0.002 * scale) * r.standard_normal((batch, V, 3)).astype(np.float32)
    return base[None] + torch.from_numpy(off)


def gt_like(pred_shape, template: torch.Tensor, seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    """Ground-truth vertices: template + 0.002*scale*N(0,1)."""
    r = _rng(seed, 107)
    noise = (0.002 * scale) * r.standard_normal(tuple(pred_shape)).astype(np.float32)
    return template + torch.from_numpy(noise)


def speech_like_windows(batch: int, n_samples: int = 11440, sample_rate: int = 22000, seed: int = 0) -> torch.Tensor:
    """[B, N] audio windows for the MFCC extractor (ref:src/dataset/vocaset.py:408-430: 0.52 s at 22 kHz = 11 440
    samples): harmonics of a per-window pitch with a formant-like envelope, broadband noise, one window with a silent
    stretch (exercises the 1e-10 clamp / top_db floor) and one quiet window (exercises the batch-global cut-off)."""
    r = _rng(seed, 108)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    out = np.zeros((batch, n_samples), dtype=np.float64)
    for b in range(batch):
        f0 = r.uniform(90.0, 260.0)
        amp = r.uniform(0.05, 0.4)
        for h in range(1, 30):
            fh = f0 * h
            if fh > sample_rate / 2:
                break
            env = 1.0 / (1.0 + ((fh - 700.0) / 900.0) ** 2) + 0.5 / (1.0 + ((fh - 2400.0) / 1200.0) ** 2)
            out[b] += amp * env / h ** 0.5 * np.sin(2 * np.pi * fh * t + r.uniform(0, 2 * np.pi))
        out[b] += 0.01 * r.standard_normal(n_samples)
    if batch > 1:
        out[1, : n_samples // 3] = 0.0
    if batch > 2:
        out[2] *= 1e-3
    return torch.from_numpy(out.astype(np.float32))


def pcm16_clip(n_samples: int = 127600, sample_rate: int = 22000, seed: int = 0) -> np.ndarray:
    """int16 mono clip shaped like ref:assets/audio_sample.npy (22 kHz, 127 600 samples = 5.8 s; the asset itself is not
    redistributed): speech-like harmonics with a slow amplitude envelope plus noise, scaled to ~60 % of full scale."""
    r = _rng(seed, 109)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    f0 = 120.0 + 30.0 * np.sin(2 * np.pi * 0.7 * t)
    phase = 2 * np.pi * np.cumsum(f0) / sample_rate
    x = np.zeros(n_samples)
    for h in range(1, 25):
        x += np.sin(h * phase + r.uniform(0, 2 * np.pi)) / h
    x *= 0.5 * (1.0 + np.sin(2 * np.pi * 2.3 * t)) * (t > 0.2)
    x += 0.02 * r.standard_normal(n_samples)
    x *= 0.6 / np.abs(x).max()
    return np.round(x * 32767.0).astype(np.int16)
