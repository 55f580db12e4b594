"""TEST INFRASTRUCTURE (oracle side) -- CPU restatement of the reference's MFCC feature extractor.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

Follows ref:src/model/extractor.py:10-60 (MFCCExtractor) whose arithmetic lives in torchaudio (third-party, not under
/root/reference, version unpinned by ref:requirements.txt; restated from the installed torchaudio 2.11.0):
  transforms.MFCC.forward            mel_specgram -> amplitude_to_DB("power", top_db 80) -> (x^T @ dct_mat)^T
  transforms.Spectrogram / F.spectrogram   torch.stft(n_fft, hop, win_length, hann periodic, center=True, reflect,
                                           onesided) ; power 2 ; normalized False
  transforms.MelScale                (spec^T @ fb)^T, fb = F.melscale_fbanks(n_freqs, 0, sr//2, 128, sr, None, "htk")
  F.amplitude_to_DB                  10*log10(clamp(x, 1e-10)) - 10*log10(max(1e-10, 1.0)); top_db: for a 3-D input ONE
                                     cut-off max(x) - 80 for the whole batch (the reshape packs the batch into channels)
  F.create_dct(n_mfcc, 128, "ortho")
then ref:extractor.py:49-59: transpose(1, 2) and, when frames != out_dim, F.interpolate(size=(out_dim, n_mfcc), bilinear).

Pinned by tests/golden/mfcc.npz, produced by the live reference MFCCExtractor (tests/golden/make_golden_mfcc.py).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

N_MELS = 128
CONFIGS = {
    # name: (sample_rate, n_feature, out_dim, win_length, hop_length, n_fft)   SURVEY.md 8(d) configs 1 and 2
    "voca": (22000, 16, 29, 790, None, 1024),
    "audio2mesh": (22000, 32, 52, 440, None, 1024),
}


def make_buffers(sample_rate: int, n_mfcc: int, win_length: int, n_fft: int) -> "OrderedDict[str, torch.Tensor]":
    """The three persistent buffers of the reference module, by torchaudio's formulas."""
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    hz2mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)                     # noqa: E731  (F._hz_to_mel, htk)
    m_pts = torch.linspace(hz2mel(0.0), hz2mel(float(sample_rate // 2)), N_MELS + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)                               # F._mel_to_hz, htk
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    fb = torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
    n = torch.arange(float(N_MELS))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(N_MELS) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(N_MELS))
    sd = OrderedDict()
    sd["T.dct_mat"] = dct.t().contiguous()
    sd["T.MelSpectrogram.spectrogram.window"] = torch.hann_window(win_length, periodic=True)
    sd["T.MelSpectrogram.mel_scale.fb"] = fb
    return sd


def mfcc_forward(sd, x: torch.Tensor, out_dim: int, win_length: int, hop_length=None, n_fft=None) -> torch.Tensor:
    """x [B, N] fp32 -> [B, out_dim, n_mfcc]."""
    hop = hop_length if hop_length else win_length // 2
    n_fft = n_fft if n_fft else win_length
    window = sd["T.MelSpectrogram.spectrogram.window"]
    wp = torch.zeros(n_fft)
    lo = (n_fft - win_length) // 2
    wp[lo:lo + win_length] = window                                   # torch.stft centres a short window in the frame
    xp = F.pad(x[:, None], (n_fft // 2, n_fft // 2), mode="reflect")[:, 0]
    frames = xp.unfold(1, n_fft, hop) * wp                            # [B, F, n_fft]
    X = torch.fft.rfft(frames, dim=-1)
    power = X.real * X.real + X.imag * X.imag                         # [B, F, n_freqs]
    mel = power @ sd["T.MelSpectrogram.mel_scale.fb"]                 # [B, F, 128]
    db = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
    db = torch.max(db, db.max() - 80.0)                               # single cut-off for the whole 3-D batch
    m = db @ sd["T.dct_mat"]                                          # [B, F, n_mfcc]
    if m.shape[1] != out_dim:
        m = F.interpolate(m[:, None], size=(out_dim, m.shape[2]), mode="bilinear")[:, 0]
    return m
