"""Feature extractors in front of the conv models: drop-in for the reference's MFCCExtractor.

ref:src/model/extractor.py:10-60 wraps torchaudio.transforms.MFCC (MelSpectrogram -> AmplitudeToDB(top_db=80) -> ortho
DCT-II), transposes to [B, frames, n_mfcc] and bilinearly resizes the time axis to out_dim.  Here the same arithmetic
runs as four launches of liba2f_sm100.so: frame gather, DFT as a GEMM against a window-folded (cos | sin) basis on the
a2f_gemm back ends (fp32 SIMT, or tcgen05 on an error-compensated bf16x3 split), mel + dB + batch-global maximum, and
clamp + DCT + resize.  Constructor arguments, output shape and the persistent buffers (`T.dct_mat`,
`T.MelSpectrogram.spectrogram.window`, `T.MelSpectrogram.mel_scale.fb`) match the reference module, so a Lightning
checkpoint's `feature_extractor.*` entries load strictly.  CUDA (sm_100a) only -- there is no CPU fallback.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import lib as L
from . import ops
from .modules import _PackCache

N_MELS = 128        # torchaudio.transforms.MFCC default melkwargs["n_mels"]
TOP_DB = 80.0       # torchaudio.transforms.MFCC: AmplitudeToDB("power", 80.0)


def hann_periodic(win_length: int) -> torch.Tensor:
    """The buffer torchaudio's Spectrogram registers: torch.hann_window(win_length, periodic=True), fp32 arithmetic
    (an fp64 evaluation differs from it by up to 2e-7, and checkpoints carry the fp32 one)."""
    return torch.hann_window(win_length, periodic=True)


def htk_mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") -> [n_freqs, n_mels] fp32 (same fp32 op order)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def dct2_ortho(n_mfcc: int, n_mels: int) -> torch.Tensor:
    """torchaudio.functional.create_dct(n_mfcc, n_mels, norm="ortho") -> [n_mels, n_mfcc]."""
    n = torch.arange(float(n_mels))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n_mels))
    return dct.t()


class _Buffers(nn.Module):
    """Empty container used to reproduce torchaudio's buffer names."""


class MFCCExtractor(nn.Module):
    """
    Input shape: (batch, time)
    Output shape: (batch, out_dim, n_mfcc)
    """

    def __init__(self, sample_rate: int, n_feature: int, out_dim: int, win_length: int, hop_length: int = None,
                 n_fft: int = None):
        super().__init__()
        self.sample_rate = sample_rate
        self.n_mfcc = n_feature
        self.out_dim = out_dim
        self.win_length = win_length
        self.hop_length = hop_length if hop_length else win_length // 2
        self.n_fft = n_fft if n_fft else win_length
        if self.win_length > self.n_fft:
            raise ValueError("win_length must not exceed n_fft")
        self.n_freq = self.n_fft // 2 + 1
        # torchaudio's module tree: MFCC.dct_mat, MFCC.MelSpectrogram.spectrogram.window, MFCC.MelSpectrogram.mel_scale.fb
        self.T = _Buffers()
        self.T.register_buffer("dct_mat", dct2_ortho(self.n_mfcc, N_MELS))
        self.T.MelSpectrogram = _Buffers()
        self.T.MelSpectrogram.spectrogram = _Buffers()
        self.T.MelSpectrogram.spectrogram.register_buffer("window", hann_periodic(self.win_length))
        self.T.MelSpectrogram.mel_scale = _Buffers()
        self.T.MelSpectrogram.mel_scale.register_buffer(
            "fb", htk_mel_filterbank(self.n_freq, 0.0, float(sample_rate // 2), N_MELS, sample_rate))
        self.precision = "fp32"
        self._cache = _PackCache()

    def set_precision(self, precision: str):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        return self

    # ---- derived operands (rebuilt when a buffer changes, e.g. after load_state_dict) -------------------------------
    def _basis(self):
        window = self.T.MelSpectrogram.spectrogram.window
        bf = self.precision == "bf16"

        def build():
            win, n_fft, nf = self.win_length, self.n_fft, self.n_freq
            kpad = (win + 63) // 64 * 64
            lo = (n_fft - win) // 2
            j = torch.arange(win, dtype=torch.float64) + lo
            ang = 2.0 * math.pi * torch.outer(torch.arange(nf, dtype=torch.float64), j) / n_fft
            w = window.detach().double().cpu()
            npad = (2 * nf + 7) // 8 * 8
            basis = torch.zeros((npad, kpad), dtype=torch.float64)
            basis[:nf, :win] = torch.cos(ang) * w
            basis[nf:2 * nf, :win] = torch.sin(ang) * w
            basis = basis.float().to(window.device)
            return {"kpad": kpad, "npad": npad, "w": ops.split_bf16x3(basis, True) if bf else basis}
        return self._cache.get("basis_bf16" if bf else "basis_f32", (window,), build)

    def _bands(self):
        fb = self.T.MelSpectrogram.mel_scale.fb

        def build():
            nz = (fb.detach().cpu() != 0)
            band = torch.zeros((fb.shape[1], 2), dtype=torch.int32)
            for m in range(fb.shape[1]):
                idx = torch.nonzero(nz[:, m]).flatten()
                if idx.numel():
                    band[m, 0], band[m, 1] = int(idx[0]), int(idx[-1]) + 1
            return band.to(fb.device)
        return self._cache.get("bands", (fb,), build)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise L.A2FError("the a2f_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")
        if x.dim() != 2:
            raise L.A2FError("MFCCExtractor expects (batch, time) audio")
        if not self.T.dct_mat.is_cuda:
            raise L.A2FError("move the extractor to the GPU first (.to(device))")
        x = x.contiguous().float()
        B = x.shape[0]
        bs = self._basis()
        bf = self.precision == "bf16"
        gmax = torch.empty(1, dtype=torch.float32, device=x.device)
        frames, F_ = ops.mfcc_frames(x, self.win_length, self.hop_length, self.n_fft, bs["kpad"],
                                     torch.bfloat16 if bf else torch.float32, gmax)
        spec = torch.empty((B * F_, bs["npad"]), dtype=torch.float32, device=x.device)
        ops.gemm(frames, bs["w"], spec, backend=L.TCGEN05 if bf else L.SIMT_F32)
        fb = self.T.MelSpectrogram.mel_scale.fb
        db = ops.mfcc_mel_db(spec, self.n_freq, fb.detach().contiguous(), self._bands(), gmax)
        return ops.mfcc_dct_resize(db, gmax, TOP_DB, self.T.dct_mat.detach().contiguous(), B, F_, self.out_dim)


class ExtractAndPredict(nn.Module):
    """extractor -> model, the composition ref:src/model/lightning_model.py:111-117 (Audio2FaceModel.forward) performs:
    `feature = self.feature_extractor(x).detach(); return self.model(feature, one_hot, template)`."""

    def __init__(self, feature_extractor: nn.Module, model: nn.Module):
        super().__init__()
        self.feature_extractor = feature_extractor
        self.model = model

    def set_precision(self, precision: str):
        self.feature_extractor.set_precision(precision)
        self.model.set_precision(precision)
        return self

    def forward(self, x, one_hot, template, **kwargs):
        if self.feature_extractor is None:
            return self.model(x, one_hot, template, **kwargs)
        feature = self.feature_extractor(x).detach()
        return self.model(feature, one_hot, template, **kwargs)

    def graphed(self, x, one_hot, template, **kwargs):
        from .modules import GraphedForward
        return GraphedForward(self, x, one_hot, template, **kwargs)


class ClipToVerts(nn.Module):
    """One clip -> one mesh per video frame: the reference's dataset + Lightning composition for the conv models
    (ref:src/dataset/vocaset.py:408-430 get_audio_fragment per frame, :64-69 normalize_audio;
    ref:src/model/lightning_model.py:111-117 extractor -> model).  forward(clip, one_hot, template): clip 1-D int16 or
    float32 at `sample_rate`, one_hot [F,n], template [F,5023,3] with F = len(clip) * fps // sample_rate frames."""

    def __init__(self, feature_extractor: nn.Module, model: nn.Module, *, fps: int = 60, sample_rate: int = 22000,
                 length: float = 0.52):
        super().__init__()
        self.net = ExtractAndPredict(feature_extractor, model)
        self.fps, self.sample_rate, self.length = int(fps), int(sample_rate), float(length)

    def set_precision(self, precision: str):
        self.net.set_precision(precision)
        return self

    def n_frames(self, n_samples: int) -> int:
        return int(n_samples) * self.fps // self.sample_rate

    def forward(self, clip, one_hot, template, **kwargs):
        win = audio_fragments(clip, self.n_frames(clip.numel()), fps=self.fps, sample_rate=self.sample_rate, length=self.length)
        return self.net(win, one_hot, template, **kwargs)

    def graphed(self, clip, one_hot, template, **kwargs):
        from .modules import GraphedForward
        return GraphedForward(self, clip, one_hot, template, **kwargs)


# ------------------------------------------------------------------------------------------------ audio preparation
def audio_fragments(audio: torch.Tensor, n_frames: int, *, fps: int = 60, sample_rate: int = 22000, length: float = 0.52,
                    shift: int = 0, first_frame: int = 0) -> torch.Tensor:
    """All per-frame windows of one clip in one launch: row f = ref:src/dataset/vocaset.py:408-430
    get_audio_fragment(audio, first_frame + f, fps=..., sample_rate=..., length=..., shift=...), int16 clips scaled by
    1/32768 (ref:vocaset.py:64-69).  audio: 1-D CUDA tensor (float32 or int16) -> [n_frames, 2*int(sample_rate*length/2)]."""
    if not audio.is_cuda:
        raise L.A2FError("the a2f_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")
    if audio.dim() != 1 or audio.dtype not in (torch.float32, torch.int16):
        raise L.A2FError("audio_fragments expects a 1-D float32 or int16 clip")
    audio = audio.contiguous()
    n_pad = int(sample_rate * length / 2)
    if n_frames > 0:
        # the reference returns None ("Audio is not long enough to get fragment", ref:vocaset.py:424-428) when a window
        # ends beyond the padded clip; the batched kernel would zero-fill there, so refuse instead of inventing data
        end = (int(first_frame) + int(n_frames) - 1) * int(sample_rate) // int(fps) + 2 * n_pad
        if end > (n_pad + int(shift)) + audio.numel() + 2 * n_pad or first_frame < 0:
            raise L.A2FError(f"audio_fragments: frame {first_frame + n_frames - 1} ends at padded sample {end}, beyond the clip "
                             "(the reference's get_audio_fragment returns None here)")
    out = torch.empty((n_frames, 2 * n_pad), dtype=torch.float32, device=audio.device)
    L.check(L.load().a2f_audio_fragments(audio.data_ptr(), L.I16 if audio.dtype == torch.int16 else L.F32, audio.numel(),
                                         int(first_frame), int(n_frames), int(sample_rate), int(fps), n_pad, int(shift),
                                         out.data_ptr(), torch.cuda.current_stream().cuda_stream), "a2f_audio_fragments")
    return out


_RESAMPLE_KERNELS = {}


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio.functional.functional._get_sinc_resample_kernel for a float32 waveform ("sinc_interp_hann"): the same
    float32 operation sequence, so the filter bank is the one torchaudio applies.  -> (kernel [new, 2*width+orig], width,
    orig, new) with the frequencies already divided by their gcd."""
    gcd = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // gcd, int(new_freq) // gcd
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float32)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float32)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    return kernels.reshape(new, -1).contiguous(), width, orig, new


def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """torchaudio.functional.resample(waveform, orig_freq, new_freq) with its default parameters, on (..., time) CUDA fp32."""
    if not waveform.is_cuda:
        raise L.A2FError("the a2f_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")
    if orig_freq <= 0 or new_freq <= 0:
        raise ValueError("Original frequency and desired frequecy should be positive")
    if orig_freq == new_freq:
        return waveform
    key = (int(orig_freq), int(new_freq), waveform.device)
    if key not in _RESAMPLE_KERNELS:
        k, width, orig, new = sinc_resample_kernel(orig_freq, new_freq)
        _RESAMPLE_KERNELS[key] = (k.to(waveform.device), width, orig, new)
    k, width, orig, new = _RESAMPLE_KERNELS[key]
    shape = waveform.shape
    x = waveform.reshape(-1, shape[-1]).contiguous().float()
    B, N = x.shape
    target = int(math.ceil(new * N / orig))
    out = torch.empty((B, target), dtype=torch.float32, device=x.device)
    L.check(L.load().a2f_resample_sinc(x.data_ptr(), B, N, orig, new, k.data_ptr(), k.shape[1], width, out.data_ptr(), target,
                                       torch.cuda.current_stream().cuda_stream), "a2f_resample_sinc")
    return out.view(shape[:-1] + (target,))


# ------------------------------------------------------------------------------------------------ wav2vec extractor
class Wav2VecExtractor(nn.Module):
    """Drop-in for ref:src/model/extractor.py:63-96: resample to 16 kHz, the HF processor's zero-mean / unit-variance
    normalisation (computed JOINTLY over the whole [B, N] tensor: a torch tensor is not "batched" for the HF feature
    extractor, so it is treated as one array -- ref:extractor.py:88-91, HF feature_extraction_wav2vec2.py), the
    wav2vec2-base encoder (`model.*` parameters, HF Wav2Vec2Model names), transpose and bilinear resize of the
    [768, frames] map to (out_dim, n_feature).  Output [B, out_dim, n_feature] fp32.

    The encoder is the sm_100a path of modules.Faceformer (conv stack as implicit GEMMs, tcgen05 GEMMs, flash
    attention); with frame_num = the conv stack's own length its interpolation is an identity, i.e. the plain HF model."""

    def __init__(self, sample_rate: int, n_feature: int, out_dim: int, *args, **kwargs):
        super().__init__()
        from .modules import Faceformer
        self.ori_sample_rate = sample_rate
        self.sample_rate = 16000
        self.out_dim = out_dim
        self.n_feature = n_feature
        ff = Faceformer(15069, 12)
        object.__setattr__(self, "_ff", ff)          # compute engine; only its encoder is part of this module's state
        self.model = ff.audio_encoder                # registered: state_dict keys model.* (HF Wav2Vec2Model layout)
        self.precision = "fp32"

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self._ff._apply(fn, *a, **k)                 # decoder / head tensors of the engine follow the device moves
        return self

    def set_precision(self, precision: str):
        self._ff.set_precision(precision)
        self.precision = precision
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise L.A2FError("the a2f_b200 modules run on CUDA (sm_100a) only; there is no CPU fallback")
        if x.dim() != 2:
            raise L.A2FError("Wav2VecExtractor expects (batch, time) audio")
        with torch.no_grad():
            x = resample(x.contiguous().float(), self.ori_sample_rate, self.sample_rate).contiguous()
            B, N = x.shape
            joint = ops.audio_stats(x.view(1, -1))                    # one (mean, rstd) for the whole tensor
            stats = joint.expand(B, 2).contiguous()
            n = (N - 10) // 5 + 1
            for k in (3, 3, 3, 3, 2, 2):
                n = (n - k) // 2 + 1
            if n < 1:
                raise L.A2FError("audio too short for one wav2vec2 frame")
            h = self._ff.encode(x, n, stats=stats)                    # [B*n, 768]
            if self.out_dim == 768:
                # ref:extractor.py:92-96 resizes only `if self.out_dim != x.shape[1]`, and that axis is the 768 channels
                # after the transpose: with out_dim == 768 the reference returns the transposed hidden states [B,768,n]
                if h.dtype != torch.float32:
                    h32 = torch.empty(h.shape, dtype=torch.float32, device=h.device)
                    L.check(L.load().a2f_cast_bf16_to_f32(h.data_ptr(), h32.data_ptr(), h.numel(),
                                                          torch.cuda.current_stream().cuda_stream), "a2f_cast_bf16_to_f32")
                    h = h32
                return ops.transpose_batched(h.view(B, n, 768))
            out = torch.empty((B, self.out_dim, self.n_feature), dtype=torch.float32, device=x.device)
            L.check(L.load().a2f_bilinear_cl(h.data_ptr(), L.BF16 if h.dtype == torch.bfloat16 else L.F32, B, n, 768, self.out_dim,
                                             self.n_feature, out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                    "a2f_bilinear_cl")
        return out
