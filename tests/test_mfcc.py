"""MFCC feature extractor (ref:src/model/extractor.py:10-60).

CPU: the oracle restatement (oracle/ref_mfcc.py) against the fixture produced by the LIVE reference MFCCExtractor
(tests/golden/make_golden_mfcc.py), and the host logic of the drop-in module (buffers, state_dict keys, band supports).
GPU (-m gpu): the CUDA path (frame gather -> DFT GEMM -> mel/dB/global max -> clamp/DCT/resize) through the C-ABI against
the oracle on the same inputs and against the fixture.

Tolerances (absolute, on coefficients of magnitude up to ~570, plus 2e-6 relative to the largest coefficient -- c0 sums
128 same-sign dB values, so the reference's own fp32 matmul carries that much summation-order noise; the kernel
accumulates the DCT in fp64): oracle vs live reference 1e-4 (fp32 FFT rounding); fp32 path 1e-3 (the DFT is an fp32
GEMM: measured 2e-4 on B200); tensor-core path 1e-2 (bf16x3 split: measured 2e-3).  A plain bf16 DFT would be off by
0.3 -- that is why the split is used.
"""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_mfcc as omf

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mfcc.npz")


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


@pytest.mark.parametrize("name", ["voca", "audio2mesh"])
def test_oracle_matches_reference_fixture(name):
    z = np.load(G)
    sr, nf, od, win, hop, nfft = omf.CONFIGS[name]
    sd = omf.make_buffers(sr, nf, win, nfft)
    x = oin.speech_like_windows(int(z["batch"]), seed=int(z["seed_in"]))
    y = omf.mfcc_forward(sd, x, od, win, hop, nfft)
    assert tuple(y.shape) == z[f"{name}_out"].shape
    np.testing.assert_allclose(y.numpy(), z[f"{name}_out"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("name", ["voca", "audio2mesh"])
def test_module_buffers_and_keys_match_reference(name):
    """The drop-in module's persistent buffers (own formulas) equal the live reference's; same state_dict keys."""
    from a2f_b200 import features
    z = np.load(G)
    sr, nf, od, win, hop, nfft = omf.CONFIGS[name]
    m = features.MFCCExtractor(sr, nf, od, win, hop, nfft)
    sd = m.state_dict()
    assert list(sd.keys()) == ["T.dct_mat", "T.MelSpectrogram.spectrogram.window", "T.MelSpectrogram.mel_scale.fb"]
    np.testing.assert_array_equal(sd["T.dct_mat"].numpy(), z[f"{name}_dct"])
    np.testing.assert_array_equal(sd["T.MelSpectrogram.spectrogram.window"].numpy(), z[f"{name}_window"])
    np.testing.assert_allclose(sd["T.MelSpectrogram.mel_scale.fb"].double().sum(0).numpy(), z[f"{name}_fb_sum"], rtol=0, atol=0)
    ref = omf.make_buffers(sr, nf, win, nfft)
    for k in sd:
        assert torch.equal(sd[k], ref[k]), k
    m.load_state_dict(ref, strict=True)
    assert m.hop_length == win // 2 and m.n_fft == 1024 and m.n_freq == 513


def test_band_supports_cover_the_filterbank():
    from a2f_b200 import features
    m = features.MFCCExtractor(22000, 16, 29, 790, None, 1024)
    band = m._bands()
    fb = m.T.MelSpectrogram.mel_scale.fb
    for k in range(fb.shape[1]):
        f0, f1 = int(band[k, 0]), int(band[k, 1])
        col = fb[:, k]
        assert float(col[:f0].abs().sum()) == 0.0 and float(col[f1:].abs().sum()) == 0.0
    assert int((band[:, 1] - band[:, 0]).sum()) < 2 * 513 + 128        # triangular bands: ~2 non-zeros per frequency bin


def test_cpu_tensors_are_refused():
    from a2f_b200 import features, A2FError
    m = features.MFCCExtractor(22000, 16, 29, 790, None, 1024)
    with pytest.raises(A2FError):
        m(torch.zeros(2, 11440))


# ---------------------------------------------------------------------------------------------------------------- GPU
def _close(got, want, tol):
    err = float((got - want).abs().max())
    assert err < tol + 2e-6 * float(want.abs().max()), err


def _run(dev, name, x, precision, cfg=None):
    from a2f_b200 import features
    sr, nf, od, win, hop, nfft = cfg or omf.CONFIGS[name]
    m = features.MFCCExtractor(sr, nf, od, win, hop, nfft).to(dev).set_precision(precision)
    y = m(x.to(dev))
    torch.cuda.synchronize()
    return y.cpu()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["voca", "audio2mesh"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 1e-2)])
def test_mfcc_gpu_matches_oracle_and_fixture(a2f_lib, dev, name, precision, tol):
    z = np.load(G)
    sr, nf, od, win, hop, nfft = omf.CONFIGS[name]
    x = oin.speech_like_windows(int(z["batch"]), seed=int(z["seed_in"]))
    want = omf.mfcc_forward(omf.make_buffers(sr, nf, win, nfft), x, od, win, hop, nfft)
    got = _run(dev, name, x, precision)
    assert tuple(got.shape) == tuple(want.shape)
    _close(got, want, tol)
    _close(got, torch.from_numpy(z[f"{name}_out"]), tol)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,cfg", [
    (1, 11440, (22000, 16, 29, 790, None, 1024)),        # one window
    (3, 5000, (22000, 13, 20, 400, 160, 512)),           # explicit hop, n_fft 512, resize 32 -> 20
    (2, 11440, (22000, 32, 53, 440, None, 1024)),        # out_dim == frames: no resize
    (5, 16000, (16000, 40, 64, 1024, 256, None)),        # n_fft defaults to win_length, upsampling resize 63 -> 64
    (70, 11440, (22000, 16, 29, 790, None, 1024)),       # more rows than one tile
])
def test_mfcc_gpu_geometries(a2f_lib, dev, B, N, cfg):
    sr, nf, od, win, hop, nfft = cfg
    x = oin.speech_like_windows(B, n_samples=N, sample_rate=sr, seed=5)
    want = omf.mfcc_forward(omf.make_buffers(sr, nf, win, nfft or win), x, od, win, hop, nfft)
    for precision, tol in (("fp32", 1e-3), ("bf16", 1e-2)):
        got = _run(dev, None, x, precision, cfg)
        assert tuple(got.shape) == (B, od, nf)
        _close(got, want, tol)


@pytest.mark.gpu
def test_mfcc_gpu_silence_and_global_cutoff(a2f_lib, dev):
    """All-zero audio sits on the 1e-10 clamp (-100 dB everywhere); one loud window among silent ones raises the
    batch-global cut-off for ALL windows (torchaudio's 3-D behaviour), which a per-window cut-off would get wrong."""
    cfg = omf.CONFIGS["voca"]
    sd = omf.make_buffers(cfg[0], cfg[1], cfg[3], cfg[5])
    x = torch.zeros(3, 11440)
    got = _run(dev, "voca", x, "fp32")
    want = omf.mfcc_forward(sd, x, cfg[2], cfg[3], cfg[4], cfg[5])
    _close(got, want, 1e-3)
    x = 1e-4 * oin.speech_like_windows(4, seed=9)
    x[0] *= 1e4
    got = _run(dev, "voca", x, "fp32")
    want = omf.mfcc_forward(sd, x, cfg[2], cfg[3], cfg[4], cfg[5])
    _close(got, want, 1e-3)
    alone = omf.mfcc_forward(sd, x[1:2], cfg[2], cfg[3], cfg[4], cfg[5])
    assert float((want[1:2] - alone).abs().max()) > 1.0          # the cut-off really is batch-global in the reference


@pytest.mark.gpu
def test_mfcc_then_voca_matches_oracle_chain(a2f_lib, dev):
    """ref:src/model/lightning_model.py:111-117: feature = extractor(x).detach(); model(feature, one_hot, template)."""
    from a2f_b200 import features, modules
    from oracle import ref_models as orm, weights as ow
    cfg = omf.CONFIGS["voca"]
    B = 8
    x = oin.speech_like_windows(B, seed=3)
    oh, tp = oin.one_hot(B, 12, 3), oin.batch_templates(B, 3)
    sd = ow.make_state_dict("voca", seed=11)
    feat = omf.mfcc_forward(omf.make_buffers(cfg[0], cfg[1], cfg[3], cfg[5]), x, cfg[2], cfg[3], cfg[4], cfg[5])
    want = orm.voca_forward(sd, feat, oh, tp)
    ext = features.MFCCExtractor(*cfg).to(dev)
    model = modules.Voca(15069, 12).to(dev)
    model.load_state_dict(sd, strict=True)
    got = model(ext(x.to(dev)).detach(), oh.to(dev), tp.to(dev)).cpu()
    scale = float((want - tp).abs().max())
    assert float((got - want).abs().max()) < 1e-5 * max(1.0, scale)
