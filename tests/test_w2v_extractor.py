"""Wav2VecExtractor (ref:src/model/extractor.py:63-96): resample -> joint normalisation ->
wav2vec2-base encoder -> transpose + bilinear resize to (out_dim, n_feature).

CPU: the oracle (oracle/ref_w2v_extractor.py) against the fixture of the LIVE reference class
(tests/golden/make_golden_w2v_extractor.py); state_dict key parity of the drop-in module.
GPU (-m gpu): the CUDA path against the oracle and the fixture -- fp32 path 2e-4 absolute on features of magnitude ~2
(12 encoder layers of fp32 GEMMs), bf16 tensor-core path 6e-2 (bf16 activations through 12 layers).
"""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_w2v_extractor as owx, weights as ow

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "w2v_extractor.npz")


def _sd(seed):
    ff = ow.make_state_dict("faceformer", seed)
    return {"model." + k[len("audio_encoder."):]: v for k, v in ff.items() if k.startswith("audio_encoder.")}


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def test_oracle_matches_reference_fixture():
    z = np.load(G)
    x = oin.speech_like_windows(int(z["batch"]), seed=int(z["seed_in"]))
    y = owx.w2v_extractor_forward(_sd(int(z["seed_w"])), x, 22000, 32, 52)
    assert tuple(y.shape) == z["out"].shape == (3, 52, 32)
    np.testing.assert_allclose(y.numpy(), z["out"], rtol=0, atol=5e-5)


def test_module_state_dict_keys_match_reference():
    from a2f_b200 import features, get_extractor
    z = np.load(G)
    m = get_extractor("wav2vec")(22000, 32, 52)
    assert isinstance(m, features.Wav2VecExtractor)
    sd = _sd(13)
    assert len(m.state_dict()) == int(z["n_keys"]) == len(sd)
    m.load_state_dict(sd, strict=True)                      # same names / shapes as the reference's `model.*`


def test_joint_normalisation_and_groupnorm():
    """The reference normalises the whole [B, N] tensor with ONE mean / variance (HF treats a tensor as a single array),
    so the oracle must too; wav2vec2's first layer is followed by a per-(utterance, channel) GroupNorm, which removes
    almost all of the difference again -- both facts are pinned here."""
    sd = _sd(13)
    x = oin.speech_like_windows(2, n_samples=6000, seed=5)
    x[1] *= 4.0
    both = owx.w2v_extractor_forward(sd, x, 22000, 8, 16)
    alone = owx.w2v_extractor_forward(sd, x[:1], 22000, 8, 16)
    d = float((both[:1] - alone).abs().max())
    assert 0.0 < d < 5e-3
    xr = owx.ora.resample(x, 22000, 16000)
    joint = (xr - xr.mean()) / torch.sqrt(xr.var(unbiased=False) + 1e-7)
    per = (xr - xr.mean(1, keepdim=True)) / torch.sqrt(xr.var(1, unbiased=False, keepdim=True) + 1e-7)
    assert float((joint - per).abs().max()) > 0.1            # the normalised audio itself differs a lot
    assert torch.equal(owx.ora.resample(x, 16000, 16000), x)  # equal rates: untouched, like torchaudio


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 6e-2)])
def test_w2v_extractor_gpu_matches_oracle_and_fixture(a2f_lib, dev, precision, tol):
    from a2f_b200 import features
    z = np.load(G)
    sd = _sd(int(z["seed_w"]))
    x = oin.speech_like_windows(int(z["batch"]), seed=int(z["seed_in"]))
    want = owx.w2v_extractor_forward(sd, x, 22000, 32, 52)
    m = features.Wav2VecExtractor(22000, 32, 52)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).set_precision(precision)
    got = m(x.to(dev)).cpu()
    assert tuple(got.shape) == (3, 52, 32)
    err = float((got - want).abs().max())
    print(f"wav2vec extractor {precision}: max|err| {err:.3e} on features of magnitude {float(want.abs().max()):.2f}")
    assert err < tol
    assert float((got.numpy() - z["out"]).__abs__().max()) < tol


@pytest.mark.gpu
def test_w2v_extractor_gpu_other_geometry(a2f_lib, dev):
    """16 kHz input (no resampling), a batch large enough for the CTA-pair GEMMs, upsampling resize along time."""
    from a2f_b200 import features
    sd = _sd(13)
    x = oin.speech_like_windows(12, n_samples=8000, sample_rate=16000, seed=6)
    want = owx.w2v_extractor_forward(sd, x, 16000, 40, 64)
    m = features.Wav2VecExtractor(16000, 40, 64)
    m.load_state_dict(sd, strict=True)
    got = m.to(dev)(x.to(dev)).cpu()
    assert tuple(got.shape) == (12, 64, 40)
    assert float((got - want).abs().max()) < 2e-4


@pytest.mark.gpu
def test_w2v_extractor_gpu_out_dim_768_is_not_resized(a2f_lib, dev):
    """ref:src/model/extractor.py:92-96 resizes only `if self.out_dim != x.shape[1]` (the 768 channels after the transpose):
    with out_dim == 768 the reference returns the transposed hidden states [B, 768, frames]."""
    from a2f_b200 import features
    sd = _sd(13)
    x = oin.speech_like_windows(2, n_samples=8000, sample_rate=16000, seed=8)
    want = owx.w2v_extractor_forward(sd, x, 16000, 32, 768)
    assert tuple(want.shape) == (2, 768, 24)
    m = features.Wav2VecExtractor(16000, 32, 768)
    m.load_state_dict(sd, strict=True)
    got = m.to(dev)(x.to(dev)).cpu()
    assert tuple(got.shape) == tuple(want.shape)
    assert float((got - want).abs().max()) < 2e-4
