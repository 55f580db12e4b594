"""Audio preparation kernels: per-frame window slicing (ref:src/dataset/vocaset.py:408-430) and
22 kHz -> 16 kHz sinc resampling (torchaudio.functional.resample at ref:vocaset.py:279-283, ref:extractor.py:88).

CPU: the oracle restatement (oracle/ref_audio.py) against the fixture made from the reference's own function / the live
torchaudio (tests/golden/make_golden_audio_prep.py); the drop-in filter bank against the oracle's.
GPU (-m gpu): the kernels through the C-ABI against the oracle (fragments: bit-exact -- it is a gather and an exact
scale by 2^-15; resample: 2e-6 absolute on |x| <= 1, fp32 summation order) and against the fixture.
"""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_audio as ora

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "audio_prep.npz")


def _clip():
    z = np.load(G)
    clip = oin.pcm16_clip(seed=int(z["seed"]))
    assert int(np.abs(clip.astype(np.int64)).sum()) == int(z["clip_abs_sum"])     # same synthetic clip on every machine
    return z, clip


def _resample_cases(clip):
    x = torch.from_numpy((clip / 32768).astype(np.float32))
    xb = torch.stack([x[:50000], x[30000:80000] * 0.5, torch.flip(x[:50000], [0])])
    return (("clip", x, (22000, 16000)), ("batch", xb, (22000, 16000)), ("up", xb[:, :7001], (16000, 22050)))


def test_oracle_fragments_match_reference_fixture():
    z, clip = _clip()
    n = int(z["n_frames"])
    assert n == 348
    for shift in (0, 137, -200):
        got = ora.fragments(clip, n, shift=shift).numpy()
        assert got.shape == (348, 11440)
        np.testing.assert_array_equal(got[::29, ::97], z[f"frag_{shift}_sub"])
        np.testing.assert_allclose(got.astype(np.float64).sum(1), z[f"frag_{shift}_rowsum"], rtol=0, atol=1e-9)


def test_oracle_resample_matches_torchaudio_fixture():
    z, clip = _clip()
    for name, wav, (fo, fn) in _resample_cases(clip):
        got = ora.resample(wav, fo, fn)
        assert got.shape[-1] == int(z[f"rs_{name}_len"])
        np.testing.assert_allclose(got.reshape(-1)[::53].numpy(), z[f"rs_{name}_sub"], rtol=0, atol=1e-6)


def test_fragment_past_the_clip_is_refused_like_the_reference():
    clip = oin.pcm16_clip(n_samples=22000, seed=1)
    assert ora.get_audio_fragment(clip, 100, 60, 22000, 0.52, 0) is None      # ref:vocaset.py:425-428


def test_drop_in_filter_bank_equals_the_oracles():
    from a2f_b200 import features
    k, width, orig, new = features.sinc_resample_kernel(22000, 16000)
    assert (orig, new, width) == (11, 8, 9) and tuple(k.shape) == (8, 29)
    x = torch.zeros(1, 64)
    x[0, 20] = 1.0                                                             # impulse response = the filter taps
    y = ora.resample(x, 22000, 16000)
    for p in range(8):
        i = 2
        taps = torch.tensor([float(k[p, kk]) if (i * 11 + kk - 9) == 20 else 0.0 for kk in range(29)]).sum()
        assert abs(float(y[0, i * 8 + p]) - float(taps)) < 1e-7


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shift", [0, 137, -200])
def test_fragments_gpu_bit_exact(a2f_lib, dev, shift):
    from a2f_b200 import features
    z, clip = _clip()
    n = int(z["n_frames"])
    want = ora.fragments(clip, n, shift=shift)
    got = features.audio_fragments(torch.from_numpy(clip).to(dev), n, shift=shift).cpu()
    assert torch.equal(got, want)
    np.testing.assert_array_equal(got.numpy()[::29, ::97], z[f"frag_{shift}_sub"])
    gotf = features.audio_fragments(torch.from_numpy((clip / 32768).astype(np.float32)).to(dev), 40, shift=shift, first_frame=300).cpu()
    assert torch.equal(gotf, want[300:340])                                    # fp32 input, frame sub-range


@pytest.mark.gpu
def test_fragments_gpu_refuses_a_frame_past_the_clip(a2f_lib, dev):
    from a2f_b200 import features, A2FError
    clip = torch.from_numpy(oin.pcm16_clip(n_samples=22000, seed=1)).to(dev)
    features.audio_fragments(clip, 60)
    with pytest.raises(A2FError):
        features.audio_fragments(clip, 101)


@pytest.mark.gpu
def test_resample_gpu_matches_oracle_and_fixture(a2f_lib, dev):
    from a2f_b200 import features
    z, clip = _clip()
    for name, wav, (fo, fn) in _resample_cases(clip):
        want = ora.resample(wav, fo, fn)
        got = features.resample(wav.to(dev), fo, fn).cpu()
        assert got.shape == want.shape
        assert float((got - want).abs().max()) < 2e-6
        np.testing.assert_allclose(got.reshape(-1)[::53].numpy(), z[f"rs_{name}_sub"], rtol=0, atol=2e-6)
    same = torch.randn(2, 100, device=dev)
    assert features.resample(same, 16000, 16000) is same


@pytest.mark.gpu
def test_clip_to_vertices_chain(a2f_lib, dev):
    """SURVEY.md 8(d) config 1 end to end on the GPU: int16 clip -> 348 windows -> MFCC -> VOCA -> [348, 5023, 3], against
    the same chain on the oracle."""
    from a2f_b200 import features, modules
    from oracle import ref_mfcc as omf, ref_models as orm, weights as ow
    z, clip = _clip()
    n = int(z["n_frames"])
    cfg = omf.CONFIGS["voca"]
    sd = ow.make_state_dict("voca", seed=11)
    oh = torch.eye(12)[:1].repeat(n, 1)
    tp = oin.flame_like_template(0)[None].repeat(n, 1, 1)
    feat = omf.mfcc_forward(omf.make_buffers(cfg[0], cfg[1], cfg[3], cfg[5]), ora.fragments(clip, n), cfg[2], cfg[3], cfg[4], cfg[5])
    want = orm.voca_forward(sd, feat, oh, tp)
    ext = features.MFCCExtractor(*cfg).to(dev)
    model = modules.Voca(15069, 12).to(dev)
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        win = features.audio_fragments(torch.from_numpy(clip).to(dev), n)
        got = model(ext(win), oh.to(dev), tp.to(dev)).cpu()
    assert tuple(got.shape) == (348, 5023, 3)
    scale = float((want - tp).abs().max())
    assert float((got - want).abs().max()) < 1e-5 * max(1.0, scale)
