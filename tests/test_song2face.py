"""Song2Face (ref:src/model/song2face.py; registry entry "song2face"): conv stack -> two LSTMs over
the channel axis -> bilinear resize -> regression convs -> output MLP + vertex head.

CPU: the oracle (oracle/ref_models.song2face_forward, LSTM written out) against the fixture of the LIVE reference module
(tests/golden/make_golden_song2face.py); state_dict key order of the drop-in module.
GPU (-m gpu): the CUDA path against the oracle and the fixture: fp32 path 1e-5 (offsets of magnitude ~1), tensor-core path
(bf16x3 split GEMMs, fp32 LSTM recurrence) 2e-4.
"""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_models as orm, weights as ow

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "song2face.npz")


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def test_oracle_matches_reference_fixture():
    z = np.load(G)
    sd = ow.make_state_dict("song2face", seed=int(z["seed_w"]))
    assert len(sd) == int(z["n_keys"])
    B, s = int(z["batch"]), int(z["seed_in"])
    y = orm.song2face_forward(sd, oin.a2m_features(B, s), oin.one_hot(B, 12, s), oin.batch_templates(B, s))
    np.testing.assert_allclose(y.reshape(-1)[:: int(z["step"])].numpy(), z["out"], rtol=0, atol=5e-6)


def test_lstm_restatement_matches_torch_lstm():
    g = torch.Generator().manual_seed(3)
    lstm = torch.nn.LSTM(24, 32, 1, batch_first=True)
    x = torch.randn(3, 17, 24, generator=g)
    want, _ = lstm(x)
    got = orm.lstm_forward(x, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0)
    assert float((got - want).abs().max()) < 1e-6


def test_registry_and_state_dict_keys():
    import a2f_b200
    cls = a2f_b200.get_model("song2face")
    m = cls(15069, 12)
    sd = ow.make_state_dict("song2face", seed=14)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd, strict=True)
    with pytest.raises(a2f_b200.A2FError):
        m(torch.zeros(1, 52, 32), torch.zeros(1, 12), torch.zeros(1, 5023, 3))       # CPU tensors: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 5, 9])
def test_song2face_gpu_matches_oracle(a2f_lib, dev, B):
    from a2f_b200 import modules
    sd = ow.make_state_dict("song2face", seed=14)
    x, oh, tp = oin.a2m_features(B, 41), oin.one_hot(B, 12, 41), oin.batch_templates(B, 41)
    want = orm.song2face_forward(sd, x, oh, tp)
    m = modules.Song2Face(15069, 12).to(dev).eval()
    m.load_state_dict(sd, strict=True)
    got32 = m.set_precision("fp32")(x.to(dev), oh.to(dev), tp.to(dev)).cpu()
    got16 = m.set_precision("bf16")(x.to(dev), oh.to(dev), tp.to(dev)).cpu()
    e32, e16 = float((got32 - want).abs().max()), float((got16 - want).abs().max())
    print(f"song2face B={B}: fp32 max|err| {e32:.3e}, bf16x3 max|err| {e16:.3e}, |offset|max {float((want - tp).abs().max()):.3f}")
    assert got32.shape == (B, 5023, 3)
    assert e32 < 1e-5
    assert e16 < 2e-4
    if B == 5:
        z = np.load(G)
        np.testing.assert_allclose(got32.reshape(-1)[:: int(z["step"])].numpy(), z["out"], rtol=0, atol=1e-5)


@pytest.mark.gpu
def test_lstm_recurrence_kernel(a2f_lib, dev):
    """a2f_lstm_recurrence against torch.nn.LSTM on the CPU (hidden 256, ragged batch vs the 4-per-CTA tiling)."""
    from a2f_b200 import ops
    g = torch.Generator().manual_seed(5)
    lstm = torch.nn.LSTM(40, 256, 1, batch_first=True)
    for B, T in ((1, 7), (6, 33)):
        x = torch.randn(B, T, 40, generator=g)
        want, _ = lstm(x)
        xp = torch.nn.functional.linear(x, lstm.weight_ih_l0, lstm.bias_ih_l0 + lstm.bias_hh_l0).contiguous()
        got = ops.lstm_recurrence(xp.to(dev), lstm.weight_hh_l0.detach().t().contiguous().to(dev), B, T, 256).cpu()
        assert float((got - want).abs().max()) < 2e-6
