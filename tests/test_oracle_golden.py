"""The oracle restatement (oracle/ref_models.py) against the golden fixtures produced by the LIVE reference modules
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_models as orm, weights as ow

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def _sub(t, step):
    return t.reshape(-1)[::step].numpy()


def test_voca_oracle_matches_reference_fixture():
    z = np.load(os.path.join(G, "voca.npz"))
    sd = ow.make_state_dict("voca", seed=int(z["seed_w"]))
    B, s = int(z["batch"]), int(z["seed_in"])
    y = orm.voca_forward(sd, oin.voca_features(B, s), oin.one_hot(B, 12, s), oin.batch_templates(B, s))
    np.testing.assert_allclose(_sub(y, int(z["step"])), z["out"], rtol=0, atol=2e-6)


def test_audio2mesh_oracle_matches_reference_fixture():
    z = np.load(os.path.join(G, "audio2mesh.npz"))
    sd = ow.make_state_dict("audio2mesh", seed=int(z["seed_w"]))
    B, s = int(z["batch"]), int(z["seed_in"])
    x, oh, tp = oin.a2m_features(B, s), oin.one_hot(B, 12, s), oin.batch_templates(B, s)
    np.testing.assert_allclose(_sub(orm.audio2mesh_forward(sd, x, oh, tp), int(z["step"])), z["out_eval"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(_sub(orm.audio2mesh_forward(sd, x, oh, tp, train_bn=True), int(z["step"])),
                               z["out_train"], rtol=0, atol=5e-6)


def test_loss_oracle_matches_reference_fixture():
    z = np.load(os.path.join(G, "loss.npz"))
    rows = 6
    tp = oin.batch_templates(rows, 3)
    lv = orm.voca_loss(oin.gt_like((rows, 5023, 3), tp, 31), oin.gt_like((rows, 5023, 3), tp, 32))
    got = np.array([float(lv["loss"]), float(lv["rec_loss"]), float(lv["vel_loss"])])
    np.testing.assert_allclose(got, z["voca"], rtol=1e-6)
    lf = orm.faceformer_loss(oin.gt_like((1, 7, 5023, 3), tp[:1, None], 33), oin.gt_like((1, 7, 5023, 3), tp[:1, None], 34))
    got = np.array([float(lf["loss"]), float(lf["rec_loss"]), float(lf["vel_loss"])])
    np.testing.assert_allclose(got, z["faceformer"], rtol=1e-6)
    # ref lightning_model.py:119-125: err == rec_loss / 3
    p, g = oin.gt_like((rows, 5023, 3), tp, 31), oin.gt_like((rows, 5023, 3), tp, 32)
    assert abs(float(orm.mse_error(p, g)) - float(lv["rec_loss"]) / 3) < 1e-9


@pytest.mark.parametrize("tag", ["a", "b"])
def test_faceformer_oracle_matches_reference_fixture(tag):
    z = np.load(os.path.join(G, "faceformer.npz"))
    sd = ow.make_state_dict("faceformer", seed=int(z["seed_w"]))
    n, s = int(z[f"n_{tag}"]), int(z[f"seed_{tag}"])
    audio, oh, tp = oin.audio(1, n, s), oin.one_hot(1, 12, s), oin.batch_templates(1, s, scale=100.0)
    y, parts = orm.faceformer_forward(sd, audio, oh, tp, return_parts=True)
    assert y.shape == (1, n * 60 // 16000, 5023, 3)
    np.testing.assert_allclose(_sub(parts["encoder"], int(z["step_enc"])), z[f"enc_{tag}"], rtol=0, atol=1e-5)
    # centimetre units (template x100, ref lightning_model.py:145-148): 1e-5 here = 1e-7 m
    np.testing.assert_allclose(_sub(y, int(z["step_out"])), z[f"out_{tag}"], rtol=0, atol=1e-5)


def test_state_dict_key_sets():
    assert len(ow.voca_shapes()) == 16
    assert len(ow.audio2mesh_shapes()) == 78
    assert len(ow.faceformer_shapes()) == 237
    n = sum(int(np.prod(s)) for k, s in ow.faceformer_shapes().items() if k != "PPE.pe")
    assert n == 96_415_901


def test_biased_mask_closed_form_and_ppe():
    m = orm.init_biased_mask(4, 600, 60)
    assert m.shape == (4, 600, 600)
    assert m[0, 0, 0] == 0 and m[0, 59, 0] == 0 and m[0, 60, 0] == -0.25 and m[3, 120, 0] == -2 * 2.0 ** -8
    assert torch.isinf(m[0, 0, 1]) and m[0, 0, 1] < 0
    big = orm.init_biased_mask(4, 1200, 60)
    assert torch.equal(big[:, :600, :600], m)
    pe = ow.ppe_table()
    assert pe.shape == (1, 660, 64)
    assert torch.equal(pe[0, :60], pe[0, 60:120])


def test_decode_incremental_equals_prefix_recompute():
    """The KV-cached O(T) form the CUDA decoder uses equals the reference's O(T^2) prefix recomputation."""
    sd = ow.make_state_dict("faceformer", seed=13)
    T = 12
    g = torch.Generator().manual_seed(3)
    mem = torch.randn(1, T, 64, generator=g)
    oh = oin.one_hot(1, 12, 9)
    full = orm.faceformer_decode(sd, mem, oh, T)
    for t in (1, 5, T):
        part = orm.faceformer_decode(sd, mem, oh, t)      # memory longer than the prefix: diagonal mask only
        assert float((part[0, t - 1] - full[0, t - 1]).abs().max()) < 2e-5
