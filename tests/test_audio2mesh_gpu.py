"""Audio2Mesh drop-in module on the GPU vs the oracle / the live-reference golden fixture (eval-mode BatchNorm)."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_models as orm, weights as ow

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(dev, seed):
    from a2f_b200 import modules
    sd = ow.make_state_dict("audio2mesh", seed=seed)
    m = modules.Audio2Mesh(15069, 12).to(dev)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def test_a2m_matches_golden_fixture(a2f_lib, dev):
    z = np.load(os.path.join(G, "audio2mesh.npz"))
    m, _ = _model(dev, int(z["seed_w"]))
    B, s = int(z["batch"]), int(z["seed_in"])
    with torch.no_grad():
        got = m.set_precision("fp32")(oin.a2m_features(B, s).to(dev), oin.one_hot(B, 12, s).to(dev),
                                      oin.batch_templates(B, s).to(dev)).cpu()
    np.testing.assert_allclose(got.reshape(-1)[:: int(z["step"])].numpy(), z["out_eval"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("B", [1, 64, 130])
def test_a2m_fp32_and_bf16_match_oracle(a2f_lib, dev, B):
    m, sd = _model(dev, 12)
    x, oh, tp = oin.a2m_features(B, 3), oin.one_hot(B, 12, 3), oin.batch_templates(B, 3)
    want = orm.audio2mesh_forward(sd, x, oh, tp)
    with torch.no_grad():
        got = m.set_precision("fp32")(x.to(dev), oh.to(dev), tp.to(dev)).cpu()
        got16 = m.set_precision("bf16")(x.to(dev), oh.to(dev), tp.to(dev)).cpu()
    assert got.shape == (B, 5023, 3)
    assert float((got - want).abs().max()) < 1e-5
    # precision "bf16": conv stack AND head on tcgen05 with the error-compensated bf16x3 split (im2col hi|lo|hi x weights
    # hi|hi|lo); the oracle's random-init offsets are O(4 m), so 2e-4 is 5e-5 relative (north-star budget 5e-4 m; a plain
    # bf16 trunk would be off by 4e-2 here)
    assert float((got16 - want).abs().max()) < 2e-4


def test_a2m_eval_mode_autograd_is_refused_loudly(a2f_lib, dev):
    """Gradients with frozen (eval-mode) BatchNorm are not built: the module must say so instead of silently
    returning a tensor without a graph.  (Train-mode forward/backward: tests/test_conv_train_gpu.py.)"""
    import a2f_b200
    m, _ = _model(dev, 12)
    m.eval()
    with pytest.raises(a2f_b200.A2FError):
        m(oin.a2m_features(2, 3).to(dev), oin.one_hot(2, 12, 3).to(dev), oin.batch_templates(2, 3).to(dev))
