"""VOCA drop-in module on the GPU vs the oracle (which is pinned to the live reference by tests/golden/voca.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_models as orm, weights as ow

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(dev, seed):
    from a2f_b200 import modules
    sd = ow.make_state_dict("voca", seed=seed)
    m = modules.Voca(15069, 12).to(dev)
    m.load_state_dict(sd, strict=True)
    return m, sd


@pytest.mark.parametrize("B", [1, 6, 64, 348])
def test_voca_fp32_matches_oracle(a2f_lib, dev, B):
    m, sd = _model(dev, 11)
    x, oh, tp = oin.voca_features(B, 1), oin.one_hot(B, 12, 1), oin.batch_templates(B, 1)
    want = orm.voca_forward(sd, x, oh, tp)
    with torch.no_grad():
        got = m.set_precision("fp32")(x.to(dev), oh.to(dev), tp.to(dev)).cpu()
    assert got.shape == (B, 5023, 3)
    assert float((got - want).abs().max()) < 1e-5          # north_star: 1e-5 m on the fp32 path


def test_voca_fp32_matches_golden_fixture(a2f_lib, dev):
    z = np.load(os.path.join(G, "voca.npz"))
    m, _ = _model(dev, int(z["seed_w"]))
    B, s = int(z["batch"]), int(z["seed_in"])
    with torch.no_grad():
        got = m.set_precision("fp32")(oin.voca_features(B, s).to(dev), oin.one_hot(B, 12, s).to(dev),
                                      oin.batch_templates(B, s).to(dev)).cpu()
    np.testing.assert_allclose(got.reshape(-1)[:: int(z["step"])].numpy(), z["out"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("B", [6, 130])
def test_voca_bf16_matches_oracle(a2f_lib, dev, B):
    m, sd = _model(dev, 11)
    x, oh, tp = oin.voca_features(B, 2), oin.one_hot(B, 12, 2), oin.batch_templates(B, 2)
    want = orm.voca_forward(sd, x, oh, tp)
    with torch.no_grad():
        got = m.set_precision("bf16")(x.to(dev), oh.to(dev), tp.to(dev)).cpu()
    # raw model units (|offset| ~ 0.6 at this init).  The tensor-core head runs on an error-compensated bf16 split
    # and the trunk is fp32, so the "bf16" VOCA path is far inside the 5e-4 bar.
    assert float((got - want).abs().max()) < 5e-5


def test_voca_shared_template_and_linearity(a2f_lib, dev):
    """Size-independent property at a large batch: out - template does not depend on the template."""
    m, _ = _model(dev, 11)
    B = 4096
    x, oh = oin.voca_features(B, 3).to(dev), oin.one_hot(B, 12, 3).to(dev)
    t0 = torch.zeros(B, 5023, 3, device=dev)
    t1 = oin.flame_like_template(0).to(dev)[None].expand(B, -1, -1).contiguous()
    with torch.no_grad():
        y0 = m.set_precision("fp32")(x, oh, t0)
        y1 = m(x, oh, t1)
    assert float(((y1 - t1) - y0).abs().max()) < 1e-6
