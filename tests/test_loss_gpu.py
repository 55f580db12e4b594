"""Fused loss kernels vs the oracle (ref:src/loss/loss.py) incl. gradients via autograd on the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_models as orm

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_voca_loss_matches_golden_and_oracle(a2f_lib, dev):
    from a2f_b200 import modules
    z = np.load(os.path.join(G, "loss.npz"))
    rows = 6
    tp = oin.batch_templates(rows, 3)
    pred, gt = oin.gt_like((rows, 5023, 3), tp, 31), oin.gt_like((rows, 5023, 3), tp, 32)
    out = modules.VocaLoss()(pred.to(dev), gt.to(dev))
    got = np.array([float(out["loss"]), float(out["rec_loss"]), float(out["vel_loss"])])
    np.testing.assert_allclose(got, z["voca"], rtol=1e-4)            # north_star: losses to 1e-4 relative
    err = float(modules.mse_error(pred.to(dev), gt.to(dev)))
    assert abs(err - float(orm.mse_error(pred, gt))) <= 1e-4 * abs(err)


def test_faceformer_loss_odd_frames(a2f_lib, dev):
    from a2f_b200 import modules
    z = np.load(os.path.join(G, "loss.npz"))
    tp = oin.batch_templates(6, 3)
    pred = oin.gt_like((1, 7, 5023, 3), tp[:1, None], 33)
    gt = oin.gt_like((1, 7, 5023, 3), tp[:1, None], 34)
    out = modules.FaceFormerLoss()(pred.to(dev), gt.to(dev))
    got = np.array([float(out["loss"]), float(out["rec_loss"]), float(out["vel_loss"])])
    np.testing.assert_allclose(got, z["faceformer"], rtol=1e-4)


def test_voca_loss_gradient(a2f_lib, dev):
    from a2f_b200 import modules
    rows = 8
    tp = oin.batch_templates(rows, 4)
    pred, gt = oin.gt_like((rows, 5023, 3), tp, 41), oin.gt_like((rows, 5023, 3), tp, 42)
    with torch.enable_grad():
        p_ref = pred.clone().requires_grad_(True)
        (orm.voca_loss(p_ref, gt)["loss"] * 3.0).backward()
        p_gpu = pred.to(dev).requires_grad_(True)
        (modules.VocaLoss()(p_gpu, gt.to(dev))["loss"] * 3.0).backward()
    g_ref, g_gpu = p_ref.grad, p_gpu.grad.cpu()
    assert float((g_ref - g_gpu).abs().max()) <= 1e-4 * float(g_ref.abs().max())


def test_odd_rows_rejected(a2f_lib, dev):
    import a2f_b200
    from a2f_b200 import modules
    with pytest.raises(a2f_b200.A2FError):
        modules.VocaLoss()(torch.zeros(3, 5023, 3, device=dev), torch.zeros(3, 5023, 3, device=dev))


@pytest.mark.parametrize("rows", [1, 2, 7, 300])
def test_mse_error_any_frame_count(a2f_lib, dev, rows):
    """ref:src/model/lightning_model.py:119-125 takes any number of frames (odd-length FaceFormer clips reach it)."""
    from a2f_b200 import modules
    g = torch.Generator().manual_seed(70 + rows)
    pred, gt = torch.randn(rows, 5023, 3, generator=g), torch.randn(rows, 5023, 3, generator=g)
    want = orm.mse_error(pred, gt)
    got = modules.mse_error(pred.to(dev), gt.to(dev)).cpu()
    assert abs(float(got) - float(want)) < 1e-5 * abs(float(want))
