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


@pytest.mark.parametrize("B,T", [(1, 2), (2, 24), (3, 150)])
def test_vertex_head_with_fused_loss(a2f_lib, dev, B, T):
    """a2f_vertex_head_loss (head + template add + rec / vel loss + dL/dy in ONE kernel) vs the separate kernels
    (a2f_gemm head, a2f_voca_loss_fwd, a2f_voca_loss_bwd) and vs the oracle loss of the same prediction."""
    from a2f_b200 import ops, lib as L
    V3, M = 15069, B * T
    g = torch.Generator().manual_seed(900 + M)
    z = torch.zeros(M, 64)
    z[:, :64] = torch.randn(M, 64, generator=g)
    w = 0.02 * torch.randn(V3, 64, generator=g)
    bias = 0.01 * torch.randn(V3, generator=g)
    tmpl = torch.randn(B, V3, generator=g)
    gt = (tmpl[:, None] + 0.2 * torch.randn(B, T, V3, generator=g)).reshape(M, V3).contiguous()
    z3, w3 = ops.split_bf16x3(z.to(dev), False), ops.split_bf16x3(w.to(dev), True)
    # separate kernels
    pred = torch.empty((M, V3), device=dev)
    ops.gemm(z3, w3, pred, bias=bias.to(dev), tmpl=tmpl.to(dev), rows_per_tmpl=T, backend=L.TCGEN05, K=192)
    want3 = ops.voca_loss_fwd(pred, gt.to(dev), M, V3, 1.0, 10.0)
    dpred = torch.empty_like(pred)
    ops.voca_loss_bwd(pred, gt.to(dev), M, V3, 1.0, 10.0, None, dpred)
    # fused
    dy = torch.zeros((M, 15072), dtype=torch.bfloat16, device=dev)
    pred2 = torch.empty((M, V3), device=dev)
    got3 = ops.vertex_head_loss(z3, w3, bias.to(dev), tmpl.to(dev), T, gt.to(dev), dy, 1.0, 10.0, pred=pred2)
    got3_nopred = ops.vertex_head_loss(z3, w3, bias.to(dev), tmpl.to(dev), T, gt.to(dev), torch.zeros_like(dy), 1.0, 10.0)
    torch.cuda.synchronize()
    assert torch.equal(pred2, pred)                                     # same y, bit for bit
    assert torch.equal(got3, got3_nopred)                               # deterministic, independent of the optional store
    for j in range(3):
        assert abs(float(got3[j]) - float(want3[j])) < 2e-6 * abs(float(want3[j]))
    ref = orm.voca_loss(pred.cpu().view(M, -1, 3), gt.view(M, -1, 3))
    assert abs(float(got3[0]) - float(ref["loss"])) < 1e-5 * abs(float(ref["loss"]))
    d = dy[:, :V3].float()
    assert bool(((d - dpred).abs() <= 2.0 ** -8 * dpred.abs() + 1e-12).all())
    assert float(dy[:, V3:].abs().max()) == 0.0                         # pad columns untouched
