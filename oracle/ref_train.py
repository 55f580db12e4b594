"""TEST INFRASTRUCTURE -- CPU fp32 restatement (oracle) of the reference's TRAINING step for FaceFormer.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

The reference trains through Lightning (ref:src/model/lightning_model.py:145-161): `verts*=100; template*=100;
pred = model(x, one_hot, template); loss = FaceFormerLoss()(pred, gt)["loss"]`, Adam(lr, weight_decay=lr/10)
(:209-213), and gets every gradient from torch.autograd.  This file restates that with torch.autograd over the
oracle forward (oracle/ref_models.py), in eval-mode semantics (dropout / LayerDrop / SpecAugment inactive: the
stochastic ops make train-mode parity meaningless).  Batches are the documented extension: the
loss of a batch is the mean of the per-utterance FaceFormerLoss values.

PINNING: tests/golden/make_golden_train.py runs the LIVE reference Faceformer (eval mode, autograd) on the same
weights / inputs and stores the loss and sub-sampled gradients of every parameter in
tests/golden/faceformer_train.npz; tests/test_oracle_golden.py checks this restatement against them.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import ref_models as orm


def faceformer_loss_and_grads(sd: Dict[str, torch.Tensor], audio: torch.Tensor, one_hot: torch.Tensor,
                              template: torch.Tensor, gt: torch.Tensor, fps: int = 60, spec_mask=None
                              ) -> Tuple[Dict[str, float], Dict[str, torch.Tensor]]:
    """audio [B,N], one_hot [B,12], template [B,5023,3], gt [B,T,5023,3] (already in training units) ->
    ({"loss","rec_loss","vel_loss"} batch means, {param name: grad})."""
    params = {}
    for k, v in sd.items():
        if v.is_floating_point() and k != "PPE.pe":
            params[k] = v.detach().clone().requires_grad_(True)
        else:
            params[k] = v
    B = audio.shape[0]
    tot = {"loss": 0.0, "rec_loss": 0.0, "vel_loss": 0.0}
    loss_sum = None
    with torch.enable_grad():
        for b in range(B):
            out = orm.faceformer_forward(params, audio[b:b + 1], one_hot[b:b + 1], template[b:b + 1], fps,
                                         spec_mask=None if spec_mask is None else spec_mask[b:b + 1])
            l = orm.faceformer_loss(out, gt[b:b + 1])
            for k in tot:
                tot[k] += float(l[k]) / B
            loss_sum = l["loss"] if loss_sum is None else loss_sum + l["loss"]
        (loss_sum / B).backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()
             if isinstance(p, torch.Tensor) and p.requires_grad}
    return tot, grads


def conv_loss_and_grads(kind: str, sd: Dict[str, torch.Tensor], x: torch.Tensor, one_hot: torch.Tensor,
                        template: torch.Tensor, gt: torch.Tensor):
    """Training step of Voca / Audio2Mesh (ref:src/model/lightning_model.py:150-161 with VocaLoss, ref:src/loss/loss.py:24-55):
    `pred = model(x, one_hot, template)` in TRAIN mode (Audio2Mesh: batch-statistics BatchNorm), `VocaLoss()(pred, gt)`,
    autograd.  -> (losses, {param: grad}, running-stat dict after the step (Audio2Mesh) or None)."""
    params = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone())
              for k, v in sd.items()}
    running = None
    with torch.enable_grad():
        if kind == "voca":
            out = orm.voca_forward(params, x, one_hot, template)
        else:
            running = {k: v for k, v in params.items() if "running_" in k or "num_batches" in k}
            out = orm.audio2mesh_forward(params, x, one_hot, template, train_bn=True, running=running)
        l = orm.voca_loss(out, gt)
        l["loss"].backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()
             if isinstance(p, torch.Tensor) and p.requires_grad}
    return {k: float(v) for k, v in l.items()}, grads, running


def adam_reference(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], lr: float, steps_state=None):
    """One torch.optim.Adam(lr, weight_decay=lr/10) step (ref:src/model/lightning_model.py:99,209-213) on copies."""
    ps = {k: torch.nn.Parameter(v.detach().clone()) for k, v in params.items() if k in grads}
    opt = torch.optim.Adam(list(ps.values()), lr=lr, weight_decay=lr / 10)
    for k, p in ps.items():
        p.grad = grads[k].clone()
    opt.step()
    return {k: p.detach() for k, p in ps.items()}
