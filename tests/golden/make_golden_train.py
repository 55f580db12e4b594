"""Golden fixture for the TRAINING step: loss and gradients of the LIVE reference Faceformer (imported from
/root/reference; eval-mode semantics, torch.autograd) on oracle.weights / oracle.inputs data.

    python tests/golden/make_golden_train.py      (build container only: needs /root/reference)

Stored per parameter: its gradient's L2 norm and a strided sub-sample (<= 256 values)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import inputs as oin, ref_train as ort, weights as ow      # noqa: E402
from make_golden import build_reference_faceformer, check_keys        # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N_SAMPLES, SEED_IN, SEED_W, NSUB = 8000, 41, 13, 256


def train_inputs(n_samples=N_SAMPLES, seed=SEED_IN, B=1):
    audio = oin.audio(B, n_samples, seed)
    oh = oin.one_hot(B, 12, seed)
    tp = oin.batch_templates(B, seed, scale=100.0)                    # ref lightning_model.py:145-148 (x100)
    T = n_samples * 60 // 16000
    gt = oin.gt_like((B, T, 5023, 3), tp[:, None], seed + 1, scale=100.0)
    return audio, oh, tp, gt


def subsample(g: torch.Tensor) -> np.ndarray:
    flat = g.reshape(-1)
    step = max(1, flat.numel() // NSUB)
    return flat[::step][:NSUB].contiguous().numpy().astype(np.float32)


def main():
    torch.manual_seed(0)
    model, _ = build_reference_faceformer()
    model.eval()                      # dropout / LayerDrop / SpecAugment off; autograd still works
    sd = ow.make_state_dict("faceformer", seed=SEED_W)
    check_keys(model, sd, "faceformer")
    from src.loss import FaceFormerLoss
    audio, oh, tp, gt = train_inputs()
    with torch.enable_grad():
        pred = model(audio, oh, tp)
        loss = FaceFormerLoss()(pred, gt)
        loss["loss"].backward()
    ref_grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in model.named_parameters()}
    tot, or_grads = ort.faceformer_loss_and_grads(sd, audio, oh, tp, gt)
    worst = 0.0
    for k, g in ref_grads.items():
        d = float((g - or_grads[k]).norm() / (g.norm() + 1e-30))
        worst = max(worst, d if float(g.norm()) > 0 else 0.0)
    print("train  ref loss", float(loss["loss"]), "oracle", tot["loss"], " worst per-tensor rel grad diff", worst)
    fx = {"loss": np.array([float(loss["loss"]), float(loss["rec_loss"]), float(loss["vel_loss"])], dtype=np.float64),
          "n_samples": N_SAMPLES, "seed_in": SEED_IN, "seed_w": SEED_W, "nsub": NSUB}
    names = sorted(ref_grads.keys())
    fx["names"] = np.array(names)
    fx["norms"] = np.array([float(ref_grads[k].norm()) for k in names], dtype=np.float64)
    # bf16 noise floor: the LIVE reference under torch.autocast(bf16) -- what Lightning's mixed precision
    # (ref:train.py:49 precision="16-mixed") does to these gradients -- against its own fp32 gradients.  Stored per
    # parameter as a relative L2 error; tests scale the bf16 tolerance of the CUDA path by it.
    model.zero_grad(set_to_none=True)
    with torch.enable_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        pred16 = model(audio, oh, tp)
        loss16 = FaceFormerLoss()(pred16.float(), gt)
        loss16["loss"].backward()
    noise = []
    for k in names:
        g16 = dict(model.named_parameters())[k].grad
        g16 = g16 if g16 is not None else torch.zeros_like(ref_grads[k])
        noise.append(float((g16.double() - ref_grads[k].double()).norm()) / max(float(ref_grads[k].norm()), 1e-30))
    fx["bf16_noise"] = np.array(noise, dtype=np.float64)
    fx["bf16_loss"] = np.array([float(loss16["loss"])], dtype=np.float64)
    print("autocast-bf16 reference: loss", float(loss16["loss"]), " median per-tensor rel grad noise", float(np.median(noise)))
    for i, k in enumerate(names):
        fx[f"g{i}"] = subsample(ref_grads[k])
    np.savez_compressed(os.path.join(OUT, "faceformer_train.npz"), **fx)
    print("written", os.path.join(OUT, "faceformer_train.npz"), len(names), "parameters")


if __name__ == "__main__":
    main()
