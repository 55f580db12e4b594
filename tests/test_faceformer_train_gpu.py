"""FaceFormer training step on the GPU vs the oracle's autograd (oracle/ref_train.py) and the live-reference gradient
fixture (tests/golden/faceformer_train.npz): loss to 1e-4 relative (north_star), every parameter gradient by relative
L2 error -- fp32 path tight, bf16 tensor-core path at bf16 accuracy."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_train as ort, weights as ow

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _train_inputs(n_samples, seed, B=1):
    audio = oin.audio(B, n_samples, seed)
    oh = oin.one_hot(B, 12, seed)
    tp = oin.batch_templates(B, seed, scale=100.0)
    T = n_samples * 60 // 16000
    gt = oin.gt_like((B, T, 5023, 3), tp[:, None], seed + 1, scale=100.0)
    return audio, oh, tp, gt


def _run_gpu(dev, sd, precision, audio, oh, tp, gt):
    from a2f_b200 import modules
    m = modules.Faceformer(15069, 12).to(dev)
    m.load_state_dict(sd, strict=True)
    m.eval().set_precision(precision)                 # eval-mode semantics; autograd enabled
    loss_fn = modules.FaceFormerLoss()
    pred = m(audio.to(dev), oh.to(dev), tp.to(dev))
    loss = loss_fn(pred, gt.to(dev))
    loss["loss"].backward()
    torch.cuda.synchronize()
    grads = {k: (p.grad.detach().cpu() if p.grad is not None else torch.zeros_like(p).cpu()) for k, p in m.named_parameters()}
    return {k: float(v) for k, v in loss.items()}, grads


def _compare(grads, want, tol, skip_tiny=1e-6, zero_slack=10.0):
    gmax = max(float(g.norm()) for g in want.values())
    worst, worst_k = 0.0, None
    for k, g in want.items():
        n = float(g.norm())
        if n < skip_tiny * gmax:                      # mathematically-zero gradients (k_proj.bias: softmax shift invariance)
            assert float(grads[k].norm()) < zero_slack * skip_tiny * gmax, k
            continue
        rel = float((grads[k].double() - g.double()).norm()) / n
        if rel > worst:
            worst, worst_k = rel, k
    assert worst < tol, (worst_k, worst)
    return worst, worst_k


@pytest.fixture(scope="module")
def ff_sd():
    return ow.make_state_dict("faceformer", seed=13)


def test_train_step_fp32_vs_oracle_and_golden(a2f_lib, dev, ff_sd):
    z = np.load(os.path.join(G, "faceformer_train.npz"))
    audio, oh, tp, gt = _train_inputs(int(z["n_samples"]), int(z["seed_in"]))
    loss, grads = _run_gpu(dev, ff_sd, "fp32", audio, oh, tp, gt)
    # loss vs the live reference
    assert abs(loss["loss"] - float(z["loss"][0])) < 1e-4 * abs(float(z["loss"][0]))
    assert abs(loss["rec_loss"] - float(z["loss"][1])) < 1e-4 * abs(float(z["loss"][1]))
    assert abs(loss["vel_loss"] - float(z["loss"][2])) < 1e-4 * abs(float(z["loss"][2]))
    # gradients vs the live-reference fixture (sub-sampled) ...
    names = [str(n) for n in z["names"]]
    nsub = int(z["nsub"])
    gmax = float(z["norms"].max())
    for i, k in enumerate(names):
        if float(z["norms"][i]) < 1e-6 * gmax:
            continue
        flat = grads[k].reshape(-1)
        step = max(1, flat.numel() // nsub)
        got = flat[::step][:nsub].numpy()
        ref = z[f"g{i}"]
        denom = max(float(np.abs(ref).max()), 1e-3 * float(z["norms"][i]) / np.sqrt(flat.numel()))
        assert float(np.abs(got - ref).max()) < 2e-3 * denom + 1e-7 * gmax, (k, float(np.abs(got - ref).max()), denom)
    # ... and every element vs the oracle's autograd
    tot, want = ort.faceformer_loss_and_grads(ff_sd, audio, oh, tp, gt)
    assert abs(loss["loss"] - tot["loss"]) < 1e-4 * abs(tot["loss"])
    # 3e-3: the worst tensor is always the last layer's k_proj.weight (a tiny-difference gradient, softmax shift invariance),
    # measured 1.4e-3 .. 2.0e-3 from run to run (the fp32 SIMT weight-gradient / split-K kernels accumulate with atomics);
    # every other tensor is below 1e-3
    worst, k = _compare(grads, want, 3e-3)
    print(f"fp32 train step: loss {loss['loss']:.6f} (oracle {tot['loss']:.6f}); worst per-tensor rel grad err {worst:.2e} ({k})")


def test_train_step_bf16_vs_oracle(a2f_lib, dev, ff_sd):
    """bf16 tensor-core training step.  At random init this 12-layer post-LN encoder amplifies bf16 rounding strongly:
    the LIVE reference under torch.autocast(bf16) (Lightning "16-mixed", ref:train.py:49) is itself 8 % (median) to
    > 400 % (late-layer q/k projections, whose gradients are tiny differences) away from its own fp32 gradients.
    The fixture stores that per-parameter noise floor; the CUDA path must stay within 3x of it (or 8e-2), its loss
    within the north star's 1e-4, and the whole gradient within 0.97 cosine of the fp32 gradient."""
    z = np.load(os.path.join(G, "faceformer_train.npz"))
    audio, oh, tp, gt = _train_inputs(int(z["n_samples"]), int(z["seed_in"]))
    loss, grads = _run_gpu(dev, ff_sd, "bf16", audio, oh, tp, gt)
    tot, want = ort.faceformer_loss_and_grads(ff_sd, audio, oh, tp, gt)
    assert abs(loss["loss"] - tot["loss"]) < 1e-4 * abs(tot["loss"])
    noise = {str(n): float(v) for n, v in zip(z["names"], z["bf16_noise"])}
    gmax = max(float(g.norm()) for g in want.values())
    worst, worst_k, dot, n1, n2 = 0.0, None, 0.0, 0.0, 0.0
    for k, g in want.items():
        a, b = grads[k].double().reshape(-1), g.double().reshape(-1)
        dot, n1, n2 = dot + float(a @ b), n1 + float(a @ a), n2 + float(b @ b)
        n = float(g.norm())
        if n < 1e-6 * gmax:                           # mathematically-zero gradients (k_proj.bias)
            assert float(grads[k].norm()) < 3e-4 * gmax, k
            continue
        rel = float((a - b).norm()) / n
        ratio = rel / max(8e-2, 3.0 * noise[k])
        if ratio > worst:
            worst, worst_k = ratio, (k, rel, noise[k])
    cos = dot / (n1 * n2) ** 0.5
    print(f"bf16 train step: loss {loss['loss']:.6f} (oracle {tot['loss']:.6f}); gradient cosine {cos:.4f}; "
          f"worst (err / allowed) {worst:.2f} at {worst_k}")
    assert worst < 1.0, worst_k
    assert cos > 0.97, cos


def test_train_step_batch_fp32(a2f_lib, dev, ff_sd):
    """Batch extension: gradient of the mean of per-utterance losses; B=2, T=24 (even)."""
    audio, oh, tp, gt = _train_inputs(6400, 43, B=2)
    loss, grads = _run_gpu(dev, ff_sd, "fp32", audio, oh, tp, gt)
    tot, want = ort.faceformer_loss_and_grads(ff_sd, audio, oh, tp, gt)
    assert abs(loss["loss"] - tot["loss"]) < 1e-4 * abs(tot["loss"])
    worst, k = _compare(grads, want, 3e-3)          # same run-to-run spread as above
    print(f"fp32 batch train step: worst per-tensor rel grad err {worst:.2e} ({k})")
