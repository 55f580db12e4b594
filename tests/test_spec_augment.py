"""SpecAugment (ref:src/model/wav2vec.py:25-72,149-162).

CPU: the two draw-for-draw restatements of `_compute_mask_indices` (oracle/ref_models.py and the product's host code
audio2face-pytorch_b200/spec_augment.py) against masks recorded from the LIVE reference
(tests/golden/spec_augment.npz, made by tests/golden/make_golden_specaug.py), and the oracle's train-branch forward /
autograd against the live reference Faceformer in train mode with all dropout probabilities zeroed.
GPU: the CUDA training step with `model.spec_augment = True` against the same fixture and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import inputs as oin, ref_models as orm, ref_train as ort, weights as ow

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture():
    return np.load(os.path.join(G, "spec_augment.npz"))


def _inputs(z):
    n, s = int(z["n_samples"]), int(z["seed_in"])
    audio, oh, tp = oin.audio(1, n, s), oin.one_hot(1, 12, s), oin.batch_templates(1, s, scale=100.0)
    T = n * 60 // 16000
    gt = oin.gt_like((1, T, 5023, 3), tp[:, None], s + 1, scale=100.0)
    return audio, oh, tp, gt, T


def test_mask_restatements_match_live_reference_masks():
    import a2f_b200
    from a2f_b200 import spec_augment as sa
    z = _fixture()
    for i, (B, T, seed) in enumerate(z["mask_cases"].tolist()):
        want = np.unpackbits(z[f"mask{i}"], axis=1)[:, :T].astype(bool)
        np.random.seed(seed)
        got_oracle = orm.spec_augment_time_mask(B, T)
        np.random.seed(seed)
        got_product = sa.time_mask(B, T)
        assert got_oracle.shape == want.shape and bool((got_oracle == want).all()), (B, T, seed)
        assert got_product.shape == want.shape and bool((got_product == want).all()), (B, T, seed)
        assert want.sum(1).min() == want.sum(1).max()          # every utterance masks the same number of frames
        # the draw sequence is consumed identically: the generators are in the same state afterwards
        np.random.seed(seed); orm.spec_augment_time_mask(B, T); a = np.random.rand()
        np.random.seed(seed); sa.time_mask(B, T); b = np.random.rand()
        assert a == b


def test_mask_edge_cases():
    from a2f_b200 import spec_augment as sa
    rng = np.random.RandomState(0)
    m = sa.time_mask(3, 12, rng=rng)                           # shorter than span + spans: shrunk start range
    assert m.shape == (3, 12) and m.any(axis=1).all()
    with pytest.raises(ValueError):
        sa.time_mask(1, 2, rng=np.random.RandomState(0))       # two spans cannot be placed in two frames


def test_oracle_train_branch_matches_live_reference():
    z = _fixture()
    sd = ow.make_state_dict("faceformer", seed=int(z["seed_w"]))
    audio, oh, tp, gt, T = _inputs(z)
    np.random.seed(int(z["np_seed"]))
    mask = orm.spec_augment_time_mask(1, T)
    assert bool((mask == np.unpackbits(z["model_mask"], axis=1)[:, :T].astype(bool)).all())
    with torch.no_grad():
        out = orm.faceformer_forward(sd, audio, oh, tp, spec_mask=mask)
    np.testing.assert_allclose(out.reshape(-1)[::int(z["out_step"])].numpy(), z["out"], rtol=0, atol=5e-5)
    tot, grads = ort.faceformer_loss_and_grads(sd, audio, oh, tp, gt, spec_mask=mask)
    assert abs(tot["loss"] - float(z["loss"][0])) < 1e-5 * abs(float(z["loss"][0]))
    g = grads["audio_encoder.masked_spec_embed"].numpy()
    assert np.linalg.norm(g - z["g_embed"]) < 1e-4 * np.linalg.norm(z["g_embed"])
    for k, n in zip(z["grad_names"], z["grad_norms"]):
        assert abs(float(grads[str(k)].norm()) - float(n)) < 1e-4 * float(n), k


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cuda_train_step_with_spec_augment(a2f_lib, dev, precision):
    from a2f_b200 import modules
    z = _fixture()
    sd = ow.make_state_dict("faceformer", seed=int(z["seed_w"]))
    audio, oh, tp, gt, T = _inputs(z)
    m = modules.Faceformer(15069, 12).to(dev)
    m.load_state_dict(sd, strict=True)
    m.eval().set_precision(precision)
    m.spec_augment = True
    np.random.seed(int(z["np_seed"]))
    pred = m(audio.to(dev), oh.to(dev), tp.to(dev))
    loss = modules.FaceFormerLoss()(pred, gt.to(dev))
    loss["loss"].backward()
    torch.cuda.synchronize()
    got = pred.detach().cpu().reshape(-1)[::int(z["out_step"])].numpy()
    tol_cm = 1e-3 if precision == "fp32" else 5e-2            # 1e-5 m / 5e-4 m (north_star), centimetre units
    assert float(np.abs(got - z["out"]).max()) < tol_cm
    assert abs(float(loss["loss"]) - float(z["loss"][0])) < 1e-4 * abs(float(z["loss"][0]))
    g = m.audio_encoder.masked_spec_embed.grad
    assert g is not None
    g = g.cpu().numpy()
    rel = np.linalg.norm(g - z["g_embed"]) / np.linalg.norm(z["g_embed"])
    assert rel < (2e-3 if precision == "fp32" else 0.25), rel
    if precision == "fp32":
        np.random.seed(int(z["np_seed"]))
        mask = orm.spec_augment_time_mask(1, T)
        _, want = ort.faceformer_loss_and_grads(sd, audio, oh, tp, gt, spec_mask=mask)
        for k in ("audio_encoder.feature_projection.projection.weight", "audio_encoder.feature_projection.projection.bias",
                  "audio_encoder.feature_extractor.conv_layers.0.conv.weight", "audio_encoder.encoder.layers.0.attention.q_proj.weight"):
            p = dict(m.named_parameters())[k]
            r = float((p.grad.cpu().double() - want[k].double()).norm() / want[k].double().norm())
            assert r < 2e-3, (k, r)
    # without SpecAugment the parameter receives no gradient (reference behaviour)
    m2 = modules.Faceformer(15069, 12).to(dev)
    m2.load_state_dict(sd, strict=True)
    m2.eval().set_precision(precision)
    modules.FaceFormerLoss()(m2(audio.to(dev), oh.to(dev), tp.to(dev)), gt.to(dev))["loss"].backward()
    assert m2.audio_encoder.masked_spec_embed.grad is None
