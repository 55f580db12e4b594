"""Golden fixture for SpecAugment, from the LIVE reference (build container only):

    python tests/golden/make_golden_specaug.py

1. masks: `_compute_mask_indices((B,T), 0.05, 10, None, min_masks=2)` of ref:src/model/wav2vec.py:25-72 under
   `np.random.seed(s)` for several (B, T, s) -- pins the draw-for-draw restatements (oracle + product host code);
2. model: the reference Faceformer in TRAIN mode with every dropout probability and LayerDrop set to 0 (so that
   SpecAugment is the only active stochastic op), `np.random.seed(7)`, one 1 s utterance: sub-sampled output, the
   FaceFormerLoss and the gradient of `audio_encoder.masked_spec_embed` (None without SpecAugment).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import inputs as oin, ref_models as orm, ref_train as ort, weights as ow      # noqa: E402
from make_golden import check_keys                                                          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MASK_CASES = [(1, 60, 0), (1, 300, 1), (8, 300, 2), (3, 150, 3), (2, 25, 4), (4, 600, 5), (5, 41, 6)]
N_SAMPLES, SEED_IN, SEED_W, NP_SEED = 16000, 43, 13, 7


def build_reference_faceformer_no_dropout():
    from transformers import Wav2Vec2Config, Wav2Vec2FeatureExtractor
    import src.model.faceformer as ff
    from src.model.wav2vec import Wav2Vec2Model as RefW2V

    class _P:
        @staticmethod
        def from_pretrained(name):
            return Wav2Vec2FeatureExtractor()

    ff.Wav2Vec2Processor = _P

    def _mk(cls, name):
        cfg = Wav2Vec2Config(hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0,
                             layerdrop=0.0, final_dropout=0.0)
        cfg._attn_implementation = "eager"
        return cls(cfg)

    RefW2V.from_pretrained = classmethod(_mk)
    m = ff.Faceformer(15069, 12)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, torch.nn.MultiheadAttention):
            mod.dropout = 0.0
    return m


def main():
    from src.model.wav2vec import _compute_mask_indices
    fx = {"mask_cases": np.array(MASK_CASES, dtype=np.int64)}
    for i, (B, T, seed) in enumerate(MASK_CASES):
        np.random.seed(seed)
        ref = _compute_mask_indices((B, T), 0.05, 10, None, min_masks=2)
        np.random.seed(seed)
        mine = orm.spec_augment_time_mask(B, T)
        assert ref.shape == mine.shape and bool((ref == mine).all()), (B, T, seed)
        fx[f"mask{i}"] = np.packbits(ref, axis=1)
    print("masks: oracle restatement identical to the live reference on", len(MASK_CASES), "cases")

    torch.manual_seed(0)
    model = build_reference_faceformer_no_dropout()
    sd = ow.make_state_dict("faceformer", seed=SEED_W)
    check_keys(model, sd, "faceformer")
    model.train()
    from src.loss import FaceFormerLoss
    audio, oh = oin.audio(1, N_SAMPLES, SEED_IN), oin.one_hot(1, 12, SEED_IN)
    tp = oin.batch_templates(1, SEED_IN, scale=100.0)
    T = N_SAMPLES * 60 // 16000
    gt = oin.gt_like((1, T, 5023, 3), tp[:, None], SEED_IN + 1, scale=100.0)
    np.random.seed(NP_SEED)
    with torch.enable_grad():
        pred = model(audio, oh, tp)
        loss = FaceFormerLoss()(pred, gt)
        loss["loss"].backward()
    g_embed = model.audio_encoder.masked_spec_embed.grad
    assert g_embed is not None and float(g_embed.norm()) > 0
    np.random.seed(NP_SEED)
    mask = orm.spec_augment_time_mask(1, T)
    tot, grads = ort.faceformer_loss_and_grads(sd, audio, oh, tp, gt, spec_mask=mask)
    out_or = orm.faceformer_forward(sd, audio, oh, tp, spec_mask=mask)
    d_out = float((out_or - pred.detach()).abs().max())
    d_g = float((grads["audio_encoder.masked_spec_embed"] - g_embed).norm() / g_embed.norm())
    print(f"model: T={T}, {int(mask.sum())} masked frames; ref loss {float(loss['loss']):.6f} oracle {tot['loss']:.6f}; "
          f"max|out diff| {d_out:.3e} cm; rel diff of d(masked_spec_embed) {d_g:.3e}")
    step = max(1, pred.numel() // 4096)
    fx.update(n_samples=N_SAMPLES, seed_in=SEED_IN, seed_w=SEED_W, np_seed=NP_SEED, out_step=step,
              out=pred.detach().reshape(-1)[::step].numpy().astype(np.float32),
              loss=np.array([float(loss["loss"]), float(loss["rec_loss"]), float(loss["vel_loss"])], dtype=np.float64),
              g_embed=g_embed.numpy().astype(np.float32), model_mask=np.packbits(mask, axis=1))
    names = ["audio_encoder.feature_projection.projection.weight", "audio_encoder.encoder.layers.0.attention.q_proj.weight",
             "audio_encoder.feature_extractor.conv_layers.0.conv.weight", "vertice_map_r.weight"]
    ref_named = dict(model.named_parameters())
    fx["grad_names"] = np.array(names)
    fx["grad_norms"] = np.array([float(ref_named[k].grad.norm()) for k in names], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "spec_augment.npz"), **fx)
    print("written", os.path.join(OUT, "spec_augment.npz"))


if __name__ == "__main__":
    main()
