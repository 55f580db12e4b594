"""Golden fixture for the wav2vec feature extractor: the LIVE reference Wav2VecExtractor (ref:src/model/extractor.py:63-96,
imported from /root/reference -- build container only) with the offline shims of SURVEY.md App. B (the two
from_pretrained calls return a default-config Wav2Vec2FeatureExtractor / a random-init Wav2Vec2Model with eager attention),
loaded with the encoder part of oracle.weights' FaceFormer state_dict.

    python tests/golden/make_golden_w2v_extractor.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import inputs as oin, ref_w2v_extractor as owx, weights as ow      # noqa: E402
import src.model.extractor as ext_mod                                           # noqa: E402  (the live reference)
from transformers import Wav2Vec2Config, Wav2Vec2FeatureExtractor               # noqa: E402

torch.set_grad_enabled(False)


class _P:
    @staticmethod
    def from_pretrained(name):
        return Wav2Vec2FeatureExtractor()


def _mk(cls, name):
    cfg = Wav2Vec2Config()
    cfg._attn_implementation = "eager"
    return cls(cfg)


ext_mod.Wav2Vec2Processor = _P
ext_mod.Wav2Vec2Model.from_pretrained = classmethod(_mk)

ff_sd = ow.make_state_dict("faceformer", 13)
sd_model = {"model." + k[len("audio_encoder."):]: v for k, v in ff_sd.items() if k.startswith("audio_encoder.")}
ref = ext_mod.Wav2VecExtractor(22000, 32, 52).eval()
missing, unexpected = ref.load_state_dict(sd_model, strict=False)
print("reference keys not in the oracle state_dict:", missing, "| unexpected:", unexpected)
assert not unexpected
x = oin.speech_like_windows(3, seed=31)
y = ref(x)
got = owx.w2v_extractor_forward(sd_model, x, 22000, 32, 52)
print(f"output {tuple(y.shape)}, max|.| {float(y.abs().max()):.3f}, max|oracle - reference| = {float((got - y).abs().max()):.3e}")
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "w2v_extractor.npz"), out=y.numpy().astype(np.float32),
                    seed_in=np.int64(31), batch=np.int64(3), seed_w=np.int64(13), n_keys=np.int64(len(ref.state_dict())))
print("wrote tests/golden/w2v_extractor.npz")
