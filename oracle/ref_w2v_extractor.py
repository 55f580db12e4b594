"""TEST INFRASTRUCTURE (oracle side) -- CPU restatement of the reference's Wav2VecExtractor.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

Follows ref:src/model/extractor.py:63-96:
  :88     torchaudio.functional.resample(x, sample_rate, 16000)                    -> oracle/ref_audio.resample
  :89-91  Wav2Vec2Processor(x, return_tensors="pt", padding=True, sampling_rate=16000).input_values[0]: the HF feature
          extractor does not treat a torch tensor as a batch, so the [B, N] tensor is normalised as ONE array,
          (x - mean(x)) / sqrt(var(x) + 1e-7) over all B*N samples (HF feature_extraction_wav2vec2.py:78-98,223-231)
  :92     Wav2Vec2Model(x).last_hidden_state: feature extractor -> feature projection -> encoder, eval mode
          (HF modeling_wav2vec2.py; restated in oracle/ref_models.py, here WITHOUT the reference's frame interpolation)
  :93-96  transpose(1, 2) and, when out_dim != 768, F.interpolate(x[:, None], size=(out_dim, n_feature), mode="bilinear")
Pinned by tests/golden/w2v_extractor.npz (tests/golden/make_golden_w2v_extractor.py, the live reference class with the
offline from_pretrained shims of SURVEY.md App. B).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ref_audio as ora
from . import ref_models as orm


def rekey(sd_model):
    """state_dict of the extractor (`model.*`) -> the `audio_encoder.*` keys oracle/ref_models.py reads."""
    return {"audio_encoder." + k[len("model."):]: v for k, v in sd_model.items() if k.startswith("model.")}


def w2v_extractor_forward(sd_model, x: torch.Tensor, sample_rate: int, n_feature: int, out_dim: int) -> torch.Tensor:
    sd = rekey(sd_model)
    x = ora.resample(x, sample_rate, 16000)
    x = (x - x.mean()) / torch.sqrt(x.var(unbiased=False) + 1e-7)          # one statistic for the whole tensor
    h = orm.feature_extractor(sd, x).transpose(1, 2)
    h = orm.feature_projection(sd, h)
    h = orm.encoder(sd, h)                                                # [B, T', 768]
    h = h.transpose(1, 2)
    if out_dim != h.shape[1]:
        h = F.interpolate(h.unsqueeze(1), size=(out_dim, n_feature), mode="bilinear").squeeze(1)
    return h
