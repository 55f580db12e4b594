"""Golden fixture for Song2Face: the LIVE reference module (ref:src/model/song2face.py, imported from /root/reference --
build container only) with the oracle.weights state_dict loaded strict=True, eval mode.

    python tests/golden/make_golden_song2face.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import inputs as oin, ref_models as orm, weights as ow      # noqa: E402
from src.model.song2face import Song2Face                                # noqa: E402  (the live reference)

torch.set_grad_enabled(False)
sd = ow.make_state_dict("song2face", seed=14)
ref = Song2Face(15069, 12).eval()
assert list(ref.state_dict().keys()) == list(sd.keys()), "state_dict key order / set differs from the reference"
for k, v in ref.state_dict().items():
    assert tuple(v.shape) == tuple(sd[k].shape), k
ref.load_state_dict(sd, strict=True)
B, s = 5, 41
x, oh, tp = oin.a2m_features(B, s), oin.one_hot(B, 12, s), oin.batch_templates(B, s)
y = ref(x, oh, tp)
got = orm.song2face_forward(sd, x, oh, tp)
print(f"output {tuple(y.shape)}, |offset|max {float((y - tp).abs().max()):.3f}, max|oracle - reference| = {float((got - y).abs().max()):.3e}")
step = 37
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "song2face.npz"),
                    out=y.reshape(-1)[::step].numpy().astype(np.float32), step=np.int64(step), batch=np.int64(B),
                    seed_in=np.int64(s), seed_w=np.int64(14), n_keys=np.int64(len(sd)))
print("wrote tests/golden/song2face.npz")
