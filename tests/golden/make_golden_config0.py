"""Golden fixture for BASELINE.json configs[0] at its LITERAL shape, from the live reference:

    assets/audio_sample.npy (int16, 22 kHz, 127 600 samples)
      -> 348 windows of 11 440 samples  (get_audio_fragment + normalize_audio, taken out of
         /root/reference/src/dataset/vocaset.py with `ast`: the module itself needs `lightning`)
      -> MFCCExtractor(22000, 16, 29, 790, None, 1024)          (ref:src/model/extractor.py, imported)
      -> Voca(15069, 12) with oracle.weights seed 11            (ref:src/model/voca.py, imported)
      -> [348, 5023, 3]

`assets/verts_sample.npy`, which BASELINE names as the check, is absent from the reference checkout
(ref:.MISSING_LARGE_BLOBS) and is a renderer input, not a model output -- the live modules are the pin.
The fixture carries the clip (the GPU box has no /root/reference), the reference's MFCC features and vertices
sub-sampled, and per-window checksums.          python tests/golden/make_golden_config0.py
"""
from __future__ import annotations

import ast
import logging
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import inputs as oin, ref_audio as ora, ref_mfcc as omf, ref_models as orm, weights as ow   # noqa: E402
from src.model.extractor import MFCCExtractor               # noqa: E402  (the live reference)
from src.model.voca import Voca                             # noqa: E402  (the live reference)

logging.disable(logging.WARNING)
torch.set_grad_enabled(False)

src = open("/root/reference/src/dataset/vocaset.py").read()
ns = {"np": np, "Unpack": lambda x: x, "AduioParams": dict}
for node in ast.parse(src).body:
    if isinstance(node, ast.FunctionDef) and node.name in ("get_audio_fragment", "normalize_audio"):
        node.returns = None
        for a in node.args.args + ([node.args.kwarg] if node.args.kwarg else []):
            a.annotation = None
        exec(compile(ast.Module([node], []), "vocaset.py", "exec"), ns)

clip = np.load("/root/reference/assets/audio_sample.npy")
assert clip.dtype == np.int16 and clip.shape == (127600,)
n_frames = clip.shape[0] * 60 // 22000
assert n_frames == 348
win = np.stack([ns["normalize_audio"](ns["get_audio_fragment"](clip, i, fps=60, sample_rate=22000, length=0.52, shift=0))
                for i in range(n_frames)]).astype(np.float32)
x = torch.from_numpy(win)
cfg = omf.CONFIGS["voca"]
feat = MFCCExtractor(*cfg)(x)                                # [348, 29, 16]
assert tuple(feat.shape) == (348, 29, 16)
sd = ow.make_state_dict("voca", seed=11)
model = Voca(15069, 12)
model.load_state_dict(sd, strict=True)
model.eval()
oh = torch.zeros(n_frames, 12)
oh[:, 0] = 1.0                                               # one_hot = e0
tp = oin.flame_like_template(3)[None].expand(n_frames, -1, -1).contiguous()
verts = model(feat, oh, tp)                                  # [348, 5023, 3]

# the oracle chain on the same clip (pins the restatement at this shape)
o_win = ora.fragments(clip, n_frames)
o_feat = omf.mfcc_forward(omf.make_buffers(cfg[0], cfg[1], cfg[3], cfg[5]), o_win, cfg[2], cfg[3], cfg[4], cfg[5])
o_verts = orm.voca_forward(sd, o_feat, oh, tp)
print(f"windows  max|oracle - reference| = {float((o_win - x).abs().max()):.2e}")
print(f"features max|oracle - reference| = {float((o_feat - feat).abs().max()):.2e} (|feat| max {float(feat.abs().max()):.1f})")
print(f"vertices max|oracle - reference| = {float((o_verts - verts).abs().max()):.2e} "
      f"(|offset| max {float((verts - tp).abs().max()):.3f})")

out = {
    "clip": clip, "n_frames": np.int64(n_frames), "template_seed": np.int64(3), "weight_seed": np.int64(11),
    "feat_sub": feat.reshape(-1)[::7].numpy().copy(), "feat_step": np.int64(7),
    "verts_sub": verts.reshape(-1)[::487].numpy().copy(), "verts_step": np.int64(487),
    "verts_rowsum": verts.double().reshape(n_frames, -1).sum(1).numpy(),
    "offset_absmax": np.float64(float((verts - tp).abs().max())),
}
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config0_voca.npz"), **out)
print("wrote tests/golden/config0_voca.npz")
