"""Golden fixture for the audio preparation kernels, from the reference itself:
  * get_audio_fragment / normalize_audio are taken out of /root/reference/src/dataset/vocaset.py with `ast` and executed
    (the module cannot be imported: it needs `lightning`), on oracle.inputs.pcm16_clip, a synthetic clip shaped like the
    reference's assets/audio_sample.npy (int16, 22 kHz, 127 600 samples -> 348 frames at 60 fps);
  * torchaudio.functional.resample(., 22000, 16000) is the live library the reference calls.
Stores sub-sampled outputs plus checksums.       python tests/golden/make_golden_audio_prep.py
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import inputs as oin, ref_audio as ora          # noqa: E402

src = open("/root/reference/src/dataset/vocaset.py").read()
tree = ast.parse(src)
ns = {"np": np, "Unpack": lambda x: x, "AduioParams": dict}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in ("get_audio_fragment", "normalize_audio"):
        node.returns = None
        for a in node.args.args + ([node.args.kwarg] if node.args.kwarg else []):
            a.annotation = None
        exec(compile(ast.Module([node], []), "vocaset.py", "exec"), ns)

clip = oin.pcm16_clip(seed=4)
real = np.load("/root/reference/assets/audio_sample.npy")
assert clip.dtype == real.dtype == np.int16 and clip.shape == real.shape == (127600,)
n_frames = clip.shape[0] * 60 // 22000
out = {"n_frames": np.int64(n_frames), "seed": np.int64(4), "clip_abs_sum": np.int64(int(np.abs(clip.astype(np.int64)).sum()))}
for shift in (0, 137, -200):
    ref = np.stack([ns["normalize_audio"](ns["get_audio_fragment"](clip, i, fps=60, sample_rate=22000, length=0.52, shift=shift))
                    for i in range(n_frames)])
    got = ora.fragments(clip, n_frames, shift=shift).numpy()
    print(f"fragments shift {shift}: {ref.shape}, max|oracle - reference| = {np.abs(got - ref).max():.1e}")
    out[f"frag_{shift}_rowsum"] = ref.astype(np.float64).sum(1)
    out[f"frag_{shift}_sub"] = ref[::29, ::97].copy()
x = torch.from_numpy((clip / 32768).astype(np.float32))
xb = torch.stack([x[:50000], x[30000:80000] * 0.5, torch.flip(x[:50000], [0])])
for name, wav, (fo, fn) in (("clip", x, (22000, 16000)), ("batch", xb, (22000, 16000)), ("up", xb[:, :7001], (16000, 22050))):
    ref = torchaudio.functional.resample(wav, fo, fn)
    got = ora.resample(wav, fo, fn)
    print(f"resample {name} {fo}->{fn}: {tuple(ref.shape)}, max|oracle - torchaudio| = {float((got - ref).abs().max()):.2e}")
    out[f"rs_{name}_len"] = np.int64(ref.shape[-1])
    out[f"rs_{name}_sub"] = ref.reshape(-1)[::53].numpy().copy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "audio_prep.npz"), **out)
print("wrote tests/golden/audio_prep.npz")
