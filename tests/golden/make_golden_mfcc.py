"""Golden fixture for the MFCC extractor: the LIVE reference MFCCExtractor (ref:src/model/extractor.py:10-60, imported
from /root/reference -- build container only) on oracle.inputs.speech_like_windows, for the two configurations the
reference uses.  Also stores the reference module's three buffers.

    python tests/golden/make_golden_mfcc.py
"""
from __future__ import annotations

import logging
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import inputs as oin, ref_mfcc as omf          # noqa: E402
from src.model.extractor import MFCCExtractor               # noqa: E402  (the live reference)

logging.disable(logging.WARNING)
torch.set_grad_enabled(False)
out = {}
for name, cfg in omf.CONFIGS.items():
    ext = MFCCExtractor(*cfg)
    x = oin.speech_like_windows(6, seed=21)
    y = ext(x)
    sd_ref = ext.state_dict()
    sd = omf.make_buffers(cfg[0], cfg[1], cfg[3], cfg[5])
    assert list(sd_ref.keys()) == list(sd.keys()), (list(sd_ref.keys()), list(sd.keys()))
    for k in sd:
        d = float((sd[k] - sd_ref[k]).abs().max())
        print(f"{name}: buffer {k} {tuple(sd[k].shape)} max|oracle - reference| = {d:.3e}")
        assert d <= 1e-6, (k, d)
    got = omf.mfcc_forward(sd, x, cfg[2], cfg[3], cfg[4], cfg[5])
    print(f"{name}: output {tuple(y.shape)}, max|.| {float(y.abs().max()):.1f}, max|oracle - reference| = {float((got - y).abs().max()):.3e}")
    out[f"{name}_out"] = y.numpy().astype(np.float32)
    out[f"{name}_fb_sum"] = sd_ref["T.MelSpectrogram.mel_scale.fb"].double().sum(0).numpy()
    out[f"{name}_dct"] = sd_ref["T.dct_mat"].numpy()
    out[f"{name}_window"] = sd_ref["T.MelSpectrogram.spectrogram.window"].numpy()
out["seed_in"] = np.int64(21)
out["batch"] = np.int64(6)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mfcc.npz"), **out)
print("wrote tests/golden/mfcc.npz")
