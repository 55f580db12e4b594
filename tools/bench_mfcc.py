#!/usr/bin/env python
"""MFCC extractor: per-stage device time (CUDA events around each C-ABI call) and error against the oracle.

    python tools/bench_mfcc.py [B]          (default 4096 windows of 11440 samples, VOCA configuration)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from a2f_b200 import features, lib as L, ops
from oracle import inputs as oin, ref_mfcc as omf

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
for name, cfg in omf.CONFIGS.items():
    sr, nf, od, win, hop, nfft = cfg
    xs = oin.speech_like_windows(8, seed=2)
    want = omf.mfcc_forward(omf.make_buffers(sr, nf, win, nfft), xs, od, win, hop, nfft)
    x = oin.speech_like_windows(64, seed=3).repeat((B + 63) // 64, 1)[:B].contiguous().to(dev)
    for prec in ("fp32", "bf16"):
        m = features.MFCCExtractor(*cfg).to(dev).set_precision(prec)
        with torch.no_grad():
            err = float((m(xs.to(dev)).cpu() - want).abs().max())
            for _ in range(3):
                m(x)
            torch.cuda.synchronize()
            bs = m._basis()
            gmax = torch.empty(1, device=dev)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
            frames, F_ = ops.mfcc_frames(x, win, m.hop_length, nfft, bs["kpad"], torch.bfloat16 if prec == "bf16" else torch.float32, gmax)
            ev[1].record()
            spec = torch.empty((B * F_, bs["npad"]), dtype=torch.float32, device=dev)
            ops.gemm(frames, bs["w"], spec, backend=L.TCGEN05 if prec == "bf16" else L.SIMT_F32)
            ev[2].record()
            db = ops.mfcc_mel_db(spec, m.n_freq, m.T.MelSpectrogram.mel_scale.fb, m._bands(), gmax)
            ev[3].record()
            ops.mfcc_dct_resize(db, gmax, 80.0, m.T.dct_mat, B, F_, od)
            ev[4].record()
            torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(4)]
        tot = sum(t)
        alg = 2.0 * B * F_ * 1026 * win
        print(f"{name:10s} {prec}: max|err| vs oracle {err:.2e}; B={B}: frames {t[0]:.0f} us, DFT GEMM {t[1]:.0f} us "
              f"({alg / t[1] * 1e-6:.0f} algorithmic TFLOP/s), mel+dB {t[2]:.0f} us, DCT+resize {t[3]:.0f} us; "
              f"{B / tot * 1e6 / 1e6:.2f} M windows/s")
