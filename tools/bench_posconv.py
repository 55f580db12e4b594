#!/usr/bin/env python
"""Bring-up + timing of the positional-conv kernels (tcgen05 backend): posconv_tc.cu (kpad 8) vs the generic-GEMM path
(kpad 64, debug field 9), error against an fp64 torch conv on the bf16-rounded operands.   python tools/bench_posconv.py [B,T ...]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import a2f_b200  # noqa: E402
from a2f_b200 import lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(2, 150), (32, 150), (8, 300), (2, 3600)]
gen = torch.Generator().manual_seed(7)
g = (0.5 + torch.rand(128, generator=gen)).to(dev)
v = (torch.randn(768, 48, 128, generator=gen) * (48 * 128) ** -0.5).to(dev)
bias = (torch.randn(768, generator=gen) * 0.05).to(dev)
norm = torch.empty(128, device=dev)
wfull = torch._weight_norm(v.cpu(), g.cpu().view(1, 1, 128), 2).to(torch.bfloat16).double()
for B, T in shapes:
    h = torch.randn(B, T, 768, generator=gen).to(dev).to(torch.bfloat16)
    want = None
    if B * T <= 8000:
        hh = h.double().cpu()
        pos = F.conv1d(hh.transpose(1, 2), wfull, bias.double().cpu(), padding=64, groups=16)[:, :, :-1]
        want = hh + F.gelu(pos).transpose(1, 2)
    for name, kpad, impl, swap in (("posconv_tc", 8, 0, 0), ("gemm_tc mode 2", 64, 1, 0)):
        lib.a2f_debug_set_umma_field(9, impl)
        lib.a2f_debug_set_umma_field(10, swap)
        wp = torch.zeros((16, 128, 6, 48, 8) if kpad == 8 else (16, 48, 128, 64), device=dev, dtype=torch.bfloat16)
        L.check(lib.a2f_pack_posconv_weight(g.data_ptr(), v.data_ptr(), wp.data_ptr(), 1, kpad, norm.data_ptr(), st))
        out = torch.zeros((B, T, 768), device=dev, dtype=torch.bfloat16)
        try:
            for _ in range(3):
                L.check(lib.a2f_posconv(h.data_ptr(), 1, wp.data_ptr(), bias.data_ptr(), out.data_ptr(), 1, B, T, L.TCGEN05, st))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"B={B} T={T} {name}: FAILED {e}")
            continue
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(20):
            L.check(lib.a2f_posconv(h.data_ptr(), 1, wp.data_ptr(), bias.data_ptr(), out.data_ptr(), 1, B, T, L.TCGEN05, st))
        ev[1].record()
        torch.cuda.synchronize()
        us = ev[0].elapsed_time(ev[1]) * 1e3 / 20
        err = float((out.cpu().double() - want).abs().max()) if want is not None else float("nan")
        print(f"B={B} T={T} {name}: {us:.1f} us/call, {2.0 * B * T * 768 * 48 * 128 / us * 1e-6:.1f} TFLOP/s, max err {err:.4f}")
lib.a2f_debug_set_umma_field(9, 0)
lib.a2f_debug_set_umma_field(10, 0)
