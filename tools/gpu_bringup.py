"""Bring-up / measurement helper for gpurun sessions (not a pytest file).  Each stage runs in its own process so that
a trapped kernel (poisoned CUDA context) cannot hide later stages:

    python tools/gpu_bringup.py all          # runs every stage in subprocesses, prints a summary
    python tools/gpu_bringup.py <stage>
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _rand(shape, dev, seed, dtype=None, scale=1.0):
    import torch
    g = torch.Generator().manual_seed(seed)
    t = (scale * torch.randn(shape, generator=g)).to(dev)
    return t.to(dtype) if dtype is not None else t


def stage_tc_basic():
    import torch
    from a2f_b200 import ops, lib as L
    dev = torch.device("cuda:0")
    lib = L.load()
    L.check(lib.a2f_device_check())
    for (M, N, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 256), (300, 768, 512), (1000, 100, 72)]:
        a = _rand((M, K), dev, 1, torch.bfloat16)
        w = _rand((N, K), dev, 2, torch.bfloat16, K ** -0.5)
        out = torch.full((M, N), float("nan"), device=dev)
        ops.gemm(a, w, out, backend=L.TCGEN05)
        torch.cuda.synchronize()
        want = a.double().cpu() @ w.double().cpu().T
        err = (out.cpu().double() - want).abs()
        print(f"tc_basic M={M} N={N} K={K}: max err {float(err.max()):.3e} nan {int(torch.isnan(out).sum())}")


def stage_tc_scan():
    """If tc_basic is wrong: scan smem-descriptor hypotheses on a single-tile problem."""
    import itertools
    import torch
    from a2f_b200 import ops, lib as L
    dev = torch.device("cuda:0")
    lib = L.load()
    M, N, K = 128, 256, 128
    a = _rand((M, K), dev, 1, torch.bfloat16)
    w = _rand((N, K), dev, 2, torch.bfloat16, K ** -0.5)
    want = a.double().cpu() @ w.double().cpu().T
    for lbo, sbo, ver, lay in itertools.product((1, 0, 64), (64, 1, 8), (1, 0), (2, 1, 6)):
        for f, v in zip(range(4), (lbo, sbo, ver, lay)):
            lib.a2f_debug_set_umma_field(f, v)
        out = torch.zeros((M, N), device=dev)
        try:
            ops.gemm(a, w, out, backend=L.TCGEN05)
            torch.cuda.synchronize()
            err = float((out.cpu().double() - want).abs().max())
        except Exception as e:  # noqa: BLE001
            print(f"scan lbo={lbo} sbo={sbo} ver={ver} lay={lay}: EXC {e}")
            return
        print(f"scan lbo={lbo} sbo={sbo} ver={ver} lay={lay}: max err {err:.3e}")


def stage_tmap_odd_batch_stride():
    """Does the driver accept a batch stride that is not a multiple of the row stride? (conv over odd L_in)"""
    import torch
    from a2f_b200 import ops, lib as L
    dev = torch.device("cuda:0")
    C, taps, L_in, B = 512, 3, 41, 2
    L_out = (L_in - taps) // 2 + 1
    x = _rand((B, L_in, C), dev, 3, torch.bfloat16)
    wp = _rand((C, taps * C), dev, 4, torch.bfloat16, (taps * C) ** -0.5)
    out = torch.zeros((B * L_out, C), device=dev)
    try:
        ops.gemm(x, wp, out, backend=L.TCGEN05, M=B * L_out, K=taps * C, a_row_stride=2 * C, a_batch_stride=L_in * C,
                 rows_per_batch=L_out)
        torch.cuda.synchronize()
        rows = torch.stack([x[b, 2 * t: 2 * t + taps].reshape(-1) for b in range(B) for t in range(L_out)])
        want = rows.double().cpu() @ wp.double().cpu().T
        print("odd batch stride: accepted, max err", float((out.cpu().double() - want).abs().max()))
    except Exception as e:  # noqa: BLE001
        print("odd batch stride: rejected:", e)


def _time(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def _time_graph(fn, reps=10, iters=5):
    """Device time per call with host launch overhead removed: `reps` calls captured in one CUDA graph."""
    import torch
    fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (iters * reps) * 1e-3


def stage_tc_perf():
    import torch
    from a2f_b200 import ops, lib as L
    dev = torch.device("cuda:0")
    res = {}
    for (M, N, K) in [(9600, 3072, 768), (9600, 768, 3072), (9600, 2304, 768), (9600, 768, 768), (8192, 8192, 8192),
                      (4800, 3072, 768)]:
        a = _rand((M, K), dev, 1, torch.bfloat16)
        w = _rand((N, K), dev, 2, torch.bfloat16, K ** -0.5)
        out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
        t = _time_graph(lambda: ops.gemm(a, w, out, backend=L.TCGEN05))
        t_ref = _time_graph(lambda: torch.matmul(a, w.T, out=out))
        res[f"{M}x{N}x{K}"] = (2 * M * N * K / t / 1e12, 2 * M * N * K / t_ref / 1e12)
        print(f"tc_perf {M}x{N}x{K}: a2f {res[f'{M}x{N}x{K}'][0]:.1f} TFLOP/s   cuBLAS {res[f'{M}x{N}x{K}'][1]:.1f} TFLOP/s")
    # vertex head: HBM-bound
    for M, rpt in ((9600, 300), (4096, 1)):
        z = _rand((M, 64), dev, 3, torch.bfloat16)
        w = _rand((15069, 64), dev, 4, torch.bfloat16, 0.02)
        b = _rand((15069,), dev, 5)
        tm = _rand(((M + rpt - 1) // rpt, 15069), dev, 6)
        out = torch.empty((M, 15069), device=dev)
        t = _time_graph(lambda: ops.gemm(z, w, out, bias=b, tmpl=tm, rows_per_tmpl=rpt, backend=L.TCGEN05))
        byts = out.numel() * 4 + (tm.numel() * 4 if rpt == 1 else 0)
        print(f"vertex_head M={M} rows_per_tmpl={rpt}: {t * 1e6:.1f} us  {byts / t / 1e9:.0f} GB/s (out{'+tmpl' if rpt == 1 else ''})")


def stage_ff_perf():
    """FaceFormer bf16 forward at the bench shape (B=32 x 5 s) with a per-section breakdown."""
    import torch
    from a2f_b200 import modules, ops, lib as L
    from oracle import inputs as oin, weights as ow
    dev = torch.device("cuda:0")
    m = modules.Faceformer(15069, 12)
    m.load_state_dict(ow.make_state_dict("faceformer", 13), strict=True)
    m = m.to(dev).eval().set_precision("bf16")
    B = 32
    audio = oin.audio(B, 80000, 1).to(dev)
    oh = oin.one_hot(B, 12, 1).to(dev)
    tp = oin.batch_templates(B, 1, scale=100.0).to(dev)
    for fps in (30, 60):
        T = 80000 * fps // 16000
        with torch.no_grad():
            t_all = _time(lambda: m(audio, oh, tp, fps=fps), iters=5, warm=2)
            t_enc = _time(lambda: m.encode(audio, T), iters=5, warm=1)
            P = m._packed()
            mem = torch.randn(B * T, 64, device=dev)
            t_dec = _time(lambda: ops.decoder_rollout(P["dec"][0], mem, oh, 60, B, T), iters=5, warm=1)
            D = torch.randn(B * T, 64, device=dev)
            t_head = _time(lambda: m._vertex_head(D, m.vertice_map_r.weight, m.vertice_map_r.bias, tp.reshape(B, -1), T, 64),
                           iters=5, warm=1)
        print(f"ff_perf fps={fps} T={T}: forward {t_all * 1e3:.2f} ms ({B * T / t_all:.0f} frames/s)  encode {t_enc * 1e3:.2f} ms"
              f"  decode {t_dec * 1e3:.2f} ms  head {t_head * 1e3:.2f} ms")
    # launch-level breakdown of one forward with the torch profiler
    from torch.profiler import profile, ProfilerActivity
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(audio, oh, tp, fps=30)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))


STAGES = {k[6:]: v for k, v in list(globals().items()) if k.startswith("stage_")}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which != "all":
        STAGES[which]()
        return
    order = sys.argv[2:] or list(STAGES)
    for name in order:
        t0 = time.time()
        r = subprocess.run(["timeout", "300", sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True)
        print(f"===== stage {name}: rc={r.returncode} {time.time() - t0:.1f}s")
        print(r.stdout[-6000:])
        if r.returncode != 0:
            print(r.stderr[-3000:])


if __name__ == "__main__":
    main()
