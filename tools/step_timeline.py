"""Per-call device timeline of one step of the hot path, measured with CUDA events around EVERY C-ABI call.

    python tools/step_timeline.py [infer|train|a2m_train|a2m] [B] [fps]

The GPU is first parked on a long spin kernel so that the host has enqueued the whole step (calls + events) before
the device starts: the event pairs then measure the kernels back to back, warm L2, real clocks -- without the
host-side gaps an eager Python loop inserts, and without ncu's cold-cache serialisation.  Output: time per entry
point (sum over the step), launches, share; plus the step's wall time on the device.  Summaries are committed under
profiles/."""
import os
import sys
from collections import OrderedDict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from a2f_b200 import lib as L, modules, trainer as tr
from oracle import inputs as oin, weights as ow


class Recorder:
    def __init__(self):
        self.lib = L.load()
        self.rows = []
        self.on = False
        for name, (res, args) in L._SIGNATURES.items():
            fn = getattr(self.lib, name, None)
            if fn is None or res is not L.c_int or not args or args[-1] is not L.c_void_p or name.startswith("a2f_debug"):
                continue
            setattr(self.lib, name, self._wrap(name, fn))

    def _wrap(self, name, fn):
        def call(*a):
            if not self.on:
                return fn(*a)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*a)
            e.record()
            tag = name
            if name == "a2f_gemm":
                g = a[0]._obj
                tag = f"a2f_gemm[{'tc' if a[1] == L.TCGEN05 else 'simt'} M{g.M} N{g.N} K{g.K}]"
            elif name == "a2f_gemm_wgrad":
                g = a[0]._obj
                tag = f"a2f_gemm_wgrad[{'tc' if a[1] == L.TCGEN05 else 'simt'} M{g.M} N{g.N} K{g.K}x{max(1, g.n_seg)}]"
            self.rows.append((tag, s, e))
            return rc
        return call


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (32 if mode == "infer" else 8)
    fps = int(sys.argv[3]) if len(sys.argv) > 3 else (30 if mode == "infer" else 60)
    rec = Recorder()
    dev = torch.device("cuda:0")
    if mode in ("a2m_train", "a2m"):
        from a2f_b200 import features
        B = int(sys.argv[2]) if len(sys.argv) > 2 else (128 if mode == "a2m_train" else 64)
        am = modules.Audio2Mesh(15069, 12)
        am.load_state_dict(ow.make_state_dict("audio2mesh", 12), strict=True)
        am = am.to(dev)
        ext = features.MFCCExtractor(22000, 32, 52, 440, None, 1024).to(dev).set_precision("bf16")
        tp = oin.batch_templates(B, 1)
        x, oh, tpl = oin.speech_like_windows(B, seed=1).to(dev), oin.one_hot(B, 12, 1).to(dev), tp.to(dev)
        if mode == "a2m_train":
            t = tr.ConvModelTrainer(am, ext, lr=1e-4)
            gt = oin.gt_like((B, 5023, 3), tp, 2).to(dev)
            step = lambda: t.step(x, oh, tpl, gt)                        # noqa: E731
        else:
            am = am.eval().set_precision("bf16")

            def step():
                with torch.no_grad():
                    return am(ext(x), oh, tpl)
        fps, T = 0, 1
        return run(rec, step, mode, B, fps, T)
    m = modules.Faceformer(15069, 12)
    m.load_state_dict(ow.make_state_dict("faceformer", 13), strict=True)
    m = m.to(dev).eval().set_precision("bf16")
    n = 80000
    T = n * fps // 16000
    tp = oin.batch_templates(B, 1, scale=100.0)
    audio, oh, tpl = oin.audio(B, n, 1).to(dev), oin.one_hot(B, 12, 1).to(dev), tp.to(dev)
    if mode == "train":
        t = tr.FaceformerTrainer(m, fps=fps)
        gt = oin.gt_like((B, T, 5023, 3), tp[:, None], 2, scale=100.0).to(dev)
        step = lambda: t.step(audio, oh, tpl, gt)                    # noqa: E731
    else:
        def step():
            with torch.no_grad():
                return m(audio, oh, tpl, fps=fps)
    return run(rec, step, mode, B, fps, T)


def run(rec, step, mode, B, fps, T):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(1.9e9 * (0.06 if mode in ("infer", "a2m") else 0.25)))     # park the GPU while the host enqueues
    rec.on = True
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    step()
    e0.record()
    rec.on = False
    torch.cuda.synchronize()
    agg = OrderedDict()
    tot = 0.0
    for tag, s, e in rec.rows:
        us = s.elapsed_time(e) * 1e3
        a = agg.setdefault(tag, [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    wall = s0.elapsed_time(e0) * 1e3
    print(f"# {mode} step, B={B}, 5 s audio, {fps} fps (T={T}); CUDA-event pairs around every C-ABI call, host enqueued ahead")
    print(f"# step wall time on device {wall:.1f} us; sum of calls {tot:.1f} us; {len(rec.rows)} calls")
    print(f"{'entry point':64s} {'calls':>5s} {'total_us':>10s} {'share':>7s} {'us/call':>8s}")
    for tag, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{tag:64s} {c:5d} {us:10.1f} {100 * us / tot:6.1f}% {us / c:8.1f}")


if __name__ == "__main__":
    main()
