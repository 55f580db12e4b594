"""Where one step of the autoregressive decoder goes: clock cycles between the block barriers of decoder_rollout_kernel
(a2f_debug_set_decoder_timing), for CTA 0, per phase and per step.

    python tools/decoder_phases.py [B] [T]            (default 32 x 150: the bench shape; 8 x 300 = the training shape)

The instrumented instantiation is slower than the production kernel (its 12 counters spill: ~7.7 k cycles per step against
~5.3 k), so read the SHARES: attention ~48 % (its fixed part -- query load, score loop set-up, the 4-warp merge -- dominates;
the share barely moves between T = 40 and T = 300), the two phases after the FFN's first layer ~37 %, and the matvecs themselves
little: the step is a chain of short dependent phases, each dominated by its own issue latency, not by the key count.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from a2f_b200 import lib as L, modules, ops
from oracle import inputs as oin, weights as ow

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 150
dev = torch.device("cuda:0")
m = modules.Faceformer(15069, 12)
m.load_state_dict(ow.make_state_dict("faceformer", 13), strict=True)
m = m.to(dev).eval().set_precision("bf16")
P = m._packed()
mem = torch.randn(B, T, 64, device=dev)
oh = oin.one_hot(B, 12, 1).to(dev)
lib = L.load()
buf = torch.zeros(16, dtype=torch.int64, device=dev)
names = ["attention: q load, scores (q.k, bias), row max candidates", "attention: warp max (redux), exp, row sums",
         "attention: P.V over the warp's keys, shuffle reductions, partials to shared memory",
         "attention: wait for the other 3 warps of the head (named barrier)", "attention: merge of the 4 warp partials (16 threads)",
         "wait at the block barrier that ends the attention phase", "out_proj + residual (+ barrier)",
         "LN1, +cross-attn, LN2, linear1 + ReLU (+ barrier)", "linear2 + residual (+ barrier)", "LN3 (each warp), copy of d_i",
         "feedback / next in-projection matvec, stores", "wait at the block barrier that ends the step"]
for label, fn in (("inference rollout", lambda: ops.decoder_rollout(P["dec"][0], mem, oh, 60, B, T, memory_is_ca=True)),
                  ("training rollout (saves activations)", lambda: ops.decoder_rollout_train(P["dec"][0], mem, oh, 60, B, T))):
    fn()
    torch.cuda.synchronize()
    L.check(lib.a2f_debug_set_decoder_timing(buf.data_ptr()))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    L.check(lib.a2f_debug_set_decoder_timing(None))
    c = buf.cpu().tolist()
    tot = sum(c[:12])
    print(f"# {label}: B={B}, T={T}, kernel {1e3 * s.elapsed_time(e):.1f} us; thread 0 of CTA 0: {tot / c[12]:.0f} cycles per step")
    for n, v in zip(names, c[:12]):
        print(f"  {v / c[12]:8.0f} cycles/step  {100 * v / tot:5.1f} %  {n}")
    buf.zero_()
