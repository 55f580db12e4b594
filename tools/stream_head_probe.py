"""Rollout + vertex head at the bench shape (B=32, T=150), eager, CUDA events: one after the other vs concurrently
(a2f_decoder_rollout_stream on the current stream, a2f_vertex_head_stream on a side stream), plus the two streaming kernels
serialised on one stream (cost of the hand-over inside the rollout) and smaller head grids.
    python tools/stream_head_probe.py"""
import os
import sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from a2f_b200 import modules, ops, lib as L
from oracle import inputs as oin, weights as ow
dev = torch.device("cuda:0")
B, T = 32, 150
sd = ow.make_state_dict("faceformer", seed=13)
m = modules.Faceformer(15069, 12); m.load_state_dict(sd, strict=True); m = m.to(dev).eval().set_precision("bf16")
P = m._packed()
g = torch.Generator().manual_seed(0)
memory = torch.randn(B * T, 64, generator=g).to(dev)
oh = oin.one_hot(B, 12, 100).to(dev)
tmpl = oin.batch_templates(B, 100, scale=100.0).reshape(B, -1).contiguous().float().to(dev)
out = torch.empty((B * T, 15069), dtype=torch.float32, device=dev)
w3 = m._head_operand(m.vertice_map_r.weight, 64)
bias = m.vertice_map_r.bias.detach()
def seq():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    D = ops.decoder_rollout(P["dec"][0], memory, oh, m.period, B, T, memory_is_ca=True)
    e[1].record()
    m._vertex_head(D.view(B * T, 64), m.vertice_map_r.weight, m.vertice_map_r.bias, tmpl, T, 64, out=out)
    e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]) * 1e3, e[1].elapsed_time(e[2]) * 1e3, e[0].elapsed_time(e[2]) * 1e3
def par(concurrent=True, reserve=None):
    cur = torch.cuda.current_stream()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    # re-implementation of ops.rollout_and_head_stream with an event behind the rollout launch
    lib = L.load()
    import ctypes as C
    rows = lib.a2f_vertex_head_stream_rows(B, T)
    nbytes = lib.a2f_decoder_workspace_bytes(B, T)
    ws = torch.empty((nbytes + 15) // 16 * 4, dtype=torch.float32, device=dev)
    D = torch.empty((B, T, 64), dtype=torch.float32, device=dev)
    z3 = torch.zeros((rows, 192), dtype=torch.bfloat16, device=dev)
    done = torch.zeros(T, dtype=torch.int32, device=dev)
    side = ops._StreamPair.side(dev) if concurrent else cur
    if concurrent: side.wait_stream(cur)
    L.check(lib.a2f_decoder_rollout_stream(C.byref(P["dec"][0]), memory.data_ptr(), oh.data_ptr(), oh.shape[1], m.period,
                                           D.data_ptr(), B, T, ws.data_ptr(), ws.numel() * 4, z3.data_ptr(), done.data_ptr(), cur.cuda_stream), "r")
    e[1].record()
    with torch.cuda.stream(side):
        L.check(lib.a2f_vertex_head_stream(z3.data_ptr(), w3.data_ptr(), 192, bias.data_ptr(), tmpl.data_ptr(), B, T, 15069,
                                           out.data_ptr(), done.data_ptr(), B if reserve is None else reserve, side.cuda_stream), "h")
    if concurrent: cur.wait_stream(side)
    e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]) * 1e3, e[1].elapsed_time(e[2]) * 1e3, e[0].elapsed_time(e[2]) * 1e3
for name, fn in (("sequential", seq), ("streamed", par), ("stream-kernels, one stream", lambda: par(False)),
                 ("streamed, 74 head CTAs", lambda: par(True, 74)), ("streamed, 32 head CTAs", lambda: par(True, 116)), ("sequential", seq)):
    for _ in range(3): fn()
    r = sorted(fn() for _ in range(9))[4]
    print(f"{name:28s} rollout(+setup) {r[0]:7.1f} us   rest {r[1]:7.1f} us   total {r[2]:7.1f} us")
