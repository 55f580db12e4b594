"""Host-side logic of the data-parallel trainer (audio2face-pytorch_b200/trainer.py FlatBuffers) on CPU:
layout of the flat buffers and the staged gradient all-reduce over a world_size-2 gloo group.  (The compute path is
CUDA-only; its N>1 run is tests/dp_check.py under torchrun on the GPU box.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from oracle import weights as ow


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(5, 7)          # stage 1
        self.b = nn.Linear(7, 3)          # stage 0
        self.c = nn.Parameter(torch.ones(130))   # stage 2
        self.unused = nn.Parameter(torch.ones(4))


def _stage(name):
    return {"a": 1, "b": 0, "c": 2, "u": 2}[name[0]]


def test_flat_layout_aliases_parameters():
    from a2f_b200.trainer import ALIGN, FlatBuffers
    torch.manual_seed(0)
    m = _Toy()
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    fb = FlatBuffers(m.named_parameters(), _stage, 3, skip=("unused",))
    names = [e[0] for e in fb.entries]
    assert names == ["b.weight", "b.bias", "a.weight", "a.bias", "c"]          # stage-major, original order inside
    assert all(e[2] % ALIGN == 0 for e in fb.entries)
    assert fb.stage_ranges[0][0] == 0 and fb.stage_ranges[-1][1] == fb.total
    assert all(fb.stage_ranges[i][1] == fb.stage_ranges[i + 1][0] for i in range(2))
    for name, p, off, n, st in fb.entries:
        assert torch.equal(p.detach(), before[name])                            # values preserved
        assert p.data_ptr() == fb.params[off:].data_ptr() and p.grad.data_ptr() == fb.grads[off:].data_ptr()
        lo, hi = fb.stage_ranges[st]
        assert lo <= off and off + n <= hi
    assert m.unused.grad is None and m.unused.data_ptr() != fb.params.data_ptr()
    fb.params.mul_(2.0)                                                         # an "optimizer" writing the flat buffer
    assert torch.equal(m.a.weight.detach(), 2 * before["a.weight"])
    v0 = m.a.weight._version
    fb.bump_versions()
    assert m.a.weight._version > v0
    assert list(m.state_dict().keys()) == ["c", "unused", "a.weight", "a.bias", "b.weight", "b.bias"]


def test_faceformer_stage_map_covers_every_parameter():
    from a2f_b200 import training
    shapes = ow.faceformer_shapes()
    seen = set()
    for k in shapes:
        if k == "PPE.pe":
            continue
        st = training.grad_stage_of(k)
        assert 0 <= st < training.N_GRAD_STAGES
        seen.add(st)
    assert seen == set(range(training.N_GRAD_STAGES))
    assert training.grad_stage_of("audio_encoder.encoder.layers.11.attention.q_proj.weight") == 1
    assert training.grad_stage_of("audio_encoder.encoder.layers.0.final_layer_norm.bias") == 12
    assert training.grad_stage_of("vertice_map_r.weight") == 0
    assert training.grad_stage_of("audio_encoder.feature_extractor.conv_layers.0.conv.weight") == 13


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from a2f_b200.trainer import FlatBuffers
        torch.manual_seed(rank)                       # replicas start DIFFERENT; broadcast must fix that
        m = _Toy()
        fb = FlatBuffers(m.named_parameters(), _stage, 3, skip=("unused",))
        fb.broadcast_params(0)
        ref = _Toy_params_seed0()
        ok_bcast = all(torch.equal(p.detach(), ref[n]) for n, p, *_ in fb.entries)
        fb.zero_grads()
        for name, p, off, n, st in fb.entries:        # local "gradient": (rank+1) * (stage+1)
            p.grad.fill_(float((rank + 1) * (st + 1)))
        fb.all_reduce_stage(0)                        # started "during the backward"
        fb.all_reduce_stage(2)
        fb.finish_all_reduce()                        # stage 1 is reduced here
        want = sum(r + 1 for r in range(world))
        ok_sum = all(bool((p.grad == want * (st + 1)).all()) for _, p, _, _, st in fb.entries)
        pad_zero = float(fb.grads.sum()) == sum(want * (st + 1) * n for _, _, _, n, st in fb.entries)
        q.put((rank, ok_bcast, ok_sum, pad_zero))
    finally:
        dist.destroy_process_group()


def _worker_bf16_buckets(rank, world, port, q):
    """bf16 wire + buckets: stages 0 and 1 travel together (sent when BOTH are final), stage 2 alone."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from a2f_b200.trainer import FlatBuffers
        torch.manual_seed(0)
        m = _Toy()
        fb = FlatBuffers(m.named_parameters(), _stage, 3, skip=("unused",), wire="bf16", buckets=[[0, 1], [2]])
        fb.zero_grads()
        for name, p, off, n, st in fb.entries:
            p.grad.fill_(float((rank + 1) * (st + 1)) + 0.001)          # 0.001 is below bf16 resolution at these magnitudes
        fb.all_reduce_stage(0)
        sent_early = fb._sent[0]                                        # must wait for stage 1
        fb.all_reduce_stage(1)
        sent_both = fb._sent[0] and not fb._sent[1]
        fb.finish_all_reduce()
        red = fb.reduced_grads()
        want = sum(r + 1 for r in range(world))
        ok = red.dtype == torch.bfloat16 and all(
            bool((red[off:off + n].float() == float(torch.tensor(want * (st + 1) + 0.002).bfloat16())).all())
            or bool((red[off:off + n].float() - want * (st + 1)).abs().max() < 0.05) for _, _, off, n, st in fb.entries)
        local_kept = all(bool((p.grad == float((rank + 1) * (st + 1)) + 0.001).all()) for _, p, _, _, st in fb.entries)
        q.put((rank, (not sent_early) and sent_both, ok, local_kept))
    finally:
        dist.destroy_process_group()


def test_bucketed_bf16_wire_all_reduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_bf16_buckets, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True, True), (1, True, True, True)]


class _ExplicitFn(torch.autograd.Function):
    """The pattern of training.FaceformerTrainFn / conv_training.ConvModelTrainFn on a toy model: an explicit backward that
    accumulates through training._grad, run inside training.collect_grads, every parameter an input of the Function."""

    @staticmethod
    def forward(ctx, model, x, *params):
        ctx.model = model
        ctx.save_for_backward(x)
        return x @ model.w.detach() + model.b.detach()

    @staticmethod
    def backward(ctx, dout):
        from a2f_b200 import training
        (x,) = ctx.saved_tensors
        m = ctx.model
        with training.collect_grads(list(m.parameters())) as sink:
            training._grad(m.w).add_(x.t() @ dout)           # what the kernels do: accumulate into the buffer handed out
            training._grad(m.b).add_(dout.sum(0))
        return (None, None) + sink.grads()


class _ExplicitToy(nn.Module):
    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.randn(4, 3))
        self.b = nn.Parameter(torch.zeros(3))
        self.unused = nn.Parameter(torch.ones(2))            # never written by the backward (masked_spec_embed's role)

    def forward(self, x):
        return _ExplicitFn.apply(self, x, *self.parameters())


def _worker_ddp(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        m = _ExplicitToy()
        ddp = nn.parallel.DistributedDataParallel(m, find_unused_parameters=True)
        ok = True
        for it in range(3):                                  # iteration 2+ raised under the one-anchor scheme
            torch.manual_seed(100 + 10 * it + rank)
            x = torch.randn(5, 4)
            for p_ in m.parameters():
                p_.grad = None
            ddp(x).square().sum().backward()
            xs = []
            for r in range(world):
                torch.manual_seed(100 + 10 * it + r)
                xs.append(torch.randn(5, 4))
            gw = sum(2 * xr.t() @ (xr @ m.w.detach() + m.b.detach()) for xr in xs) / world
            gb = sum(2 * (xr @ m.w.detach() + m.b.detach()).sum(0) for xr in xs) / world
            ok = ok and torch.allclose(m.w.grad, gw, atol=1e-5) and torch.allclose(m.b.grad, gb, atol=1e-5)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_autograd_glue_is_ddp_compatible_gloo_world2():
    """the module-level autograd path must hand every gradient to AccumulateGrad so that torch DDP's
    reducer averages it across ranks (the reference trains under Lightning's implicit DDP)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ddp, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_default_buckets():
    from a2f_b200.trainer import default_buckets
    assert default_buckets(14, 4) == [[0, 1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [13]]
    assert default_buckets(1, 4) == [[0]]
    assert default_buckets(14, 14) == [[i] for i in range(14)]
    for n in range(1, 20):
        for k in range(1, 8):
            b = default_buckets(n, k)
            assert [s for x in b for s in x] == list(range(n)) and len(b) <= max(1, k)


def _Toy_params_seed0():
    torch.manual_seed(0)
    return {k: v.detach().clone() for k, v in _Toy().named_parameters()}


def test_staged_all_reduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True, True), (1, True, True, True)]
