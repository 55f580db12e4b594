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
