"""GPU parity of the backward-pass kernels (csrc/train.cu, attention backward) against torch autograd in fp64 on the
CPU.  fp32 instantiations: tight tolerances; bf16 instantiations: bf16 rounding of the stored operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(shape, generator=g)


def _err(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max())


def test_act_fwd_bwd(a2f_lib, dev):
    from a2f_b200 import ops, lib as L
    z = _rand((1000, 77), 1, 2.0)
    dy = _rand((1000, 77), 2)
    r = _rand((1000, 77), 3)
    for act, fn in ((L.ACT_GELU, F.gelu), (L.ACT_RELU, F.relu), (L.ACT_TANH, torch.tanh)):
        zz = z.double().requires_grad_(True)
        y = fn(zz)
        y.backward(dy.double())
        got = ops.act_fwd(z.to(dev), act, resid=r.to(dev))
        assert _err(got, y.detach() + r.double()) < 2e-6
        gz = ops.act_bwd(dy.to(dev), z.to(dev), act)
        assert _err(gz, zz.grad) < 2e-6
    g16 = ops.act_fwd(z.to(dev).bfloat16(), L.ACT_GELU)
    assert _err(g16, F.gelu(z.bfloat16().double())) < 2e-2


def test_cast_rows_transpose_colsum(a2f_lib, dev):
    from a2f_b200 import ops
    x = _rand((37, 15069), 4)
    o = ops.cast_rows(x.to(dev), torch.bfloat16, 15072)
    assert o.shape == (37, 15072) and float(o[:, 15069:].abs().max()) == 0.0
    assert _err(o[:, :15069], x.bfloat16()) == 0.0
    w = _rand((130, 70), 5)
    assert _err(ops.transpose_cast(w.to(dev), torch.float32), w.T) == 0.0
    w3 = _rand((64, 40, 3), 6)                    # conv weight [co, ci, taps]: tap 1 transposed -> [ci, co]
    t1 = ops.transpose_cast(w3.to(dev), torch.bfloat16, R=64, Cc=40, ld_r=120, ld_c=3, offset=1)
    assert _err(t1, w3[:, :, 1].T.bfloat16()) == 0.0
    acc = torch.ones(70, device=dev)
    ops.colsum(w.to(dev), acc)
    assert _err(acc, 1.0 + w.double().sum(0)) < 1e-4
    acc16 = torch.zeros(70, device=dev)
    ops.colsum(w.to(dev).bfloat16(), acc16)
    assert _err(acc16, w.bfloat16().double().sum(0)) < 1e-3


@pytest.mark.parametrize("C,rows", [(768, 301), (512, 77)])
def test_layernorm_bwd(a2f_lib, dev, C, rows):
    from a2f_b200 import ops
    x = _rand((rows, C), 7, 2.0) + 0.5
    dy = _rand((rows, C), 8)
    gamma = torch.rand(C, generator=torch.Generator().manual_seed(9)) + 0.5
    xx = x.double().requires_grad_(True)
    gg = gamma.double().requires_grad_(True)
    bb = torch.zeros(C, dtype=torch.double, requires_grad=True)
    F.layer_norm(xx, (C,), gg, bb, 1e-5).backward(dy.double())
    dg, db, dbias = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dx = ops.layernorm_bwd(dy.to(dev), x.to(dev), gamma.to(dev), dg, db, dbias)
    assert _err(dx, xx.grad) < 2e-5
    assert _err(dg, gg.grad) < 2e-4
    assert _err(db, bb.grad) < 2e-4
    assert _err(dbias, xx.grad.sum(0)) < 2e-4
    dg16, db16 = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    dx16 = ops.layernorm_bwd(dy.to(dev).bfloat16(), x.to(dev).bfloat16(), gamma.to(dev), dg16, db16)
    assert _err(dx16, xx.grad) < 6e-2


@pytest.mark.parametrize("S,T", [(49, 60), (249, 300), (249, 150), (10, 1)])
def test_interp_ln_bwd(a2f_lib, dev, S, T):
    from a2f_b200 import ops
    from oracle import ref_models as orm
    x = _rand((2, S, 512), 10)
    dy = _rand((2, T, 512), 11)
    gamma = torch.rand(512, generator=torch.Generator().manual_seed(12)) + 0.5
    # fp32 reference: the interpolation weights follow ATen's fp32 formula, an fp64 reference differs
    # from it by ~1e-4 in the weights themselves
    xx = x.clone().requires_grad_(True)
    gg = gamma.clone().requires_grad_(True)
    bb = torch.zeros(512, requires_grad=True)
    F.layer_norm(orm.linear_interpolation(xx, T), (512,), gg, bb, 1e-5).backward(dy)
    dg, db = torch.zeros(512, device=dev), torch.zeros(512, device=dev)
    din = ops.interp_ln_bwd(x.to(dev), dy.to(dev), gamma.to(dev), dg, db)
    assert _err(din, xx.grad) < 5e-5
    assert _err(dg, gg.grad) < 3e-4
    assert _err(db, bb.grad) < 3e-4


@pytest.mark.parametrize("n", [4000, 11205])
def test_conv0_bwd(a2f_lib, dev, n):
    from a2f_b200 import ops
    from oracle import inputs as oin, ref_models as orm
    B = 2
    a = oin.audio(B, n, 8)
    an = torch.stack([orm.processor_normalize(a[b]) for b in range(B)]).double()
    w = _rand((512, 1, 10), 13, 0.3)
    gamma = torch.rand(512, generator=torch.Generator().manual_seed(14)) + 0.5
    beta = _rand((512,), 15, 0.1)
    ww, gg, bb = (t.double().requires_grad_(True) for t in (w, gamma, beta))
    y = F.gelu(F.group_norm(F.conv1d(an[:, None], ww, stride=5), 512, gg, bb, 1e-5)).transpose(1, 2)   # [B,L0,512]
    da = _rand(tuple(y.shape), 16, 0.05)
    y.backward(da.double())
    ad = a.to(dev)
    st = ops.audio_stats(ad)
    wd, gd, bd = w.reshape(512, 10).to(dev).contiguous(), gamma.to(dev), beta.to(dev)
    out, ws = ops.conv0_gn_gelu_train(ad, st, wd, gd, bd, torch.float32)
    assert _err(out, y.detach()) < 3e-5
    dw, dg, db = torch.zeros((512, 10), device=dev), torch.zeros(512, device=dev), torch.zeros(512, device=dev)
    ops.conv0_bwd(ad, st, wd, gd, bd, ws, da.to(dev), dw, dg, db)
    scale = float(ww.grad.abs().max())
    assert _err(dw, ww.grad.reshape(512, 10)) < 2e-4 * max(1.0, scale)
    assert _err(dg, gg.grad) < 2e-4 * max(1.0, float(gg.grad.abs().max()))
    assert _err(db, bb.grad) < 2e-4 * max(1.0, float(bb.grad.abs().max()))


@pytest.mark.parametrize("B,T", [(1, 60), (2, 150)])
def test_posconv_backward(a2f_lib, dev, B, T):
    """input gradient, weight gradient (packed layout) and weight-norm backward of the positional conv."""
    from a2f_b200 import ops, lib as L
    h = _rand((B, T, 768), 41)
    g = (0.5 + torch.rand(1, 1, 128, generator=torch.Generator().manual_seed(42)))
    v = _rand((768, 48, 128), 43, (48 * 128) ** -0.5)
    bias = _rand((768,), 44, 0.05)
    dout = _rand((B, T, 768), 45)
    hh, gg, vv, bbias = (t.double().requires_grad_(True) for t in (h, g, v, bias))
    wfull = torch._weight_norm(vv, gg, 2)
    pc = F.conv1d(hh.transpose(1, 2), wfull, bbias, padding=64, groups=16)[:, :, :-1].transpose(1, 2)
    out = hh + F.gelu(pc)
    out.backward(dout.double())
    for backend, dt, tol in ((L.SIMT_F32, torch.float32, 1e-4), (L.TCGEN05, torch.bfloat16, 8e-2)):
        wf, wb = ops.pack_posconv_weights_train(g.reshape(-1).to(dev), v.to(dev), dt)
        hd = h.to(dev).to(dt).contiguous()
        pcd = ops.posconv_pre(hd, wf, bias.to(dev), B, T, backend)
        assert _err(pcd, pc.detach()) < tol
        outd = ops.act_fwd(pcd, L.ACT_GELU, resid=hd)
        assert _err(outd, out.detach()) < tol
        doutd = dout.to(dev).to(dt)
        dpc = ops.act_bwd(doutd, pcd, L.ACT_GELU)
        dh = ops.posconv_dgrad(dpc, wb, doutd, B, T, backend)
        assert _err(dh, hh.grad) < tol * 3
        dwp = ops.posconv_wgrad(dpc, hd, B, T, backend)                      # [16][48][128][48]
        dv, dg = torch.zeros((768, 48, 128), device=dev), torch.zeros(128, device=dev)
        ops.weight_norm_bwd(dwp, v.to(dev), g.reshape(-1).to(dev), dv, dg)
        rel = lambda a, b: float((a.double().cpu() - b).norm() / b.norm())
        assert rel(dv, vv.grad) < (1e-4 if dt == torch.float32 else 3e-2)
        assert rel(dg, gg.grad.reshape(-1)) < (1e-4 if dt == torch.float32 else 3e-2)


@pytest.mark.parametrize("B,T", [(1, 60), (2, 150), (1, 333)])
def test_mha_bwd(a2f_lib, dev, B, T):
    from a2f_b200 import ops
    qkv = _rand((B, T, 2304), 9)
    dout = _rand((B, T, 768), 19)
    for dt, tol in ((torch.float32, 5e-5), (torch.bfloat16, 6e-2)):
        qd = qkv.to(dev).to(dt)
        dd = dout.to(dev).to(dt)
        qq = qd.double().cpu().requires_grad_(True)
        q, k, v = [t.view(B, T, 12, 64).transpose(1, 2) for t in qq.split(768, dim=-1)]
        o = (torch.softmax(q @ k.transpose(2, 3) * 0.125, -1) @ v).transpose(1, 2).reshape(B, T, 768)
        o.backward(dd.double().cpu())
        out = torch.empty((B, T, 768), device=dev, dtype=dt)
        lse = torch.empty((B, 12, T), device=dev)
        ops.mha_lse(qd, out, lse, B, T)
        want_lse = torch.logsumexp(q @ k.transpose(2, 3) * 0.125, -1).detach()
        assert _err(lse, want_lse) < (1e-4 if dt == torch.float32 else 2e-2)
        dqkv = ops.mha_bwd(qd, out, dd, lse, B, T)
        torch.cuda.synchronize()
        assert _err(dqkv, qq.grad) < tol, dt


def test_adam_step_matches_torch(a2f_lib, dev):
    from a2f_b200 import ops
    n = 100003
    p0, g = _rand((n,), 30), _rand((n,), 31, 0.1)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-4, weight_decay=1e-5)
    p, m, v = p0.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for step in range(1, 4):
        ref.grad = g.clone() * step
        opt.step()
        ops.adam_step(p, (g * step * 2).to(dev), m, v, 1e-4, 0.9, 0.999, 1e-8, 1e-5, step, grad_scale=0.5)
    assert _err(p, ref.detach()) < 1e-6


def test_adam_step_bf16_gradient_matches_torch(a2f_lib, dev):
    """a2f_adam_step_bf16g: the gradient as it comes off the bf16 all-reduce; everything else fp32.  Reference = torch Adam
    fed the SAME bf16-rounded gradient; n is not a multiple of 4 (vector body + scalar tail)."""
    from a2f_b200 import ops
    n = 100003
    p0, g = _rand((n,), 32), _rand((n,), 33, 0.1)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-4, weight_decay=1e-5)
    p, m, v = p0.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for step in range(1, 4):
        g16 = (g * step * 2).bfloat16()
        ref.grad = g16.float() * 0.5
        opt.step()
        ops.adam_step(p, g16.to(dev), m, v, 1e-4, 0.9, 0.999, 1e-8, 1e-5, step, grad_scale=0.5)
    assert _err(p, ref.detach()) < 1e-6


def test_pack_cross_attention_matches_fp64_fold(a2f_lib, dev):
    """out_proj(v_proj(audio_feature_map(h))) folded on the device (no library GEMM) vs the same fold in torch fp64."""
    from a2f_b200 import ops
    in_w, in_b = _rand((192, 64), 40, 0.2), _rand((192,), 41, 0.1)
    wo, bo = _rand((64, 64), 42, 0.2), _rand((64,), 43, 0.1)
    wa, ba = _rand((64, 768), 44, 0.05), _rand((64,), 45, 0.1)
    wv, bv = in_w[128:192].double(), in_b[128:192].double()
    W_ref = wo.double() @ (wv @ wa.double())
    b_ref = wo.double() @ (wv @ ba.double() + bv) + bo.double()
    for dt, tol in ((torch.float32, 2e-7), (torch.bfloat16, 4e-3)):
        W = torch.empty((64, 768), dtype=dt, device=dev)
        b = torch.empty(64, dtype=torch.float32, device=dev)
        ops.pack_cross_attention(in_w.to(dev), in_b.to(dev), wo.to(dev), bo.to(dev), wa.to(dev), ba.to(dev), W, b)
        assert float((W.double().cpu() - W_ref).abs().max()) <= tol * float(W_ref.abs().max())
        assert float((b.double().cpu() - b_ref).abs().max()) <= 2e-7 * float(b_ref.abs().max())
