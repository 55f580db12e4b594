"""GPU parity of the backward-pass GEMM modes (explicit K segments, activation-backward epilogue, a2f_gemm_wgrad)
on both back ends against fp64 CPU references built from torch autograd / matmuls."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(shape, dev, seed, dtype=torch.float32, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(shape, generator=g)).to(dev).to(dtype)


def _backends():
    from a2f_b200 import lib as L
    return ((L.SIMT_F32, torch.float32, 3e-5), (L.TCGEN05, torch.bfloat16, 4e-2))


@pytest.mark.parametrize("M,N,K", [(300, 768, 3072), (77, 96, 64), (513, 512, 512)])
def test_dgrad_with_dact_epilogue(a2f_lib, dev, M, N, K):
    """dX = (dY W) * gelu'(z): the data-gradient GEMM with the activation backward fused in."""
    from a2f_b200 import ops, lib as L
    for backend, dt, tol in _backends():
        dy = _rand((M, K), dev, 1, dt)
        wt = _rand((N, K), dev, 2, dt, scale=K ** -0.5)     # = W^T of the forward layer, [in, out]
        z = _rand((M, N), dev, 3, dt)
        out = torch.empty((M, N), device=dev, dtype=dt)
        ops.gemm(dy, wt, out, act=L.ACT_GELU, resid=z, resid_mode=L.RESID_DACT, backend=backend)
        torch.cuda.synchronize()
        zz = z.double().cpu().requires_grad_(True)
        F.gelu(zz).backward(dy.double().cpu() @ wt.double().cpu().T)
        err = float((out.double().cpu() - zz.grad).abs().max())
        assert err < tol, (backend, err)


@pytest.mark.parametrize("taps,L_in,B", [(3, 41, 2), (3, 40, 3), (2, 38, 2), (2, 39, 1), (3, 515, 2)])
def test_conv1d_dgrad_gather(a2f_lib, dev, taps, L_in, B):
    """Data gradient of Conv1d(512->512, k taps, stride 2) over channels-last activations as gather GEMMs
    (even / odd input rows), with gelu'(pre-activation of the producing layer) fused; vs torch autograd."""
    from a2f_b200 import ops, lib as L
    C = 512
    L_out = (L_in - taps) // 2 + 1
    for backend, dt, tol in _backends():
        w = _rand((C, C, taps), dev, 5, scale=(C * taps) ** -0.5)           # [co, ci, tap]
        dpre = _rand((B, L_out, C), dev, 6, dt)                               # gradient wrt this conv's output
        zprev = _rand((B, L_in, C), dev, 7, dt)                               # pre-activation of the previous layer
        wq = w.to(dt)
        # reference
        zz = zprev.double().cpu().requires_grad_(True)
        y = F.conv1d(F.gelu(zz).transpose(1, 2), wq.double().cpu(), stride=2).transpose(1, 2)
        y.backward(dpre.double().cpu())
        want = zz.grad
        dx = torch.zeros((B, L_in, C), device=dev, dtype=dt)
        if taps == 3:
            # even rows l=2u: dpre[u-1] W2 + dpre[u] W0 ; odd rows l=2u+1: dpre[u] W1
            w_even = torch.cat([wq[:, :, 2].T, wq[:, :, 0].T], dim=1).contiguous()      # [ci, 2*co]
            w_odd = wq[:, :, 1].T.contiguous()
            U = min(L_out + 1, (L_in + 1) // 2)
            ops.gemm(dpre, w_even, dx, backend=backend, M=B * U, K=2 * C, N=C, a_row_stride=C, a_batch_stride=L_out * C,
                     rows_per_batch=U, a_rows=L_out, segs=[(-1, 0), (0, 0)], ldc=2 * C, c_batch_stride=L_in * C,
                     act=L.ACT_GELU, resid=zprev, resid_mode=L.RESID_DACT, ldr=2 * C, r_batch_stride=L_in * C)
            ops.gemm(dpre, w_odd, dx, backend=backend, M=B * L_out, K=C, N=C, a_row_stride=C, a_batch_stride=L_out * C,
                     rows_per_batch=L_out, ldc=2 * C, c_batch_stride=L_in * C, c_offset=C,
                     act=L.ACT_GELU, resid=zprev, resid_mode=L.RESID_DACT, ldr=2 * C, r_batch_stride=L_in * C, r_offset=C)
        else:
            w_both = torch.cat([wq[:, :, 0].T, wq[:, :, 1].T], dim=0).contiguous()      # [(tap,ci), co]
            ops.gemm(dpre, w_both, dx, backend=backend, M=B * L_out, K=C, N=2 * C, a_row_stride=C, a_batch_stride=L_out * C,
                     rows_per_batch=L_out, ldc=2 * C, c_batch_stride=L_in * C,
                     act=L.ACT_GELU, resid=zprev, resid_mode=L.RESID_DACT, ldr=2 * C, r_batch_stride=L_in * C)
        torch.cuda.synchronize()
        err = float((dx.double().cpu() - want).abs().max())
        assert err < tol, (backend, taps, err)


@pytest.mark.parametrize("M,N,K", [(300, 768, 3072), (2400, 3072, 768), (77, 96, 64), (1000, 15069, 64), (130, 64, 72)])
def test_wgrad_plain(a2f_lib, dev, M, N, K):
    from a2f_b200 import ops
    for backend, dt, tol in _backends():
        if backend == 1 and (N % 8 or K % 8):
            Np = (N + 7) // 8 * 8
        else:
            Np = N
        dy_full = _rand((M, Np), dev, 11, dt, scale=M ** -0.5)
        x = _rand((M, K), dev, 12, dt)
        dw = torch.ones((N, K), device=dev)                                    # accumulate semantics: starts at 1
        ops.gemm_wgrad(dy_full, x, dw, backend=backend, N=N)
        torch.cuda.synchronize()
        want = 1.0 + dy_full[:, :N].double().cpu().T @ x.double().cpu()
        err = float((dw.double().cpu() - want).abs().max())
        assert err < tol, (backend, err)


@pytest.mark.parametrize("taps,L_in,B", [(3, 41, 2), (2, 38, 3), (3, 1030, 2)])
def test_conv1d_wgrad(a2f_lib, dev, taps, L_in, B):
    """dW[co, tap, ci] = sum_{b,t} dpre[b,t,co] x[b, 2t+tap, ci] with x read in place (row stride 2C view)."""
    from a2f_b200 import ops
    C = 512
    L_out = (L_in - taps) // 2 + 1
    L_pad = L_in + (L_in % 2)
    for backend, dt, tol in _backends():
        x = _rand((B, L_pad, C), dev, 21, dt)
        dpre = _rand((B, L_out, C), dev, 22, dt, scale=(B * L_out) ** -0.5)
        dw = torch.zeros((C, taps * C), device=dev)
        segs = [(0, 0), (0, C), (1, 0)][:taps]
        ops.gemm_wgrad(dpre, x, dw, backend=backend, M=B * L_out, N=C, K=C, dy_row_stride=C, dy_batch_stride=L_out * C,
                       x_row_stride=2 * C, x_batch_stride=L_pad * C, rows_per_batch=L_out, x_rows=L_pad // 2, segs=segs)
        torch.cuda.synchronize()
        xx = x[:, :L_in].double().cpu().transpose(1, 2)
        w = torch.zeros(C, C, taps, dtype=torch.double, requires_grad=True)
        F.conv1d(xx, w, stride=2).transpose(1, 2).backward(dpre.double().cpu())
        want = w.grad.permute(0, 2, 1).reshape(C, taps * C)                  # [co, tap*C + ci]
        err = float((dw.double().cpu() - want).abs().max())
        assert err < tol, (backend, taps, err)
