"""Thin torch-tensor wrappers over the C-ABI (lib.py): argument checking, struct filling, stream plumbing.

torch is used for device memory and streams only; every op here launches kernels of liba2f_sm100.so and raises
A2FError on any non-zero status (no fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as L


# bench.py's roofline pass: when set to a list, every tcgen05 GEMM launch is bracketed by CUDA events on the launch
# stream and recorded as (kind, algorithmic_flops, start_event, end_event).  None (default) = no instrumentation.
PROFILE = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.F32
    if t.dtype == torch.bfloat16:
        return L.BF16
    raise L.A2FError(f"unsupported dtype {t.dtype}")


def _dev(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.A2FError("a2f ops need CUDA tensors (there is no CPU path)")


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
         act: int = L.ACT_NONE, resid: Optional[torch.Tensor] = None, tmpl: Optional[torch.Tensor] = None,
         rows_per_tmpl: int = 1, backend: int = L.SIMT_F32, M: Optional[int] = None, K: Optional[int] = None,
         a_row_stride: Optional[int] = None, a_batch_stride: int = 0, rows_per_batch: Optional[int] = None,
         N: Optional[int] = None, ldw: Optional[int] = None, ldc: Optional[int] = None, c_batch_stride: int = 0,
         c_offset: int = 0, a_rows: int = 0, segs=None, resid_mode: int = 0, ldr: Optional[int] = None,
         r_batch_stride: int = 0, r_offset: int = 0) -> torch.Tensor:
    """out[m,n] = act(sum_k a[m,k] w[n,k] + bias[n]) + resid[m,n] + tmpl[m // rows_per_tmpl, n]  (a2f_gemm)."""
    _dev(a, w, out, bias, resid, tmpl)
    lib = L.load()
    g = L.GemmArgs()
    g.M = int(M if M is not None else a.shape[0])
    g.N = int(N if N is not None else w.shape[0])
    g.K = int(K if K is not None else a.shape[-1])
    g.A, g.a_dtype = a.data_ptr(), _dt(a)
    g.a_row_stride = int(a_row_stride if a_row_stride is not None else a.stride(0))
    g.a_batch_stride = int(a_batch_stride)
    g.rows_per_batch = int(rows_per_batch if rows_per_batch is not None else g.M)
    g.W, g.ldw = w.data_ptr(), int(ldw if ldw is not None else w.stride(0))
    g.bias = L.ptr(bias)
    g.act = act
    if resid is not None:
        g.resid, g.resid_dtype = resid.data_ptr() + int(r_offset) * resid.element_size(), _dt(resid)
        g.ldr = int(ldr if ldr is not None else resid.stride(0))
        g.r_batch_stride = int(r_batch_stride)
        g.resid_mode = int(resid_mode)
    g.a_rows = int(a_rows)
    if segs is not None:                       # [(row_off, col_off), ...]: explicit K segments (a2f.h)
        g.n_seg = len(segs)
        for i, (ro, co) in enumerate(segs):
            g.seg_row_off[i], g.seg_col_off[i] = int(ro), int(co)
    if tmpl is not None:
        g.tmpl, g.rows_per_tmpl = tmpl.data_ptr(), int(rows_per_tmpl)
    g.C, g.c_dtype = out.data_ptr() + int(c_offset) * out.element_size(), _dt(out)
    g.ldc = int(ldc if ldc is not None else out.stride(0))
    g.c_batch_stride = int(c_batch_stride)
    if bias is not None and bias.dtype != torch.float32:
        raise L.A2FError("bias must be fp32")
    if tmpl is not None and tmpl.dtype != torch.float32:
        raise L.A2FError("tmpl must be fp32")
    if a.dtype != w.dtype:
        raise L.A2FError("A and W must share a dtype")
    if PROFILE is not None and backend == L.TCGEN05:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        L.check(lib.a2f_gemm(C.byref(g), backend, _stream()), "a2f_gemm")
        e.record()
        PROFILE.append(("gemm_tc", 2.0 * g.M * g.N * g.K, s, e))
        return out
    L.check(lib.a2f_gemm(C.byref(g), backend, _stream()), "a2f_gemm")
    return out


def gemm_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, *, backend: int, M: Optional[int] = None,
               N: Optional[int] = None, K: Optional[int] = None, dy_row_stride: Optional[int] = None,
               dy_batch_stride: int = 0, x_row_stride: Optional[int] = None, x_batch_stride: int = 0,
               rows_per_batch: Optional[int] = None, x_rows: int = 0, segs=None, ldw: Optional[int] = None,
               x_offset: int = 0, dy_offset: int = 0) -> torch.Tensor:
    """dw[n, s*K+k] += sum_m dy[m,n] * x[(b, r+roff_s), coff_s+k]   (a2f_gemm_wgrad; dw fp32, accumulated into)."""
    _dev(dy, x, dw)
    if dw.dtype != torch.float32:
        raise L.A2FError("dW must be fp32")
    if dy.dtype != x.dtype:
        raise L.A2FError("dY and X must share a dtype")
    g = L.WgradArgs()
    g.M = int(M if M is not None else dy.shape[0])
    g.N = int(N if N is not None else dy.shape[-1])
    g.K = int(K if K is not None else x.shape[-1])
    g.dtype = _dt(dy)
    g.dY = dy.data_ptr() + int(dy_offset) * dy.element_size()
    g.dy_row_stride = int(dy_row_stride if dy_row_stride is not None else dy.stride(-2))
    g.dy_batch_stride = int(dy_batch_stride)
    g.X = x.data_ptr() + int(x_offset) * x.element_size()
    g.x_row_stride = int(x_row_stride if x_row_stride is not None else x.stride(-2))
    g.x_batch_stride = int(x_batch_stride)
    g.rows_per_batch = int(rows_per_batch if rows_per_batch is not None else g.M)
    g.x_rows = int(x_rows)
    if segs is not None:
        g.n_seg = len(segs)
        for i, (ro, co) in enumerate(segs):
            g.x_row_off[i], g.x_col_off[i] = int(ro), int(co)
    g.dW, g.ldw = dw.data_ptr(), int(ldw if ldw is not None else dw.stride(0))
    lib = L.load()
    if PROFILE is not None and backend == L.TCGEN05:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        L.check(lib.a2f_gemm_wgrad(C.byref(g), backend, _stream()), "a2f_gemm_wgrad")
        e.record()
        PROFILE.append(("wgrad_tc", 2.0 * g.M * g.N * g.K * max(1, g.n_seg), s, e))
        return dw
    L.check(lib.a2f_gemm_wgrad(C.byref(g), backend, _stream()), "a2f_gemm_wgrad")
    return dw


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _dev(x)
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    L.check(L.load().a2f_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "a2f_cast_f32_to_bf16")
    return out


def voca_trunk(wptrs: "L.VocaWeights", x: torch.Tensor, one_hot: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    _dev(x, one_hot, z)
    L.check(L.load().a2f_voca_trunk(C.byref(wptrs), x.data_ptr(), one_hot.data_ptr(), one_hot.shape[1], z.data_ptr(),
                                    _dt(z), z.stride(0), x.shape[0], _stream()), "a2f_voca_trunk")
    return z


def loss_workspace(device) -> torch.Tensor:
    n = L.load().a2f_voca_loss_workspace_bytes()
    return torch.empty((n + 7) // 8, dtype=torch.float64, device=device)


def voca_loss_fwd(pred: torch.Tensor, gt: torch.Tensor, rows: int, v3: int, k_rec: float, k_vel: float,
                  ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    _dev(pred, gt)
    out3 = torch.empty(3, dtype=torch.float32, device=pred.device)
    if ws is None:
        ws = loss_workspace(pred.device)
    L.check(L.load().a2f_voca_loss_fwd(pred.data_ptr(), gt.data_ptr(), rows, v3, k_rec, k_vel, out3.data_ptr(),
                                       ws.data_ptr(), ws.numel() * 8, _stream()), "a2f_voca_loss_fwd")
    return out3


def voca_loss_bwd(pred: torch.Tensor, gt: torch.Tensor, rows: int, v3: int, k_rec: float, k_vel: float,
                  gscale: Optional[torch.Tensor], dpred: torch.Tensor) -> torch.Tensor:
    _dev(pred, gt, dpred, gscale)
    L.check(L.load().a2f_voca_loss_bwd(pred.data_ptr(), gt.data_ptr(), rows, v3, k_rec, k_vel, L.ptr(gscale),
                                       dpred.data_ptr(), _stream()), "a2f_voca_loss_bwd")
    return dpred


def split_bf16x3(x: torch.Tensor, is_weight: bool) -> torch.Tensor:
    """fp32 [rows,K] -> bf16 [rows,3K] error-compensated split (a2f_split_bf16x3)."""
    _dev(x)
    rows, K = x.shape
    out = torch.empty((rows, 3 * K), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().a2f_split_bf16x3(x.data_ptr(), x.stride(0), out.data_ptr(), rows, K, 1 if is_weight else 0, _stream()),
            "a2f_split_bf16x3")
    return out


def pack_conv1d_weight(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _dev(w)
    cout, cin, taps = w.shape
    out = torch.empty((cout, taps * cin), dtype=dtype, device=w.device)
    L.check(L.load().a2f_pack_conv1d_weight(w.contiguous().data_ptr(), out.data_ptr(), _dt(out), cout, cin, taps, _stream()),
            "a2f_pack_conv1d_weight")
    return out


def pack_posconv_weight(g: torch.Tensor, v: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _dev(g, v)
    kpad = 64 if dtype == torch.bfloat16 else 48
    out = torch.empty((16, 48, 128, kpad), dtype=dtype, device=v.device)
    norm = torch.empty(128, dtype=torch.float32, device=v.device)
    L.check(L.load().a2f_pack_posconv_weight(g.contiguous().data_ptr(), v.contiguous().data_ptr(), out.data_ptr(), _dt(out),
                                             kpad, norm.data_ptr(), _stream()), "a2f_pack_posconv_weight")
    return out


def posconv(h: torch.Tensor, wp: torch.Tensor, bias: torch.Tensor, out: torch.Tensor, B: int, T: int, backend: int):
    _dev(h, wp, bias, out)
    prof = PROFILE is not None and backend == L.TCGEN05
    if prof:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
    L.check(L.load().a2f_posconv(h.data_ptr(), _dt(h), wp.data_ptr(), bias.data_ptr(), out.data_ptr(), _dt(out), B, T,
                                 backend, _stream()), "a2f_posconv")
    if prof:
        e.record()
        PROFILE.append(("gemm_tc", 2.0 * B * T * 768 * 48 * 128, s, e))
    return out


def audio_stats(audio: torch.Tensor) -> torch.Tensor:
    _dev(audio)
    B, N = audio.shape
    stats = torch.empty((B, 2), dtype=torch.float32, device=audio.device)
    L.check(L.load().a2f_audio_stats(audio.data_ptr(), B, N, stats.data_ptr(), _stream()), "a2f_audio_stats")
    return stats


def conv0_gn_gelu(audio: torch.Tensor, stats: torch.Tensor, w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                  dtype: torch.dtype) -> torch.Tensor:
    """-> channels-last [B, L0, 512]"""
    _dev(audio, stats, w, gamma, beta)
    B, N = audio.shape
    L0 = (N - 10) // 5 + 1
    lib = L.load()
    nbytes = lib.a2f_conv0_workspace_bytes(B, N)
    ws = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=audio.device)
    out = torch.empty((B, L0, 512), dtype=dtype, device=audio.device)
    L.check(lib.a2f_conv0_gn_gelu(audio.data_ptr(), stats.data_ptr(), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                  out.data_ptr(), _dt(out), B, N, ws.data_ptr(), ws.numel() * 8, _stream()),
            "a2f_conv0_gn_gelu")
    return out


def interp_ln(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, T: int, out_dtype: torch.dtype, eps: float = 1e-5):
    """[B,S,512] -> LayerNorm(linear_interpolation to T frames) [B,T,512]"""
    _dev(x, gamma, beta)
    B, S, Cc = x.shape
    out = torch.empty((B, T, Cc), dtype=out_dtype, device=x.device)
    L.check(L.load().a2f_interp_ln(x.data_ptr(), _dt(x), gamma.data_ptr(), beta.data_ptr(), eps, out.data_ptr(), _dt(out),
                                   B, S, T, Cc, _stream()), "a2f_interp_ln")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor, eps: float = 1e-5,
              out2: Optional[torch.Tensor] = None) -> torch.Tensor:
    _dev(x, gamma, beta, out, out2)
    rows, Cc = x.shape
    L.check(L.load().a2f_layernorm(x.data_ptr(), _dt(x), gamma.data_ptr(), beta.data_ptr(), eps, out.data_ptr(), _dt(out),
                                   L.ptr(out2), _dt(out2) if out2 is not None else 0, rows, Cc, _stream()), "a2f_layernorm")
    return out


def mha(qkv: torch.Tensor, out: torch.Tensor, B: int, T: int, H: int = 12, D: int = 64, scale: float = 0.125):
    _dev(qkv, out)
    L.check(L.load().a2f_mha_fwd(qkv.data_ptr(), out.data_ptr(), _dt(qkv), B, T, H, D, scale, _stream()), "a2f_mha_fwd")
    return out


def a2m_assemble(x: torch.Tensor, one_hot: torch.Tensor) -> torch.Tensor:
    """x [B,52,32], one_hot [B,n] -> zero-left-padded single-channel input [B,64,33] (a2f_a2m_assemble)."""
    _dev(x, one_hot)
    B = x.shape[0]
    out = torch.empty((B, 64, 33), dtype=torch.float32, device=x.device)
    L.check(L.load().a2f_a2m_assemble(x.data_ptr(), one_hot.data_ptr(), one_hot.shape[1], out.data_ptr(), B, _stream()),
            "a2f_a2m_assemble")
    return out


def channel_affine(buf: torch.Tensor, offset: int, scale: torch.Tensor, shift: torch.Tensor, C_: int, rows_per_batch: int,
                   ld: int, batch_stride: int, batches: int) -> None:
    _dev(buf, scale, shift)
    L.check(L.load().a2f_channel_affine(buf.data_ptr() + offset * buf.element_size(), _dt(buf), scale.data_ptr(),
                                        shift.data_ptr(), C_, rows_per_batch, ld, batch_stride, batches, _stream()),
            "a2f_channel_affine")


def pack_feedback(vm_w, vm_b, vmr_w, vmr_b):
    _dev(vm_w, vm_b, vmr_w, vmr_b)
    wc = torch.empty((64, 64), dtype=torch.float32, device=vm_w.device)
    bc = torch.empty((64,), dtype=torch.float32, device=vm_w.device)
    L.check(L.load().a2f_pack_feedback(vm_w.data_ptr(), vm_b.data_ptr(), vmr_w.data_ptr(), vmr_b.data_ptr(),
                                       vmr_w.shape[0], wc.data_ptr(), bc.data_ptr(), _stream()), "a2f_pack_feedback")
    return wc, bc


def decoder_rollout(wstruct: "L.DecoderWeights", memory: torch.Tensor, one_hot: torch.Tensor, period: int, B: int, T: int):
    """memory [B,T,64] fp32 -> decoder states D [B,T,64] fp32 (a2f_decoder_rollout)."""
    _dev(memory, one_hot)
    lib = L.load()
    nbytes = lib.a2f_decoder_workspace_bytes(B, T)
    ws = torch.empty((nbytes + 15) // 16 * 4, dtype=torch.float32, device=memory.device)
    D = torch.empty((B, T, 64), dtype=torch.float32, device=memory.device)
    L.check(lib.a2f_decoder_rollout(C.byref(wstruct), memory.data_ptr(), one_hot.data_ptr(), one_hot.shape[1], period,
                                    D.data_ptr(), B, T, ws.data_ptr(), ws.numel() * 4, _stream()), "a2f_decoder_rollout")
    return D
