"""Single-kernel entry points of the C ABI as torch-tensor functions (kernel parity tests
and micro-benchmarks use these; the model graphs call the same kernels from C++)."""
from __future__ import annotations

import math

import torch

from . import _lib

_UG = {torch.float16: _lib.UG_F16, torch.bfloat16: _lib.UG_BF16}


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk16(*ts):
    dt = ts[0].dtype
    assert dt in _UG, "16-bit tensors only"
    for t in ts:
        assert t is None or (t.is_cuda and t.is_contiguous() and t.dtype == dt)
    return _UG[dt]


def linear(x, W, bias=None, res=None, geglu=False, out_fp32=False):
    """x [M,K], W [N,K] (geglu: rows pre-interleaved, see geglu_interleave) -> [M,N] (or [M,N/2])."""
    d = _chk16(x, W, res)
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty((M, N // 2 if geglu else N), device=x.device, dtype=torch.float32 if out_fp32 else x.dtype)
    _lib.check(_lib.load().ug_op_linear(d, x.data_ptr(), M, K, W.data_ptr(), N, _p(bias), _p(res), int(geglu),
                                        int(out_fp32), y.data_ptr(), _s()))
    return y


def linear_blend(x, W, blend, alpha, bias=None, res=None):
    """alpha * blend + (1 - alpha) * (x @ W.T + bias (+res)); blend [M,N] (see ug_op_linear_blend)."""
    d = _chk16(x, W, res, blend)
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty((M, N), device=x.device, dtype=x.dtype)
    _lib.check(_lib.load().ug_op_linear_blend(d, x.data_ptr(), M, K, W.data_ptr(), N, _p(bias), _p(res), blend.data_ptr(),
                                              float(alpha), y.data_ptr(), _s()))
    return y


def geglu_interleave(W, b):
    """[2H,K] (value rows | gate rows) -> 256-row tiles of [128 value | 128 gate] (what ug_ctx_finalize builds)."""
    H = W.shape[0] // 2
    Wv, Wg = W[:H].reshape(H // 128, 128, -1), W[H:].reshape(H // 128, 128, -1)
    Wi = torch.cat([Wv, Wg], dim=1).reshape(2 * H, -1).contiguous()
    bv, bg = b[:H].reshape(H // 128, 128), b[H:].reshape(H // 128, 128)
    return Wi, torch.cat([bv, bg], dim=1).reshape(2 * H).contiguous()


def conv3x3(x, Wt, bias=None, res=None, stride=1, asym_pad=False):
    """x [N,H,W,C], Wt [9,Cout,C] -> [N,H/stride,W/stride,Cout]."""
    d = _chk16(x, Wt, res)
    N, H, W, Cc = x.shape
    Cout = Wt.shape[1]
    y = torch.empty((N, H // stride, W // stride, Cout), device=x.device, dtype=x.dtype)
    _lib.check(_lib.load().ug_op_conv3x3(d, x.data_ptr(), N, H, W, Cc, Wt.data_ptr(), Cout, stride, int(asym_pad),
                                         _p(bias), _p(res), y.data_ptr(), _s()))
    return y


def tconv3(x, Wt, bias=None, res=None, blend=None, alpha=0.0, chunk=None):
    """x [T,P,C], Wt [3,Cout,C] -> [T,P,Cout]; zero padding at the ends of every ``chunk`` frames."""
    d = _chk16(x, Wt, res, blend)
    T, P, Cc = x.shape
    Cout = Wt.shape[1]
    y = torch.empty((T, P, Cout), device=x.device, dtype=x.dtype)
    _lib.check(_lib.load().ug_op_tconv3(d, x.data_ptr(), T, P, Cc, Wt.data_ptr(), Cout, chunk or T, _p(bias),
                                        _p(res), _p(blend), float(alpha), y.data_ptr(), _s()))
    return y


def groupnorm(x1, gamma, beta, rows_per_set, groups=32, eps=1e-5, silu=True, x2=None):
    """x1 [rows,C1] (| x2 [rows,C2]) -> [rows,C1+C2]."""
    d = _chk16(x1, x2)
    rows, C1 = x1.shape
    C2 = 0 if x2 is None else x2.shape[1]
    y = torch.empty((rows, C1 + C2), device=x1.device, dtype=x1.dtype)
    _lib.check(_lib.load().ug_op_groupnorm(d, x1.data_ptr(), C1, _p(x2), C2, rows, rows_per_set, groups,
                                           gamma.data_ptr(), beta.data_ptr(), float(eps), int(silu), y.data_ptr(),
                                           _s()))
    return y


def layernorm(x, gamma, beta, eps=1e-5, add=None, add_div=1):
    d = _chk16(x)
    rows, Cc = x.shape
    y = torch.empty_like(x)
    _lib.check(_lib.load().ug_op_layernorm(d, x.data_ptr(), rows, Cc, gamma.data_ptr(), beta.data_ptr(), float(eps),
                                           _p(add), add_div, y.data_ptr(), _s()))
    return y


def ln_linear(x, gamma, beta, W, bias=None, eps=1e-5, geglu=False, producer=None):
    """LayerNorm(x) @ W.T (+bias) in the folded form (see ug_op_ln_linear).  x [M,K], W [N,K].

    producer = (x0 [M,K0], W0 [K,K0], bias0 | None, res0 [M,K] | None): x is computed as x0 @ W0.T (+bias0) (+res0)
    first and its row statistics come out of that GEMM's epilogue; returns (x, y) then."""
    d = _chk16(x, W)
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty((M, N // 2 if geglu else N), device=x.device, dtype=x.dtype)
    if producer is None:
        _lib.check(_lib.load().ug_op_ln_linear(d, None, 0, None, None, None, x.data_ptr(), M, K, gamma.data_ptr(),
                                               beta.data_ptr(), float(eps), W.data_ptr(), N, _p(bias), int(geglu),
                                               y.data_ptr(), _s()))
        return y
    x0, W0, b0, r0 = producer
    _chk16(x0, W0, r0)
    _lib.check(_lib.load().ug_op_ln_linear(d, x0.data_ptr(), x0.shape[1], W0.data_ptr(), _p(b0), _p(r0), x.data_ptr(),
                                           M, K, gamma.data_ptr(), beta.data_ptr(), float(eps), W.data_ptr(), N,
                                           _p(bias), int(geglu), y.data_ptr(), _s()))
    return x, y


def spatial_attention(qkv, F, N, C, head_dim=64):
    """qkv [F*N,3C] -> [F*N,C]."""
    d = _chk16(qkv)
    y = torch.empty((F * N, C), device=qkv.device, dtype=qkv.dtype)
    _lib.check(_lib.load().ug_op_spatial_attention(d, qkv.data_ptr(), F, N, C, head_dim, y.data_ptr(), _s()))
    return y


def temporal_attention(qkv, T, P, C):
    """qkv [T,P,3C] -> [T,P,C]; heads of 64."""
    d = _chk16(qkv)
    y = torch.empty((T, P, C), device=qkv.device, dtype=qkv.dtype)
    _lib.check(_lib.load().ug_op_temporal_attention(d, qkv.data_ptr(), T, P, C, y.data_ptr(), _s()))
    return y


def cross_attention(q, kv, F, N, C, Lk, kv_per_frame=False):
    """q [F*N,C], kv [Fk*Lk,2C] (K | V; Fk = F if kv_per_frame else 1) -> [F*N,C]; heads of 64, Lk <= 128."""
    d = _chk16(q, kv)
    y = torch.empty((F * N, C), device=q.device, dtype=q.dtype)
    _lib.check(_lib.load().ug_op_cross_attention(d, q.data_ptr(), kv.data_ptr(), F, N, C, Lk, int(kv_per_frame),
                                                 y.data_ptr(), _s()))
    return y
