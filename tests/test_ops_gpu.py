"""Kernel parity: every C-ABI op against a plain PyTorch fp32 restatement of the same op,
fed the same 16-bit inputs.  Tolerances: rel-L2 <= 5e-3 (fp16) / 2e-2 (bf16)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {torch.float16: 5e-3, torch.bfloat16: 2e-2}
DTYPES = [torch.float16, torch.bfloat16]


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def check(name, got, ref, dtype, tol_scale=1.0):
    assert torch.isfinite(got.float()).all(), f"{name}: non-finite output"
    e = rel_l2(got, ref)
    m = (got.float() - ref.float()).abs().max().item()
    assert e <= TOL[dtype] * tol_scale, f"{name}: rel-L2 {e:.3e} (max abs {m:.3e}) > {TOL[dtype] * tol_scale:.1e}"


def rnd(shape, dtype, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev).to(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,K,N", [(256, 64, 128), (300, 320, 320), (1000, 1280, 64), (77, 72, 200),
                                   (128, 64, 16), (25, 1024, 256), (4096, 256, 1024)])
def test_linear(cuda, dtype, M, K, N):
    from unigeo_b200 import ops
    x = rnd((M, K), dtype, cuda, 1)
    W = rnd((N, K), dtype, cuda, 2, 1 / math.sqrt(K))
    b = rnd((N,), torch.float32, cuda, 3)
    r = rnd((M, N), dtype, cuda, 4)
    ref = x.float() @ W.float().t()
    check("plain", ops.linear(x, W), ref, dtype)
    check("bias+res", ops.linear(x, W, bias=b, res=r), ref + b + r.float(), dtype)
    y32 = ops.linear(x, W, bias=b, out_fp32=True)
    assert y32.dtype == torch.float32
    check("fp32 out", y32, ref + b, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,K,N", [(300, 1280, 320), (6144, 1280, 320), (2500, 2560, 640), (1000, 640, 1280),
                                   (77, 72, 200), (256, 64, 32), (4096, 256, 96)])
def test_linear_blend_separate_tensors(cuda, dtype, M, K, N):
    """AlphaBlender fused into the temporal block's last linear layer with the blend input a DIFFERENT tensor than the
    residual: both tiles arrive by TMA (the epilogue's chunk ring as two buffer pairs) where the shapes allow, by
    per-thread loads otherwise (N = 200: rows not 16-byte aligned)."""
    from unigeo_b200 import ops
    x = rnd((M, K), dtype, cuda, 1)
    W = rnd((N, K), dtype, cuda, 2, 1 / math.sqrt(K))
    b = rnd((N,), torch.float32, cuda, 3)
    r = rnd((M, N), dtype, cuda, 4)
    bl = rnd((M, N), dtype, cuda, 5, 2.0)
    for alpha in (0.3, 0.85):
        ref = alpha * bl.float() + (1 - alpha) * (x.float() @ W.float().t() + b + r.float())
        check(f"blend+res a={alpha}", ops.linear_blend(x, W, bl, alpha, bias=b, res=r), ref, dtype)
    ref = 0.5 * bl.float() + 0.5 * (x.float() @ W.float().t() + b)
    check("blend only", ops.linear_blend(x, W, bl, 0.5, bias=b), ref, dtype)


def _ln_ref(x, gamma, beta, eps):
    return F.layer_norm(x.float(), (x.shape[1],), gamma, beta, eps)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,K,N", [(256, 64, 192), (300, 320, 960), (1000, 1280, 320), (77, 72, 200), (6144, 320, 960),
                                   (1200, 1280, 3840)])
def test_ln_linear_folded(cuda, dtype, M, K, N):
    """LayerNorm folded into the consuming GEMM (row_stats pass + epilogue correction) against LayerNorm -> linear in
    fp32; rows with a large common offset check the cancellation acc - mean * colsum."""
    from unigeo_b200 import ops
    x = rnd((M, K), dtype, cuda, 1) * 2.0 + rnd((M, 1), dtype, cuda, 5) * 3.0
    gamma = 1.0 + rnd((K,), torch.float32, cuda, 6, 0.2)
    beta = rnd((K,), torch.float32, cuda, 7, 0.2)
    W = rnd((N, K), dtype, cuda, 2, 1 / math.sqrt(K))
    b = rnd((N,), torch.float32, cuda, 3)
    ref = _ln_ref(x, gamma, beta, 1e-5) @ W.float().t()
    check("lnfold", ops.ln_linear(x, gamma, beta, W), ref, dtype, 1.5)
    check("lnfold+bias", ops.ln_linear(x, gamma, beta, W, bias=b), ref + b, dtype, 1.5)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,K,H", [(200, 64, 128), (513, 128, 512), (6144, 320, 1280)])
def test_ln_linear_folded_geglu(cuda, dtype, M, K, H):
    from unigeo_b200 import ops
    x = rnd((M, K), dtype, cuda, 1) * 1.5 + rnd((M, 1), dtype, cuda, 5)
    gamma = 1.0 + rnd((K,), torch.float32, cuda, 6, 0.2)
    beta = rnd((K,), torch.float32, cuda, 7, 0.2)
    W = rnd((2 * H, K), dtype, cuda, 2, 1 / math.sqrt(K))
    b = rnd((2 * H,), torch.float32, cuda, 3, 0.1)
    h = _ln_ref(x, gamma, beta, 1e-5) @ W.float().t() + b
    ref = h[:, :H] * F.gelu(h[:, H:])
    Wi, bi = ops.geglu_interleave(W, b)
    check("lnfold geglu", ops.ln_linear(x, gamma, beta, Wi, bias=bi, geglu=True), ref, dtype, 1.5)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,K0,K,N,res", [(300, 320, 320, 960, True), (6144, 1280, 320, 2560, True),
                                          (6144, 320, 320, 960, False), (1000, 256, 640, 1920, True),
                                          (77, 64, 72, 200, True), (2400, 5120, 1280, 3840, True)])
def test_ln_linear_folded_producer_stats(cuda, dtype, M, K0, K, N, res):
    """Row statistics left behind by the epilogue of the GEMM that produces the rows (sum / sum-of-squares partials per
    N tile and column group) instead of a statistics pass; the consumer is checked against LayerNorm of the rows the
    producer actually stored."""
    from unigeo_b200 import ops
    x0 = rnd((M, K0), dtype, cuda, 11)
    W0 = rnd((K, K0), dtype, cuda, 12, 1 / math.sqrt(K0))
    b0 = rnd((K,), torch.float32, cuda, 13)
    r0 = rnd((M, K), dtype, cuda, 14, 2.0) if res else None
    gamma = 1.0 + rnd((K,), torch.float32, cuda, 6, 0.2)
    beta = rnd((K,), torch.float32, cuda, 7, 0.2)
    W = rnd((N, K), dtype, cuda, 2, 1 / math.sqrt(K))
    b = rnd((N,), torch.float32, cuda, 3)
    x = torch.empty((M, K), device=cuda, dtype=dtype)
    x, y = ops.ln_linear(x, gamma, beta, W, bias=b, producer=(x0, W0, b0, r0))
    xr = x0.float() @ W0.float().t() + b0 + (r0.float() if res else 0.0)
    check("producer rows", x, xr, dtype)
    check("lnfold after producer", y, _ln_ref(x, gamma, beta, 1e-5) @ W.float().t() + b, dtype, 1.5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_geglu_gate_function_pointwise(cuda, dtype):
    """The gate function of the fused GEGLU epilogue against the exact erf GELU, point by point over gate values in
    [-12, 12] (value = 1, gate = one input channel, so the GEMM part is exact).  The epilogue evaluates x Phi(x) with one
    MUFU.TANH per element and a fitted quartic (max |error| 2.5e-5 + 2^-12 |x| from the hardware tanh); the bound below
    is that plus the 16-bit rounding of the output."""
    from unigeo_b200 import ops
    M, K, H = 4096, 64, 128
    x = torch.zeros((M, K), device=cuda, dtype=dtype)
    x[:, 0] = torch.linspace(-12, 12, M, device=cuda).to(dtype)
    W = torch.zeros((2 * H, K), device=cuda, dtype=dtype)
    W[H:, 0] = 1.0                                      # every gate column = x[:, 0]
    b = torch.zeros(2 * H, device=cuda)
    b[:H] = 1.0                                         # every value column = 1
    Wi, bi = ops.geglu_interleave(W, b)
    y = ops.linear(x, Wi, bias=bi, geglu=True).float()
    g = x[:, 0].float()
    ref = F.gelu(g)[:, None].expand(M, H)
    ulp = 2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8
    bound = 2.5e-5 + ulp * ref.abs() + 2.0 ** -12 * g.abs()[:, None] + 1e-6
    assert ((y - ref).abs() <= bound).all(), f"max excess {((y - ref).abs() - bound).max().item():.3e}"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,K,H", [(200, 64, 128), (513, 128, 512)])
def test_linear_geglu(cuda, dtype, M, K, H):
    from unigeo_b200 import ops
    x = rnd((M, K), dtype, cuda, 1)
    W = rnd((2 * H, K), dtype, cuda, 2, 1 / math.sqrt(K))
    b = rnd((2 * H,), torch.float32, cuda, 3, 0.1)
    h = x.float() @ W.float().t() + b
    ref = h[:, :H] * F.gelu(h[:, H:])
    Wi, bi = ops.geglu_interleave(W, b)
    check("geglu", ops.linear(x, Wi, bias=bi, geglu=True), ref, dtype)


def conv_ref(x, Wt, b, stride, asym):
    # x [N,H,W,C] 16-bit, Wt [9,Cout,C]
    Cout, C = Wt.shape[1], Wt.shape[2]
    w = Wt.float().reshape(3, 3, Cout, C).permute(2, 3, 0, 1)
    xi = x.float().permute(0, 3, 1, 2)
    if stride == 2 and asym:
        xi = F.pad(xi, (0, 1, 0, 1))
        y = F.conv2d(xi, w, b, stride=2, padding=0)
    else:
        y = F.conv2d(xi, w, b, stride=stride, padding=1)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("N,H,W,C,Cout,stride,asym", [
    (2, 16, 16, 64, 64, 1, 0), (3, 48, 64, 32, 128, 1, 0), (25, 6, 8, 128, 128, 1, 0), (2, 4, 4, 64, 64, 1, 0),
    (2, 8, 256, 64, 32, 1, 0), (1, 16, 16, 8, 64, 1, 0), (2, 16, 16, 320, 4, 1, 0), (2, 16, 16, 64, 64, 2, 0),
    (3, 12, 16, 128, 128, 2, 0), (2, 16, 16, 64, 64, 2, 1), (2, 32, 32, 32, 32, 2, 1), (5, 2, 2, 64, 64, 1, 0)])
def test_conv3x3(cuda, dtype, N, H, W, C, Cout, stride, asym):
    from unigeo_b200 import ops
    x = rnd((N, H, W, C), dtype, cuda, 1)
    Wt = rnd((9, Cout, C), dtype, cuda, 2, 1 / math.sqrt(9 * C))
    b = rnd((Cout,), torch.float32, cuda, 3)
    ref = conv_ref(x, Wt, b, stride, asym)
    got = ops.conv3x3(x, Wt, bias=b, stride=stride, asym_pad=bool(asym))
    check("conv", got, ref, dtype)
    if stride == 1:
        r = rnd((N, H, W, Cout), dtype, cuda, 4)
        check("conv+res", ops.conv3x3(x, Wt, bias=b, res=r), ref + r.float(), dtype)


def tconv_ref(x, Wt, b, chunk):
    T, P, C = x.shape
    outs = []
    for t0 in range(0, T, chunk):
        xc = x[t0:t0 + chunk].float()
        xp = F.pad(xc, (0, 0, 0, 0, 1, 1))
        y = sum(xp[k:k + xc.shape[0]] @ Wt[k].float().t() for k in range(3))
        outs.append(y + b)
    return torch.cat(outs, 0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("T,P,C,Cout,chunk", [(25, 48, 128, 128, 25), (8, 256, 64, 64, 8), (25, 100, 64, 64, 8),
                                              (5, 3072, 64, 128, 5), (3, 16, 8, 3, 3), (49, 4, 64, 64, 49)])
def test_tconv3(cuda, dtype, T, P, C, Cout, chunk):
    from unigeo_b200 import ops
    x = rnd((T, P, C), dtype, cuda, 1)
    Wt = rnd((3, Cout, C), dtype, cuda, 2, 1 / math.sqrt(3 * C))
    b = rnd((Cout,), torch.float32, cuda, 3)
    ref = tconv_ref(x, Wt, b, chunk)
    check("tconv", ops.tconv3(x, Wt, bias=b, chunk=chunk), ref, dtype)
    s = rnd((T, P, Cout), dtype, cuda, 5)
    alpha = 0.3
    ref2 = alpha * s.float() + (1 - alpha) * (ref + s.float())
    check("tconv+res+blend", ops.tconv3(x, Wt, bias=b, res=s, blend=s, alpha=alpha, chunk=chunk), ref2, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sets,rps,C1,C2,silu", [(4, 48, 320, 0, 1), (1, 500, 64, 0, 1), (3, 64, 128, 64, 0),
                                                 (25, 12, 1280, 640, 1), (2, 3072, 32, 0, 1), (2, 7, 2560, 0, 1)])
def test_groupnorm(cuda, dtype, sets, rps, C1, C2, silu):
    from unigeo_b200 import ops
    rows, C = sets * rps, C1 + C2
    x1 = rnd((rows, C1), dtype, cuda, 1) * 2 + 0.5
    x2 = rnd((rows, C2), dtype, cuda, 2) if C2 else None
    g = rnd((C,), torch.float32, cuda, 3) * 0.1 + 1
    b = rnd((C,), torch.float32, cuda, 4) * 0.1
    xc = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], 1)
    ref = F.group_norm(xc.view(sets, rps, C).permute(0, 2, 1), 32, g, b, 1e-6).permute(0, 2, 1).reshape(rows, C)
    if silu:
        ref = F.silu(ref)
    check("gn", ops.groupnorm(x1, g, b, rps, 32, 1e-6, bool(silu), x2), ref, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,C,div", [(100, 320, 0), (75, 1280, 25), (64, 64, 16), (33, 2048, 0)])
def test_layernorm(cuda, dtype, rows, C, div):
    from unigeo_b200 import ops
    x = rnd((rows, C), dtype, cuda, 1) * 2 + 0.3
    g = rnd((C,), torch.float32, cuda, 3) * 0.1 + 1
    b = rnd((C,), torch.float32, cuda, 4) * 0.1
    add = rnd(((rows + div - 1) // div, C), torch.float32, cuda, 5) if div else None
    xin = x.float()
    if div:
        xin = xin + add[torch.arange(rows, device=cuda) // div]
    ref = F.layer_norm(xin, (C,), g, b, 1e-5)
    check("ln", ops.layernorm(x, g, b, 1e-5, add, div or 1), ref, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("Fr,N,C,dh", [(2, 256, 128, 64), (3, 48, 128, 64), (1, 192, 64, 64), (2, 128, 128, 128),
                                       (2, 1024, 64, 64), (1, 16, 64, 64), (1, 144, 512, 512),
                                       (2, 3072, 320, 64), (3, 200, 128, 64), (1, 2304, 64, 64)])
def test_spatial_attention(cuda, dtype, Fr, N, C, dh):
    from unigeo_b200 import ops
    qkv = rnd((Fr * N, 3 * C), dtype, cuda, 1)
    heads = C // dh
    q, k, v = [t.float().view(Fr, N, heads, dh).transpose(1, 2) for t in qkv.split(C, dim=1)]
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), -1)
    ref = (p @ v).transpose(1, 2).reshape(Fr * N, C)
    check("attn", ops.spatial_attention(qkv, Fr, N, C, dh), ref, dtype, tol_scale=2.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("T,P,C", [(25, 37, 128), (49, 5, 64), (8, 300, 320), (1, 9, 64)])
def test_temporal_attention(cuda, dtype, T, P, C):
    from unigeo_b200 import ops
    qkv = rnd((T, P, 3 * C), dtype, cuda, 1)
    heads = C // 64
    q, k, v = [t.float().view(T, P, heads, 64).permute(1, 2, 0, 3) for t in qkv.split(C, dim=2)]
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1)
    ref = (p @ v).permute(2, 0, 1, 3).reshape(T, P, C)
    check("tattn", ops.temporal_attention(qkv, T, P, C), ref, dtype, tol_scale=2.0)


def test_balanced_tile_lists_do_not_change_a_bit(cuda):
    """tapgemm deals ragged-width tiles to the CTAs from a host-built balanced list (UG_SCHED, read once per process);
    every tile is computed independently, so the outputs must equal the round-robin order's bit for bit."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("0", "1"):
        env = dict(os.environ, UG_SCHED=flag)
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "ab_sched.py"), "--quick"], env=env,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    a, b = outs
    assert [x["op"] for x in a["rows"]] == [x["op"] for x in b["rows"]] and len(a["rows"]) >= 5
    for x, y in zip(a["rows"], b["rows"]):
        assert x["sha"] == y["sha"], (x, y)
