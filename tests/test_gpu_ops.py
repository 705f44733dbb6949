"""Kernel-level parity on a B200: every hand-written kernel, called through the C ABI, against a plain PyTorch fp32
restatement of the same op on the same (bf16-rounded) operands."""
import ctypes
import math
from ctypes import c_void_p

import numpy as np
import pytest
import torch

from chunkformer_b200 import lib as cflib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


_VARIANT = [-1]


def _gemm(A, B, epi, out, bias, act=0, resid=None, alpha=1.0, row_range=None, rows_per_chunk=1, parts=(None, None, None)):
    L = cflib.load()
    M, K = A.shape
    N = B.shape[0]
    rc = L.cf_op_gemm(_p(A), A.stride(0), _p(B), B.stride(0), M, N, K, epi, act, _p(bias),
                      _p(resid), resid.stride(0) if resid is not None else 0, alpha, _p(row_range), rows_per_chunk,
                      _p(out), out.stride(0) if out is not None else 0, _p(parts[0]), _p(parts[1]), _p(parts[2]), _VARIANT[0],
                      _stream())
    cflib.check(rc, None, "cf_op_gemm")
    torch.cuda.synchronize()


@pytest.fixture(params=[0, 1], ids=["gemm1cta", "gemm2cta"])
def gemm_variant(request):
    """Run a GEMM test on the 1-CTA kernel and on the 2-CTA pair (cta_group::2) kernel."""
    _VARIANT[0] = request.param
    yield request.param
    _VARIANT[0] = -1


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (200, 512, 512), (1000, 2048, 512), (777, 512, 2048), (64, 256, 4608),
                                   (40000, 512, 512)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_bias_act_bf16(M, N, K, act, gemm_variant):
    A = _rand((M, K), 1.0, 1).bfloat16()
    B = _rand((N, K), 1.0 / math.sqrt(K), 2).bfloat16()
    bias = _rand((N,), 0.5, 3)
    out = torch.full((M, N), 7.0, device=DEV, dtype=torch.bfloat16)
    _gemm(A, B, cflib.EPI_BF16, out, bias, act=act)
    ref = A.float() @ B.float().T + bias
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = ref * torch.sigmoid(ref)
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-2, err
    assert (out.float() - ref).abs().mean().item() < 3e-3


def test_gemm_glu(gemm_variant):
    M, d = 300, 512
    A = _rand((M, d), 1.0, 1).bfloat16()
    W = _rand((2 * d, d), 1.0 / math.sqrt(d), 2)
    b = _rand((2 * d,), 0.5, 3)
    Wi = torch.empty_like(W)
    bi = torch.empty_like(b)
    Wi[0::2], Wi[1::2] = W[:d], W[d:]
    bi[0::2], bi[1::2] = b[:d], b[d:]
    out = torch.zeros((M, d), device=DEV, dtype=torch.bfloat16)
    _gemm(A, Wi.bfloat16().contiguous(), cflib.EPI_GLU, out, bi.contiguous())
    h = A.float() @ W.bfloat16().float().T + b
    ref = h[:, :d] * torch.sigmoid(h[:, d:])
    assert (out.float() - ref).abs().max().item() < 3e-2


def test_gemm_f32_residual_rowmask(gemm_variant):
    c, n, d, K = 64, 5, 512, 2048
    M = n * c
    A = _rand((M, K), 1.0, 1).bfloat16()
    B = _rand((d, K), 1.0 / math.sqrt(K), 2).bfloat16()
    bias = _rand((d,), 0.5, 3)
    x = _rand((M, d), 1.0, 4)
    rng = torch.tensor([[0, 64], [3, 60], [0, 0], [10, 64], [0, 17]], dtype=torch.int32, device=DEV)
    out = x.clone()
    _gemm(A, B, cflib.EPI_F32, out, bias, resid=out, alpha=0.5, row_range=rng, rows_per_chunk=c)
    keep = torch.zeros(M, dtype=torch.bool, device=DEV)
    for g, (lo, hi) in enumerate(rng.tolist()):
        keep[g * c + lo: g * c + hi] = True
    ref = x + keep.unsqueeze(1) * 0.5 * (A.float() @ B.float().T + bias)
    assert (out - ref).abs().max().item() < 2e-3
    assert torch.equal(out[~keep], x[~keep])          # masked rows: exactly the residual
    # no residual, scale only, ragged N (vocab-like)
    Nv = 5000
    Bv = _rand((Nv, 512), 0.05, 5).bfloat16()
    bv = _rand((Nv,), 0.5, 6)
    Av = _rand((300, 512), 1.0, 7).bfloat16()
    outv = torch.full((300, Nv), -1.0, device=DEV)
    _gemm(Av, Bv, cflib.EPI_F32, outv, bv, alpha=2.0)
    refv = 2.0 * (Av.float() @ Bv.float().T + bv)
    assert (outv - refv).abs().max().item() < 2e-3


def test_gemm_strided_output_view(gemm_variant):
    """The fused [Q+u | Q+v | K | V] projection writes into a row-offset view of the flat QKV buffer."""
    M, d, lead = 333, 512, 128
    A = _rand((M, d), 1.0, 1).bfloat16()
    W = _rand((4 * d, d), 1.0 / math.sqrt(d), 2).bfloat16()
    b = _rand((4 * d,), 0.5, 3)
    buf = torch.full((lead + M + 64, 4 * d), 3.0, device=DEV, dtype=torch.bfloat16)
    _gemm(A, W, cflib.EPI_BF16, buf[lead:lead + M], b)
    ref = A.float() @ W.float().T + b
    assert (buf[lead:lead + M].float() - ref).abs().max().item() < 3e-2
    assert bool((buf[:lead] == 3.0).all()) and bool((buf[lead + M:] == 3.0).all())   # rows outside [0, M) untouched


def test_gemm_argmax_partials(gemm_variant):
    M, d, V = 500, 512, 5000
    A = _rand((M, d), 1.0, 1).bfloat16()
    W = _rand((V, d), 1.0 / math.sqrt(d), 2).bfloat16()
    b = _rand((V,), 0.5, 3)
    nt = 2 * ((V + 255) // 256)
    best = torch.zeros((M, nt), device=DEV)
    second = torch.zeros((M, nt), device=DEV)
    index = torch.zeros((M, nt), device=DEV, dtype=torch.int32)
    _gemm(A, W, cflib.EPI_ARGMAX, None, b, parts=(best, second, index))
    logits = A.float() @ W.float().T + b
    top2 = logits.topk(2, -1)
    got_best, pos = best.max(-1)
    got_idx = index.gather(1, pos.unsqueeze(1)).squeeze(1).long()
    margin = top2.values[:, 0] - top2.values[:, 1]
    ok = (got_idx == top2.indices[:, 0]) | (margin < 1e-3)
    assert bool(ok.all())
    assert (got_best - top2.values[:, 0]).abs().max().item() < 2e-3


@pytest.mark.parametrize("d", [256, 512])
def test_layernorm_modes(d):
    L = cflib.load()
    rows = 1031
    x = _rand((rows, d), 2.0, 1) + 0.3
    w1, b1 = 1 + _rand((d,), 0.1, 2), _rand((d,), 0.1, 3)
    w2, b2 = 1 + _rand((d,), 0.1, 4), _rand((d,), 0.1, 5)
    ln1 = torch.nn.functional.layer_norm(x, (d,), w1, b1, 1e-5)
    ln2 = torch.nn.functional.layer_norm(ln1, (d,), w2, b2, 1e-5)
    y = torch.zeros((rows, d), device=DEV, dtype=torch.bfloat16)
    cflib.check(L.cf_op_layernorm(0, d, _p(x), None, _p(y), _p(w1), _p(b1), None, None, rows, _stream()))
    torch.cuda.synchronize()
    assert (y.float() - ln1).abs().max().item() < 3e-2
    xo = x.clone()
    cflib.check(L.cf_op_layernorm(1, d, _p(xo), _p(xo), _p(y), _p(w1), _p(b1), _p(w2), _p(b2), rows, _stream()))
    torch.cuda.synchronize()
    assert (xo - ln1).abs().max().item() < 1e-4
    assert (y.float() - ln2).abs().max().item() < 3e-2
    o32 = torch.zeros((rows, d), device=DEV)
    cflib.check(L.cf_op_layernorm(2, d, _p(x), _p(o32), _p(y), _p(w1), _p(b1), _p(w2), _p(b2), rows, _stream()))
    torch.cuda.synchronize()
    assert (o32 - ln2).abs().max().item() < 1e-4


@pytest.mark.parametrize("d,c", [(512, 64), (256, 16), (512, 8), (256, 4)])
def test_dwconv_ln_silu(d, c):
    L = cflib.load()
    n, lo = 7, 7
    rows = n * c
    g = torch.zeros((rows + 2 * lo + 64, d), device=DEV, dtype=torch.bfloat16)
    g[: rows + 2 * lo] = _rand((rows + 2 * lo, d), 1.0, 1).bfloat16()
    w, b = _rand((d, 15), 0.3, 2), _rand((d,), 0.2, 3)
    lw, lb = 1 + _rand((d,), 0.1, 4), _rand((d,), 0.1, 5)
    rng_np = np.array([[0, c + 14], [7, c + 14], [0, c + 7], [3, c + 2], [0, 0], [5, 9], [0, c + 14]], dtype=np.int32)
    rng = torch.from_numpy(rng_np).to(DEV)
    z = torch.zeros((rows, d), device=DEV, dtype=torch.bfloat16)
    cflib.check(L.cf_op_dwconv(d, 15, _p(g), _p(z), _p(w), _p(b), _p(lw), _p(lb), _p(rng), c, n, _stream()))
    torch.cuda.synchronize()
    gf = g.float()
    for ch in range(n):
        win = gf[ch * c: ch * c + c + 14].clone()
        m = torch.zeros(c + 14, dtype=torch.bool, device=DEV)
        m[rng_np[ch, 0]: rng_np[ch, 1]] = True
        win = win * m.unsqueeze(1)
        conv = torch.stack([(win[i:i + 15] * w.T).sum(0) for i in range(c)], 0) + b
        ref = torch.nn.functional.layer_norm(conv, (d,), lw, lb, 1e-5)
        ref = ref * torch.sigmoid(ref)
        assert (z[ch * c:(ch + 1) * c].float() - ref).abs().max().item() < 3e-2, ch


def _attention_reference(qkv, pos, rng, n, c, l, r, d, H):
    dk = d // H
    W = l + c + r
    q = qkv.float()
    out = torch.zeros((n * c, d), device=qkv.device)
    idx = (c - 1) - torch.arange(c, device=qkv.device).unsqueeze(1) + torch.arange(W, device=qkv.device).unsqueeze(0)
    for g in range(n):
        lo, hi = rng[g]
        for h in range(H):
            qu = q[l + g * c: l + (g + 1) * c, h * dk:(h + 1) * dk]
            qv = q[l + g * c: l + (g + 1) * c, d + h * dk: d + (h + 1) * dk]
            k = q[g * c: g * c + W, 2 * d + h * dk: 2 * d + (h + 1) * dk]
            v = q[g * c: g * c + W, 3 * d + h * dk: 3 * d + (h + 1) * dk]
            p = pos.float()[:, h * dk:(h + 1) * dk]
            s = (qu @ k.T + (qv @ p.T).gather(1, idx)) / math.sqrt(dk)
            m = torch.zeros(W, dtype=torch.bool, device=qkv.device)
            m[lo:hi] = True
            s = s.masked_fill(~m, float("-inf"))
            if hi > lo:
                a = torch.softmax(s, -1)
            else:
                a = torch.zeros_like(s)
            out[g * c:(g + 1) * c, h * dk:(h + 1) * dk] = a @ v
    return out


def _attention_case(impl, c, l, r, d, H, n, seed=0, prescaled=0):
    L = cflib.load()
    W = l + c + r
    R = 2 * c + l + r - 1
    rows = l + n * c + r + 2 * c + 128
    qkv = torch.zeros((rows, 4 * d), device=DEV, dtype=torch.bfloat16)
    qkv[: l + n * c] = _rand((l + n * c, 4 * d), 1.0, seed).bfloat16()
    Rpad = (R + 127) // 128 * 128
    pos = torch.zeros((Rpad, d), device=DEV, dtype=torch.bfloat16)
    pos[:R] = _rand((R, d), 1.0, seed + 1).bfloat16()
    rs = np.random.RandomState(seed)
    rng_np = np.zeros((n + 16, 2), dtype=np.int32)   # phantom chunks of the last 128-row attention tile stay empty
    for g in range(n):
        kind = g % 4
        if kind == 0:
            rng_np[g] = (0, W)
        elif kind == 1:
            rng_np[g] = (rs.randint(0, l + 1), W)
        elif kind == 2:
            rng_np[g] = (0, rs.randint(l + 1, W + 1))
        else:
            rng_np[g] = (rs.randint(0, l + 1), rs.randint(l + 1, W + 1))
    if n > 5:
        rng_np[5] = (0, 0)          # chunk with no valid key: zero context, not NaN
    rng = torch.from_numpy(rng_np).to(DEV)
    ctx = torch.full((n * c, d), 9.0, device=DEV, dtype=torch.bfloat16)
    qkv_ref = qkv
    if prescaled:
        # the product path stores (Q+u), (Q+v) multiplied by (1/sqrt(d_k)) * log2(e); undo it for the reference
        alpha = (1.0 / math.sqrt(d // H)) * 1.4426950408889634
        qkv = qkv.clone()
        qkv[:, : 2 * d] = (qkv[:, : 2 * d].float() * alpha).bfloat16()
        qkv_ref = qkv.float()
        qkv_ref[:, : 2 * d] /= alpha
    cflib.check(L.cf_op_attention(impl, _p(qkv), _p(pos), _p(rng), _p(ctx), n, c, l, r, d, H, prescaled, _stream()), None, "attention")
    torch.cuda.synchronize()
    ref = _attention_reference(qkv_ref, pos, rng_np, n, c, l, r, d, H)
    assert torch.isfinite(ctx.float()).all()
    err = (ctx.float() - ref).abs().max().item()
    assert err < 4e-2, err


@pytest.mark.parametrize("c,l,r,d,H,n", [(64, 128, 128, 512, 8, 7), (16, 64, 0, 256, 4, 9), (8, 16, 16, 256, 2, 6),
                                         (64, 64, 64, 512, 4, 4)])
@pytest.mark.parametrize("prescaled", [0, 1])
def test_attention_generic(c, l, r, d, H, n, prescaled):
    _attention_case(0, c, l, r, d, H, n, prescaled=prescaled)


@pytest.mark.parametrize("l,r,n", [(128, 128, 7), (128, 128, 12), (64, 64, 5), (128, 0, 6), (0, 0, 3), (192, 64, 9)])
@pytest.mark.parametrize("prescaled", [0, 1])
@pytest.mark.parametrize("version", [1])
def test_attention_tcgen05(l, r, n, prescaled, version):
    """Chunk-pair tcgen05 kernels (c=64, d_k=64), odd and even chunk counts,
    several window shapes."""
    _attention_case(version, 64, l, r, 512, 8, n, seed=3, prescaled=prescaled)


@pytest.mark.parametrize("c,l,r,d,H,n", [(16, 64, 0, 512, 8, 37), (16, 64, 0, 256, 4, 9), (32, 64, 64, 512, 8, 11),
                                         (8, 16, 16, 256, 4, 21), (64, 100, 30, 512, 8, 7), (16, 50, 10, 512, 8, 8),
                                         (32, 128, 96, 256, 4, 5), (8, 40, 0, 512, 8, 16)])
@pytest.mark.parametrize("prescaled", [0, 1])
def test_attention_tcgen05_small_chunks(c, l, r, d, H, n, prescaled):
    """The tcgen05 kernel with tiles of 128 / c chunks (streaming presets such as 16/64/0) and context sizes that are not
    multiples of 64."""
    _attention_case(1, c, l, r, d, H, n, seed=7, prescaled=prescaled)


@pytest.mark.parametrize("c,l,r,d,H,n", [(64, 128, 128, 512, 8, 7), (16, 64, 0, 512, 8, 19), (128, 128, 128, 512, 8, 5),
                                         (256, 64, 64, 512, 4, 3), (374, 0, 0, 512, 4, 3), (64, 256, 256, 512, 8, 9),
                                         (128, 0, 0, 256, 4, 4), (200, 37, 11, 512, 8, 3), (32, 300, 20, 256, 2, 7),
                                         (1000, 0, 0, 512, 8, 2)])
@pytest.mark.parametrize("prescaled", [0, 1])
def test_attention_tcgen05_ring(c, l, r, d, H, n, prescaled):
    """The ring kernel (64-key blocks, streamed position table): d_k 64 and 128, multi-chunk tiles, long chunks (chunk
    128 / 256, full attention = one chunk per utterance with a ragged last tile), long and odd contexts."""
    _attention_case(3, c, l, r, d, H, n, seed=9, prescaled=prescaled)


@pytest.mark.parametrize("l,r,n", [(128, 128, 7), (128, 128, 12), (64, 64, 5), (128, 0, 6), (0, 0, 3), (192, 64, 9)])
@pytest.mark.parametrize("prescaled", [0, 1])
def test_attention_tcgen05_dk128(l, r, n, prescaled):
    """d_k = 128 chunk-pair tcgen05 kernel (rnnt-large / classification geometry: d 512, 4 heads)."""
    _attention_case(1, 64, l, r, 512, 4, n, seed=4, prescaled=prescaled)


def test_attention_tcgen05_dk128_matches_generic_large():
    L = cflib.load()
    c, l, r, d, H, n = 64, 128, 128, 512, 4, 301
    rows = l + n * c + r + 2 * c + 128
    qkv = torch.zeros((rows, 4 * d), device=DEV, dtype=torch.bfloat16)
    qkv[: l + n * c] = _rand((l + n * c, 4 * d), 1.0, 5).bfloat16()
    R = 2 * c + l + r - 1
    pos = torch.zeros(((R + 127) // 128 * 128, d), device=DEV, dtype=torch.bfloat16)
    pos[:R] = _rand((R, d), 1.0, 6).bfloat16()
    rng = torch.zeros((n + 16, 2), dtype=torch.int32)
    rng[:n, 1] = l + c + r
    rng[0, 0] = l
    rng[n - 1, 1] = l + 40
    rng = rng.to(DEV)
    a = torch.zeros((n * c, d), device=DEV, dtype=torch.bfloat16)
    b = torch.zeros((n * c, d), device=DEV, dtype=torch.bfloat16)
    cflib.check(L.cf_op_attention(0, _p(qkv), _p(pos), _p(rng), _p(a), n, c, l, r, d, H, 0, _stream()))
    cflib.check(L.cf_op_attention(1, _p(qkv), _p(pos), _p(rng), _p(b), n, c, l, r, d, H, 0, _stream()))
    torch.cuda.synchronize()
    assert (a.float() - b.float()).abs().max().item() < 4e-2


@pytest.mark.parametrize("version", [1])
def test_attention_tcgen05_matches_generic_large(version):
    L = cflib.load()
    c, l, r, d, H, n = 64, 128, 128, 512, 8, 301
    rows = l + n * c + r + 2 * c + 128
    qkv = torch.zeros((rows, 4 * d), device=DEV, dtype=torch.bfloat16)
    qkv[: l + n * c] = _rand((l + n * c, 4 * d), 1.0, 5).bfloat16()
    R = 2 * c + l + r - 1
    pos = torch.zeros(((R + 127) // 128 * 128, d), device=DEV, dtype=torch.bfloat16)
    pos[:R] = _rand((R, d), 1.0, 6).bfloat16()
    rng = torch.zeros((n + 16, 2), dtype=torch.int32)
    rng[:n, 1] = l + c + r
    rng[0, 0] = l
    rng[n - 1, 1] = l + 40
    rng = rng.to(DEV)
    a = torch.zeros((n * c, d), device=DEV, dtype=torch.bfloat16)
    b = torch.zeros((n * c, d), device=DEV, dtype=torch.bfloat16)
    cflib.check(L.cf_op_attention(0, _p(qkv), _p(pos), _p(rng), _p(a), n, c, l, r, d, H, 0, _stream()))
    cflib.check(L.cf_op_attention(version, _p(qkv), _p(pos), _p(rng), _p(b), n, c, l, r, d, H, 0, _stream()))
    torch.cuda.synchronize()
    assert (a.float() - b.float()).abs().max().item() < 4e-2


@pytest.mark.parametrize("impl", [0, 2])
@pytest.mark.parametrize("d,c,cmvn", [(512, 64, False), (256, 16, True), (512, 8, True)])
def test_frontend_conv0_dw1(impl, d, c, cmvn):
    """conv0 + ReLU + depthwise conv1 (subsampling.py:70-92) on ragged chunks: CUDA-core (0) and channel-major tcgen05 (2)
    versions."""
    import ctypes
    from ctypes import POINTER, c_int32, c_int64
    L = cflib.load()
    size = 8 * (c - 1) + 15
    in_lens = [size, size, size - 37, 5, size]              # chunks 2 and 3 are zero padded at the end
    n = len(in_lens)
    feats = _rand((n * size + 11, 80), 1.0, 1)
    rows = np.array([k * size + (3 if k else 0) for k in range(n)], dtype=np.int64)
    w0, b0 = _rand((d, 1, 3, 3), 0.3, 2), _rand((d,), 0.3, 3)
    w1, b1 = _rand((d, 1, 3, 3), 0.3, 4), _rand((d,), 0.3, 5)
    wpack = torch.cat([w0.view(d, 9), b0.view(d, 1), w1.view(d, 9), b1.view(d, 1)], 1).contiguous()
    mean = _rand((80,), 0.5, 6) if cmvn else None
    istd = (1.0 / (1.0 + 0.2 * torch.rand(80, device=DEV))) if cmvn else None
    T2, F2 = 2 * c + 1, 19
    out = torch.zeros((n * T2 * F2, d), device=DEV, dtype=torch.bfloat16)
    lens_a = np.array(in_lens, dtype=np.int32)
    cflib.check(L.cf_op_frontend_conv(impl, d, _p(feats), rows.ctypes.data_as(POINTER(c_int64)),
                                      lens_a.ctypes.data_as(POINTER(c_int32)), n, c, 80, _p(wpack), _p(mean), _p(istd),
                                      _p(out), _stream()), None, "frontend")
    torch.cuda.synchronize()
    x = torch.zeros((n, size, 80), device=DEV)
    for k in range(n):
        m = min(in_lens[k], size)
        x[k, :m] = feats[rows[k]: rows[k] + m]
    if cmvn:
        x = (x - mean) * istd
    if impl >= 1:
        x = x.bfloat16().float()       # the tensor-core version rounds inputs and conv0 weights to bf16
        w0, b0 = w0.bfloat16().float(), b0.bfloat16().float()
    y = torch.relu(torch.nn.functional.conv2d(x.unsqueeze(1), w0, b0, stride=2))
    y = torch.nn.functional.conv2d(y, w1, b1, stride=2, groups=d)                  # (n, d, T2, F2)
    ref = y.permute(0, 2, 3, 1).reshape(n * T2 * F2, d)
    err = (out.float() - ref).abs().max().item()
    assert err < 5e-2, err
    assert (out.float() - ref).abs().mean().item() < 5e-3


# ------------------------------------------------------------------------------------------------ residual GEMM + fused LayerNorm(s)
def _ln_ref(x, w, b):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


@pytest.mark.parametrize("M,N,K", [(128, 512, 512), (1000, 512, 512), (333, 256, 256), (5000, 512, 2048), (40000, 512, 512),
                                   (777, 256, 2304), (20000, 256, 2048)])
@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("with_resid,with_mask", [(True, False), (True, True), (False, False)])
@pytest.mark.parametrize("variant", [64, 16, 32])
def test_gemm_ln_pair_kernel(M, N, K, mode, with_resid, with_mask, variant):
    """gemm_ln_split_kernel (variant 64: normalisation passes on their own warps), gemm_ln_quad_kernel (32: cluster of four with
    cta_group::2 MMAs) and gemm_ln_kernel (16)
    - CTA pair, row statistics over DSMEM - against fp32 torch: x_new = resid + rowmask * alpha * (A W^T + b)
    followed by one or two LayerNorms, every output the kernel writes; ragged last row block, row masks, zeroed rows."""
    L = cflib.load()
    A = _rand((M, K), 1.0, 1).bfloat16()
    W = _rand((N, K), 1.0 / math.sqrt(K), 2).bfloat16()
    bias = _rand((N,), 0.5, 3)
    resid = (_rand((M, N), 2.0, 4) + 0.7) if with_resid else None       # non-zero row mean: exercises the variance merge
    alpha = 0.5 if with_resid else math.sqrt(N)
    w1, b1 = 1.0 + _rand((N,), 0.1, 5), _rand((N,), 0.1, 6)
    w2, b2 = 1.0 + _rand((N,), 0.1, 7), _rand((N,), 0.1, 8)
    rpc = 16
    n_chunks = (M + rpc - 1) // rpc
    rng = None
    keep = torch.ones(M, dtype=torch.bool, device=DEV)
    if with_mask:
        g = torch.Generator().manual_seed(9)
        lo = torch.randint(0, 4, (n_chunks,), generator=g)
        hi = torch.randint(10, rpc + 1, (n_chunks,), generator=g)
        rng = torch.stack([lo, hi], 1).to(torch.int32).to(DEV)
        rr = torch.arange(M, device=DEV) % rpc
        ch = torch.arange(M, device=DEV) // rpc
        keep = (rr >= rng[ch, 0]) & (rr < rng[ch, 1])
    rows_per_seq = 50
    limit = None
    zero = torch.zeros(M, dtype=torch.bool, device=DEV)
    if mode == 1 and with_mask:
        n_seq = (M + rows_per_seq - 1) // rows_per_seq
        limit = torch.randint(30, rows_per_seq + 1, (n_seq,), generator=torch.Generator().manual_seed(10)).to(torch.int32).to(DEV)
        zero = (torch.arange(M, device=DEV) % rows_per_seq) >= limit[torch.arange(M, device=DEV) // rows_per_seq]
    x_out = torch.full((M, N), 7.0, device=DEV)
    y_out = torch.full((M, N), 7.0, device=DEV, dtype=torch.bfloat16)
    rc = L.cf_op_gemm_ln(_p(A), K, _p(W), K, M, N, K, _p(bias), _p(resid), N if resid is not None else 0, alpha, _p(rng), rpc,
                         mode + variant, _p(w1), _p(b1), _p(w2), _p(b2), _p(x_out), N, _p(y_out), N, _p(limit), rows_per_seq, _stream())
    cflib.check(rc, None, "cf_op_gemm_ln")
    torch.cuda.synchronize()
    upd = alpha * (A.float() @ W.float().T + bias) * keep.unsqueeze(1)
    x_new = upd + (resid if resid is not None else 0.0)
    if mode == 1:
        want_x, want_y = x_new, _ln_ref(x_new, w1, b1)
        want_y = want_y * (~zero).unsqueeze(1)
    elif mode == 2:
        want_x = _ln_ref(x_new, w1, b1)
        want_y = _ln_ref(want_x, w2, b2)
    else:
        want_x = _ln_ref(_ln_ref(x_new, w1, b1), w2, b2)
        want_y = want_x
    assert (x_out - want_x).abs().max().item() < 2e-3 * max(1.0, float(want_x.abs().max()))
    assert (y_out.float() - want_y).abs().max().item() < 3e-2


@pytest.mark.parametrize("M,d,F", [(128, 512, 2048), (1000, 512, 2048), (333, 256, 2048), (20000, 512, 2048), (5000, 256, 512),
                                   (40001, 512, 1024)])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_ffn_fused_kernel(M, d, F, mode):
    """ffn_fused_kernel (w_1 -> SiLU -> w_2 -> residual -> LayerNorm(s) in one kernel, hidden activation exchanged between the
    two CTAs of a cluster over DSMEM) against fp32 torch with the hidden activation rounded to bf16 as the kernel does."""
    L = cflib.load()
    y = _rand((M, d), 1.0, 1).bfloat16()
    W1 = _rand((F, d), 1.0 / math.sqrt(d), 2).bfloat16()
    b1 = _rand((F,), 0.3, 3)
    W2 = _rand((d, F), 1.0 / math.sqrt(F), 4).bfloat16()
    b2 = _rand((d,), 0.3, 5)
    resid = _rand((M, d), 2.0, 6) + 0.5
    w1, bb1 = 1.0 + _rand((d,), 0.1, 7), _rand((d,), 0.1, 8)
    w2, bb2 = 1.0 + _rand((d,), 0.1, 9), _rand((d,), 0.1, 10)
    x_out = torch.full((M, d), 7.0, device=DEV)
    y_out = torch.full((M, d), 7.0, device=DEV, dtype=torch.bfloat16)
    rc = L.cf_op_ffn(_p(y), d, _p(W1), _p(b1), _p(W2), _p(b2), M, d, F, _p(resid), d, 0.5, mode, _p(w1), _p(bb1), _p(w2), _p(bb2),
                     _p(x_out), d, _p(y_out), d, _stream())
    cflib.check(rc, None, "cf_op_ffn")
    torch.cuda.synchronize()
    hid = torch.nn.functional.silu(y.float() @ W1.float().T + b1).bfloat16().float()
    x_new = resid + 0.5 * (hid @ W2.float().T + b2)
    if mode == 1:
        want_x, want_y = x_new, _ln_ref(x_new, w1, bb1)
    elif mode == 2:
        want_x = _ln_ref(x_new, w1, bb1)
        want_y = _ln_ref(want_x, w2, bb2)
    else:
        want_x = _ln_ref(_ln_ref(x_new, w1, bb1), w2, bb2)
        want_y = want_x
    # tanh.approx SiLU + bf16 hidden activation: a few 1e-3 on O(1) outputs
    assert (x_out - want_x).abs().max().item() < 1e-2 * max(1.0, float(want_x.abs().max()))
    assert (y_out.float() - want_y).abs().max().item() < 4e-2
