"""Unit tests of individual sm_100a kernels against plain PyTorch fp32 references
of the same op (on bf16-rounded operands), through the C-ABI debug hooks."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _norm(v, kind, g, b):
    mean = v.mean(-1, keepdim=True)
    var = v.var(-1, keepdim=True, unbiased=(kind == 1))
    return (v - mean) * torch.rsqrt(var + 1e-5) * g + b


def _rnn_tc(kind, x, w_ih, w_hh, b_ih, b_hh, groups=0):
    """x: (nseq, T, 256) bf16 cuda. Returns (nseq, T, 256) bf16."""
    from voiceactivityprojection_b200 import _lib

    lib = _lib.load()
    wcat = torch.empty((1024, 512), dtype=torch.float32)
    bias = torch.empty(1024, dtype=torch.float32)
    rc = lib.vapb_debug_rnn_pack(kind, w_ih.contiguous().data_ptr(), w_hh.contiguous().data_ptr(),
                                 b_ih.contiguous().data_ptr(), b_hh.contiguous().data_ptr(), wcat.data_ptr(),
                                 bias.data_ptr())
    assert rc == 0
    wcat_d, bias_d = wcat.cuda().bfloat16().contiguous(), bias.cuda()
    nseq, T, _ = x.shape
    out = torch.zeros((nseq, T, 256), device="cuda", dtype=torch.bfloat16)
    err = C.create_string_buffer(512)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vapb_debug_rnn_tc(st, kind, x.data_ptr(), T * 256, 256, wcat_d.data_ptr(), bias_d.data_ptr(),
                               out.data_ptr(), T * 256, nseq, T, err, 512, None, groups)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("kind,nseq,T,groups", [(0, 2, 5, 0), (0, 37, 233, 1602), (1, 16, 64, 3202), (1, 70, 200, 1604),
                                                (0, 64, 2000, 0), (0, 300, 50, 3202), (0, 512, 40, 0),
                                                (1, 512, 33, 0)])
def test_rnn_tc_matches_torch(kind, nseq, T, groups):
    """tcgen05 cluster recurrence vs nn.LSTM / nn.GRU (fp32, on the bf16-rounded weights and inputs)."""
    torch.manual_seed(100 * kind + nseq + T)
    cell = (torch.nn.LSTM if kind == 0 else torch.nn.GRU)(256, 256, batch_first=True)
    with torch.no_grad():
        for p in cell.parameters():
            p.copy_(p.bfloat16().float() if p.ndim == 2 else p)
    x = (torch.randn(nseq, T, 256) * 0.7).bfloat16()
    with torch.no_grad():
        ref, _ = cell(x.float())
    out = _rnn_tc(kind, x.cuda().contiguous(), cell.weight_ih_l0.detach(), cell.weight_hh_l0.detach(),
                  cell.bias_ih_l0.detach(), cell.bias_hh_l0.detach(), groups).float().cpu()
    err = (out - ref).abs().max().item()
    # h is rounded to bf16 every step (|h| < 1 -> 2^-9 abs) and fed back; tanh/sigmoid are MUFU approximations
    assert err <= 2e-2, f"max-abs {err}"
    assert (out - ref).abs().mean().item() <= 2e-3


def _attn_ref(q, k, v, slopes, cross, dtype=torch.float32):
    """fp32 torch statement of vap/modules.py:82-110,169-202 on (nseq, T, 256) inputs."""
    nseq, T, _ = q.shape
    H = slopes.numel()
    if cross:
        idx = (torch.arange(nseq, device=q.device) + nseq // 2) % nseq
        k, v = k[idx], v[idx]
    qh, kh, vh = (t.to(dtype).reshape(nseq, T, H, 64).transpose(1, 2) for t in (q, k, v))
    slopes = slopes.to(dtype)
    att = qh @ kh.transpose(-1, -2) * (1.0 / 16.0)
    j = torch.arange(T, device=q.device, dtype=dtype)
    bias = 1.0 + slopes.view(1, H, 1, 1) * j.view(1, 1, 1, T)
    mask = torch.ones(T, T, device=q.device, dtype=torch.bool).tril()
    att = (att + bias).masked_fill(~mask, float("-inf")).softmax(-1)
    return (att @ vh).transpose(1, 2).reshape(nseq, T, 256)


@pytest.mark.parametrize("nseq,T,cross,packed,scale", [(2, 117, 0, True, 1.0), (4, 128, 1, False, 1.0),
                                                       (2, 1000, 0, True, 1.0), (6, 500, 1, False, 4.0),
                                                       (80, 300, 0, True, 8.0), (2, 1250, 1, False, 1.0)])
def test_attention_tc_matches_torch(nseq, T, cross, packed, scale):
    """tcgen05 fused causal ALiBi attention vs a plain fp32 softmax(QK^T/16 + 1 + m*j) V."""
    from voiceactivityprojection_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(nseq * 1000 + T)
    slopes = torch.tensor([0.25, 0.0625, 0.015625, 0.00390625], device="cuda")
    if packed:  # self-attention layout: q | k | v in one (nseq*T, 768) buffer
        buf = (torch.randn((nseq, T, 768), device="cuda", generator=g) * scale).bfloat16()
        q, k, v = buf[..., :256], buf[..., 256:512], buf[..., 512:]
        qs = ks = 768
    else:  # cross-attention layout: q alone, k | v in one (nseq*T, 512) buffer
        q = (torch.randn((nseq, T, 256), device="cuda", generator=g) * scale).bfloat16()
        kv = (torch.randn((nseq, T, 512), device="cuda", generator=g) * scale).bfloat16()
        k, v = kv[..., :256], kv[..., 256:]
        qs, ks = 256, 512
    out = torch.full((nseq, T, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    err = C.create_string_buffer(512)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vapb_debug_attn_tc(st, q.data_ptr(), qs, k.data_ptr(), v.data_ptr(), ks, out.data_ptr(), nseq, T, 4,
                                slopes.data_ptr(), cross, err, 512, None)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, slopes, cross)
    d = (out.float() - ref).abs()
    assert torch.isfinite(out.float()).all()
    # P and the output are rounded to bf16 (2^-9 relative); V is O(scale)
    # P and the output are bf16: allow 2.5 ulp (2^-8 relative) of the largest output; the mean bound is the tight one
    assert d.max().item() <= 2.5 * 2.0 ** -8 * ref.abs().max().item() + 1e-2, d.max().item()
    assert d.mean().item() <= 2e-3 * scale


@pytest.mark.parametrize("nseq,T,cross,scale", [(2, 117, 0, 1.0), (4, 128, 1, 1.0), (2, 1000, 0, 1.0), (6, 500, 1, 3.0),
                                                (40, 300, 0, 6.0), (2, 1250, 1, 1.0)])
def test_attention_x3_matches_torch(nseq, T, cross, scale):
    """fp32-class tensor-core attention of mode fp32_tc (fp16 hi / lo operands, three MMA series per contraction)
    vs a plain fp32 softmax(QK^T/16 + 1 + m*j) V: relative error of the output at fp32 level."""
    from voiceactivityprojection_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(nseq * 1000 + T + 7)
    slopes = torch.tensor([0.25, 0.0625, 0.015625, 0.00390625], device="cuda")
    if not cross:  # q | k | v in one (nseq*T, 768) buffer
        buf = (torch.randn((nseq, T, 768), device="cuda", generator=g) * scale).contiguous()
        q, k, v = buf[..., :256], buf[..., 256:512], buf[..., 512:]
        qb, kvb, qc, kvc, ko, vo = buf, buf, 768, 768, 256, 512
        planes = torch.empty(buf.numel() * 4, device="cuda", dtype=torch.uint8)
    else:  # q alone, k | v in one (nseq*T, 512) buffer
        qb = (torch.randn((nseq, T, 256), device="cuda", generator=g) * scale).contiguous()
        kvb = (torch.randn((nseq, T, 512), device="cuda", generator=g) * scale).contiguous()
        q, k, v = qb, kvb[..., :256], kvb[..., 256:]
        qc, kvc, ko, vo = 256, 512, 0, 256
        planes = torch.empty((qb.numel() + kvb.numel()) * 4, device="cuda", dtype=torch.uint8)
    out = torch.full((nseq, T, 256), float("nan"), device="cuda")
    err = C.create_string_buffer(512)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vapb_debug_attn_x3(st, qb.data_ptr(), qc, kvb.data_ptr(), kvc, ko, vo, planes.data_ptr(), out.data_ptr(),
                                nseq, T, slopes.data_ptr(), cross, err, 512)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, slopes, cross, dtype=torch.float64).float()
    assert torch.isfinite(out).all()
    d = (out - ref).abs()
    # operands carry 22 bits, exp2 is the MUFU approximation (2 ulp), accumulation is fp32
    assert d.max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), d.max().item()
    assert d.mean().item() <= 2e-6 * scale


def test_attention_tc_fp16_far_key_dominates(monkeypatch):
    """fp16 operands, steepest ALiBi head: one key far back in the tile out-scores everything by more than the bias can
    take away. The softmax reference must stay near the TRUE row maximum: a bound that is the tile's largest raw score
    plus the bias of the row's own key sits up to 46 binades above it here, and p (fp16) would underflow to zero."""
    from voiceactivityprojection_b200 import _lib

    monkeypatch.setenv("VAPB_DEBUG_FP16", "1")
    lib = _lib.load()
    nseq, T = 2, 256
    g = torch.Generator(device="cuda").manual_seed(5)
    buf = torch.randn((nseq, T, 768), device="cuda", generator=g) * 0.5
    u = torch.zeros(64, device="cuda")
    u[:8] = 1.0
    for h in range(4):  # every head: key 0 and key 130 carry a large component along u, every query too
        buf[:, :, h * 64:(h + 1) * 64] += 6.0 * u                      # q
        buf[:, 0, 256 + h * 64:256 + (h + 1) * 64] += 18.0 * u         # k of key 0
        buf[:, 130, 256 + h * 64:256 + (h + 1) * 64] += 18.0 * u       # k of key 130
    buf = buf.half().contiguous()
    q, k, v = buf[..., :256], buf[..., 256:512], buf[..., 512:]
    out = torch.full((nseq, T, 256), float("nan"), device="cuda", dtype=torch.float16)
    slopes = torch.tensor([0.25, 0.0625, 0.015625, 0.00390625], device="cuda")
    err = C.create_string_buffer(512)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vapb_debug_attn_tc(st, q.data_ptr(), 768, k.data_ptr(), v.data_ptr(), 768, out.data_ptr(), nseq, T, 4,
                                slopes.data_ptr(), 0, err, 512, None)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, slopes, 0, dtype=torch.float64)
    assert torch.isfinite(out.float()).all()
    d = (out.double() - ref).abs()
    assert d.max().item() <= 4 * 2.0 ** -11 * ref.abs().max().item() + 2e-3, d.max().item()


def _to_blocked(x):
    """(M, 256) fp32 -> row-blocked [ceil(M/128)][64][128][4] flat buffer (k_gemm_lin.cu layout)."""
    M = x.shape[0]
    Mp = (M + 127) // 128 * 128
    xp = torch.zeros((Mp, 256), device=x.device, dtype=torch.float32)
    xp[:M] = x
    return xp.view(Mp // 128, 128, 64, 4).permute(0, 2, 1, 3).contiguous()


def _from_blocked(b, M):
    return b.permute(0, 2, 1, 3).reshape(-1, 256)[:M]


def _gemm_lin(A, W, M, N, K, bias=None, norm1=0, g1=None, b1=None, act=0, resid=None, accumulate=False, out1_init=None,
              f32_mode=1, want_f32=True, want_bf16=False, norm2=0, g2=None, b2=None, nseq=1, a_seq_stride=0,
              a_row_stride=None):
    from voiceactivityprojection_b200 import _lib

    lib = _lib.load()
    dev = A.device
    rps = M // nseq
    out1 = None
    if want_f32:
        if f32_mode == 1:
            out1 = _to_blocked(out1_init if out1_init is not None else torch.full((M, N), float("nan"), device=dev))
        else:
            out1 = torch.full((M, N), float("nan"), device=dev)
    out1b = torch.zeros((M, N), device=dev, dtype=torch.bfloat16) if want_bf16 else None
    out2 = torch.zeros((M, N), device=dev, dtype=torch.bfloat16) if norm2 else None
    rb = _to_blocked(resid) if resid is not None else None
    err = C.create_string_buffer(512)
    p = lambda t: None if t is None else t.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vapb_debug_gemm_lin(st, A.data_ptr(), a_seq_stride, a_row_stride or K, W.data_ptr(), nseq, rps, N, K, p(bias),
                                 norm1, p(g1), p(b1), act, p(rb), int(accumulate), p(out1), f32_mode, p(out1b), norm2,
                                 p(g2), p(b2), p(out2), err, 512)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    if want_f32 and f32_mode == 1:
        out1 = _from_blocked(out1, M)
    return out1, out1b, out2


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 768, 256), (4096 + 77, 512, 256), (300, 256, 768)])
def test_gemm_lin_plain_bf16_out(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    W = (torch.randn((N, K), device="cuda", generator=g) * 0.05).bfloat16()
    _, outb, _ = _gemm_lin(A, W, M, N, K, want_f32=False, want_bf16=True)
    ref = A.float() @ W.float().T
    assert (outb.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())


def test_gemm_lin_rowmajor_f32_with_bias():
    M, N, K = 1000 + 37, 256, 256
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    W = (torch.randn((N, K), device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out, _, _ = _gemm_lin(A, W, M, N, K, bias=bias, f32_mode=2)
    ref = A.float() @ W.float().T + bias
    assert (out - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def test_gemm_lin_residual_layernorm2_blocked():
    M, K = 2000 + 5, 768
    g = torch.Generator(device="cuda").manual_seed(12)
    A = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    W = (torch.randn((256, K), device="cuda", generator=g) * 0.05).bfloat16()
    resid = torch.randn((M, 256), device="cuda", generator=g)
    g2 = torch.randn(256, device="cuda", generator=g)
    b2 = torch.randn(256, device="cuda", generator=g)
    out1, out1b, out2 = _gemm_lin(A, W, M, 256, K, resid=resid, want_bf16=True, norm2=2, g2=g2, b2=b2)
    v = A.float() @ W.float().T + resid
    assert (out1 - v).abs().max().item() <= 2e-3 * v.abs().max().item()
    assert (out1b.float() - v).abs().max().item() <= 1e-2 * v.abs().max().item()
    assert (out2.float() - _norm(v, 2, g2, b2)).abs().max().item() <= 5e-2


def test_gemm_lin_norm1_gelu_accumulate():
    """Combinator branch: GELU(LN(x W^T)) accumulated onto the first branch (vap/modules.py:446-449)."""
    M, K = 900, 256
    g = torch.Generator(device="cuda").manual_seed(13)
    A = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    W = (torch.randn((256, K), device="cuda", generator=g) * 0.08).bfloat16()
    g1 = torch.randn(256, device="cuda", generator=g)
    b1 = torch.randn(256, device="cuda", generator=g)
    prev = torch.randn((M, 256), device="cuda", generator=g)
    out1, out1b, _ = _gemm_lin(A, W, M, 256, K, norm1=2, g1=g1, b1=b1, act=2, accumulate=True, out1_init=prev,
                               want_bf16=True)
    v = F.gelu(_norm(A.float() @ W.float().T, 2, g1, b1)) + prev
    assert (out1 - v).abs().max().item() <= 3e-3   # tanh-form GELU on tanh.approx: <= ~1e-3 abs
    assert (out1b.float() - v).abs().max().item() <= 3e-2


def test_gemm_lin_implicit_conv_multi_seq():
    """Downsample conv: k=5, s=2 over left-padded sequences, LayerNorm + GELU, blocked fp32 + bf16 + LN2 outputs."""
    g = torch.Generator(device="cuda").manual_seed(14)
    nseq, L, k, s = 3, 233, 5, 2
    T = (L - 1) // 2 + 1
    Lpad = 4 + L + (L & 1)
    x = torch.randn((nseq, L, 256), device="cuda", generator=g).bfloat16()
    buf = torch.zeros((nseq, Lpad, 256), device="cuda", dtype=torch.bfloat16)
    buf[:, 4:4 + L] = x
    w = torch.randn((256, 256, k), device="cuda", generator=g) * 0.03
    Wp = w.permute(0, 2, 1).reshape(256, k * 256).contiguous().bfloat16()
    bias = torch.randn(256, device="cuda", generator=g)
    g1 = torch.randn(256, device="cuda", generator=g)
    b1 = torch.randn(256, device="cuda", generator=g)
    g2 = torch.randn(256, device="cuda", generator=g)
    b2 = torch.randn(256, device="cuda", generator=g)
    out1, out1b, out2 = _gemm_lin(buf, Wp, nseq * T, 256, k * 256, bias=bias, norm1=2, g1=g1, b1=b1, act=2,
                                  want_bf16=True, norm2=2, g2=g2, b2=b2, nseq=nseq, a_seq_stride=Lpad * 256,
                                  a_row_stride=s * 256)
    wr = Wp.float().reshape(256, k, 256).permute(0, 2, 1)
    y = F.conv1d(F.pad(x.float().transpose(1, 2), (4, 0)), wr, bias, stride=s).transpose(1, 2).reshape(nseq * T, 256)
    v = F.gelu(_norm(y, 2, g1, b1))
    assert (out1 - v).abs().max().item() <= 2e-2
    assert (out1b.float() - v).abs().max().item() <= 4e-2
    assert (out2.float() - _norm(v, 2, g2, b2)).abs().max().item() <= 8e-2


@pytest.mark.parametrize("M,with_ln", [(128, True), (1000 + 77, True), (5000, False)])
def test_ffn_fused_matches_torch(M, with_ln):
    """x + W2 GELU(W1 z) with the hidden activation kept on chip, + bf16 shadow + next LayerNorm."""
    from voiceactivityprojection_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M)
    z = torch.randn((M, 256), device="cuda", generator=g).bfloat16()
    w1 = (torch.randn((768, 256), device="cuda", generator=g) * 0.06).bfloat16()
    w2 = (torch.randn((256, 768), device="cuda", generator=g) * 0.04).bfloat16()
    resid = torch.randn((M, 256), device="cuda", generator=g)
    g2 = torch.randn(256, device="cuda", generator=g)
    b2 = torch.randn(256, device="cuda", generator=g)
    rb = _to_blocked(resid)
    xo = _to_blocked(torch.full((M, 256), float("nan"), device="cuda"))
    xs = torch.zeros((M, 256), device="cuda", dtype=torch.bfloat16)
    zn = torch.zeros((M, 256), device="cuda", dtype=torch.bfloat16) if with_ln else None
    err = C.create_string_buffer(512)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vapb_debug_ffn_fused(st, z.data_ptr(), w1.data_ptr(), w2.data_ptr(), rb.data_ptr(), xo.data_ptr(),
                                  xs.data_ptr(), zn.data_ptr() if with_ln else None, g2.data_ptr(), b2.data_ptr(), M,
                                  err, 512, None)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    h = F.gelu(z.float() @ w1.float().T).bfloat16().float()  # the kernel rounds the hidden activation to bf16
    ref = resid + h @ w2.float().T
    out = _from_blocked(xo, M)
    assert (out - ref).abs().max().item() <= 1e-2, (out - ref).abs().max().item()
    assert (out - ref).abs().mean().item() <= 1e-3
    assert (xs.float() - ref).abs().max().item() <= 3e-2
    if with_ln:
        assert (zn.float() - _norm(ref, 2, g2, b2)).abs().max().item() <= 6e-2


@pytest.mark.parametrize("nseq,L,k,s,pad", [(1, 1024, 4, 2, 1), (3, 1000, 8, 4, 2), (5, 4000, 4, 2, 1)])
def test_gemm_2sm_conv_channelnorm_relu(nseq, L, k, s, pad):
    """CTA-pair (cta_group::2) implicit-GEMM conv + ChannelNorm + ReLU vs torch."""
    from voiceactivityprojection_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(nseq * 100 + k)
    Lout = (L + 2 * pad - k) // s + 1
    Lpad = ((max(s * (Lout - 1) + k, pad + L) + s - 1) // s) * s
    x = torch.randn((nseq, L, 256), device="cuda", generator=g).bfloat16()
    buf = torch.zeros((nseq, Lpad, 256), device="cuda", dtype=torch.bfloat16)
    buf[:, pad:pad + L] = x
    w = torch.randn((256, 256, k), device="cuda", generator=g) * 0.03
    Wp = w.permute(0, 2, 1).reshape(256, k * 256).contiguous().bfloat16()
    bias = torch.randn(256, device="cuda", generator=g)
    g1 = torch.randn(256, device="cuda", generator=g)
    b1 = torch.randn(256, device="cuda", generator=g)
    out = torch.full((nseq * Lout, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    err = C.create_string_buffer(512)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vapb_debug_gemm_2sm(st, buf.data_ptr(), Lpad * 256, s * 256, Wp.data_ptr(), nseq, Lout, k * 256,
                                 bias.data_ptr(), 1, g1.data_ptr(), b1.data_ptr(), 1, out.data_ptr(), err, 512)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    wr = Wp.float().reshape(256, k, 256).permute(0, 2, 1)
    y = F.conv1d(x.float().transpose(1, 2), wr, bias, stride=s, padding=pad).transpose(1, 2).reshape(nseq * Lout, 256)
    ref = F.relu(_norm(y, 1, g1, b1))
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 4e-2


def _conv01_case(B, S, fp16, seed, dbg=None):
    """Fused conv0 -> conv1 kernel (k_conv01.cu) against torch fp32 on the same 16-bit-rounded hand-off."""
    from voiceactivityprojection_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    dt16 = torch.float16 if fp16 else torch.bfloat16
    wav = torch.randn((B, 2, S), device="cuda", generator=g) * 0.05
    w0 = torch.randn((256, 1, 10), device="cuda", generator=g) * 0.3
    b0 = torch.randn(256, device="cuda", generator=g) * 0.1
    g0 = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    be0 = 0.1 * torch.randn(256, device="cuda", generator=g)
    w1 = torch.randn((256, 256, 8), device="cuda", generator=g) * 0.03
    W1p = w1.permute(0, 2, 1).reshape(256, 8 * 256).contiguous().to(dt16)  # [N][tap*256 + cin]
    b1 = torch.randn(256, device="cuda", generator=g) * 0.1
    g1 = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    be1 = 0.1 * torch.randn(256, device="cuda", generator=g)
    L0 = (S + 6 - 10) // 5 + 1
    L1 = (L0 + 4 - 8) // 4 + 1
    pad, rows = 1, 1 + (L1 + 127) // 128 * 128 + 3
    out = torch.full((2 * B, rows, 256), float("nan"), device="cuda", dtype=dt16)
    err = C.create_string_buffer(512)
    st = torch.cuda.current_stream().cuda_stream
    host = [t.detach().cpu().contiguous() for t in (w0, b0, g0, be0)]
    rc = lib.vapb_debug_conv01(st, wav.data_ptr(), B, S, host[0].data_ptr(), host[1].data_ptr(), host[2].data_ptr(),
                               host[3].data_ptr(), W1p.data_ptr(), b1.data_ptr(), g1.data_ptr(), be1.data_ptr(),
                               out.data_ptr(), rows * 256, pad, int(fp16), err, 512,
                               None if dbg is None else dbg.data_ptr())
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    x = wav.transpose(0, 1).reshape(2 * B, 1, S)  # sequences in channel-major order c * B + item
    y0 = F.conv1d(x, w0, b0, stride=5, padding=3).transpose(1, 2)  # (2B, L0, 256)
    y0 = F.relu(_norm(y0, 1, g0, be0)).to(dt16).float()
    wr = W1p.float().reshape(256, 8, 256).permute(0, 2, 1)
    y1 = F.conv1d(y0.transpose(1, 2), wr, b1, stride=4, padding=2).transpose(1, 2)  # (2B, L1, 256)
    ref = F.relu(_norm(y1, 1, g1, be1))
    assert y1.shape[1] == L1
    return out, ref, pad, L1


@pytest.mark.parametrize("B,S,fp16", [(1, 2000, 0), (1, 40000, 0), (2, 37392, 1), (3, 320000, 1), (2, 81920 * 5 + 7, 0)])
def test_conv01_fused_matches_torch(B, S, fp16):
    out, ref, pad, L1 = _conv01_case(B, S, fp16, seed=B * 1000 + S % 997)
    got = out[:, pad:pad + L1].float()
    assert torch.isfinite(got).all()
    err = (got - ref).abs()
    tol = 8e-3 if fp16 else 5e-2
    assert err.max().item() <= tol, f"max-abs {err.max().item()} at {divmod(int(err.argmax()), 256)}"
    assert err.mean().item() <= tol / 10
    tail = out[:, pad + L1:pad + (L1 + 127) // 128 * 128].float()
    assert (tail == 0).all()                      # rows past the sequence end are the next layer's zero padding
    assert torch.isnan(out[:, :pad].float()).all()  # rows before the sequence are not touched
    assert torch.isnan(out[:, pad + (L1 + 127) // 128 * 128:].float()).all()


@pytest.mark.parametrize("B,S,fp16", [(1, 2000, 0), (2, 37392, 1), (3, 320000, 0)])
def test_conv0_tc_standalone_matches_torch(B, S, fp16):
    """The stand-alone tf32 conv0 + ChannelNorm + ReLU kernel (k_conv0_tc.cu, used when VAPB_CONV01=0) against torch fp32."""
    from voiceactivityprojection_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(B + S)
    dt16 = torch.float16 if fp16 else torch.bfloat16
    wav = torch.randn((B, 2, S), device="cuda", generator=g) * 0.05
    w0 = torch.randn((256, 1, 10), device="cuda", generator=g) * 0.3
    b0 = torch.randn(256, device="cuda", generator=g) * 0.1
    g0 = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    be0 = 0.1 * torch.randn(256, device="cuda", generator=g)
    L0 = (S + 6 - 10) // 5 + 1
    out = torch.full((2 * B, L0, 256), float("nan"), device="cuda", dtype=dt16)
    err = C.create_string_buffer(512)
    host = [t.detach().cpu().contiguous() for t in (w0, b0, g0, be0)]
    rc = lib.vapb_debug_conv0_tc(torch.cuda.current_stream().cuda_stream, wav.data_ptr(), B, S, host[0].data_ptr(),
                                 host[1].data_ptr(), host[2].data_ptr(), host[3].data_ptr(), out.data_ptr(), int(fp16),
                                 err, 512)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    x = wav.transpose(0, 1).reshape(2 * B, 1, S)
    ref = F.relu(_norm(F.conv1d(x, w0, b0, stride=5, padding=3).transpose(1, 2), 1, g0, be0))
    assert torch.isfinite(out.float()).all()
    d = (out.float() - ref).abs()
    # tf32 operands (2^-11) and a 16-bit result: absolute error scales with the output magnitude (|ref| up to ~6)
    assert d.max().item() <= (6e-3 if fp16 else 4e-2), d.max().item()
    assert d.mean().item() <= (5e-4 if fp16 else 3e-3)


def _gemm_x3(A, a_seq_stride, a_row_stride, Wt, nseq, rps, N, K, bias=None, norm1=0, g1=None, b1=None, act=0, resid=None,
             accumulate=False, out1_init=None, norm2=0, g2=None, b2=None):
    from voiceactivityprojection_b200 import _lib

    lib = _lib.load()
    M = nseq * rps
    out1 = out1_init.clone() if out1_init is not None else torch.full((M, N), float("nan"), device="cuda")
    out2 = torch.full((M, N), float("nan"), device="cuda") if norm2 else None
    wt = Wt.detach().cpu().contiguous()
    err = C.create_string_buffer(512)
    p = lambda t: None if t is None else t.data_ptr()
    rc = lib.vapb_debug_gemm_x3(torch.cuda.current_stream().cuda_stream, A.data_ptr(), a_seq_stride, a_row_stride,
                                wt.data_ptr(), nseq, rps, N, K, p(bias), norm1, p(g1), p(b1), act, p(resid),
                                int(accumulate), out1.data_ptr(), norm2, p(g2), p(b2), p(out2), err, 512)
    assert rc == 0, err.value.decode()
    torch.cuda.synchronize()
    return out1, out2


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (1000, 768, 256), (4096 + 77, 256, 768), (333, 1024, 256)])
def test_gemm_x3_plain_is_fp32_class(M, N, K):
    """Split-fp16 tensor-core GEMM of the parity mode against an fp64 reference: error at the level of an fp32 GEMM."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((M, K), device="cuda", generator=g) * 3
    Wt = torch.randn((K, N), device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    out, _ = _gemm_x3(A, 0, K, Wt, 1, M, N, K, bias=bias)
    ref = (A.double() @ Wt.double() + bias.double())
    scale = (A.abs().double() @ Wt.abs().double()).max().item()
    assert (out.double() - ref).abs().max().item() <= 2e-6 * scale
    f32 = (A @ Wt + bias).double()
    assert (out.double() - ref).abs().max().item() <= 4 * max((f32 - ref).abs().max().item(), 1e-7 * scale)


def test_gemm_x3_implicit_conv_norm_relu_and_residual_layernorm2():
    g = torch.Generator(device="cuda").manual_seed(11)
    nseq, L, k, s, pad = 3, 1000, 8, 4, 2
    Lout = (L + 2 * pad - k) // s + 1
    Lpad = ((max(s * (Lout - 1) + k, pad + L) + s - 1) // s) * s
    x = torch.randn((nseq, L, 256), device="cuda", generator=g)
    buf = torch.zeros((nseq, Lpad, 256), device="cuda")
    buf[:, pad:pad + L] = x
    w = torch.randn((256, 256, k), device="cuda", generator=g) * 0.03
    Wt = w.permute(2, 1, 0).reshape(k * 256, 256).contiguous()  # [(tap*256 + cin)][out]
    bias = torch.randn(256, device="cuda", generator=g) * 0.1
    g1 = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    b1 = 0.1 * torch.randn(256, device="cuda", generator=g)
    out, _ = _gemm_x3(buf, Lpad * 256, s * 256, Wt, nseq, Lout, 256, k * 256, bias=bias, norm1=1, g1=g1, b1=b1, act=1)
    y = F.conv1d(x.double().transpose(1, 2), w.double(), bias.double(), stride=s, padding=pad).transpose(1, 2)
    ref = F.relu(_norm(y, 1, g1.double(), b1.double())).reshape(-1, 256)
    assert (out.double() - ref).abs().max().item() <= 5e-5  # outputs up to ~5; K = 2048 products of 22-bit operands
    # Linear + residual + LayerNorm2, then GELU + accumulate
    M, K = 777, 256
    A = torch.randn((M, K), device="cuda", generator=g)
    W2 = torch.randn((K, 256), device="cuda", generator=g) * 0.05
    resid = torch.randn((M, 256), device="cuda", generator=g) * 2
    g2 = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    b2 = 0.1 * torch.randn(256, device="cuda", generator=g)
    o1, o2 = _gemm_x3(A, 0, K, W2, 1, M, 256, K, resid=resid, norm2=2, g2=g2, b2=b2)
    v = A.double() @ W2.double() + resid.double()
    assert (o1.double() - v).abs().max().item() <= 1e-5
    assert (o2.double() - _norm(v, 2, g2.double(), b2.double())).abs().max().item() <= 1e-5
    o3, _ = _gemm_x3(A, 0, K, W2, 1, M, 256, K, norm1=2, g1=g1, b1=b1, act=2, accumulate=True, out1_init=o1)
    ref3 = v + F.gelu(_norm(A.double() @ W2.double(), 2, g1.double(), b1.double()))
    assert (o3.double() - ref3).abs().max().item() <= 3e-5  # LayerNorm divides the GEMM's error by sigma ~ 0.8; |v| up to ~10
