"""Per-kernel parity of the CUDA path (through the C ABI) against plain PyTorch fp32 on the same bf16-rounded inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from hipt_abmil_atec23_b200 import _lib as L
    return L


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype)


@pytest.mark.parametrize("M,N,K", [(128, 384, 384), (257, 1152, 384), (1000, 1536, 384), (514, 384, 1536),
                                   (65792, 1152, 384), (257, 576, 192), (257, 192, 768), (300, 768, 192)])
def test_gemm_bias_bf16(M, N, K):
    L = _lib()
    a = _rand((M, K), 1).cuda().bfloat16()
    w = _rand((N, K), 2, 0.05).cuda().bfloat16()
    b = _rand((N,), 3, 0.1).cuda()
    out = L.gemm_bf16(a, w, b, L.HB_EPI_BIAS_BF16)
    ref = a.float() @ w.float().t() + b
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("M,N,K", [(257, 1536, 384), (4112, 768, 192), (65792, 1536, 384)])
def test_gemm_gelu_bf16(M, N, K, fast):
    """exact-erf GELU epilogue and the tanh-form GELU fitted to erf that the MLP hot path uses, both against
    F.gelu (exact erf) in fp32: the fast form must stay inside bf16 output rounding (2^-9 relative) + 3e-4 |x|."""
    L = _lib()
    a = _rand((M, K), 4, 2.0).cuda().bfloat16()
    w = _rand((N, K), 5, 0.05).cuda().bfloat16()
    b = _rand((N,), 6, 0.1).cuda()
    out = L.gemm_bf16(a, w, b, L.HB_EPI_BIAS_GELU_FAST_BF16 if fast else L.HB_EPI_BIAS_GELU_BF16)
    pre = a.float() @ w.float().t() + b
    ref = F.gelu(pre)
    assert pre.abs().max().item() > 4.0                      # the tails of the approximation are exercised
    err = (out.float() - ref).abs()
    bound = ref.abs() * 2.0 ** -8 + 3e-4 * pre.abs() + 1e-5
    assert bool((err <= bound).all()), (err - bound).max().item()


@pytest.mark.parametrize("M,N,K", [(257, 384, 384), (1000, 384, 1536), (65792, 384, 1536), (257, 192, 768)])
def test_gemm_resadd_f32(M, N, K):
    L = _lib()
    a = _rand((M, K), 7).cuda().bfloat16()
    w = _rand((N, K), 8, 0.05).cuda().bfloat16()
    b = _rand((N,), 9, 0.1).cuda()
    x0 = _rand((M, N), 10).cuda()
    x = x0.clone()
    L.gemm_bf16(a, w, b, L.HB_EPI_BIAS_RESADD_F32, out=x)
    ref = x0 + a.float() @ w.float().t() + b
    err = (x - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M,N,K,gelu", [(257, 1152, 384, 0), (65792, 1152, 384, 0), (65792, 1536, 384, 1),
                                        (65792, 1536, 384, 2), (1000, 576, 192, 0), (300, 768, 192, 1), (300, 768, 192, 2)])
def test_gemm_layernorm_folded(M, N, K, gelu):
    """LayerNorm(eps 1e-6) + Linear (+GELU) as one GEMM over the un-normalised bf16 rows, against fp32
    F.layer_norm -> F.linear -> F.gelu on the same fp32 rows.  M = 65792 rows = 257 CTA-pair tiles also covers the
    one-tile-ahead prefetch of the per-tile side inputs and the rotating chunk assignment of the epilogue warpgroups."""
    L = _lib()
    x = (_rand((M, K), 50, 1.5) + 0.3 * _rand((1, K), 51)).cuda()          # per-channel offsets: non-zero row means
    gamma = (1.0 + 0.2 * _rand((K,), 52)).cuda()
    beta = (0.1 * _rand((K,), 53)).cuda()
    W = _rand((N, K), 54, 0.05).cuda()
    b = _rand((N,), 55, 0.1).cuda()
    wg = (W * gamma[None, :]).bfloat16()
    c = wg.float().sum(1)
    d = W @ beta + b
    stats = torch.stack([x.sum(1), (x * x).sum(1)], dim=1).contiguous()
    out = L.gemm_lnfold_bf16(x.bfloat16(), wg, c, d, stats, 1e-6, gelu=gelu)
    ref = F.linear(F.layer_norm(x, (K,), gamma, beta, 1e-6), W, b)
    if gelu:
        ref = F.gelu(ref) * float(gelu)          # gelu = 2: twice the GELU (0.5 folded into the consumer's weights)
    err = (out.float() - ref).abs().max().item()
    assert err <= 3e-2 * max(1.0, ref.abs().max().item()), err
    assert F.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item() > 0.9999


@pytest.mark.parametrize("M,N,K", [(257, 384, 384), (1000, 384, 1536), (65792, 384, 1536), (65792, 384, 384), (257, 192, 768)])
def test_gemm_residual_with_row_stats(M, N, K):
    L = _lib()
    a = _rand((M, K), 60).cuda().bfloat16()
    w = _rand((N, K), 61, 0.05).cuda().bfloat16()
    b = _rand((N,), 62, 0.1).cuda()
    x0 = _rand((M, N), 63).cuda()
    x = x0.clone()
    xb = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros((M, 2), device="cuda")
    other = torch.full((M, 2), 5.0, device="cuda")
    L.gemm_resid_stats(a, w, b, x, xb, stats, other)
    ref = x0 + a.float() @ w.float().t() + b
    assert (x - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
    assert torch.equal(xb, x.bfloat16())
    assert torch.allclose(stats[:, 0], x.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(stats[:, 1], (x * x).sum(1), rtol=1e-4, atol=1e-2)
    assert torch.all(other == 0)
    # run-to-run determinism of the statistics (one atomic per (row, n-tile), two n-tiles: order independent)
    x2 = x0.clone(); stats2 = torch.zeros((M, 2), device="cuda")
    L.gemm_resid_stats(a, w, b, x2, xb, stats2, None)
    assert torch.equal(stats2, stats) and torch.equal(x2, x)


@pytest.mark.parametrize("M,N,K", [(257, 384, 384), (1000, 384, 1536), (65792, 384, 384), (65792, 384, 1536), (257, 192, 768),
                                   (4112, 192, 192)])
def test_gemm_residual_bf16_stream(M, N, K):
    """x = bf16(x + a W^T + b) in place on the bf16 residual stream, with the per-64-column partial row statistics of
    the unrounded values (bit-exact run to run: plain stores, no atomics)."""
    L = _lib()
    a = _rand((M, K), 60).cuda().bfloat16()
    w = _rand((N, K), 61, 0.05).cuda().bfloat16()
    b = _rand((N,), 62, 0.1).cuda()
    x0 = _rand((M, N), 63).cuda().bfloat16()
    x = x0.clone()
    _, stats = L.gemm_resid_bf16(a, w, b, x)
    ref = x0.float() + a.float() @ w.float().t() + b
    err = (x.float() - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -8 + 1e-3).all()), err.max().item()
    part = ref.view(M, N // 64, 64)
    assert torch.allclose(stats[:, :, 0].t(), part.sum(2), rtol=1e-4, atol=2e-2)
    assert torch.allclose(stats[:, :, 1].t(), (part * part).sum(2), rtol=1e-4, atol=2e-2)
    x2 = x0.clone()
    _, stats2 = L.gemm_resid_bf16(a, w, b, x2)
    assert torch.equal(x2, x) and torch.equal(stats2, stats)


def test_gemm_residual_bf16_strided_source():
    """The CLS rows of a [n_seq, 257, 384] stream as the residual source, compact destination (last-block tail)."""
    L = _lib()
    n_seq, S, N, K = 256, 257, 384, 384
    a = _rand((n_seq, K), 70).cuda().bfloat16()
    w = _rand((N, K), 71, 0.05).cuda().bfloat16()
    b = _rand((N,), 72, 0.1).cuda()
    stream = _rand((n_seq * S, N), 73).cuda().bfloat16()
    out = torch.zeros((n_seq, N), dtype=torch.bfloat16, device="cuda")
    before = stream.clone()
    L.gemm_resid_bf16(a, w, b, stream, out=out, res_pitch_bytes=S * N * 2)
    ref = stream[::S].float() + a.float() @ w.float().t() + b
    assert (out.float() - ref).abs().max().item() <= 2.0 ** -8 * ref.abs().max().item() + 1e-3
    assert torch.equal(stream, before)


def test_gemm_layernorm_folded_partial_planes():
    """LN-folded GEMM reading the row statistics as 6 partial planes with a plane stride larger than M."""
    L = _lib()
    M, N, K = 1000, 1152, 384
    x = (_rand((M, K), 80, 1.5) + 0.3 * _rand((1, K), 81)).cuda()
    gamma = (1.0 + 0.2 * _rand((K,), 82)).cuda(); beta = (0.1 * _rand((K,), 83)).cuda()
    W = _rand((N, K), 84, 0.05).cuda(); b = _rand((N,), 85, 0.1).cuda()
    wg = (W * gamma[None, :]).bfloat16(); c = wg.float().sum(1); d = W @ beta + b
    planes = torch.zeros((6, 1024, 2), device="cuda")
    part = x.view(M, 6, 64)
    planes[:, :M, 0] = part.sum(2).t(); planes[:, :M, 1] = (part * part).sum(2).t()
    out = L.gemm_lnfold_bf16(x.bfloat16(), wg, c, d, planes, 1e-6)
    ref = F.linear(F.layer_norm(x, (K,), gamma, beta, 1e-6), W, b)
    assert F.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item() > 0.9999
    assert (out.float() - ref).abs().max().item() <= 3e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M", [256, 1000, 65792])
def test_mlp_fused(M):
    """fc1 (norm2 folded) -> GELU -> fc2 -> residual as one kernel against fp32 torch on the same bf16 stream."""
    L = _lib()
    D, H = 384, 1536
    x = (_rand((M, D), 100, 1.5) + 0.3 * _rand((1, D), 101)).cuda()
    xb0 = x.bfloat16()
    gamma = (1.0 + 0.2 * _rand((D,), 102)).cuda(); beta = (0.1 * _rand((D,), 103)).cuda()
    W1 = _rand((H, D), 104, 0.05).cuda(); b1 = _rand((H,), 105, 0.1).cuda()
    W2 = _rand((D, H), 106, 0.03).cuda(); b2 = _rand((D,), 107, 0.1).cuda()
    w1g = (W1 * gamma[None, :]).bfloat16(); c1 = w1g.float().sum(1); d1 = W1 @ beta + b1
    w2h = (0.5 * W2).bfloat16()
    part = x.view(M, 6, 64)
    stats_in = torch.stack([part.sum(2).t(), (part * part).sum(2).t()], dim=2).contiguous()      # [6, M, 2]
    xb = xb0.clone()
    stats_out = L.mlp_fused_bf16(xb, w1g, c1, d1, w2h, b2, stats_in)
    torch.cuda.synchronize()
    hid = F.gelu(F.linear(F.layer_norm(x, (D,), gamma, beta, 1e-6), W1, b1))
    ref = xb0.float() + F.linear(hid, W2, b2)
    err = (xb.float() - ref).abs()
    assert F.cosine_similarity(xb.float().flatten(), ref.flatten(), dim=0).item() > 0.9999
    assert err.max().item() <= 4e-2 * max(1.0, ref.abs().max().item()), err.max().item()
    # statistics of the produced rows (compare with the kernel's own rounded output, loosely: they are of the unrounded values)
    po = xb.float().view(M, 6, 64)
    assert torch.allclose(stats_out[:, :M, 0].t(), po.sum(2), rtol=2e-2, atol=0.5)
    assert torch.allclose(stats_out[:, :M, 1].t(), (po * po).sum(2), rtol=2e-2, atol=0.5)
    xb2 = xb0.clone()
    stats2 = L.mlp_fused_bf16(xb2, w1g, c1, d1, w2h, b2, stats_in)
    assert torch.equal(xb2, xb) and torch.equal(stats2[:, :M], stats_out[:, :M])


@pytest.mark.parametrize("n_seq,T,N,K,gelu", [(3, 256, 384, 768, False), (2, 256, 192, 384, True), (3, 35, 192, 384, True)])
def test_gemm_tokens(n_seq, T, N, K, gelu):
    L = _lib()
    M = n_seq * T
    a = _rand((M, K), 11).cuda().bfloat16()
    w = _rand((N, K), 12, 0.05).cuda().bfloat16()
    b = _rand((N,), 13, 0.1).cuda()
    tab = _rand((T + 1, N), 14, 0.02).cuda()
    out = torch.full((n_seq * (T + 1), N), 7.0, device="cuda")
    L.gemm_bf16(a, w, b, L.HB_EPI_TOKENS_GELU_F32 if gelu else L.HB_EPI_TOKENS_F32, out=out, tok_table=tab,
                tokens_per_seq=T)
    y = a.float() @ w.float().t() + b
    if gelu:
        y = F.gelu(y)
    y = y.view(n_seq, T, N) + tab[1:]
    o = out.view(n_seq, T + 1, N)
    assert torch.all(o[:, 0] == 7.0)          # CLS slots untouched
    err = (o[:, 1:] - y).abs().max().item()
    assert err <= 1e-3 * max(1.0, y.abs().max().item()), err


@pytest.mark.parametrize("rows,dim", [(1, 384), (257, 384), (65792, 384), (257, 192), (5, 192)])
def test_layernorm(rows, dim):
    L = _lib()
    x = _rand((rows, dim), 20, 3.0).cuda() + 0.5
    g = _rand((dim,), 21).cuda()
    b = _rand((dim,), 22).cuda()
    ob, of = L.layernorm(x, g, b, 1e-6, rows, dim, want_bf16=True, want_f32=True)
    ref = F.layer_norm(x, (dim,), g, b, 1e-6)
    assert (of - ref).abs().max().item() < 2e-5
    assert (ob.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


def test_layernorm_bf16_rows_strided():
    """Final norm as the plan runs it: bf16 CLS rows of a [n, S, dim] stream."""
    L = _lib()
    n, S, dim = 256, 257, 384
    x = _rand((n * S, dim), 90, 2.0).cuda().bfloat16()
    g = (1.0 + 0.1 * _rand((dim,), 91)).cuda(); b = (0.1 * _rand((dim,), 92)).cuda()
    _, of = L.layernorm(x, g, b, 1e-6, n, dim, row_stride=S * dim, want_bf16=False, want_f32=True)
    ref = F.layer_norm(x[::S].float(), (dim,), g, b, 1e-6)
    assert (of - ref).abs().max().item() < 1e-4


def test_layernorm_strided_cls_rows():
    L = _lib()
    n, S, dim = 7, 257, 384
    x = _rand((n * S, dim), 23).cuda()
    g = _rand((dim,), 24).cuda()
    b = _rand((dim,), 25).cuda()
    _, of = L.layernorm(x, g, b, 1e-6, n, dim, row_stride=S * dim, want_bf16=False, want_f32=True)
    ref = F.layer_norm(x.view(n, S, dim)[:, 0], (dim,), g, b, 1e-6)
    assert (of - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("n_seq,S,heads,hd", [(4, 257, 6, 64), (2, 257, 6, 32), (3, 64, 6, 64), (2, 100, 6, 32),
                                              (1, 17, 6, 64), (2, 128, 6, 64), (1, 257, 6, 64), (256, 257, 6, 64),
                                              (37, 257, 6, 64), (3, 257, 2, 64)])
def test_attention(n_seq, S, heads, hd):
    L = _lib()
    D = heads * hd
    qkv = _rand((n_seq * S, 3 * D), 30).cuda().bfloat16()
    scale = hd ** -0.5
    out = L.attention(qkv, n_seq, S, heads, hd, scale)
    q, k, v = qkv.float().view(n_seq, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = ((q @ k.transpose(-2, -1)) * scale).softmax(-1)
    ref = (att @ v).transpose(1, 2).reshape(n_seq * S, D)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err


@pytest.mark.parametrize("n_seq", [5, 160])
def test_attention_tc_dominant_keys_and_large_scores(n_seq):
    """The tcgen05 kernel merges two key halves with separate (max, sum) and handles key / query 256 outside the 128-wide
    tiles: put the dominant key of a row in either half, at key 256, or nowhere (flat rows), with scores up to ~+-60."""
    L = _lib()
    S, heads, hd = 257, 6, 64
    D = heads * hd
    g = torch.Generator().manual_seed(31)
    qkv = torch.randn((n_seq, S, 3, heads, hd), generator=g)
    qkv[:, :, 0] *= 2.0                                            # larger logits
    for s in range(n_seq):                                         # sequence s: key (s * 37) % 257 dominates every row of head s % 6
        kd = (s * 37) % S
        qkv[s, kd, 1, s % heads] = 0.0
        qkv[s, kd, 1, s % heads, : hd // 2] = 3.0
        qkv[s, :, 0, s % heads, : hd // 2] += 2.5
    qkv[0, :, 1, 1] = 0.0                                          # head 1 of sequence 0: all scores equal (flat softmax)
    qkv = qkv.reshape(n_seq * S, 3 * D).cuda().bfloat16()
    scale = hd ** -0.5
    out = L.attention(qkv, n_seq, S, heads, hd, scale)
    q, k, v = qkv.float().view(n_seq, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = ((q @ k.transpose(-2, -1)) * scale).softmax(-1)
    ref = (att @ v).transpose(1, 2).reshape(n_seq * S, D)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-2, err
    # row 256 (the tail tile) and rows 0 / 128 (first rows of the two full tiles) separately
    o3, r3 = out.float().view(n_seq, S, D), ref.view(n_seq, S, D)
    for row in (0, 127, 128, 255, 256):
        assert (o3[:, row] - r3[:, row]).abs().max().item() < 3e-2, row


def test_attention_is_bit_identical_run_to_run():
    """300 launches over one seeded qkv (256 sequences = one region) must return the same bits: a hand-off race inside
    the kernel (found in round 2: the last word of the row-statistics record arriving stale about once per 10^5 items)
    shows up as a handful of rows differing along v_256."""
    L = _lib()
    n_seq = 256
    qkv = _rand((n_seq * 257, 1152), 33).cuda().bfloat16()
    ref = L.attention(qkv, n_seq, 257, 6, 64, 0.125).clone()
    bad = sum(0 if torch.equal(L.attention(qkv, n_seq, 257, 6, 64, 0.125), ref) else 1 for _ in range(300))
    assert bad == 0, f"{bad} of 300 runs differ"


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_im2col(dtype):
    L = _lib()
    g = torch.Generator().manual_seed(40)
    H, W = 512, 768
    if dtype == torch.uint8:
        img = torch.randint(0, 256, (3, H, W), dtype=torch.uint8, generator=g).cuda()
    else:
        img = torch.randn((3, H, W), generator=g).cuda()
    grid_cols = W // 256
    n = (H // 256) * grid_cols
    a = L.im2col_patches(img, 0, n)
    x = img.float().unsqueeze(0)
    patches = x.unfold(2, 256, 256).unfold(3, 256, 256)            # [1,3,p1,p2,256,256]
    patches = patches.permute(0, 2, 3, 1, 4, 5).reshape(n, 3, 256, 256)
    cols = F.unfold(patches, kernel_size=16, stride=16)            # [n, 768, 256], K order (c,i,j)
    ref = cols.transpose(1, 2).reshape(n * 256, 768)
    if dtype == torch.uint8:
        assert torch.equal(a.float(), ref)
    else:
        assert torch.equal(a, ref.bfloat16())
    # a sub-range of patches
    a2 = L.im2col_patches(img, 2, 3)
    assert torch.equal(a2, a[2 * 256:5 * 256])
    # batch mode: separate [B,3,256,256] patches
    a3 = L.im2col_patches(patches.to(dtype).contiguous(), 1, 4)
    assert torch.equal(a3, a[256:5 * 256])
