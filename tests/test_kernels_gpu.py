"""Kernel-level parity on the GPU: every C-ABI primitive against a plain torch fp32/fp64 restatement
of the same operator on the same seeded inputs (tolerances written next to each check)."""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize(
    "M,N,K",
    [(128, 128, 64), (300, 1024, 768), (257, 72, 200), (1000, 512, 1024), (12608, 3072, 1024), (4097, 1536, 512)],
)
def test_gemm_store16(cuda_device, M, N, K, dtype):
    from stamp_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(cuda_device, dtype)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device, dtype)
    bias = torch.randn(N, generator=g).to(cuda_device)
    out = torch.full((M, N), float("nan"), device=cuda_device, dtype=dtype)
    ops.gemm_tn(a, w, out=out, bias=bias)
    ref = a.float() @ w.float().T + bias
    tol = 2e-3 if dtype == torch.float16 else 1e-2  # output rounding of the 16-bit store
    assert torch.isfinite(out).all()
    assert _rel(out.float(), ref) < tol
    assert (out.float() - ref).abs().max().item() < 0.05 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(300, 512, 512), (4097, 512, 512), (130, 72, 100)])
def test_gemm_tf32(cuda_device, M, N, K):
    """kind::tf32 path: fp32 operands pre-rounded to TF32, huge dynamic range (ALiBi outputs)."""
    from stamp_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 1e5).to(cuda_device)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device)
    # round-to-nearest to 10 explicit mantissa bits
    rnd = lambda t: ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    a, w = rnd(a), rnd(w)
    bias = torch.randn(N, generator=g).to(cuda_device)
    x = torch.zeros(M, N, device=cuda_device)
    ops.gemm_tn(a, w, out=x, bias=bias, store=ops.ST_RESID32)
    ref = (a.double() @ w.double().T + bias.double())
    assert _rel(x, ref) < 1e-5


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("M,N,K", [(300, 1024, 768), (257, 72 * 4, 200), (12608, 3072, 1024), (5000, 1024, 4096), (130, 256, 64)])
def test_gemm_tile_shapes(cuda_device, M, N, K, mode):
    """Same products through single-CTA tiles (mode 1) and CTA-pair cta_group::2 tiles (mode 2),
    with the fp32 residual epilogue (LayerScale gamma) and the fp16 store."""
    from stamp_b200 import _lib, ops

    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda_device, torch.float16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device, torch.float16)
    bias = torch.randn(N, generator=g).to(cuda_device)
    gamma = torch.rand(N, generator=g).to(cuda_device) + 0.5
    ref = a.float() @ w.float().T + bias
    _lib.load().stamp_b200_gemm_force_mode(mode)
    try:
        out = torch.full((M, N), float("nan"), device=cuda_device, dtype=torch.float16)
        ops.gemm_tn(a, w, out=out, bias=bias)
        x = torch.ones(M, N, device=cuda_device)
        ops.gemm_tn(a, w, out=x, bias=bias, gamma=gamma, store=ops.ST_RESID32)
    finally:
        _lib.load().stamp_b200_gemm_force_mode(0)
    assert torch.isfinite(out).all()
    assert _rel(out.float(), ref) < 2e-3
    assert _rel(x, 1.0 + gamma * ref) < 1e-5


def test_gemm_epilogues(cuda_device):
    from stamp_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 777, 640, 320
    a = torch.randn(M, K, generator=g).to(cuda_device, torch.float16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device, torch.float16)
    bias = torch.randn(N, generator=g).to(cuda_device)
    gamma = torch.rand(N, generator=g).to(cuda_device) + 0.5
    lin = a.float() @ w.float().T + bias

    out = torch.empty(M, N, device=cuda_device, dtype=torch.float16)
    ops.gemm_tn(a, w, out=out, bias=bias, act=ops.ACT_GELU)
    assert _rel(out.float(), torch.nn.functional.gelu(lin)) < 2e-3

    ops.gemm_tn(a, w, out=out, bias=bias, act=ops.ACT_RELU)
    assert _rel(out.float(), torch.relu(lin)) < 2e-3

    x = torch.randn(M, N, generator=g).to(cuda_device)
    x0 = x.clone()
    ops.gemm_tn(a, w, out=x, bias=bias, gamma=gamma, store=ops.ST_RESID32)
    assert _rel(x, x0 + gamma * lin) < 1e-5

    o32 = torch.empty(M, N, device=cuda_device)
    ops.gemm_tn(a, w, out=o32, bias=bias, store=ops.ST_32)
    assert _rel(o32, lin) < 1e-5

    # SwiGLU / gated: adjacent column pairs (x1, x2)
    o16 = torch.empty(M, N // 2, device=cuda_device, dtype=torch.float16)
    ops.gemm_tn(a, w, out=o16, bias=bias, store=ops.ST_SWIGLU16)
    ref = torch.nn.functional.silu(lin[:, 0::2]) * lin[:, 1::2]
    assert _rel(o16.float(), ref) < 2e-3
    ops.gemm_tn(a, w, out=o16, bias=bias, store=ops.ST_GATED16)
    ref = torch.tanh(lin[:, 0::2]) * torch.sigmoid(lin[:, 1::2])
    assert _rel(o16.float(), ref) < 2e-3


@pytest.mark.parametrize("M,N,K", [(20000, 512, 256), (37824, 1024, 320), (5000, 6832, 128), (333, 200, 64), (129, 2056, 96)])
def test_gemm_tma_epilogues_match_register_epilogue(cuda_device, M, N, K):
    """The TMA-store (16-bit outputs, GELU, SwiGLU) and TMA reduce-add (fp32 residual) epilogues against the
    generic register epilogue (force-mode bit 2) of the same kernel: identical 16-bit outputs, fp32 residual equal
    to round-off, hardware-clipped edges (M, N not multiples of the 128 x 256 tiles), untouched memory outside."""
    from stamp_b200 import _lib, ops

    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda_device, torch.float16)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device, torch.float16)
    bias = torch.randn(N, generator=g).to(cuda_device)
    gamma = torch.rand(N, generator=g).to(cuda_device) + 0.5
    x0 = torch.randn(M, N, generator=g).to(cuda_device)
    lib = _lib.load()

    def run(mode):
        lib.stamp_b200_gemm_force_mode(mode)
        try:
            pad = torch.full((M + 2, N + 8), 7.0, device=cuda_device, dtype=torch.float16)   # guard band around the output
            o = pad[1:M + 1, :N]
            ops.gemm_tn(a, w, out=o, bias=bias)
            og = torch.empty(M, N, device=cuda_device, dtype=torch.float16)
            ops.gemm_tn(a, w, out=og, bias=bias, act=ops.ACT_GELU)
            os_ = torch.empty(M, N // 2, device=cuda_device, dtype=torch.float16)
            ops.gemm_tn(a, w, out=os_, bias=bias, store=ops.ST_SWIGLU16)
            xp = torch.full((M + 2, N + 4), 3.0, device=cuda_device)
            x = xp[1:M + 1, :N]
            x.copy_(x0)
            ops.gemm_tn(a, w, out=x, bias=bias, gamma=gamma, store=ops.ST_RESID32)
        finally:
            lib.stamp_b200_gemm_force_mode(0)
        return pad, og, os_, xp

    pad_t, og_t, os_t, xp_t = run(0)
    pad_r, og_r, os_r, xp_r = run(4)
    assert torch.equal(pad_t, pad_r)                 # includes the guard band: nothing written outside [M, N]
    assert torch.equal(og_t, og_r) and torch.equal(os_t, os_r)
    assert (pad_t[0] == 7).all() and (pad_t[-1] == 7).all() and (pad_t[:, N:] == 7).all()
    assert (xp_t[0] == 3).all() and (xp_t[-1] == 3).all() and (xp_t[:, N:] == 3).all()
    lin = a.float() @ w.float().T + bias
    assert _rel(xp_t[1:M + 1, :N], x0 + gamma * lin) < 1e-5
    assert _rel(xp_t[1:M + 1, :N], xp_r[1:M + 1, :N]) < 1e-6
    assert _rel(og_t.float(), torch.nn.functional.gelu(lin)) < 2e-3
    assert _rel(os_t.float(), torch.nn.functional.silu(lin[:, 0::2]) * lin[:, 1::2]) < 2e-3


def test_gemm_row_remap_and_table(cuda_device):
    """Patch-embed style epilogue: rows of each 196-row group land at 197*g + 1 + r, plus a table."""
    from stamp_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(6)
    B, P, T, D, K = 5, 196, 197, 256, 768
    a = torch.randn(B * P, K, generator=g).to(cuda_device, torch.float16)
    w = (torch.randn(D, K, generator=g) / math.sqrt(K)).to(cuda_device, torch.float16)
    bias = torch.randn(D, generator=g).to(cuda_device)
    pos = torch.randn(T, D, generator=g).to(cuda_device)
    x = torch.zeros(B * T, D, device=cuda_device)
    ops.gemm_tn(a, w, out=x, bias=bias, store=ops.ST_32, table=pos[1:], gin=P, gout=T, goff=1)
    ref = (a.float() @ w.float().T + bias).view(B, P, D) + pos[1:]
    x = x.view(B, T, D)
    assert _rel(x[:, 1:], ref) < 1e-5
    assert (x[:, 0] == 0).all()


@pytest.mark.parametrize("cols,dtype", [(1024, torch.float16), (512, torch.float32), (1280, torch.float16)])
def test_layernorm(cuda_device, cols, dtype):
    from stamp_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(cols)
    x = (torch.randn(1001, cols, generator=g) * 3 + 1.5).to(cuda_device)
    w = torch.randn(cols, generator=g).to(cuda_device)
    b = torch.randn(cols, generator=g).to(cuda_device)
    out = ops.layernorm(x, w, b, 1e-6, dtype)
    ref = torch.nn.functional.layer_norm(x, (cols,), w, b, 1e-6)
    assert _rel(out.float(), ref) < (1e-6 if dtype == torch.float32 else 5e-4)


@pytest.mark.parametrize("patch", [16, 14])
def test_tiles_to_patches(cuda_device, patch):
    from stamp_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(patch)
    tiles = torch.randint(0, 256, (3, 224, 224, 3), generator=g, dtype=torch.uint8).to(cuda_device)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    out = ops.tiles_to_patches(tiles, patch, mean, std)
    x = tiles.permute(0, 3, 1, 2).float() / 255.0
    x = (x - torch.tensor(mean, device=cuda_device).view(1, 3, 1, 1)) / torch.tensor(std, device=cuda_device).view(1, 3, 1, 1)
    ref = torch.nn.functional.unfold(x, kernel_size=patch, stride=patch).transpose(1, 2).reshape(-1, 3 * patch * patch)
    k = 3 * patch * patch
    assert out.shape[1] % 8 == 0 and out.shape[1] >= k
    assert (out[:, :k].float() - ref).abs().max().item() < 2e-3  # fp16 rounding of values in [-2.2, 2.7]
    assert (out[:, k:] == 0).all()


def _attn_ref(qkv, H, coords=None, slope=None, mask=None, mask_mode=1):
    B, S, D3 = qkv.shape
    D = D3 // 3
    hd = D // H
    q, k, v = qkv.double().view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)  # [B,H,S,hd]
    logits = q @ k.transpose(-1, -2) / math.sqrt(hd)
    attn_mask = None
    if mask is not None:
        m = mask.bool()
        attn_mask = m[:, :, None] & m[:, None, :]
        attn_mask[:, 1:, 0] = True
        attn_mask = attn_mask[:, None]
        if mask_mode == 2:
            # reference quirk: (bag b, head h) uses the mask of bag (b*H + h) % B (see attention.cu)
            idx = (torch.arange(B, device=qkv.device)[:, None] * H + torch.arange(H, device=qkv.device)[None, :]) % B
            logits = logits.masked_fill(attn_mask[:, 0][idx], float("-inf"))
    w = torch.softmax(logits, -1)
    if coords is not None:
        c = coords.double()
        dist = (c[:, :, None, :] - c[:, None, :, :]).norm(dim=-1)  # exact distances
        sd = dist[:, None] * slope.double().view(1, H, 1, 1)
        if mask is not None and mask_mode == 1:
            am = torch.zeros(B, 1, S, S, dtype=torch.bool, device=qkv.device)
            am[:, :, 0, :] = True
            am[:, :, :, 0] = True
            sd = sd.masked_fill(am, 0.0)
        w = w - sd
    if mask is not None and mask_mode == 1:
        w = w.masked_fill(attn_mask, 0.0)
    o = w @ v
    return o.permute(0, 2, 1, 3).reshape(B, S, D)


@pytest.mark.parametrize("B,S,H,hd", [(3, 197, 4, 64), (2, 261, 2, 80), (2, 64, 2, 64), (1, 1, 2, 64), (1, 1000, 2, 32)])
def test_attention_plain(cuda_device, B, S, H, hd):
    from stamp_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(S + hd)
    qkv = torch.randn(B, S, 3 * H * hd, generator=g).to(cuda_device, torch.float16)
    out = ops.attention(qkv, H)
    ref = _attn_ref(qkv, H)
    assert torch.isfinite(out).all()
    assert _rel(out.float(), ref) < 2e-3  # fp16 P and fp16 output rounding


@pytest.mark.parametrize("S", [65, 513, 1500])
@pytest.mark.parametrize("masked", [False, True])
def test_attention_alibi(cuda_device, S, masked):
    from stamp_b200 import ops

    B, H, hd = 2, 8, 64
    g = torch.Generator(device="cpu").manual_seed(S)
    qkv = torch.randn(B, S, 3 * H * hd, generator=g).to(cuda_device, torch.float16)
    coords = (torch.randint(0, 100, (B, S, 2), generator=g).float() * 256.0).to(cuda_device)
    coords[:, 0] = 0
    slope = torch.rand(H, generator=g).to(cuda_device)
    mask = None
    if masked:
        mask = (torch.rand(B, S, generator=g) < 0.3).to(cuda_device)
        mask[:, 0] = False
    out = ops.attention(qkv, H, coords=coords, slope=slope, mask=mask, mask_mode=1)
    ref = _attn_ref(qkv, H, coords, slope, mask, 1)
    assert out.dtype == torch.float32 and torch.isfinite(out).all()
    # the ALiBi term dominates (|out| ~ 1e5): fp16 Dist operand (2^-11 rel) + fp16 V + fp16 output
    assert _rel(out.float(), ref) < 1e-3


def test_attention_mask_mode2(cuda_device):
    from stamp_b200 import ops

    B, S, H, hd = 2, 300, 4, 64
    g = torch.Generator(device="cpu").manual_seed(11)
    qkv = torch.randn(B, S, 3 * H * hd, generator=g).to(cuda_device, torch.float16)
    mask = (torch.rand(B, S, generator=g) < 0.3).to(cuda_device)
    mask[:, 0] = False
    out = ops.attention(qkv, H, mask=mask, mask_mode=2)
    ref = _attn_ref(qkv, H, None, None, mask, 2)
    assert _rel(out.float(), ref) < 2e-3


@pytest.mark.parametrize("B,S,H", [(2, 16, 2), (3, 100, 3), (2, 128, 2), (2, 129, 2), (5, 197, 16), (2, 256, 4), (1, 7, 1),
                                   (40, 197, 16), (3, 64, 2), (2, 65, 1), (300, 50, 4)])
def test_attention_tcgen05_vs_general(cuda_device, B, S, H):
    """Short unmasked head_dim-64 sequences take the tcgen05 kernel; same result as the general one."""
    from stamp_b200 import _lib, ops

    g = torch.Generator(device="cpu").manual_seed(S * 3 + H)
    qkv = torch.randn(B, S, 3 * H * 64, generator=g).to(cuda_device, torch.float16)
    ref = _attn_ref(qkv, H)
    out_tc = ops.attention(qkv, H)          # default: persistent streaming tcgen05 kernel
    try:
        _lib.load().stamp_b200_attention_tc_enable(9)
        out_eager = ops.attention(qkv, H)   # same, accumulator rescaled whenever a row maximum grows
        _lib.load().stamp_b200_attention_tc_enable(65)
        out_one = ops.attention(qkv, H)     # one-shot tcgen05 kernel (first round), all keys in one pass
        _lib.load().stamp_b200_attention_tc_enable(0)
        out_gen = ops.attention(qkv, H)     # general (legacy tensor path) kernel
    finally:
        _lib.load().stamp_b200_attention_tc_enable(1)
    for o in (out_tc, out_eager, out_one, out_gen):
        assert torch.isfinite(o).all()
        assert _rel(o.float(), ref) < 2e-3


@pytest.mark.parametrize("B,S,H", [(2, 261, 16), (3, 265, 2), (2, 64, 3), (40, 261, 16), (2, 130, 1), (5, 16, 2)])
def test_attention_head_dim_80_tcgen05_vs_general(cuda_device, B, S, H):
    """ViT-H/14 heads (Virchow2: 261 tokens, 16 heads of 80): the streaming tcgen05 kernel handles the head as two
    64-column sub-tiles (5 contraction steps, one N = 128 product for P V); same result as the general kernel."""
    from stamp_b200 import _lib, ops

    g = torch.Generator(device="cpu").manual_seed(S * 5 + H)
    qkv = torch.randn(B, S, 3 * H * 80, generator=g).to(cuda_device, torch.float16)
    ref = _attn_ref(qkv, H)
    out_tc = ops.attention(qkv, H)
    try:
        _lib.load().stamp_b200_attention_tc_enable(9)
        out_eager = ops.attention(qkv, H)
        _lib.load().stamp_b200_attention_tc_enable(0)
        out_gen = ops.attention(qkv, H)
    finally:
        _lib.load().stamp_b200_attention_tc_enable(1)
    for o in (out_tc, out_eager, out_gen):
        assert torch.isfinite(o).all()
        assert _rel(o.float(), ref) < 2e-3
    assert _rel(out_tc.float(), out_gen.float()) < 1e-3


@pytest.mark.parametrize("S", [257, 1000, 4097])
@pytest.mark.parametrize("alibi", [False, True])
def test_attention_long_bag_tcgen05_vs_general(cuda_device, S, alibi):
    """Unmasked long bags take the two-pass tcgen05 kernel (plain and ALiBi); both kernels vs fp64."""
    from stamp_b200 import _lib, ops

    B, H = 2, 8
    g = torch.Generator(device="cpu").manual_seed(S + alibi)
    qkv = torch.randn(B, S, 3 * H * 64, generator=g).to(cuda_device, torch.float16)
    coords = slope = None
    if alibi:
        coords = (torch.randint(0, 100, (B, S, 2), generator=g).float() * 256.0).to(cuda_device)
        coords[:, 0] = 0
        slope = torch.rand(H, generator=g).to(cuda_device)
    ref = _attn_ref(qkv, H, coords, slope)
    out_tc = ops.attention(qkv, H, coords=coords, slope=slope)
    _lib.load().stamp_b200_attention_tc_enable(0)
    try:
        out_gen = ops.attention(qkv, H, coords=coords, slope=slope)
    finally:
        _lib.load().stamp_b200_attention_tc_enable(1)
    assert torch.isfinite(out_tc).all()
    tol = 1e-3 if alibi else 2e-3
    assert _rel(out_tc.float(), ref) < tol, _rel(out_tc.float(), ref)
    assert _rel(out_gen.float(), ref) < tol


def test_launch_counter(cuda_device):
    from stamp_b200 import _lib, ops

    _lib.reset_launch_count()
    x = torch.randn(64, 512, device=cuda_device)
    ops.layernorm(x, torch.ones(512, device=cuda_device), torch.zeros(512, device=cuda_device), 1e-5, torch.float16)
    assert _lib.launch_count() == 1
