"""Parity of every CUDA entry point (called through the C ABI) against a plain PyTorch fp32 CPU restatement of the
same reference operation (oracle/imagen_oracle.py for module-level math).

Tolerances: fp16 tensor-core kernels rel-L2 <= 1e-2 (north_star), typically <= 1e-3 from operand / output rounding;
fp32 elementwise / sampler kernels bit-exact or <= 1e-6 where transcendental functions differ.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bf(x):
    return x.to(torch.float16)


def nhwc(x):  # NCHW fp32 -> NHWC fp16 on device
    return bf(x.permute(0, 2, 3, 1).contiguous()).to(DEV)


def from_nhwc(y):
    return y.float().cpu().permute(0, 3, 1, 2)


def pack_w(w):  # [Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin] fp16
    return bf(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()).to(DEV)


def rb(x):  # round-trip through fp16 so the reference sees the same operand values
    return x.to(torch.float16).float()


# ------------------------------------------------------------------------------------------------ conv_gemm
CONV_CASES = [
    # B, H, W, Cin, Cout, k
    (1, 16, 16, 64, 128, 3),
    (2, 32, 32, 128, 128, 3),
    (1, 64, 64, 256, 256, 3),
    (3, 8, 8, 128, 64, 3),       # tile spans two images, batch tail
    (1, 24, 40, 64, 192, 3),     # non power-of-two spatial size -> partial tiles, Cout not a multiple of 128
    (2, 16, 16, 192, 128, 1),
    (1, 4, 4, 64, 128, 3),       # tiny spatial
    (1, 32, 32, 1024, 128, 3),   # long K loop (144 k-blocks): ring wrap-around many times
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", CONV_CASES)
def test_conv_gemm_matches_conv2d(cuda_lib, B, H, W, Cin, Cout, k):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout + k)
    x = rb(torch.randn(B, Cin, H, W, generator=g))
    w = rb(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k))
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w, b, padding=k // 2)
    out = ops.conv_gemm(nhwc(x), pack_w(w), b.to(DEV), ksize=k)
    torch.cuda.synchronize()
    err = rel_l2(from_nhwc(out), ref)
    print(f"conv {B}x{H}x{W} {Cin}->{Cout} k{k}: rel_l2={err:.3e}")
    assert err < 5e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(1, 64, 64, 256, 256, 3), (3, 24, 40, 128, 640, 1), (2, 16, 16, 64, 128, 3), (1, 8, 8, 1024, 1024, 3),
                                               (5, 8, 8, 128, 384, 3), (1, 128, 128, 128, 128, 3), (3, 40, 20, 192, 256, 3),
                                               (1, 16, 8, 64, 128, 3), (2, 33, 17, 128, 384, 3)])
def test_conv_pair_kernel_bit_identical_to_single_cta_kernel(cuda_lib, B, H, W, Cin, Cout, k):
    """cta_group::2 persistent kernel vs the single-CTA kernel: same k order per output -> identical bits; odd tile counts,
    N tails (640 = 2.5 x 256) and multi-tile-per-CTA persistence are all exercised."""
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(Cin + Cout + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g)
    add = torch.randn(B, Cout, H, W, generator=g)
    dx, dw, db, da = nhwc(x), pack_w(w), b.to(DEV), nhwc(add)
    try:
        ops.set_conv_impl(1)
        ref = ops.conv_gemm(dx, dw, db, ksize=k, act=ops.ACT_SILU, addend=da)
        ops.set_conv_impl(4)   # tap-loop CTA-pair kernel: same k order as the single-CTA kernel
        out = ops.conv_gemm(dx, dw, db, ksize=k, act=ops.ACT_SILU, addend=da)
        ops.set_conv_impl(2)   # halo-reuse variant for 3x3 on >= 16 x 8 images (chunk-major k order: fp32 sums differ in the last bit)
        out_halo = ops.conv_gemm(dx, dw, db, ksize=k, act=ops.ACT_SILU, addend=da)
        torch.cuda.synchronize()
    finally:
        ops.set_conv_impl(0)
    assert rel_l2(out_halo, ref) < 2e-4 and float((out_halo.float() - ref.float()).abs().max()) < 2e-2
    assert torch.equal(out, ref)
    err = rel_l2(from_nhwc(out), F.silu(F.conv2d(rb(x), rb(w), b, padding=k // 2)) + rb(add))
    assert err < 5e-3


def test_conv_gemm_two_sources_act_addend_gate(cuda_lib):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(7)
    B, H, W, Ca, Cb, Cout = 2, 16, 16, 128, 64, 128
    xa, xb = rb(torch.randn(B, Ca, H, W, generator=g)), rb(torch.randn(B, Cb, H, W, generator=g))
    w = rb(torch.randn(Cout, Ca + Cb, 3, 3, generator=g) / math.sqrt((Ca + Cb) * 9))
    b = torch.randn(Cout, generator=g)
    add = rb(torch.randn(B, Cout, H, W, generator=g))
    gate = torch.rand(B, Cout, generator=g)
    ref = F.silu(F.conv2d(torch.cat((xa, xb), 1), w, b, padding=1)) + gate[:, :, None, None] * add
    out = ops.conv_gemm(nhwc(xa), pack_w(w), b.to(DEV), xb=nhwc(xb), ksize=3, act=ops.ACT_SILU, addend=nhwc(add),
                        addend_scale=gate.to(DEV))
    err = rel_l2(from_nhwc(out), ref)
    print("two-source conv rel_l2", err)
    assert err < 5e-3
    # fp32 output + fp32 addend, GELU
    ref2 = F.gelu(F.conv2d(torch.cat((xa, xb), 1), w, b, padding=1)) + add
    out2 = ops.conv_gemm(nhwc(xa), pack_w(w), b.to(DEV), xb=nhwc(xb), ksize=3, act=ops.ACT_GELU, out_f32=True,
                         addend=add.permute(0, 2, 3, 1).contiguous().to(DEV))
    assert out2.dtype == torch.float32
    err2 = rel_l2(out2.cpu().permute(0, 3, 1, 2), ref2)
    print("fp32-out conv rel_l2", err2)
    assert err2 < 5e-3


def test_conv_gemm_downsample_mode(cuda_lib):
    """Downsample = 'b c (h 2) (w 2) -> b (c 2 2) h w' + Conv2d(4c, cout, 1)."""
    from kidney_diffusion_b200 import ops
    from oracle.imagen_oracle import Downsample

    torch.manual_seed(3)
    C, Cout = 64, 128
    mod = Downsample(C, Cout)
    x = rb(torch.randn(2, C, 32, 32))
    with torch.no_grad():
        mod[1].weight.copy_(rb(mod[1].weight))
        ref = mod(x)
    w = mod[1].weight.detach().view(Cout, C, 2, 2)  # input channel index = c*4 + dy*2 + dx
    wp = bf(w.permute(0, 2, 3, 1).reshape(Cout, 4 * C).contiguous()).to(DEV)  # k = (dy*2+dx)*C + c
    out = ops.conv_gemm(nhwc(x), wp, mod[1].bias.detach().to(DEV), mode=1)
    assert out.shape == (2, 16, 16, Cout)
    err = rel_l2(from_nhwc(out), ref)
    print("downsample rel_l2", err)
    assert err < 5e-3


def test_conv_gemm_pixel_shuffle_mode(cuda_lib):
    """PixelShuffleUpsample = Conv2d(c, 4*cout, 1) -> SiLU -> PixelShuffle(2)."""
    from kidney_diffusion_b200 import ops
    from oracle.imagen_oracle import PixelShuffleUpsample

    torch.manual_seed(4)
    C, Cout = 128, 64
    mod = PixelShuffleUpsample(C, Cout)
    conv = mod.net[0]
    with torch.no_grad():
        conv.weight.copy_(rb(torch.randn_like(conv.weight) / math.sqrt(C)))
        conv.bias.copy_(torch.randn_like(conv.bias))
    x = rb(torch.randn(2, C, 16, 16))
    with torch.no_grad():
        ref = mod(x)
    w = conv.weight.detach().view(Cout, 4, C)  # out channel = c*4 + (dy*2+dx)
    wp = bf(w.permute(1, 0, 2).reshape(4 * Cout, C).contiguous()).to(DEV)  # rows ordered (dy*2+dx, c)
    bp = conv.bias.detach().view(Cout, 4).t().reshape(-1).contiguous().to(DEV)
    out = ops.conv_gemm(nhwc(x), wp, bp, ksize=1, act=ops.ACT_SILU, out_mode=1)
    assert out.shape == (2, 32, 32, Cout)
    err = rel_l2(from_nhwc(out), ref)
    print("pixel-shuffle rel_l2", err)
    assert err < 5e-3


def test_gemm_rows(cuda_lib):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(5)
    M, K, N = 300, 704, 128
    x, w, b = rb(torch.randn(M, K, generator=g)), rb(torch.randn(N, K, generator=g) / math.sqrt(K)), torch.randn(N, generator=g)
    out = ops.gemm_rows(bf(x).to(DEV), bf(w).to(DEV), b.to(DEV), out_f32=True)
    err = rel_l2(out, x @ w.t() + b)
    print("gemm_rows rel_l2", err)
    assert err < 2e-3


# ------------------------------------------------------------------------------------------------ conditioning towers
def test_linear_small_and_sinu(cuda_lib):
    from kidney_diffusion_b200 import ops
    from oracle.imagen_oracle import LearnedSinusoidalPosEmb

    g = torch.Generator().manual_seed(6)
    for M, K, N in [(4, 17, 512), (3, 1024, 2048), (20, 512, 1024), (1, 64, 3)]:
        x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / math.sqrt(K), torch.randn(N, generator=g)
        ref = F.silu(F.linear(F.silu(x), w, b))
        out = ops.linear_small(x.to(DEV), w.to(DEV), b.to(DEV), pre_act=ops.ACT_SILU, post_act=ops.ACT_SILU)
        err = rel_l2(out, ref)
        assert err < 1e-5, (M, K, N, err)
    torch.manual_seed(1)
    emb = LearnedSinusoidalPosEmb(16)
    t = torch.tensor([8.7692, 2.18, -0.0249, -33.891])
    ref = emb(t).detach()
    out = ops.sinu_emb(t.to(DEV), emb.weights.detach().to(DEV))
    assert float((out.cpu() - ref).abs().max()) < 2e-5


# ------------------------------------------------------------------------------------------------ GroupNorm
@pytest.mark.parametrize("B,H,W,Ca,Cb", [(2, 16, 16, 128, 0), (1, 32, 32, 256, 128), (2, 8, 8, 1024, 512), (1, 64, 64, 64, 64)])
def test_groupnorm_scale_shift_silu(cuda_lib, B, H, W, Ca, Cb):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(Ca + Cb + H)
    C, G = Ca + Cb, 8
    skip_scale = 2 ** -0.5
    xa = rb(torch.randn(B, Ca, H, W, generator=g) * 2 + 0.5)
    xb = rb(torch.randn(B, Cb, H, W, generator=g)) if Cb else None
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ss = torch.randn(B, 2 * C, generator=g) * 0.3
    x = xa if xb is None else torch.cat((xa, xb * skip_scale), 1)
    scale, shift = ss[:, :C, None, None], ss[:, C:, None, None]
    ref = F.silu(F.group_norm(x, G, gamma, beta, eps=1e-5) * (scale + 1) + shift)
    gs = C // G
    da = nhwc(xa)
    pa = ops.gn_stats(da, 0, gs, G)
    pb, db = None, None
    if xb is not None:
        db = nhwc(xb)
        pb = ops.gn_stats(db, Ca, gs, G)
    mr = ops.gn_finalize(pa, 1.0, pb, skip_scale, count=gs * H * W)
    kw = dict(group_size=gs, num_groups=G, scale_shift=ss.to(DEV), ctot=C)
    ya = ops.gn_apply(da, mr, gamma.to(DEV), beta.to(DEV), c_offset=0, **kw)
    out = from_nhwc(ya)
    if xb is not None:
        yb = ops.gn_apply(db, mr, gamma.to(DEV), beta.to(DEV), c_offset=Ca, src_scale=skip_scale, **kw)
        out = torch.cat((out, from_nhwc(yb)), 1)
    err = rel_l2(out, ref)
    print("groupnorm rel_l2", err)
    assert err < 4e-3  # fp16 output rounding


# ------------------------------------------------------------------------------------------------ GlobalContext
@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 128), (1, 64, 64, 256), (2, 8, 8, 1024), (1, 128, 128, 128)])
def test_global_context_gate_residual(cuda_lib, B, H, W, C):
    from kidney_diffusion_b200 import ops
    from oracle.imagen_oracle import GlobalContext

    torch.manual_seed(C + H)
    gca = GlobalContext(dim_in=C, dim_out=C)
    with torch.no_grad():
        gca.to_k.weight.mul_(4.0)  # sharper softmax
    x = rb(torch.randn(B, C, H, W))
    res = rb(torch.randn(B, C, H, W))
    with torch.no_grad():
        gate_ref = gca(x)
        ref = x * gate_ref + res
    dx = nhwc(x)
    logits = ops.rowdot(dx, gca.to_k.weight.detach().view(C).to(DEV), gca.to_k.bias.detach().to(DEV))
    pooled = ops.gca_pool(dx, logits)
    hid = ops.linear_small(pooled, gca.net[0].weight.detach().view(-1, C).to(DEV), gca.net[0].bias.detach().to(DEV), post_act=ops.ACT_SILU)
    gate = ops.linear_small(hid, gca.net[2].weight.detach().view(C, -1).to(DEV), gca.net[2].bias.detach().to(DEV), post_act=ops.ACT_SIGMOID)
    e_gate = rel_l2(gate, gate_ref.view(B, C))
    # the same tail as ONE cluster launch (kd_gca_gate: finalize + MLP + sigmoid), what the executor runs
    w0, b0 = gca.net[0].weight.detach().view(-1, C).to(DEV).contiguous(), gca.net[0].bias.detach().to(DEV)
    w1, b1 = gca.net[2].weight.detach().view(C, -1).to(DEV).contiguous(), gca.net[2].bias.detach().to(DEV)
    fused = ops.gca_gate(dx, logits, w0, b0, w1, b1)
    assert rel_l2(fused, gate_ref.view(B, C)) < 1e-4 and rel_l2(fused, gate) < 1e-5
    if B > 1:  # batch invariance: one cluster per image
        assert torch.equal(ops.gca_gate(dx[1:2].contiguous(), logits[1:2].contiguous(), w0, b0, w1, b1)[0], fused[1])
    out = ops.gate_residual(dx, gate, nhwc(res))
    err = rel_l2(from_nhwc(out), ref)
    print("gca gate rel_l2", e_gate, "out", err)
    assert e_gate < 1e-4 and err < 4e-3


def test_layernorm(cuda_lib):
    from kidney_diffusion_b200 import ops
    from oracle.imagen_oracle import LayerNorm

    torch.manual_seed(2)
    for M, C in [(100, 128), (4096, 1024), (7, 2048)]:
        ln = LayerNorm(C)
        with torch.no_grad():
            ln.g.copy_(torch.randn(C))
        x, r = rb(torch.randn(M, C) * 3 + 1), rb(torch.randn(M, C))
        ref = ln(x).detach() + r
        out = ops.layernorm_h16(bf(x).to(DEV), ln.g.detach().to(DEV), None, bf(r).to(DEV))
        assert rel_l2(out, ref) < 4e-3
    nl = torch.nn.LayerNorm(512)
    with torch.no_grad():
        nl.weight.copy_(torch.randn(512)); nl.bias.copy_(torch.randn(512))
    x = torch.randn(12, 512)
    out = ops.layernorm_f32(x.to(DEV), nl.weight.detach().to(DEV), nl.bias.detach().to(DEV))
    assert rel_l2(out, nl(x).detach()) < 1e-5


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,N,Jc", [(2, 256, 0), (1, 1024, 4), (2, 100, 36), (1, 4096, 0)])
def test_attn_mqa(cuda_lib, B, N, Jc):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(N + Jc)
    heads, d = 8, 64
    qkv = rb(torch.randn(B, N, heads * d + 2 * d, generator=g))
    ctx = torch.randn(B, Jc, 2 * d, generator=g) if Jc else None
    null_kv = torch.randn(2, d, generator=g)
    q = qkv[..., : heads * d].view(B, N, heads, d).transpose(1, 2) * d ** -0.5
    k, v = qkv[..., heads * d: heads * d + d], qkv[..., heads * d + d:]
    k = torch.cat((rb(null_kv[0]).expand(B, 1, d), k), 1)
    v = torch.cat((rb(null_kv[1]).expand(B, 1, d), v), 1)
    if Jc:
        k = torch.cat((rb(ctx[..., :d]), k), 1)
        v = torch.cat((rb(ctx[..., d:]), v), 1)
    attn = torch.einsum("bhid,bjd->bhij", q, k).softmax(-1)
    ref = torch.einsum("bhij,bjd->bhid", attn, v).transpose(1, 2).reshape(B, N, heads * d)
    dq = bf(qkv).to(DEV)
    kv = ops.kv_assemble(dq, heads * d, None if ctx is None else ctx.to(DEV), null_kv.to(DEV))
    assert kv.shape == (B, Jc + 1 + N, 128)
    out = ops.attn_mqa(dq, kv, heads, d ** -0.5)
    err = rel_l2(out, ref)
    print("attn_mqa rel_l2", err)
    assert err < 8e-3


@pytest.mark.parametrize("B,N,Jc", [(2, 256, 4), (1, 4096, 4), (1, 70, 40)])
def test_attn_cross(cuda_lib, B, N, Jc):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(N * 3 + Jc)
    heads, d = 8, 64
    q = rb(torch.randn(B, N, heads * d, generator=g))
    kv = torch.randn(B, Jc, 2 * heads * d, generator=g)
    null_kv = torch.randn(2, d, generator=g)
    qh = q.view(B, N, heads, d).transpose(1, 2) * d ** -0.5
    k = kv[..., : heads * d].view(B, Jc, heads, d).transpose(1, 2)
    v = kv[..., heads * d:].view(B, Jc, heads, d).transpose(1, 2)
    k = torch.cat((null_kv[0].expand(B, heads, 1, d), k), 2)
    v = torch.cat((null_kv[1].expand(B, heads, 1, d), v), 2)
    ref = (torch.einsum("bhid,bhjd->bhij", qh, k).softmax(-1) @ v).transpose(1, 2).reshape(B, N, heads * d)
    out = ops.attn_cross(bf(q).to(DEV), kv.to(DEV), null_kv.to(DEV), heads, d ** -0.5)
    err = rel_l2(out, ref)
    print("attn_cross rel_l2", err)
    assert err < 4e-3


# ------------------------------------------------------------------------------------------------ edge convs
def test_im2col_and_cross_embed(cuda_lib):
    """CrossEmbedLayer(k=3,7,15) == im2col(15) @ merged 15x15 weight."""
    from kidney_diffusion_b200 import ops
    from oracle.imagen_oracle import CrossEmbedLayer

    torch.manual_seed(8)
    Cin, dim = 3, 128
    mod = CrossEmbedLayer(Cin, kernel_sizes=(3, 7, 15), dim_out=dim, stride=1)
    x = rb(torch.randn(2, Cin, 40, 72))
    Wm = torch.zeros(dim, 15, 15, Cin)
    bias = torch.zeros(dim)
    o = 0
    with torch.no_grad():
        for conv in mod.convs:
            conv.weight.copy_(rb(conv.weight))
            k, co = conv.kernel_size[0], conv.out_channels
            off = (15 - k) // 2
            Wm[o:o + co, off:off + k, off:off + k, :] = conv.weight.permute(0, 2, 3, 1)
            bias[o:o + co] = conv.bias
            o += co
        ref = mod(x)
    Kp = ((15 * 15 * Cin + 63) // 64) * 64
    Wp = torch.zeros(dim, Kp)
    Wp[:, : 15 * 15 * Cin] = Wm.reshape(dim, -1)
    panel = ops.im2col_nchw(x.to(DEV), 15, Kp)
    # spot check the panel itself (exact)
    unf = F.unfold(x, 15, padding=7).view(2, Cin, 225, -1).permute(0, 3, 2, 1).reshape(2 * 40 * 72, 225 * Cin)
    assert torch.equal(panel[:, : 225 * Cin].float().cpu(), unf)
    assert float(panel[:, 225 * Cin:].float().abs().max()) == 0.0
    out = ops.gemm_rows(panel, bf(Wp).to(DEV), bias.to(DEV)).view(2, 40, 72, dim)
    err = rel_l2(from_nhwc(out), ref)
    print("cross-embed rel_l2", err)
    assert err < 5e-3


INIT_CASES = [
    # B, C, H, W, ks, Cout, addend
    (1, 3, 32, 32, 15, 128, False),
    (2, 3, 70, 200, 15, 128, True),     # two column strips, partial strip, odd row count -> single-row last pass
    (1, 1, 33, 129, 7, 64, True),       # one column past the strip boundary
    (3, 2, 16, 16, 3, 64, False),
    (1, 3, 256, 256, 15, 128, True),    # several row segments per strip: ring re-primed per work unit
]


@pytest.mark.parametrize("B,C,H,W,ks,Cout,use_add", INIT_CASES)
def test_init_conv_panel_free(cuda_lib, B, C, H, W, ks, Cout, use_add):
    """kd_init_conv (shifted-copy halo rows + SWIZZLE_NONE UMMA) against F.conv2d on the fp16-rounded operands."""
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(B * 1000 + H + W + ks)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Cout, C, ks, ks, generator=g) / math.sqrt(C * ks * ks)
    bias = torch.randn(Cout, generator=g)
    add = torch.randn(B, H, W, Cout, generator=g) if use_add else None
    Kp = ops.init_conv_kp(C, ks)
    wp = torch.zeros(Cout, ks, C, 16)
    wp[..., :ks] = w.permute(0, 2, 1, 3)  # [n, c, ky, kx] -> [n, ky, c, kx]
    W2 = torch.zeros(Cout, Kp)
    W2[:, :ks * C * 16] = wp.reshape(Cout, -1)
    out = torch.empty(B, H, W, Cout, device=DEV, dtype=torch.float16)
    ops.init_conv(x.to(DEV), ks, bf(W2).to(DEV), bias.to(DEV), None if add is None else bf(add).to(DEV), out)
    ref = F.conv2d(rb(x), rb(w), bias, padding=ks // 2)
    if add is not None:
        ref = ref + rb(add).permute(0, 3, 1, 2)
    got = from_nhwc(out)
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < 2e-3, rel_l2(got, ref)
    # chained in-place accumulation (addend aliases out), as the executor uses for > 3 fixed channels
    ops.init_conv(x.to(DEV), ks, bf(W2).to(DEV), None, out, out)
    ref2 = rb(ref) + F.conv2d(rb(x), rb(w), None, padding=ks // 2)
    assert rel_l2(from_nhwc(out), ref2) < 2e-3


def test_final_conv(cuda_lib):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(9)
    B, H, W, Ca, Cb = 2, 40, 72, 128, 3
    xa, xb = rb(torch.randn(B, Ca, H, W, generator=g)), torch.randn(B, Cb, H, W, generator=g)
    w, b = torch.randn(3, Ca + Cb, 3, 3, generator=g) * 0.05, torch.randn(3, generator=g)
    ref = F.conv2d(torch.cat((xa, xb), 1), w, b, padding=1)
    out = ops.final_conv(nhwc(xa), xb.to(DEV), w.permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV))
    err = rel_l2(out, ref)
    print("final_conv rel_l2", err)
    assert err < 1e-5
    out2 = ops.final_conv(nhwc(xa), None, w[:, :Ca].permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV))
    assert rel_l2(out2, F.conv2d(xa, w[:, :Ca], b, padding=1)) < 1e-5


# ------------------------------------------------------------------------------------------------ sampler update (K6 / K7)
def _sched_scalars(sched, t, t_next):
    from kidney_diffusion_b200.schedule import step_scalars

    return step_scalars(sched, t, t_next)


@pytest.mark.parametrize("objective", ["noise", "v"])
@pytest.mark.parametrize("n_side", [16, 64, 250])
def test_dynthresh_and_ddpm_step_exact(cuda_lib, objective, n_side):
    """K7 radix-select == torch.quantile bit for bit; K6 == oracle p_sample bit for bit (same fp32 expression order)."""
    from kidney_diffusion_b200 import ops
    from kidney_diffusion_b200.schedule import step_scalars
    from oracle.imagen_oracle import GaussianDiffusionContinuousTimes, Imagen, NullUnet

    g = torch.Generator().manual_seed(n_side)
    B = 3
    sched = GaussianDiffusionContinuousTimes(noise_schedule="cosine", timesteps=100)
    x = torch.randn(B, 3, n_side, n_side, generator=g)
    pred = torch.randn(B, 3, n_side, n_side, generator=g) * 1.7
    noise = torch.randn(B, 3, n_side, n_side, generator=g)
    for t, t_next in [(1.0, 0.99), (0.5, 0.49), (0.01, 0.0)]:
        tt, tn = torch.full((B,), t), torch.full((B,), t_next)
        # oracle
        if objective == "noise":
            x0 = sched.predict_start_from_noise(x, tt, pred)
        else:
            x0 = sched.predict_start_from_v(x, tt, pred)
        s_ref = torch.quantile(x0.flatten(1).abs(), 0.95, dim=-1).clamp(min=1.0)
        x0c = x0.clamp(-s_ref.view(B, 1, 1, 1), s_ref.view(B, 1, 1, 1)) / s_ref.view(B, 1, 1, 1)
        mean, _, logvar = sched.q_posterior(x0c, x, tt, t_next=tn)
        nonzero = (1 - (tn == 0).float()).view(B, 1, 1, 1)
        ref = mean + nonzero * (0.5 * logvar).exp() * noise
        # device
        sc = step_scalars("cosine", t, t_next)
        s = ops.dynthresh(x.to(DEV), pred.to(DEV), objective, sc["alpha"], sc["sigma"])
        x0_out = torch.empty_like(x, device=DEV)
        out = ops.ddpm_step(x.to(DEV), pred.to(DEV), noise.to(DEV), s, objective, sc, x0_out=x0_out)
        torch.cuda.synchronize()
        assert torch.equal(s.cpu(), s_ref), (s.cpu(), s_ref)
        assert torch.equal(x0_out.cpu(), x0c)
        md = float((out.cpu() - ref).abs().max())
        assert md <= 1e-6 * float(ref.abs().max()), md


def test_dynthresh_duplicates_and_full_size(cuda_lib):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 1024, 1024, generator=g)
    x[0, :, :512] = 0.25  # massive duplicates
    pred = torch.zeros_like(x)
    # objective x_start -> x0 = pred; use v with alpha=1, sigma=0 to get x0 = x
    s = ops.dynthresh(x.to(DEV), pred.to(DEV), "v", 1.0, 0.0)
    ref = torch.quantile(x.flatten(1).abs(), 0.95, dim=-1).clamp(min=1.0)
    assert torch.equal(s.cpu(), ref), (s.cpu(), ref)


def test_inpaint_blend_finalize_qsample(cuda_lib):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(11)
    B, S = 2, 64
    img, inp, noise = (torch.randn(B, 3, S, S, generator=g) for _ in range(3))
    mask = torch.rand(B, S, S, generator=g) > 0.5
    alpha, sigma = 0.8186915, 0.5742431
    m4 = mask[:, None]
    ref = img * ~m4 + (alpha * inp + sigma * noise) * m4
    out = ops.inpaint_blend(img.clone().to(DEV), inp.to(DEV), mask.to(torch.uint8).to(DEV), noise.to(DEV), alpha, sigma)
    assert torch.equal(out.cpu(), ref)
    ref2 = (img.clamp(-1, 1) * ~m4 + inp * m4 + 1) * 0.5
    out2 = ops.finalize_image(img.clone().to(DEV), inp.to(DEV), mask.to(torch.uint8).to(DEV))
    assert torch.equal(out2.cpu(), ref2)
    out3 = ops.finalize_image(img.clone().to(DEV))
    assert torch.equal(out3.cpu(), (img.clamp(-1, 1) + 1) * 0.5)
    out4 = ops.q_sample(inp.to(DEV), noise.to(DEV), alpha, sigma)
    assert torch.equal(out4.cpu(), alpha * inp + sigma * noise)


def test_randn_statistics_and_determinism(cuda_lib):
    from kidney_diffusion_b200 import ops

    a = ops.randn((3, 1024, 1024), 1234, 77, DEV)
    b = ops.randn((3, 1024, 1024), 1234, 77, DEV)
    c = ops.randn((3, 1024, 1024), 1234, 78, DEV)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(float(a.mean())) < 3e-3 and abs(float(a.std()) - 1) < 3e-3
    assert abs(float((a ** 4).mean()) - 3.0) < 0.05  # kurtosis of N(0,1)
    assert abs(float((a.flatten()[:-1] * a.flatten()[1:]).mean())) < 3e-3


@pytest.mark.parametrize("orientation", [-1, 1])
def test_border_pack_matches_reference_layout(cuda_lib, orientation):
    """Restates sample_ultra_res.py:149-170 on the CPU and compares bit-exactly."""
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(12)
    S, ov = 64, 16
    above, side, corner = (torch.rand(3, S, S, generator=g) for _ in range(3))
    for use in [(1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 0), (1, 1, 0)]:
        a, n, c = (t if u else None for t, u in zip((above, side, corner), use))
        ip, im = torch.zeros(3, S, S), torch.zeros(S, S)
        if a is not None:
            ip[:, :ov, :] = a[:, -ov:, :]
            im[:ov, :] = 1
        if n is not None:
            if orientation == -1:
                ip[:, :, :ov] = n[:, :, -ov:]
                im[:, :ov] = 1
            else:
                ip[:, :, -ov:] = n[:, :, :ov]
                im[:, -ov:] = 1
        if c is not None:
            if orientation == -1:
                ip[:, :ov, :ov] = c[:, -ov:, -ov:]
            else:
                ip[:, :ov, -ov:] = c[:, -ov:, :ov]
        dev = lambda t: None if t is None else t.to(DEV)
        # (1) strips as views into full resident patches
        sa, ss, sc = ops.neighbour_strips(S, ov, orientation, dev(a), dev(n), dev(c))
        op, om = ops.border_pack(S, ov, orientation, sa, ss, sc, DEV)
        assert torch.equal(op.cpu(), ip) and torch.equal(om.cpu().float(), im)
        # (2) contiguous strips, as received from another rank
        cont = lambda st: None if st is None else (st[0].contiguous(), st[0].shape[1] * st[0].shape[2], st[0].shape[2])
        op2, om2 = ops.border_pack(S, ov, orientation, cont(sa), cont(ss), cont(sc), DEV)
        assert torch.equal(op2.cpu(), ip) and torch.equal(om2.cpu().float(), im)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,mode", [(2, 32, 32, 128, 128, 3, 0), (3, 8, 8, 128, 256, 3, 0), (1, 40, 24, 64, 384, 1, 0), (2, 32, 32, 64, 128, 2, 1),
                                                  (1, 128, 64, 128, 128, 3, 0)])
def test_fused_groupnorm_statistics_match_standalone(cuda_lib, B, H, W, Cin, Cout, k, mode):
    """Octet statistics emitted by the conv epilogue (tap-loop and halo kernels, TB = 1 and 2, partial tiles) == a standalone
    pass over the stored tensor == exact fp64 sums of the stored values."""
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(H + Cout)
    Hin, Win = (2 * H, 2 * W) if mode == 1 else (H, W)
    taps = 4 if mode == 1 else k * k
    x = torch.randn(B, Cin, Hin, Win, generator=g)
    w = bf(torch.randn(Cout, taps * Cin, generator=g) / math.sqrt(taps * Cin)).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    out = ops.conv_gemm(nhwc(x), w, b, mode=mode, ksize=k, want_stats=True)
    st = getattr(out, "_kd_stats", None)
    assert st is not None
    fused = st.reduced().double().cpu().sum(1)
    alone = ops.oct_stats(out).reduced().double().cpu().sum(1)
    o = out.double().cpu().view(B, H * W, Cout // 8, 8)
    exact = torch.stack((o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))), dim=-1)
    for name, got in (("fused", fused), ("standalone", alone)):
        err = float(((got - exact).abs() / (exact.abs() + 1.0)).max())
        print(f"{name} octet statistics max rel err {err:.2e}")
        assert err < 2e-5
    # gate_residual epilogue statistics
    gate = torch.rand(B, Cout, generator=g).to(DEV)
    res = bf(torch.randn(B, H, W, Cout, generator=g)).to(DEV)
    o2 = ops.gate_residual(out, gate, res, want_stats=True)
    got = o2._kd_stats.reduced().double().cpu().sum(1)
    oo = o2.double().cpu().view(B, H * W, Cout // 8, 8)
    exact2 = torch.stack((oo.sum(dim=(1, 3)), (oo * oo).sum(dim=(1, 3))), dim=-1)
    assert float(((got - exact2).abs() / (exact2.abs() + 1.0)).max()) < 2e-5
    # finalize from octets == GroupNorm statistics of the stored tensor
    mr = ops.gn_finalize_oct(st, 1.0, None, 1.0, Cout // 8, 8, count=(Cout // 8) * H * W).cpu()
    of = out.float().cpu().permute(0, 3, 1, 2).reshape(B, 8, -1)
    assert torch.allclose(mr[..., 0], of.mean(-1), atol=1e-4, rtol=1e-4)
    assert torch.allclose(mr[..., 1], (of.var(-1, unbiased=False) + 1e-5).rsqrt(), atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 32, 32, 128, 128, 3), (3, 8, 8, 128, 256, 3), (1, 40, 24, 64, 512, 3), (1, 128, 64, 128, 128, 3)])
def test_fused_global_context_logits_match_rowdot(cuda_lib, B, H, W, Cin, Cout, k):
    """GlobalContext to_k logits emitted by the conv epilogue (one partial per 64 output channels) == kd_rowdot over the
    stored tensor, and the pooled context computed from either is the same."""
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(H * 7 + Cout)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = bf(torch.randn(Cout, k * k * Cin, generator=g) / math.sqrt(k * k * Cin)).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    wk = torch.randn(Cout, generator=g).to(DEV)
    out = ops.conv_gemm(nhwc(x), w, b, ksize=k, logit_w=wk, want_stats=True)
    parts = getattr(out, "_kd_logits", None)
    assert parts is not None and parts.shape == (Cout // 64, B, H * W)
    assert getattr(out, "_kd_stats", None) is not None
    alone = ops.rowdot(out, wk, None)
    exact = (out.double().cpu().view(B, H * W, Cout) * wk.double().cpu()).sum(-1)
    scale = float(exact.abs().max())
    assert float((parts.double().cpu().sum(0) - exact).abs().max()) < 1e-5 * scale + 1e-5
    assert float((alone.double().cpu() - exact).abs().max()) < 1e-5 * scale + 1e-5
    pa, pb = ops.gca_pool(out, parts), ops.gca_pool(out, alone)
    assert torch.allclose(pa, pb, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("B,H,W,Ca,Cb,Cout,use_ss,use_add", [
    (2, 32, 32, 128, 0, 128, True, False),
    (1, 40, 24, 64, 64, 256, False, True),     # two sources (skip concat with scale), partial tiles, TMA-fed addend
    (3, 16, 8, 128, 0, 192, True, True),       # single tile per image, Cout not a multiple of 128, odd tile count (idle pair half)
    (1, 128, 64, 256, 0, 128, False, False),
])
def test_conv_fused_groupnorm_preactivation(cuda_lib, B, H, W, Ca, Cb, Cout, use_ss, use_add):
    """conv3x3 with the GroupNorm affine + SiLU applied to the halo tile in shared memory (pre_coef) == the separate
    gn_apply pass followed by the same convolution."""
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(H * 3 + Cout + Ca)
    C, G = Ca + Cb, 8
    gs = C // G
    b_scale = 2 ** -0.5
    xa = nhwc(torch.randn(B, Ca, H, W, generator=g) * 2 + 0.3)
    xb = nhwc(torch.randn(B, Cb, H, W, generator=g)) if Cb else None
    w = bf(torch.randn(Cout, 9 * C, generator=g) / math.sqrt(9 * C)).to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    gamma, beta = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    ss = (torch.randn(B, 2 * C + 5, generator=g) * 0.3).to(DEV)[:, 3:3 + 2 * C] if use_ss else None  # strided column slice
    add = nhwc(torch.randn(B, Cout, H, W, generator=g)) if use_add else None
    assert ops.conv_pre_supported(B, H, W, Ca, Cb, Cout)
    sa, sb = ops.oct_stats(xa), (ops.oct_stats(xb) if Cb else None)
    mr, coef = ops.gn_finalize_oct(sa, 1.0, sb, b_scale, gs, G, count=gs * H * W, gamma=gamma, beta=beta, scale_shift=ss, want_coef=True)
    kw = dict(group_size=gs, num_groups=G, scale_shift=ss, ctot=C)
    ya = ops.gn_apply(xa, mr, gamma, beta, c_offset=0, **kw)
    yb = ops.gn_apply(xb, mr, gamma, beta, c_offset=Ca, src_scale=b_scale, **kw) if Cb else None
    ref = ops.conv_gemm(ya, w, bias, xb=yb, ksize=3, addend=add, want_stats=True)
    got = ops.conv_gemm(xa, w, bias, xb=xb, ksize=3, addend=add, want_stats=True, pre_coef=coef)
    err = rel_l2(got, ref)
    print(f"fused pre-activation vs separate pass rel-L2 {err:.2e}")
    assert torch.isfinite(got.float()).all()
    assert err < 1e-3
    # against plain PyTorch: GroupNorm over the virtual concat -> (scale + 1, shift) -> SiLU -> conv
    xcat = xa.float().cpu().permute(0, 3, 1, 2)
    if Cb:
        xcat = torch.cat((xcat, xb.float().cpu().permute(0, 3, 1, 2) * b_scale), 1)
    y = F.group_norm(xcat, G, gamma.cpu(), beta.cpu(), eps=1e-5)
    if use_ss:
        sc, sh = ss.cpu()[:, :C], ss.cpu()[:, C:]
        y = y * (sc[:, :, None, None] + 1) + sh[:, :, None, None]
    y = F.silu(y)
    wt = w.float().cpu().view(Cout, 3, 3, C).permute(0, 3, 1, 2)
    o = F.conv2d(y, wt, bias.cpu(), padding=1)
    if use_add:
        o = o + add.float().cpu().permute(0, 3, 1, 2)
    assert rel_l2(from_nhwc(got), o) < 4e-3
    # the fused statistics of the output are those of the stored tensor
    st = got._kd_stats.reduced().double().cpu().sum(1)
    ov = got.double().cpu().view(B, H * W, Cout // 8, 8)
    exact = torch.stack((ov.sum(dim=(1, 3)), (ov * ov).sum(dim=(1, 3))), dim=-1)
    assert float(((st - exact).abs() / (exact.abs() + 1.0)).max()) < 2e-5


@pytest.mark.parametrize("B,H,W,Ca,Cb", [(2, 32, 32, 128, 0), (3, 40, 24, 64, 192), (1, 128, 64, 256, 128)])
def test_reduce_finalize_single_launch_bit_identical(cuda_lib, B, H, W, Ca, Cb):
    """kd_gn_reduce_finalize (last block of an image finalizes) == kd_oct_reduce + kd_gn_finalize_oct, bit for bit, on statistics
    from the conv epilogue (source a) and from the standalone pass (source b); repeated launches reuse the arrival counters."""
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(Ca + Cb + H)
    C, G = Ca + Cb, 8
    x = nhwc(torch.randn(B, 64, H, W, generator=g))
    w = bf(torch.randn(Ca, 9 * 64, generator=g) / 24).to(DEV)
    xa = ops.conv_gemm(x, w, None, ksize=3, want_stats=True)
    xb = nhwc(torch.randn(B, Cb, H, W, generator=g)) if Cb else None
    gamma, beta = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    ss = (torch.randn(B, 2 * C, generator=g) * 0.3).to(DEV)

    def run(fused):
        ops.FUSED_REDUCE = fused
        try:
            for t in (xa, xb):
                if t is not None and getattr(t, "_kd_stats", None) is not None:
                    t._kd_stats._reduced = None
            if xb is not None and not hasattr(xb, "_kd_stats"):
                xb._kd_stats = ops.oct_stats(xb)
            return ops.gn_finalize_oct(ops.stats_of(xa), 1.0, ops.stats_of(xb) if Cb else None, 0.5, C // G, G, count=(C // G) * H * W,
                                       gamma=gamma, beta=beta, scale_shift=ss, want_coef=True)
        finally:
            ops.FUSED_REDUCE = True

    mr0, cf0 = run(False)
    for _ in range(3):
        mr1, cf1 = run(True)
        assert torch.equal(mr0, mr1) and torch.equal(cf0, cf1)


# ------------------------------------------------------------------------------------------------ split-K (tiny images, long K)
SPLITK_CASES = [
    # B, H, W, Ca, Cb, Cout, k, mode   (the 8^2 / 16^2 levels of the 64^2 base UNet)
    (1, 8, 8, 1024, 0, 1024, 3, 0),
    (2, 8, 8, 1024, 768, 1024, 3, 0),     # two sources (up path), tile holds two images
    (3, 8, 8, 4608, 0, 768, 1, 0),        # 1x1, batch tail inside a tile
    (2, 4, 8, 1024, 0, 256, 3, 0),        # non-square, tile holds four images -> no fused statistics
    (2, 8, 8, 1152, 0, 512, 2, 1),        # Downsample taps 16x16 -> 8x8
    (1, 8, 8, 1024, 0, 200, 3, 0),        # Cout not a multiple of 64
]


@pytest.mark.parametrize("B,H,W,Ca,Cb,Cout,k,mode", SPLITK_CASES)
def test_conv_splitk_matches_conv2d_and_unsplit_kernel(cuda_lib, B, H, W, Ca, Cb, Cout, k, mode):
    """Split-K path (kd_conv_splitk_workspace_bytes > 0) against F.conv2d and against the un-split kernels (impl 8), with the
    fused epilogue: bias, SiLU, gated addend, GroupNorm octet statistics, GlobalContext logits."""
    import ctypes

    from kidney_diffusion_b200 import _lib, ops

    g = torch.Generator().manual_seed(B + H + Ca + Cout)
    Cin = Ca + Cb
    Hin, Win = (2 * H, 2 * W) if mode == 1 else (H, W)
    x = rb(torch.randn(B, Cin, Hin, Win, generator=g))
    taps = 4 if mode == 1 else k * k
    w = rb(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * taps))
    b = torch.randn(Cout, generator=g)
    add = rb(torch.randn(B, Cout, H, W, generator=g))
    gate = torch.rand(B, Cout, generator=g)
    d = _lib.KdConvDesc(mode, B, H, W, Ca, Cb, Cout, k if mode == 0 else 1, ops.ACT_SILU, 0, 0, 0)
    assert ops.lib().kd_conv_splitk_workspace_bytes(ctypes.byref(d)) > 0, "case must exercise the split-K path"
    xa = nhwc(x[:, :Ca])
    xb = nhwc(x[:, Ca:]) if Cb else None
    if mode == 1:  # Downsample: rearrange 'b c (h 2) (w 2) -> b (c 4) h w' then 1x1 conv; packed weight columns ordered (dy, dx, c)
        wp = bf(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()).to(DEV)
        ref = F.conv2d(x, w, b, stride=2)
    else:
        wp = pack_w(w)
        ref = F.conv2d(x, w, b, padding=k // 2)
    ref = F.silu(ref) + gate[:, :, None, None] * add
    kw = dict(mode=mode, ksize=k, act=ops.ACT_SILU, addend=nhwc(add), addend_scale=gate.to(DEV), want_stats=True)
    out = ops.conv_gemm(xa, wp, b.to(DEV), xb=xb, **kw)
    torch.cuda.synchronize()
    err = rel_l2(from_nhwc(out), ref)
    print(f"split-K conv {B}x{H}x{W} {Cin}->{Cout} k{k} mode{mode}: rel_l2={err:.3e}")
    assert err < 5e-3
    try:
        ops.set_conv_impl(8)
        plain = ops.conv_gemm(xa, wp, b.to(DEV), xb=xb, **kw)
    finally:
        ops.set_conv_impl(0)
    assert rel_l2(out, plain) < 3e-4
    # fused statistics == statistics of the stored tensor
    st = getattr(out, "_kd_stats", None)
    assert st is not None or H * W < 64, "split-K shapes must still emit fused GroupNorm statistics"
    if st is not None:
        red = st.reduced().double().sum(1)  # [B, C/8, 2]
        xo = out.double().reshape(B, H * W, Cout // 8, 8)
        assert torch.allclose(red[..., 0], xo.sum((1, 3)), rtol=1e-5, atol=1e-3) and torch.allclose(red[..., 1], (xo * xo).sum((1, 3)), rtol=1e-5, atol=1e-3)
    if Cout % 64 == 0:  # fused GlobalContext logits (no gate allowed on that path)
        lw = torch.randn(Cout, generator=g).to(DEV)
        o2 = ops.conv_gemm(xa, wp, b.to(DEV), xb=xb, mode=mode, ksize=k, logit_w=lw)
        parts = getattr(o2, "_kd_logits", None)
        assert parts is not None or H * W < 64
        if parts is not None:
            want = (o2.float().reshape(B, H * W, Cout) * lw).sum(-1)
            assert torch.allclose(parts.sum(0), want, rtol=1e-4, atol=1e-3)


def test_conv_splitk_is_batch_invariant(cuda_lib):
    """The split count depends on the per-sample shape only: a sample's result must not change with the batch it is in."""
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(3)
    x = rb(torch.randn(5, 1024, 8, 8, generator=g))
    w = rb(torch.randn(512, 1024, 3, 3, generator=g) / 96.0)
    b = torch.randn(512, generator=g).to(DEV)
    full = ops.conv_gemm(nhwc(x), pack_w(w), b, ksize=3)
    for i in (0, 3, 4):
        one = ops.conv_gemm(nhwc(x[i:i + 1]), pack_w(w), b, ksize=3)
        assert torch.equal(one[0], full[i])
    pair = ops.conv_gemm(nhwc(x[1:4]), pack_w(w), b, ksize=3)
    assert torch.equal(pair, full[1:4])


def test_conv_splitk_pixel_shuffle_and_f32_out(cuda_lib):
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(4)
    B, H, W, Cin, Cout = 2, 8, 8, 1024, 512
    x = rb(torch.randn(B, Cin, H, W, generator=g))
    w = rb(torch.randn(Cout, Cin, 1, 1, generator=g) / 32.0)
    b = torch.randn(Cout, generator=g)
    y = F.silu(F.conv2d(x, w, b))
    # PixelShuffleUpsample: weight rows ordered (dy, dx, c) -> out [B, 2H, 2W, Cout/4]
    co = Cout // 4
    ref = y.view(B, 4, co, H, W).permute(0, 2, 3, 4, 1).reshape(B, co, H, W, 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, co, 2 * H, 2 * W)
    out = ops.conv_gemm(nhwc(x), pack_w(w), b.to(DEV), ksize=1, act=ops.ACT_SILU, out_mode=1)
    assert out.shape == (B, 2 * H, 2 * W, co) and rel_l2(from_nhwc(out), ref) < 5e-3
    o32 = ops.conv_gemm(nhwc(x), pack_w(w), b.to(DEV), ksize=1, out_f32=True)
    assert o32.dtype == torch.float32 and rel_l2(from_nhwc(o32), F.conv2d(x, w, b)) < 1e-5


def test_torch_custom_ops_match_direct_wrappers(cuda_lib):
    """torch.ops.kidney_b200.* (the registered custom ops) launch the same kernels as the ctypes wrappers."""
    from kidney_diffusion_b200 import ops, torch_ops  # noqa: F401

    K = torch.ops.kidney_b200
    g = torch.Generator().manual_seed(0)
    x, pred, noise = (torch.randn(2, 3, 32, 32, generator=g).to(DEV) for _ in range(3))
    sc = dict(alpha=0.8, sigma=0.6, one_minus_c=0.9, c=0.1, alpha_next=0.85, std=0.05)
    s = K.dynthresh(x, pred, "v", sc["alpha"], sc["sigma"], 0.95, None)
    assert torch.equal(s, ops.dynthresh(x, pred, "v", sc["alpha"], sc["sigma"], 0.95))
    a = K.ddpm_step(x, pred, noise, s, "v", sc["alpha"], sc["sigma"], sc["one_minus_c"], sc["c"], sc["alpha_next"], sc["std"], None, 0.0, 0.0, 1.0)
    assert torch.equal(a, ops.ddpm_step(x, pred, noise, s, "v", sc))
    xa = bf(torch.randn(1, 16, 16, 64, generator=g)).to(DEV)
    w = bf(torch.randn(128, 576, generator=g) / 24).to(DEV)
    assert torch.equal(K.conv2d_nhwc(xa, w, None, None, 0, 3, 0, 0, False, None, None), ops.conv_gemm(xa, w, None, ksize=3))
    patch = torch.rand(3, 64, 64, generator=g).to(DEV)
    ip, m = K.border_pack(64, 16, -1, patch[:, 48:, :], patch[:, :, 48:], None, patch)
    ip2, m2 = ops.border_pack(64, 16, -1, *ops.neighbour_strips(64, 16, -1, above=patch, side=patch)[:2], None, patch.device)
    assert torch.equal(ip, ip2) and torch.equal(m, m2)
    assert torch.equal(K.randn_like(x, 5, 9), ops.randn(tuple(x.shape), 5, 9, x.device))


# ------------------------------------------------------------------------------------------------ linear attention (north_star b)
@pytest.mark.parametrize("B,H,W,dim,Jc", [(2, 16, 16, 128, 0), (1, 32, 32, 256, 6), (3, 20, 12, 64, 3)])
def test_linear_attention_block_matches_oracle(cuda_lib, B, H, W, dim, Jc):
    """LinearAttention.forward of the oracle (fp32 CPU) against the CUDA pipeline the executor runs: ChanLayerNorm -> fused q|k|v
    1x1 conv -> depthwise 3x3 -> kd_linattn_context / kd_linattn_apply -> 1x1 conv -> ChanLayerNorm."""
    from kidney_diffusion_b200 import ops
    from oracle.imagen_oracle import LinearAttention

    torch.manual_seed(dim + H)
    cd = 96
    mod = LinearAttention(dim, dim_head=64, heads=8, context_dim=cd if Jc else None).eval()
    with torch.no_grad():
        for t in (mod.to_q, mod.to_k, mod.to_v):
            t[1].weight.copy_(rb(t[1].weight * 2.0))
            t[2].weight.mul_(2.0)
        mod.to_out[0].weight.copy_(rb(mod.to_out[0].weight))
        mod.norm.g.copy_(torch.rand_like(mod.norm.g) + 0.5)
    x = rb(torch.randn(B, dim, H, W))
    ctx = torch.randn(B, Jc, cd) if Jc else None
    with torch.no_grad():
        ref = mod(x, context=ctx)
    dx = nhwc(x)
    xn = ops.layernorm_h16(dx, mod.norm.g.detach().reshape(-1).to(DEV))
    w1 = pack_w(torch.cat([t[1].weight.detach() for t in (mod.to_q, mod.to_k, mod.to_v)], 0))
    wdw = torch.cat([t[2].weight.detach() for t in (mod.to_q, mod.to_k, mod.to_v)], 0).reshape(-1, 3, 3).contiguous().to(DEV)
    qkv1 = ops.conv_gemm(xn, w1, None, ksize=1)
    qkv = ops.dwconv3x3(qkv1, wdw)
    # depthwise conv alone vs F.conv2d on the same (fp16-rounded) input
    want_dw = F.conv2d(from_nhwc(qkv1), wdw.cpu().reshape(-1, 1, 3, 3), padding=1, groups=wdw.shape[0])
    assert rel_l2(from_nhwc(qkv), want_dw) < 1e-3
    ckv = None
    if Jc:
        cn = ops.layernorm_f32(ctx.to(DEV).view(B * Jc, cd), mod.to_context[0].weight.detach().to(DEV), mod.to_context[0].bias.detach().to(DEV))
        ckv = ops.linear_small(cn, mod.to_context[1].weight.detach().to(DEV), None).view(B, Jc, -1)
    o = ops.linear_attention(qkv.view(B, H * W, -1), 8, mod.scale, ckv)
    o = ops.conv_gemm(o.view(B, H, W, -1), pack_w(mod.to_out[0].weight.detach()), None, ksize=1)
    out = ops.layernorm_h16(o, mod.to_out[1].g.detach().reshape(-1).to(DEV))
    err = rel_l2(from_nhwc(out), ref)
    print(f"linear attention {B}x{H}x{W} dim {dim} ctx {Jc}: rel_l2 = {err:.3e}")
    assert err < 5e-3
    if B > 1:  # batch invariance
        one = ops.linear_attention(qkv.view(B, H * W, -1)[1:2].contiguous(), 8, mod.scale, None if ckv is None else ckv[1:2].contiguous())
        full = ops.linear_attention(qkv.view(B, H * W, -1), 8, mod.scale, ckv)
        assert torch.equal(one[0], full[1])


@pytest.mark.parametrize("B,N,Jc", [(2, 256, 0), (1, 1000, 5), (1, 4096, 0), (3, 384, 36)])
def test_attn_mqa_tcgen05_matches_mma_sync_kernel(cuda_lib, B, N, Jc):
    """kd_attn_mqa_tc (tcgen05 / TMEM / TMA, N >= 256) against the legacy mma.sync flash kernel and the fp32 definition: ragged
    query tiles (N = 1000), key counts that are not multiples of 8 or 128 (null + context rows), batch invariance."""
    from kidney_diffusion_b200 import ops

    g = torch.Generator().manual_seed(N + Jc + B)
    heads, d = 8, 64
    qkv = rb(torch.randn(B, N, heads * d + 2 * d, generator=g) * 1.5)
    ctx = torch.randn(B, Jc, 2 * d, generator=g) if Jc else None
    null_kv = torch.randn(2, d, generator=g)
    dq = bf(qkv).to(DEV)
    kv = ops.kv_assemble(dq, heads * d, None if ctx is None else ctx.to(DEV), null_kv.to(DEV))
    out_tc = ops.attn_mqa(dq, kv, heads, d ** -0.5)
    saved = ops.ATTN_TC_MIN_TOKENS
    try:
        ops.ATTN_TC_MIN_TOKENS = 1 << 30
        out_legacy = ops.attn_mqa(dq, kv, heads, d ** -0.5)
    finally:
        ops.ATTN_TC_MIN_TOKENS = saved
    torch.cuda.synchronize()
    q = qkv[..., : heads * d].view(B, N, heads, d).transpose(1, 2) * d ** -0.5
    kf, vf = kv[..., :d].float().cpu(), kv[..., d:].float().cpu()
    attn = torch.einsum("bhid,bjd->bhij", q, kf).softmax(-1)
    ref = torch.einsum("bhij,bjd->bhid", attn, vf).transpose(1, 2).reshape(B, N, heads * d)
    e_tc, e_leg = rel_l2(out_tc, ref), rel_l2(out_legacy, ref)
    print(f"attn N={N} J={kv.shape[1]}: tcgen05 {e_tc:.3e}, mma.sync {e_leg:.3e}")
    assert bool(torch.isfinite(out_tc).all()) and e_tc < 4e-3 and rel_l2(out_tc, out_legacy) < 4e-3
    if B > 1:
        one = ops.attn_mqa(dq[1:2].contiguous(), kv[1:2].contiguous(), heads, d ** -0.5)
        assert torch.equal(one[0], out_tc[1])


@pytest.mark.parametrize("B,N,J", [(2, 256, 4), (1, 1000, 38)])
def test_linear_cross_attention_matches_oracle(cuda_lib, B, N, J):
    """LinearCrossAttention.forward (Unet(use_linear_cross_attn=...)): keys / values are the null + context tokens only."""
    from kidney_diffusion_b200 import ops
    from oracle.imagen_oracle import LinearCrossAttention

    torch.manual_seed(N + J)
    dim, cd, heads = 128, 96, 8
    mod = LinearCrossAttention(dim, context_dim=cd, dim_head=64, heads=heads).eval()
    with torch.no_grad():
        mod.to_q.weight.copy_(rb(mod.to_q.weight * 3.0))
        mod.to_out[0].weight.copy_(rb(mod.to_out[0].weight))
    x = rb(torch.randn(B, N, dim))
    ctx = torch.randn(B, J, cd)
    with torch.no_grad():
        ref = mod(x, ctx)
    dx = bf(x).to(DEV).view(B, N, 1, dim)
    xn = ops.layernorm_h16(dx, mod.norm.g.detach().to(DEV))
    q = ops.conv_gemm(xn, bf(mod.to_q.weight.detach()).to(DEV), None, ksize=1).view(B, N, -1)
    kv = ops.linear_small(ctx.to(DEV).view(B * J, cd), mod.to_kv.weight.detach().to(DEV)).view(B, J, -1)
    nk = mod.null_kv.detach().to(DEV)
    null_row = torch.cat((nk[0].repeat(heads), nk[1].repeat(heads))).view(1, 1, -1).expand(B, 1, -1)
    tokens = torch.cat((null_row, kv), 1).contiguous()
    o = ops.linear_attention(q, heads, mod.scale, tokens, act=ops.ACT_NONE, pixels_kv=False)
    o = ops.conv_gemm(o.view(B, N, 1, -1), bf(mod.to_out[0].weight.detach()).to(DEV), None, ksize=1)
    out = ops.layernorm_h16(o, mod.to_out[1].g.detach().to(DEV)).view(B, N, dim)
    err = rel_l2(out, ref)
    print(f"linear cross attention N={N} J={J}: rel_l2 = {err:.3e}")
    assert err < 5e-3
