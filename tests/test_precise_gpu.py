"""The fp32 ("precise") path (csrc/kd_precise.cu, ops_f32, Unet.precision = "fp32"): kernels against torch fp64 restatements,
then UNet forward / sampler parity against the fp32 CPU oracle at BASELINE.json north_star's fp32 tolerance, rel-L2 <= 1e-4 per
UNet step and on final samples."""
import pytest
import torch
import torch.nn.functional as F

from helpers import U1_KW, U2_KW, U3_KW, KeyedNoise, make_pair, rel_l2

pytestmark = pytest.mark.gpu
TOL32 = 1e-4


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("case", [
    dict(B=2, H=12, W=10, Ca=64, Cb=0, Cout=72, ks=3),
    dict(B=1, H=9, W=17, Ca=64, Cb=32, Cout=64, ks=3, addend=True, gate=True),
    dict(B=3, H=8, W=8, Ca=128, Cb=0, Cout=192, ks=1, act="gelu"),
    dict(B=2, H=16, W=12, Ca=32, Cb=0, Cout=64, mode=1),
    dict(B=2, H=6, W=5, Ca=64, Cb=0, Cout=128, ks=1, act="silu", out_mode=1),
    dict(B=1, H=20, W=20, Ca=3, Cb=0, Cout=64, ks=15),          # CrossEmbed-sized filter on image channels (scalar loader)
    dict(B=2, H=10, W=10, Ca=64, Cb=3, Cout=3, ks=3),           # final conv on cat(x, lowres)
])
def test_conv_f32_matches_torch(cuda_lib, case):
    from kidney_diffusion_b200 import ops, ops_f32 as K

    g = torch.Generator().manual_seed(case["Cout"] + case["H"])
    B, H, W, Ca, Cb, Cout = (case[k] for k in ("B", "H", "W", "Ca", "Cb", "Cout"))
    mode, ks = case.get("mode", 0), case.get("ks", 2)
    C = Ca + Cb
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Cout, C, ks, ks, generator=g) / (C * ks * ks) ** 0.5
    bias = torch.randn(Cout, generator=g)
    if mode == 1:  # Downsample: Rearrange 'b c (h s1) (w s2) -> b (c s1 s2) h w' then 1x1 conv; packed as the executor does
        w4 = torch.randn(Cout, C * 4, generator=g) / (C * 4) ** 0.5
        xr = x.view(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(B, C * 4, H // 2, W // 2)
        ref = F.conv2d(xr.double(), w4.double()[:, :, None, None], bias.double())
        wp = w4.view(Cout, C, 2, 2).permute(0, 2, 3, 1).reshape(Cout, 4 * C).contiguous()
    else:
        ref = F.conv2d(x.double(), w.double(), bias.double(), padding=ks // 2)
        wp = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    act = dict(none=ops.ACT_NONE, silu=ops.ACT_SILU, gelu=ops.ACT_GELU)[case.get("act", "none")]
    if case.get("act") == "silu":
        ref = F.silu(ref)
    if case.get("act") == "gelu":
        ref = F.gelu(ref)
    if case.get("out_mode"):
        # packed rows (dy*2+dx, c): conv channel n = q * Cout/4 + c lands at pixel (2y+dy, 2x+dx), channel c
        Cq = Cout // 4
        ref = ref.view(B, 2, 2, Cq, H, W).permute(0, 3, 4, 1, 5, 2).reshape(B, Cq, 2 * H, 2 * W)
    addend = gate = None
    if case.get("addend"):
        addend = torch.randn(ref.shape, generator=g)
        gate = torch.rand(B, Cout, generator=g)
        ref = ref + addend.double() * gate.double()[:, :, None, None]
    xa = _nhwc(x[:, :Ca]).cuda()
    xb = _nhwc(x[:, Ca:]).cuda() if Cb else None
    out = K.conv_gemm(xa, wp.cuda(), bias.cuda(), xb=xb, mode=mode, ksize=ks, act=act, out_mode=case.get("out_mode", 0),
                      addend=None if addend is None else _nhwc(addend).cuda(), addend_scale=None if gate is None else gate.cuda())
    err = rel_l2(out.permute(0, 3, 1, 2), ref)
    print(f"conv_f32 {case}: rel_l2 {err:.2e}")
    assert err < 2e-6


def test_groupnorm_f32_two_sources_with_scale_shift(cuda_lib):
    from kidney_diffusion_b200 import ops_f32 as K

    g = torch.Generator().manual_seed(4)
    B, H, W, Ca, Cb, G = 2, 40, 36, 96, 48, 8   # group size 18: groups straddle the source boundary at channel 96
    xa, xb = torch.randn(B, Ca, H, W, generator=g) * 3 + 1, torch.randn(B, Cb, H, W, generator=g)
    gamma, beta = torch.randn(Ca + Cb, generator=g), torch.randn(Ca + Cb, generator=g)
    ss = torch.randn(B, 2 * (Ca + Cb), generator=g) * 0.3
    s = 2 ** -0.5
    x = torch.cat((xa, xb * s), 1).double()
    ref = F.group_norm(x, G, gamma.double(), beta.double(), eps=1e-5)
    scale, shift = ss.double()[:, :Ca + Cb, None, None], ss.double()[:, Ca + Cb:, None, None]
    ref = F.silu(ref * (scale + 1) + shift)
    ya, yb = K.groupnorm(_nhwc(xa).cuda(), _nhwc(xb).cuda(), s, G, gamma.cuda(), beta.cuda(), ss.cuda())
    out = torch.cat((ya, yb), -1).permute(0, 3, 1, 2)
    err = rel_l2(out, ref)
    print(f"groupnorm_f32 rel_l2 {err:.2e}")
    assert err < 2e-6


def test_attention_f32_shared_and_per_head_kv(cuda_lib):
    from kidney_diffusion_b200 import ops_f32 as K

    g = torch.Generator().manual_seed(8)
    B, N, J, heads = 2, 50, 77, 4
    qkv = torch.randn(B, N, heads * 64 + 128, generator=g)
    kv = torch.randn(B, J, 128, generator=g)
    scale = 64 ** -0.5
    q = qkv[:, :, :heads * 64].view(B, N, heads, 64).permute(0, 2, 1, 3).double()
    sim = torch.einsum("bhnd,bjd->bhnj", q * scale, kv[:, :, :64].double())
    ref = torch.einsum("bhnj,bjd->bhnd", sim.softmax(-1), kv[:, :, 64:].double()).permute(0, 2, 1, 3).reshape(B, N, heads * 64)
    out = K.attn_mqa(qkv.cuda(), kv.cuda(), heads, scale)
    e1 = rel_l2(out, ref)
    # CrossAttention: per-head k / v from the tokens, one null k / v (shared by the heads) in front
    Jc = 9
    tok = torch.randn(B, Jc, 2 * heads * 64, generator=g)
    null = torch.randn(2, 64, generator=g)
    qc = torch.randn(B, N, heads * 64, generator=g)
    k = torch.cat((null[0].repeat(heads).view(1, 1, -1).expand(B, 1, -1), tok[:, :, :heads * 64]), 1).view(B, Jc + 1, heads, 64).double()
    v = torch.cat((null[1].repeat(heads).view(1, 1, -1).expand(B, 1, -1), tok[:, :, heads * 64:]), 1).view(B, Jc + 1, heads, 64).double()
    sim = torch.einsum("bnhd,bjhd->bhnj", qc.view(B, N, heads, 64).double() * scale, k)
    ref2 = torch.einsum("bhnj,bjhd->bnhd", sim.softmax(-1), v).reshape(B, N, heads * 64)
    out2 = K.attn_cross(qc.cuda(), tok.cuda(), null.cuda(), heads, scale)
    e2 = rel_l2(out2, ref2)
    print(f"attn_f32 shared kv {e1:.2e}, per-head kv {e2:.2e}")
    assert e1 < 2e-6 and e2 < 2e-6


def test_global_context_f32(cuda_lib):
    from kidney_diffusion_b200 import ops_f32 as K

    g = torch.Generator().manual_seed(2)
    B, H, W, C, hid = 2, 48, 44, 96, 48
    x = torch.randn(B, C, H, W, generator=g)
    wk, bk = torch.randn(C, generator=g) * 0.3, torch.randn(1, generator=g)
    w0, b0 = torch.randn(hid, C, generator=g) / C ** 0.5, torch.randn(hid, generator=g)
    w1, b1 = torch.randn(C, hid, generator=g) / hid ** 0.5, torch.randn(C, generator=g)
    res = torch.randn(B, C, H, W, generator=g)
    xd = x.double()
    ctx = (torch.einsum("bchw,c->bhw", xd, wk.double()) + bk.double()).flatten(1).softmax(-1)
    pooled = torch.einsum("bn,bcn->bc", ctx, xd.flatten(2))
    gate = torch.sigmoid(F.silu(pooled @ w0.double().t() + b0.double()) @ w1.double().t() + b1.double())
    ref = xd * gate[:, :, None, None] + res.double()
    xh = _nhwc(x).cuda()
    logits = K.rowdot(xh, wk.cuda(), bk.cuda())
    gt = K.gca_gate(xh, logits, w0.cuda(), b0.cuda(), w1.cuda(), b1.cuda())
    out = K.gate_residual(xh, gt, _nhwc(res).cuda())
    e_gate, e_out = rel_l2(gt, gate), rel_l2(out.permute(0, 3, 1, 2), ref)
    print(f"global context f32: gate {e_gate:.2e}, gated residual {e_out:.2e}")
    assert e_gate < 5e-6 and e_out < 2e-6


@pytest.mark.parametrize("name,kw,lowres,S,B", [
    ("u3", U3_KW, True, 128, 2),
    ("u2", U2_KW, True, 64, 2),
    ("u1", U1_KW, False, 32, 3),
    ("u3_b1_rect", U3_KW, True, 64, 1),
])
def test_unet_forward_parity_fp32(cuda_lib, name, kw, lowres, S, B):
    ou, pu = make_pair(kw, lowres_cond=lowres, seed=sum(map(ord, name)))
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 3, S, S, generator=g)
    t = torch.tensor([2.18, -0.5, 5.0])[:B]
    lr = torch.randn(B, 3, S, S, generator=g) if lowres else None
    lt = torch.full((B,), 0.7093) if lowres else None
    cond = torch.rand(B, kw.get("cond_images_channels", 0), S * 2, S * 2, generator=g) if kw.get("cond_images_channels") else None
    with torch.no_grad():
        ref = ou(x, t, lowres_cond_img=lr, lowres_noise_times=lt, cond_images=cond)
    dev = lambda v: None if v is None else v.cuda()
    fast = pu(dev(x), dev(t), lowres_cond_img=dev(lr), lowres_noise_times=dev(lt), cond_images=dev(cond))
    pu.precision = "fp32"
    out = pu(dev(x), dev(t), lowres_cond_img=dev(lr), lowres_noise_times=dev(lt), cond_images=dev(cond))
    assert pu.executor().precise
    e32, e16, cross = rel_l2(out, ref), rel_l2(fast, ref), rel_l2(fast, out)
    print(f"[{name}] fp32 path vs oracle {e32:.3e}; fp16 path vs oracle {e16:.3e}; fp16 vs fp32 path on the GPU {cross:.3e}")
    assert e32 < TOL32 and e16 < 1e-2
    # a sample's result does not depend on the batch it was computed in
    one = pu(dev(x)[:1], dev(t)[:1], lowres_cond_img=None if lr is None else dev(lr)[:1], lowres_noise_times=None if lt is None else dev(lt)[:1],
             cond_images=None if cond is None else dev(cond)[:1])
    assert torch.equal(one, out[:1])
    pu.precision = "fp16"
    assert torch.equal(pu(dev(x), dev(t), lowres_cond_img=dev(lr), lowres_noise_times=dev(lt), cond_images=dev(cond)), fast)


def test_sample_parity_fp32_base_stage(cuda_lib):
    """BASELINE config 1 shape (unconditional 64x64, dim 128, batch 4; 8 steps): final samples and every step's UNet output at 1e-4."""
    from test_unet_parity_gpu import CFG1_KW, _imagen_pair, _per_step_unet_errors

    oi, pi = _imagen_pair([CFG1_KW], (64,), (8,), ("noise",))
    pi.set_precision("fp32")
    kn = KeyedNoise(7)
    ref_steps = []
    ref = oi.sample(batch_size=4, noise_fn=kn.cpu, step_taps=ref_steps)
    pi.noise_fn = kn.dev
    out = pi.sample(batch_size=4, use_tqdm=False, device="cuda")
    assert pi.unets[0].executor().precise
    errs = _per_step_unet_errors(pi, 1, ref_steps, 1, 4)
    err = rel_l2(out, ref)
    print(f"fp32 path, base stage: final rel_l2 = {err:.3e}, per-step UNet output (identical inputs) worst {max(errs):.3e}")
    assert max(errs) < TOL32 and err < TOL32


def test_sample_parity_fp32_sr_stage_with_inpainting(cuda_lib):
    """SR stage (v objective, cond image, low-res conditioning, RePaint inpainting r = 2) on the fp32 path at 1e-4."""
    from test_unet_parity_gpu import _imagen_pair, _per_step_unet_errors

    oi, pi = _imagen_pair([None, U3_KW], (32, 128), (4, 5), ("noise", "v"))
    pi.set_precision("fp32")
    kn = KeyedNoise(99)
    g = torch.Generator().manual_seed(3)
    B = 2
    cond = torch.rand(B, 3, 256, 256, generator=g)
    start = torch.rand(B, 3, 32, 32, generator=g)
    inp = torch.rand(B, 3, 128, 128, generator=g)
    mask = torch.zeros(B, 128, 128)
    mask[:, :32, :] = 1
    mask[:, :, :32] = 1
    ref_steps = []
    kw = dict(batch_size=B, cond_images=cond, start_image_or_video=start, start_at_unet_number=2, stop_at_unet_number=2, inpaint_images=inp,
              inpaint_masks=mask, inpaint_resample_times=2)
    ref = oi.sample(**kw, noise_fn=kn.cpu, step_taps=ref_steps)
    pi.noise_fn = kn.dev
    out = pi.sample(**kw, use_tqdm=False, device="cuda")
    errs = _per_step_unet_errors(pi, 2, ref_steps, 2, B)
    err = rel_l2(out, ref)
    print(f"fp32 path, SR stage: final rel_l2 = {err:.3e}, per-step worst {max(errs):.3e}")
    assert max(errs) < TOL32 and err < TOL32


def test_conditioned_cascade_fp32(cuda_lib):
    """Mask + clinical-vector conditioned cascade with classifier-free guidance on the fp32 path."""
    from test_unet_parity_gpu import _cond_pair

    oi, pi = _cond_pair()
    pi.set_precision("fp32")
    B = 2
    g = torch.Generator().manual_seed(1)
    conds = torch.tensor([0.0, 0.5, 0.2]).reshape(1, 1, 3).repeat_interleave(B, dim=0)
    labels = torch.randint(0, 5, (B, 128, 128), generator=g)
    deep = torch.stack([(labels == k).float() for k in range(1, 5)], dim=1)
    kn = KeyedNoise(21)
    ref = oi.sample(text_embeds=conds, cond_images=deep, cond_scale=3.0, noise_fn=kn.cpu)
    pi.noise_fn = kn.dev
    out = pi.sample(text_embeds=conds, cond_images=deep, cond_scale=3.0, use_tqdm=False, device="cuda")
    err = rel_l2(out, ref)
    print(f"fp32 path, conditioned cascade (cond_scale 3): final rel_l2 = {err:.3e}")
    assert err < TOL32


def test_linear_attention_fp32(cuda_lib):
    """LinearAttentionTransformerBlock and LinearCrossAttention on the fp32 path (use_linear_attn / use_linear_cross_attn)."""
    kw = dict(U1_KW, layer_attns=(False, False, True, True), use_linear_attn=(False, True, False, False),
              use_linear_cross_attn=(False, True, False, False))
    ou, pu = make_pair(kw, lowres_cond=False, seed=77)
    g = torch.Generator().manual_seed(5)
    x, t = torch.randn(2, 3, 32, 32, generator=g), torch.tensor([2.18, -0.5])
    with torch.no_grad():
        ref = ou(x, t)
    pu.precision = "fp32"
    out = pu(x.cuda(), t.cuda())
    err = rel_l2(out, ref)
    print(f"[u1 with linear attention] fp32 path vs oracle {err:.3e}")
    assert err < TOL32
