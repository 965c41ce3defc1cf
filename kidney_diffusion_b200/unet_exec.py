"""CUDA executor of the UNet forward pass (imagen-pytorch 1.18.5 ``Unet.forward`` op order).

Everything that touches activations is a hand-written kernel reached through ``ops`` (C ABI).  torch is used for
memory, for one-time weight packing, and for the x-independent nearest resize of the conditioning images.

Data layout in HBM: activations NHWC fp16 (saturating conversions); only "raw" tensors (conv outputs / residual stream)
exist -- the GroupNorm+SiLU'd operand of a 3x3 conv is produced inside the conv kernel; statistics, softmax, time /
conditioning towers and the GlobalContext gate are fp32.  Weights are packed once: conv [Cout, kh*kw*Cin] fp16 K-major
(tap-major, channel-minor) so a (tap, 64-channel chunk) k-block is one TMA box.

Fusions relative to the reference's eager graph:
  * channel concat of the up path is never materialised (two TMA sources in the conv kernel; the 2^-0.5 skip scale is
    folded into the GroupNorm affine and into res_conv's weight columns),
  * GroupNorm statistics come out of the producing kernel's epilogue (conv / gate_residual), one launch turns them into the
    per-channel affine {A, B} (incl. the time scale / shift), and the consuming 3x3 conv applies A*x + B and SiLU to its halo
    tile in shared memory (`_norm_conv`); the separate gn_apply pass remains only for shapes the halo kernel does not take,
  * bias / SiLU / GELU / residual add / GlobalContext to_k logits / gate * h + res_conv(x) / pixel-shuffle run in the conv
    epilogue,
  * Downsample's pixel-unshuffle is a 2x2 tap pattern of the TMA loads, Parallel(3x3, 1x1) is one 3x3 conv,
  * the CrossEmbed init conv is one panel-free tensor-core kernel (three filters merged into 15x15); its cond-image /
    low-res part is x-independent and computed once per sample() call,
  * all ResnetBlock time-MLPs are one launch; to_time_cond + to_lowres_time_cond are one launch.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import ops, ops_f32
from .modules import LinearAttentionTransformerBlock, Parallel, PixelShuffleUpsample, TransformerBlock, exists

H16 = ops.ACT_DTYPE  # 16-bit activation / weight dtype (saturating fp16)
IM2COL_BUDGET_BYTES = 4 << 30


_PACK_DTYPE = [H16]  # dtype the packers below produce: fp16 for the tensor-core path, fp32 while a precise executor is being built


def _bf(t):
    return t.detach().to(_PACK_DTYPE[0]).contiguous()


def _pack_conv(weight, b_cols=0, b_scale=1.0):
    """[Cout, Cin, kh, kw] fp32 -> [Cout, kh*kw*Cin] fp16; the last `b_cols` input channels are pre-scaled by b_scale."""
    w = weight.detach().float()
    if w.dim() == 2:
        w = w[:, :, None, None]
    if b_cols and b_scale != 1.0:
        w = w.clone()
        w[:, w.shape[1] - b_cols:] *= b_scale
    return _bf(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))


class _Res:
    """Packed ResnetBlock."""

    def __init__(self, m, b_cols=0, b_scale=1.0):
        self.dim, self.dim_out, self.G = m.dim, m.dim_out, m.groups
        self.b_cols, self.b_scale = b_cols, b_scale
        self.g1, self.be1 = m.block1.groupnorm.weight.detach(), m.block1.groupnorm.bias.detach()
        self.g2, self.be2 = m.block2.groupnorm.weight.detach(), m.block2.groupnorm.bias.detach()
        self.w1, self.b1 = _pack_conv(m.block1.project.weight), m.block1.project.bias.detach()
        self.w2, self.b2 = _pack_conv(m.block2.project.weight), m.block2.project.bias.detach()
        self.wr = self.br = None
        if exists(m.res_conv):
            self.wr, self.br = _pack_conv(m.res_conv.weight, b_cols, b_scale), m.res_conv.bias.detach()
        self.gca = None
        if exists(m.gca):
            C = m.dim_out
            self.gca = dict(
                wk=m.gca.to_k.weight.detach().reshape(C).contiguous(), bk=m.gca.to_k.bias.detach(),
                w0=m.gca.net[0].weight.detach().reshape(-1, C).contiguous(), b0=m.gca.net[0].bias.detach(),
                w1=m.gca.net[2].weight.detach().reshape(C, -1).contiguous(), b1=m.gca.net[2].bias.detach(),
            )
        self.xattn = None
        if exists(m.cross_attn):
            a = m.cross_attn.fn
            self.xattn = dict(
                linear=getattr(m, "linear_cross_attn", False), heads=a.heads, scale=a.scale, norm_g=a.norm.g.detach().reshape(-1).contiguous(), null_kv=a.null_kv.detach(),
                wq=_pack_conv(a.to_q.weight), wkv=a.to_kv.weight.detach(), wo=_pack_conv(a.to_out[0].weight),
                out_g=a.to_out[1].g.detach().reshape(-1).contiguous(),
            )
        self.time_w = m.time_mlp[1].weight.detach() if exists(m.time_mlp) else None
        self.time_b = m.time_mlp[1].bias.detach() if exists(m.time_mlp) else None
        self.ss_off = None  # column offset into the all-blocks scale/shift table


class _Xf:
    """Packed TransformerBlock."""

    def __init__(self, m):
        self.layers = []
        for attn_w, ff in m.layers:
            a = attn_w.fn
            d = dict(
                heads=a.heads, scale=a.scale, norm_g=a.norm.g.detach().reshape(-1).contiguous(), null_kv=a.null_kv.detach(),
                wqkv=_bf(torch.cat((a.to_q.weight.detach(), a.to_kv.weight.detach()), 0)),
                wo=_pack_conv(a.to_out[0].weight), out_g=a.to_out[1].g.detach().reshape(-1).contiguous(),
                ctx=None,
                ff_g0=ff[0].g.detach().reshape(-1).contiguous(), ff_w1=_pack_conv(ff[1].weight),
                ff_g1=ff[3].g.detach().reshape(-1).contiguous(), ff_w2=_pack_conv(ff[4].weight),
            )
            if exists(a.to_context):
                d["ctx"] = dict(ln_w=a.to_context[0].weight.detach(), ln_b=a.to_context[0].bias.detach(),
                                w=a.to_context[1].weight.detach(), b=a.to_context[1].bias.detach())
            self.layers.append(d)


class _LinXf:
    """Packed LinearAttentionTransformerBlock: the three 1x1 convolutions of to_q / to_k / to_v become one GEMM with 3 * inner output
    channels, their depthwise 3x3 convolutions one depthwise pass over the concatenated map."""

    def __init__(self, m):
        self.layers = []
        for a, ff in m.layers:
            assert a.dim_head == 64, "linear-attention kernels are built for dim_head = 64"
            d = dict(
                heads=a.heads, scale=a.scale, norm_g=a.norm.g.detach().reshape(-1).contiguous(),
                wqkv1=_pack_conv(torch.cat([t[1].weight.detach() for t in (a.to_q, a.to_k, a.to_v)], 0)),
                wdw=torch.cat([t[2].weight.detach() for t in (a.to_q, a.to_k, a.to_v)], 0).reshape(-1, 3, 3).float().contiguous(),
                wo=_pack_conv(a.to_out[0].weight), out_g=a.to_out[1].g.detach().reshape(-1).contiguous(), ctx=None,
                ff_g0=ff[0].g.detach().reshape(-1).contiguous(), ff_w1=_pack_conv(ff[1].weight),
                ff_g1=ff[3].g.detach().reshape(-1).contiguous(), ff_w2=_pack_conv(ff[4].weight),
            )
            if exists(a.to_context):
                d["ctx"] = dict(ln_w=a.to_context[0].weight.detach(), ln_b=a.to_context[0].bias.detach(), w=a.to_context[1].weight.detach())
            self.layers.append(d)


def _pack_attn(m):
    if isinstance(m, TransformerBlock):
        return _Xf(m)
    if isinstance(m, LinearAttentionTransformerBlock):
        return _LinXf(m)
    return None


class UnetExecutor:
    def __init__(self, unet, device, stamp, precision="fp16"):
        """precision "fp16": tensor-core path (fp16 storage, fp32 accumulation).  "fp32": the precise path -- the same forward code
        over ops_f32 (fp32 weights / storage / FFMA accumulation), without the fused GroupNorm prologue and epilogue statistics."""
        assert precision in ("fp16", "fp32"), precision
        self.precise = precision == "fp32"
        self.K = ops_f32 if self.precise else ops
        self.adt = torch.float32 if self.precise else H16
        self.K.lib()  # fail loudly if the CUDA library or a B200 is missing
        self.stamp, self.device, self.u = stamp, device, unet
        _PACK_DTYPE[0] = self.adt
        try:
            self._build(unet, device)
        finally:
            _PACK_DTYPE[0] = H16

    def _build(self, unet, device):
        u = unet
        self.skip_scale = u.skip_connect_scale
        self.dim = u.init_conv.convs[0].out_channels + u.init_conv.convs[1].out_channels + u.init_conv.convs[2].out_channels
        # ---- init conv: merge the CrossEmbed kernels into one KxK matrix (K = largest kernel)
        ks = max(u.init_conv.kernel_sizes)
        self.init_ks = ks
        Cin = u.init_channels
        Wm = torch.zeros(self.dim, ks, ks, Cin, device=device)
        bias = torch.zeros(self.dim, device=device)
        o = 0
        for conv in u.init_conv.convs:
            k, co = conv.kernel_size[0], conv.out_channels
            off = (ks - k) // 2
            Wm[o:o + co, off:off + k, off:off + k, :] = conv.weight.detach().permute(0, 2, 3, 1)
            bias[o:o + co] = conv.bias.detach()
            o += co
        Cc = u.cond_images_channels
        x_idx = list(range(Cc, Cc + u.channels))
        fixed_idx = [i for i in range(Cin) if i not in x_idx]  # [cond..., lowres...]
        self.init_bias = bias

        def pack_panel_w(idx):
            K = ks * ks * len(idx)
            Kp = ((K + 63) // 64) * 64
            W = torch.zeros(self.dim, Kp, device=device)
            W[:, :K] = Wm[..., idx].reshape(self.dim, K)
            return _bf(W), Kp

        self.init_wx, self.init_kpx = pack_panel_w(x_idx)
        if self.precise:  # un-padded merged filter, K = (ky, kx, c)
            self.init_wx = Wm[..., x_idx].reshape(self.dim, -1).contiguous()

        # panel-free path (self.K.init_conv): <= 3 image channels per call, K ordered (ky, c, kx padded to 16)
        self.init_direct = self.dim in (64, 128) and not self.precise

        def pack_direct_w(idx):
            groups = []
            for g0 in range(0, len(idx), 3):
                sub = idx[g0:g0 + 3]
                Wd = torch.zeros(self.dim, ks, len(sub), 16, device=device)
                Wd[..., :ks] = Wm[..., sub].permute(0, 1, 3, 2)  # [n, ky, kx, c] -> [n, ky, c, kx]
                Kp = self.K.init_conv_kp(len(sub), ks)
                W = torch.zeros(self.dim, Kp, device=device)
                W[:, :ks * len(sub) * 16] = Wd.reshape(self.dim, -1)
                groups.append((g0, g0 + len(sub), _bf(W)))
            return groups

        if self.init_direct:
            self.init_dx = pack_direct_w(x_idx)
        # algorithmic (un-merged, un-padded) K per input channel, averaged over output channels: sum_k k^2 * cout_k / dim
        self.init_algo_k = sum(c.kernel_size[0] ** 2 * c.out_channels for c in u.init_conv.convs) / self.dim
        self.n_fixed = len(fixed_idx)
        if self.n_fixed:
            self.init_wf, self.init_kpf = pack_panel_w(fixed_idx)
            if self.precise:
                self.init_wf = Wm[..., fixed_idx].reshape(self.dim, -1).contiguous()
            if self.init_direct:
                self.init_df = pack_direct_w(fixed_idx)

        # ---- conditioning towers
        self.Tc, self.cd = u.time_cond_dim, u.cond_dim
        self.sinu_w = u.to_time_hiddens[0].weights.detach()
        self.th_w, self.th_b = u.to_time_hiddens[1].weight.detach(), u.to_time_hiddens[1].bias.detach()
        self.tok_w, self.tok_b = u.to_time_tokens[0].weight.detach(), u.to_time_tokens[0].bias.detach()
        tc_w, tc_b = [u.to_time_cond[0].weight.detach()], u.to_time_cond[0].bias.detach().clone()
        self.lowres = u.lowres_cond
        if self.lowres:
            self.lsinu_w = u.to_lowres_time_hiddens[0].weights.detach()
            self.lth_w, self.lth_b = u.to_lowres_time_hiddens[1].weight.detach(), u.to_lowres_time_hiddens[1].bias.detach()
            self.ltok_w, self.ltok_b = u.to_lowres_time_tokens[0].weight.detach(), u.to_lowres_time_tokens[0].bias.detach()
            tc_w.append(u.to_lowres_time_cond[0].weight.detach())
            tc_b += u.to_lowres_time_cond[0].bias.detach()
        self.tc_w, self.tc_b = torch.cat(tc_w, 1).contiguous(), tc_b.contiguous()  # t = [th | lowres_th] @ [Wc | Wlc]^T + (bc + blc)
        self.n_time_tokens = u.num_time_tokens * (2 if self.lowres else 1)
        self.nc_w, self.nc_b = u.norm_cond.weight.detach(), u.norm_cond.bias.detach()

        # ---- body
        s = self.skip_scale
        self.blocks = []  # every packed ResnetBlock, for the batched time-MLP

        def res(m, b_cols=0, b_scale=1.0):
            r = _Res(m, b_cols, b_scale)
            self.blocks.append(r)
            return r

        self.init_res = res(u.init_resnet_block) if exists(u.init_resnet_block) else None
        self.downs = []
        for pre, init_block, blocks, attn, post in u.downs:
            d = dict(pre=None, post=None, post_parallel=None)
            if exists(pre):
                d["pre"] = self._pack_down(pre[1])
            d["init"] = res(init_block)
            d["blocks"] = [res(b) for b in blocks]
            d["attn"] = _pack_attn(attn)
            if exists(post):
                if isinstance(post, Parallel):
                    c3, c1 = post.fns
                    w = c3.weight.detach().clone()
                    w[:, :, 1, 1] += c1.weight.detach()[:, :, 0, 0]
                    d["post_parallel"] = (_pack_conv(w), (c3.bias.detach() + c1.bias.detach()).contiguous())
                else:
                    d["post"] = self._pack_down(post[1])
            self.downs.append(d)
        self.mid1 = res(u.mid_block1)
        self.mid_attn = _Xf(u.mid_attn) if exists(u.mid_attn) else None
        self.mid2 = res(u.mid_block2)
        self.ups = []
        for init_block, blocks, attn, up in u.ups:
            skip_c = init_block.dim - init_block.dim_out
            d = dict(init=res(init_block, skip_c, s), blocks=[res(b, b.dim - b.dim_out, s) for b in blocks],
                     attn=_pack_attn(attn), up=None)
            if isinstance(up, PixelShuffleUpsample):
                conv = up.net[0]
                co4, ci = conv.weight.shape[0], conv.weight.shape[1]
                co = co4 // 4
                w = conv.weight.detach().view(co, 4, ci).permute(1, 0, 2).reshape(co4, ci)  # rows (dy*2+dx, c)
                b = conv.bias.detach().view(co, 4).t().reshape(-1).contiguous()
                d["up"] = (_bf(w), b)
            self.ups.append(d)
        self.final_res = res(u.final_res_block, 0, 1.0)
        self.final_w = u.final_conv.weight.detach().permute(0, 2, 3, 1).contiguous()  # [Cout,3,3,Cin] fp32
        self.final_b = u.final_conv.bias.detach()

        # ---- all time-MLPs as one matrix
        off = 0
        ws, bs = [], []
        for r in self.blocks:
            if r.time_w is not None:
                r.ss_off = off
                off += r.time_w.shape[0]
                ws.append(r.time_w)
                bs.append(r.time_b)
        self.ss_w, self.ss_b = torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous()

        # ---- per-sample conditioning state
        self.init_base = None
        self.lowres_img = None
        self._static = {}
        self.text_by_drop, self._text_static, self.drop = {}, {}, 0.0

    # ------------------------------------------------------------------ packing helpers
    @staticmethod
    def _pack_down(conv):
        co, c4 = conv.weight.shape[0], conv.weight.shape[1]
        c = c4 // 4
        w = conv.weight.detach().view(co, c, 2, 2)  # input channel = c*4 + dy*2 + dx
        return _bf(w.permute(0, 2, 3, 1).reshape(co, 4 * c)), conv.bias.detach()

    # ------------------------------------------------------------------ x-independent conditioning (once per sample() call)
    def set_conditioning(self, *, cond_images, lowres_cond_img, text_embeds, text_mask, cond_drop_prob, image_size):
        """Refreshes the static conditioning buffers on EVERY call (once or twice per stage of a sample() call: cheap).  There is
        deliberately no "same tensors as last time" shortcut: temporaries built by consecutive sample() calls routinely reuse
        the same device address, so pointer identity says nothing about the contents."""
        self.drop = float(cond_drop_prob)
        if exists(text_embeds) and self.u.cond_on_text:
            from .text import text_conditioning

            new = text_conditioning(self, text_embeds, text_mask, cond_drop_prob)
            skey = (self.drop,) + tuple(tuple(new[k].shape) for k in sorted(new))
            cur = self._text_static.get(skey)
            if cur is not None:
                for k in new:  # refresh in place: captured CUDA graphs of this batch size keep pointing at these buffers
                    cur[k].copy_(new[k])
            else:
                cur = self._text_static[skey] = new
            self.text_by_drop[self.drop] = cur
        else:
            self.text_by_drop.pop(self.drop, None)
        # static per-(B, S) buffers: captured CUDA graphs keep pointing at valid, refreshed conditioning
        B = (lowres_cond_img if exists(lowres_cond_img) else cond_images).shape[0] if (exists(lowres_cond_img) or exists(cond_images)) else 0
        st = self._static.setdefault((B, image_size), {})
        self.lowres_img = None
        if exists(lowres_cond_img):
            if "lowres" not in st:
                st["lowres"] = torch.empty_like(lowres_cond_img, dtype=torch.float32).contiguous()
            st["lowres"].copy_(lowres_cond_img)
            self.lowres_img = st["lowres"]
        self.init_base = None
        if self.n_fixed:
            parts = []
            if exists(cond_images):
                ci = cond_images.float()
                if ci.shape[-1] != image_size:
                    ci = F.interpolate(ci, image_size, mode="nearest")  # resize_image_to(..., mode='nearest'), x-independent
                parts.append(ci)
            if exists(lowres_cond_img):
                parts.append(self.lowres_img)
            fixed = torch.cat(parts, 1).contiguous()
            assert fixed.shape[1] == self.n_fixed
            S = fixed.shape[-1]
            if "init_base" not in st:
                st["init_base"] = torch.empty((B, S, S, self.dim), device=fixed.device, dtype=self.adt)
            self.init_base = st["init_base"]
            self._init_gemm(fixed, self.init_wf, self.init_kpf, None, None, self.init_base, getattr(self, "init_df", None))

    def _init_gemm(self, img, w, Kp, bias, addend, out, direct):
        B, _, S, S2 = img.shape
        if self.precise:
            self.K.init_conv_nchw(img, self.init_ks, w, bias, addend, out)
            return
        if self.init_direct:
            # chained <= 3-channel slices: out = conv(slice_0) + bias + addend, then out = conv(slice_i) + out
            for i, (c0, c1, wd) in enumerate(direct):
                sub = img if (c0 == 0 and c1 == img.shape[1]) else img[:, c0:c1].contiguous()
                self.K.init_conv(sub, self.init_ks, wd, bias if i == 0 else None, addend if i == 0 else out, out,
                              algo_taps=self.init_algo_k)
            return
        per = S * S2 * Kp * 2
        chunk = max(1, min(B, IM2COL_BUDGET_BYTES // per))
        for b0 in range(0, B, chunk):
            b1 = min(B, b0 + chunk)
            panel = self.K.im2col_nchw(img[b0:b1], self.init_ks, Kp)
            self.K.gemm_rows(panel, w, bias, addend=None if addend is None else addend[b0:b1].view(-1, self.dim),
                          out=out[b0:b1].view(-1, self.dim), algo_k=self.init_algo_k * img.shape[1])
            del panel

    # ------------------------------------------------------------------ blocks
    def _gn(self, xa, xb, b_scale, G, gamma, beta, ss):
        """GroupNorm(+scale/shift)+SiLU over the (virtual) concat [xa | b_scale * xb] -> activated (ya, yb)."""
        if self.precise:
            return self.K.groupnorm(xa, xb, b_scale, G, gamma, beta, ss)
        B, H, W, Ca = xa.shape
        Cb = xb.shape[3] if exists(xb) else 0
        C = Ca + Cb
        gs = C // G
        # statistics come fused from the producer kernel (conv epilogue / gate_residual) when it could emit them
        mr = self.K.gn_finalize_oct(self.K.stats_of(xa), 1.0, self.K.stats_of(xb) if exists(xb) else None, b_scale, gs, G, count=gs * H * W)
        kw = dict(group_size=gs, num_groups=G, scale_shift=ss, ctot=C)
        ya = self.K.gn_apply(xa, mr, gamma, beta, c_offset=0, **kw)
        yb = self.K.gn_apply(xb, mr, gamma, beta, c_offset=Ca, src_scale=b_scale, **kw) if exists(xb) else None
        return ya, yb

    def _cross_attn(self, P, h, c):
        B, H, W, C = h.shape
        J = c.shape[1]
        xn = self.K.layernorm_h16(h, P["norm_g"])
        q = self.K.conv_gemm(xn, P["wq"], None, ksize=1)
        kv = self.K.linear_small(c.view(B * J, -1), P["wkv"]).view(B, J, -1)
        if P["linear"]:  # LinearCrossAttention: softmax_d(q) @ (softmax_tokens(k)^T v) over the null + context tokens
            inner = P["heads"] * 64
            null_row = torch.cat((P["null_kv"][0].float().repeat(P["heads"]), P["null_kv"][1].float().repeat(P["heads"])))  # k | v, every head
            tokens = torch.cat((null_row.view(1, 1, 2 * inner).expand(B, 1, 2 * inner), kv), 1).contiguous()
            o = self.K.linear_attention(q.view(B, H * W, -1), P["heads"], P["scale"], tokens, act=self.K.ACT_NONE, pixels_kv=False)
        else:
            o = self.K.attn_cross(q.view(B, H * W, -1), kv, P["null_kv"], P["heads"], P["scale"])
        o = self.K.conv_gemm(o.view(B, H, W, -1), P["wo"], None, ksize=1)
        return self.K.layernorm_h16(o, P["out_g"], residual=h)  # to_out LayerNorm, then "+ h"

    def _gn_coef(self, xa, xb, b_scale, G, gamma, beta, ss):
        """Per-channel affine {A, B} of GroupNorm(+scale/shift) over the (virtual) concat [xa | b_scale * xb], for the
        convolution's fused pre-activation (the activated tensor is never materialised)."""
        B, H, W, Ca = xa.shape
        C = Ca + (xb.shape[3] if exists(xb) else 0)
        gs = C // G
        _, coef = self.K.gn_finalize_oct(self.K.stats_of(xa), 1.0, self.K.stats_of(xb) if exists(xb) else None, b_scale, gs, G,
                                      count=gs * H * W, gamma=gamma, beta=beta, scale_shift=ss, want_coef=True)
        return coef

    def _norm_conv(self, xa, xb, b_scale, G, gamma, beta, ss, w, bias, **kw):
        """Block.forward: conv3x3(SiLU(GroupNorm(x) * (scale + 1) + shift)); the norm is applied inside the conv kernel when
        the shape runs on the halo kernel, else by a separate gn_apply pass."""
        B, H, W, Ca = xa.shape
        Cb = xb.shape[3] if exists(xb) else 0
        if self.K.conv_pre_supported(B, H, W, Ca, Cb, w.shape[0]):
            coef = self._gn_coef(xa, xb, b_scale, G, gamma, beta, ss)
            return self.K.conv_gemm(xa, w, bias, xb=xb, ksize=3, pre_coef=coef, **kw)
        ya, yb = self._gn(xa, xb, b_scale, G, gamma, beta, ss)
        return self.K.conv_gemm(ya, w, bias, xb=yb, ksize=3, **kw)

    def _resnet(self, P, xa, xb, ss_all, c):
        h = self._norm_conv(xa, xb, P.b_scale, P.G, P.g1, P.be1, None, P.w1, P.b1, want_stats=not exists(P.xattn))
        if exists(P.xattn):
            h = self._cross_attn(P.xattn, h, c)
        ss = ss_all[:, P.ss_off:P.ss_off + 2 * P.dim_out] if P.ss_off is not None else None
        if exists(P.gca):
            g = P.gca
            h2 = self._norm_conv(h, None, 1.0, P.G, P.g2, P.be2, ss, P.w2, P.b2, logit_w=g["wk"])
            logits = getattr(h2, "_kd_logits", None)  # to_k from the conv epilogue (its bias cancels in the softmax)
            if logits is None:
                logits = self.K.rowdot(h2, g["wk"], g["bk"])
            gate = self.K.gca_gate(h2, logits, g["w0"], g["b0"], g["w1"], g["b1"])
            if exists(P.wr):  # out = res_conv(x) + gate * h2, fused in the 1x1 conv epilogue
                return self.K.conv_gemm(xa, P.wr, P.br, xb=xb, ksize=1, addend=h2, addend_scale=gate, want_stats=True)
            return self.K.gate_residual(h2, gate, xa, want_stats=True)
        if exists(P.wr):
            r = self.K.conv_gemm(xa, P.wr, P.br, xb=xb, ksize=1)
            return self._norm_conv(h, None, 1.0, P.G, P.g2, P.be2, ss, P.w2, P.b2, addend=r, want_stats=True)
        return self._norm_conv(h, None, 1.0, P.G, P.g2, P.be2, ss, P.w2, P.b2, addend=xa, want_stats=True)

    def _lin_transformer(self, P, x, c):
        """LinearAttentionTransformerBlock.forward: x = attn(x, context) + x; x = ff(x) + x."""
        B, H, W, C = x.shape
        N = H * W
        for L in P.layers:
            xn = self.K.layernorm_h16(x, L["norm_g"])
            qkv = self.K.dwconv3x3(self.K.conv_gemm(xn, L["wqkv1"], None, ksize=1), L["wdw"]).view(B, N, -1)
            ctx_kv = None
            if exists(c) and exists(L["ctx"]):
                J = c.shape[1]
                cn = self.K.layernorm_f32(c.view(B * J, -1), L["ctx"]["ln_w"], L["ctx"]["ln_b"])
                ctx_kv = self.K.linear_small(cn, L["ctx"]["w"], None).view(B, J, -1)
            o = self.K.linear_attention(qkv, L["heads"], L["scale"], ctx_kv)
            o = self.K.conv_gemm(o.view(B, H, W, -1), L["wo"], None, ksize=1)
            x = self.K.layernorm_h16(o, L["out_g"], residual=x)
            f = self.K.layernorm_h16(x, L["ff_g0"])
            f = self.K.conv_gemm(f, L["ff_w1"], None, ksize=1, act=self.K.ACT_GELU)
            f = self.K.layernorm_h16(f, L["ff_g1"])
            x = self.K.conv_gemm(f, L["ff_w2"], None, ksize=1, addend=x)
        return x

    def _transformer(self, P, x, c):
        if isinstance(P, _LinXf):
            return self._lin_transformer(P, x, c)
        B, H, W, C = x.shape
        N = H * W
        for L in P.layers:
            xn = self.K.layernorm_h16(x, L["norm_g"])
            qkv = self.K.conv_gemm(xn, L["wqkv"], None, ksize=1).view(B, N, -1)
            ctx_kv = None
            if exists(c) and exists(L["ctx"]):
                J = c.shape[1]
                cn = self.K.layernorm_f32(c.view(B * J, -1), L["ctx"]["ln_w"], L["ctx"]["ln_b"])
                ctx_kv = self.K.linear_small(cn, L["ctx"]["w"], L["ctx"]["b"]).view(B, J, -1)
            kv = self.K.kv_assemble(qkv, L["heads"] * 64, ctx_kv, L["null_kv"])
            o = self.K.attn_mqa(qkv, kv, L["heads"], L["scale"])
            o = self.K.conv_gemm(o.view(B, H, W, -1), L["wo"], None, ksize=1)
            x = self.K.layernorm_h16(o, L["out_g"], residual=x)
            f = self.K.layernorm_h16(x, L["ff_g0"])
            f = self.K.conv_gemm(f, L["ff_w1"], None, ksize=1, act=self.K.ACT_GELU)
            f = self.K.layernorm_h16(f, L["ff_g1"])
            x = self.K.conv_gemm(f, L["ff_w2"], None, ksize=1, addend=x)
        return x

    # ------------------------------------------------------------------ forward
    def forward(self, x, time, lowres_noise_times=None, taps=None, drop=None):
        u = self.u
        x = x.contiguous().float()
        B, _, S, S2 = x.shape
        dev = x.device
        Tc, cd = self.Tc, self.cd
        # --- conditioning towers (x-independent, tiny)
        nT = 2 if self.lowres else 1
        hid = torch.empty((B, nT * Tc), device=dev, dtype=torch.float32)
        J_time = self.n_time_tokens
        text = self.text_by_drop.get(self.drop if drop is None else float(drop))
        J = J_time + (text["tokens"].shape[1] if exists(text) else 0)
        c_raw = torch.empty((B, J, cd), device=dev, dtype=torch.float32)
        c_flat = c_raw.view(B, J * cd)
        self.K.linear_small(self.K.sinu_emb(time.float().contiguous(), self.sinu_w), self.th_w, self.th_b, post_act=self.K.ACT_SILU,
                         out=hid[:, :Tc], ldy=nT * Tc)
        self.K.linear_small(hid[:, :Tc], self.tok_w, self.tok_b, out=c_flat[:, : u.num_time_tokens * cd], ldy=J * cd)
        if self.lowres:
            self.K.linear_small(self.K.sinu_emb(lowres_noise_times.float().contiguous(), self.lsinu_w), self.lth_w, self.lth_b,
                             post_act=self.K.ACT_SILU, out=hid[:, Tc:], ldy=nT * Tc)
            self.K.linear_small(hid[:, Tc:], self.ltok_w, self.ltok_b,
                             out=c_flat[:, u.num_time_tokens * cd: 2 * u.num_time_tokens * cd], ldy=J * cd)
        if exists(text):
            c_raw[:, J_time:] = text["tokens"]  # x- and t-independent, prepared once per sample() call
            t = self.K.linear_small(torch.cat((hid, text["hidden_in"]), 1), text["tc_w"], text["tc_b"])
        else:
            t = self.K.linear_small(hid, self.tc_w, self.tc_b)
        c = self.K.layernorm_f32(c_raw, self.nc_w, self.nc_b)
        ss_all = self.K.linear_small(t, self.ss_w, self.ss_b, pre_act=self.K.ACT_SILU)
        if taps is not None:
            taps["t"], taps["c"] = t, c

        # --- init conv (per-step part: the 3 image channels of x; fixed part added in the epilogue)
        h = torch.empty((B, S, S2, self.dim), device=dev, dtype=self.adt)
        self._init_gemm(x, self.init_wx, self.init_kpx, self.init_bias, self.init_base, h, getattr(self, "init_dx", None))
        if taps is not None:
            taps["init_conv"] = h
        init_residual = h if u.init_conv_to_final_conv_residual else None
        if exists(self.init_res):
            h = self._resnet(self.init_res, h, None, ss_all, None)
            if taps is not None:
                taps["init_resnet_block"] = h

        hiddens = []
        for li, d in enumerate(self.downs):
            if exists(d["pre"]):
                h = self.K.conv_gemm(h, d["pre"][0], d["pre"][1], mode=1, want_stats=True)
            h = self._resnet(d["init"], h, None, ss_all, c)
            for P in d["blocks"]:
                h = self._resnet(P, h, None, ss_all, None)
                hiddens.append(h)
            if exists(d["attn"]):
                h = self._transformer(d["attn"], h, c)
            hiddens.append(h)
            if exists(d["post"]):
                h = self.K.conv_gemm(h, d["post"][0], d["post"][1], mode=1, want_stats=True)
            elif exists(d["post_parallel"]):
                h = self.K.conv_gemm(h, d["post_parallel"][0], d["post_parallel"][1], ksize=3, want_stats=True)
            if taps is not None:
                taps[f"down{li}"] = h

        h = self._resnet(self.mid1, h, None, ss_all, c)
        if taps is not None:
            taps["mid_block1"] = h
        if exists(self.mid_attn):
            h = self._transformer(self.mid_attn, h, None)
            if taps is not None:
                taps["mid_attn"] = h
        h = self._resnet(self.mid2, h, None, ss_all, c)
        if taps is not None:
            taps["mid_block2"] = h

        for li, d in enumerate(self.ups):
            h = self._resnet(d["init"], h, hiddens.pop(), ss_all, c)
            for P in d["blocks"]:
                h = self._resnet(P, h, hiddens.pop(), ss_all, None)
            if exists(d["attn"]):
                h = self._transformer(d["attn"], h, c)
            if exists(d["up"]):
                h = self.K.conv_gemm(h, d["up"][0], d["up"][1], ksize=1, act=self.K.ACT_SILU, out_mode=1)
            if taps is not None:
                taps[f"up{li}"] = h

        h = self._resnet(self.final_res, h, init_residual, ss_all, None)
        if taps is not None:
            taps["final_res_block"] = h
        return self.K.final_conv(h, self.lowres_img, self.final_w, self.final_b)
