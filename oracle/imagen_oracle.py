"""CPU oracle for the sampling hot path of jameshball/kidney-diffusion.

TEST INFRASTRUCTURE ONLY.  Nothing under ``kidney_diffusion_b200/`` imports this
file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may.

PARITY UNPINNED.  The arithmetic of the reference's hot path lives in the
un-vendored third-party dependency ``imagen-pytorch==1.18.5``
(``/root/reference/requirements.txt:37``), which is neither under
``/root/reference`` nor installable here (no network, no wheel).  This file is a
plain PyTorch fp32 restatement of the published algorithm of that package
(``imagen_pytorch/imagen_pytorch.py``: ``Unet``, ``Imagen``,
``GaussianDiffusionContinuousTimes``, ``NullUnet``), anchored on the
reference's own call sites:

* constructor contract   train_ultra_res_v_param.py:27-92, train.py:28-95,
                         train_uncond.py:28-93
* ``imagen.sample(...)``  sample_ultra_res.py:183-195, sample_cond.py:40-48,
                         sample_uncond.py:49-55

The reference ships no tests / golden vectors for this path, so
the restatement is pinned only by closed-form anchors (tests/test_oracle.py):
schedule end points, q_posterior identities, v<->x0 inversion, zero-init output.

Parameter names follow imagen-pytorch 1.18.5 so a state_dict produced here can
be loaded by ``kidney_diffusion_b200.Unet`` / ``Imagen`` unchanged.

Noise injection: every ``randn`` site of the sampler draws from
``noise_fn(site, shape)`` (sites: "lowres_aug", "init", "inpaint", "p_sample",
"renoise", each with the step / resample indices) so that the CUDA path and
this oracle consume identical tensors.
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn.functional as F
from torch import nn


# --------------------------------------------------------------------------- helpers
def exists(v):
    return v is not None


def default(v, d):
    if exists(v):
        return v
    return d() if callable(d) else d


def cast_tuple(val, length=None):
    if isinstance(val, list):
        val = tuple(val)
    out = val if isinstance(val, tuple) else ((val,) * default(length, 1))
    if exists(length):
        assert len(out) == length
    return out


def pad_tuple_to_length(t, length, fillvalue=None):
    remain = length - len(t)
    if remain <= 0:
        return t
    return (*t, *((fillvalue,) * remain))


def right_pad_dims_to(x, t):
    padding_dims = x.ndim - t.ndim
    if padding_dims <= 0:
        return t
    return t.view(*t.shape, *((1,) * padding_dims))


def resize_image_to(image, target_image_size, mode="nearest"):
    if image.shape[-1] == target_image_size:
        return image
    return F.interpolate(image, target_image_size, mode=mode)


def log(t, eps: float = 1e-12):
    return torch.log(t.clamp(min=eps))


# --------------------------------------------------------------------------- schedules (A.1)
def beta_linear_log_snr(t):
    return -torch.log(torch.expm1(1e-4 + 10 * (t ** 2)))


def alpha_cosine_log_snr(t, s: float = 0.008):
    return -log((torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2) - 1, eps=1e-5)


def log_snr_to_alpha_sigma(log_snr):
    return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))


class GaussianDiffusionContinuousTimes(nn.Module):
    def __init__(self, *, noise_schedule, timesteps=1000):
        super().__init__()
        if noise_schedule == "linear":
            self.log_snr = beta_linear_log_snr
        elif noise_schedule == "cosine":
            self.log_snr = alpha_cosine_log_snr
        else:
            raise ValueError(f"invalid noise schedule {noise_schedule}")
        self.num_timesteps = timesteps

    def get_times(self, batch_size, noise_level, *, device):
        return torch.full((batch_size,), noise_level, device=device, dtype=torch.float32)

    def get_condition(self, times):
        return self.log_snr(times) if exists(times) else None

    def get_sampling_timesteps(self, batch, *, device):
        times = torch.linspace(1.0, 0.0, self.num_timesteps + 1, device=device)
        times = times[None].expand(batch, -1)
        return [(times[:, k], times[:, k + 1]) for k in range(self.num_timesteps)]

    def q_posterior(self, x_start, x_t, t, *, t_next=None):
        t_next = default(t_next, lambda: (t - 1.0 / self.num_timesteps).clamp(min=0.0))
        log_snr = right_pad_dims_to(x_t, self.log_snr(t))
        log_snr_next = right_pad_dims_to(x_t, self.log_snr(t_next))
        alpha, sigma = log_snr_to_alpha_sigma(log_snr)
        alpha_next, sigma_next = log_snr_to_alpha_sigma(log_snr_next)
        c = -torch.expm1(log_snr - log_snr_next)
        posterior_mean = alpha_next * (x_t * (1 - c) / alpha + c * x_start)
        posterior_variance = (sigma_next ** 2) * c
        posterior_log_variance_clipped = log(posterior_variance, eps=1e-20)
        return posterior_mean, posterior_variance, posterior_log_variance_clipped

    def q_sample(self, x_start, t, noise):
        if isinstance(t, float):
            t = torch.full((x_start.shape[0],), t, device=x_start.device, dtype=x_start.dtype)
        log_snr = self.log_snr(t).type(x_start.dtype)
        alpha, sigma = log_snr_to_alpha_sigma(right_pad_dims_to(x_start, log_snr))
        return alpha * x_start + sigma * noise, log_snr, alpha, sigma

    def q_sample_from_to(self, x_from, from_t, to_t, noise):
        alpha, sigma = log_snr_to_alpha_sigma(right_pad_dims_to(x_from, self.log_snr(from_t)))
        alpha_to, sigma_to = log_snr_to_alpha_sigma(right_pad_dims_to(x_from, self.log_snr(to_t)))
        return x_from * (alpha_to / alpha) + noise * (sigma_to * alpha - sigma * alpha_to) / alpha

    def predict_start_from_v(self, x_t, t, v):
        alpha, sigma = log_snr_to_alpha_sigma(right_pad_dims_to(x_t, self.log_snr(t)))
        return alpha * x_t - sigma * v

    def predict_start_from_noise(self, x_t, t, noise):
        alpha, sigma = log_snr_to_alpha_sigma(right_pad_dims_to(x_t, self.log_snr(t)))
        return (x_t - sigma * noise) / alpha.clamp(min=1e-8)


# --------------------------------------------------------------------------- modules (A.5)
class LayerNorm(nn.Module):
    """imagen-pytorch's own LayerNorm: gain only, biased variance, eps 1e-5 in fp32."""

    def __init__(self, feats, dim=-1):
        super().__init__()
        self.dim = dim
        self.g = nn.Parameter(torch.ones(feats, *((1,) * (-dim - 1))))

    def forward(self, x):
        eps = 1e-5 if x.dtype == torch.float32 else 1e-3
        var = torch.var(x, dim=self.dim, unbiased=False, keepdim=True)
        mean = torch.mean(x, dim=self.dim, keepdim=True)
        return (x - mean) * (var + eps).rsqrt() * self.g


ChanLayerNorm = partial(LayerNorm, dim=-3)


class Always(nn.Module):
    def __init__(self, val):
        super().__init__()
        self.val = val

    def forward(self, *a, **k):
        return self.val


class Identity(nn.Module):
    def forward(self, x, *a, **k):
        return x


class Parallel(nn.Module):
    def __init__(self, *fns):
        super().__init__()
        self.fns = nn.ModuleList(fns)

    def forward(self, x):
        return sum(fn(x) for fn in self.fns)


class LearnedSinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))

    def forward(self, x):
        x = x[:, None]
        freqs = x * self.weights[None, :] * 2 * math.pi
        return torch.cat((x, freqs.sin(), freqs.cos()), dim=-1)


class CrossEmbedLayer(nn.Module):
    def __init__(self, dim_in, kernel_sizes, dim_out=None, stride=2):
        super().__init__()
        dim_out = default(dim_out, dim_in)
        kernel_sizes = sorted(kernel_sizes)
        num_scales = len(kernel_sizes)
        dim_scales = [int(dim_out / (2 ** i)) for i in range(1, num_scales)]
        dim_scales = [*dim_scales, dim_out - sum(dim_scales)]
        self.convs = nn.ModuleList(
            [nn.Conv2d(dim_in, d, k, stride=stride, padding=(k - stride) // 2) for k, d in zip(kernel_sizes, dim_scales)]
        )

    def forward(self, x):
        return torch.cat([conv(x) for conv in self.convs], dim=1)


class PixelUnshuffleRearrange(nn.Module):
    """'b c (h s1) (w s2) -> b (c s1 s2) h w' with s1 = s2 = 2."""

    def forward(self, x):
        b, c, h, w = x.shape
        x = x.view(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 3, 5, 2, 4)
        return x.reshape(b, c * 4, h // 2, w // 2)


def Downsample(dim, dim_out=None):
    dim_out = default(dim_out, dim)
    return nn.Sequential(PixelUnshuffleRearrange(), nn.Conv2d(dim * 4, dim_out, 1))


class PixelShuffleUpsample(nn.Module):
    def __init__(self, dim, dim_out=None):
        super().__init__()
        dim_out = default(dim_out, dim)
        conv = nn.Conv2d(dim, dim_out * 4, 1)
        self.net = nn.Sequential(conv, nn.SiLU(), nn.PixelShuffle(2))
        o, i, h, w = conv.weight.shape
        conv_weight = torch.empty(o // 4, i, h, w)
        nn.init.kaiming_uniform_(conv_weight)
        conv.weight.data.copy_(conv_weight.repeat_interleave(4, dim=0))
        nn.init.zeros_(conv.bias.data)

    def forward(self, x):
        return self.net(x)


class Block(nn.Module):
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.groupnorm = nn.GroupNorm(groups, dim)
        self.activation = nn.SiLU()
        self.project = nn.Conv2d(dim, dim_out, 3, padding=1)

    def forward(self, x, scale_shift=None):
        x = self.groupnorm(x)
        if exists(scale_shift):
            scale, shift = scale_shift
            x = x * (scale + 1) + shift
        x = self.activation(x)
        return self.project(x)


class GlobalContext(nn.Module):
    def __init__(self, *, dim_in, dim_out):
        super().__init__()
        self.to_k = nn.Conv2d(dim_in, 1, 1)
        hidden_dim = max(3, dim_out // 2)
        self.net = nn.Sequential(
            nn.Conv2d(dim_in, hidden_dim, 1), nn.SiLU(), nn.Conv2d(hidden_dim, dim_out, 1), nn.Sigmoid()
        )

    def forward(self, x):
        context = self.to_k(x)
        b, c = x.shape[:2]
        xf, context = x.reshape(b, c, -1), context.reshape(b, 1, -1)
        out = torch.einsum("b i n, b c n -> b c i", context.softmax(dim=-1), xf)
        return self.net(out[..., None])


class CrossAttention(nn.Module):
    def __init__(self, dim, *, context_dim=None, dim_head=64, heads=8):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        inner_dim = dim_head * heads
        context_dim = default(context_dim, dim)
        self.norm = LayerNorm(dim)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(context_dim, inner_dim * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim, bias=False), LayerNorm(dim))

    def forward(self, x, context):
        b, n, h = x.shape[0], x.shape[1], self.heads
        x = self.norm(x)
        q = self.to_q(x)
        k, v = self.to_kv(context).chunk(2, dim=-1)
        q, k, v = (t.view(b, t.shape[1], h, -1).transpose(1, 2) for t in (q, k, v))
        nk, nv = (t.view(1, 1, 1, -1).expand(b, h, 1, -1) for t in self.null_kv.unbind(dim=-2))
        k = torch.cat((nk, k), dim=-2)
        v = torch.cat((nv, v), dim=-2)
        q = q * self.scale
        sim = torch.einsum("b h i d, b h j d -> b h i j", q, k)
        attn = sim.softmax(dim=-1, dtype=torch.float32)
        out = torch.einsum("b h i j, b h j d -> b h i d", attn, v)
        out = out.transpose(1, 2).reshape(b, n, -1)
        return self.to_out(out)


class LinearCrossAttention(CrossAttention):
    """imagen_pytorch.LinearCrossAttention: same parameters as CrossAttention; softmax of q over the head dimension and of k over the
    (null + context) tokens, context = k^T v, out = q context."""

    def forward(self, x, context):
        b, n, h = x.shape[0], x.shape[1], self.heads
        x = self.norm(x)
        q = self.to_q(x)
        k, v = self.to_kv(context).chunk(2, dim=-1)
        q, k, v = (t.view(b, t.shape[1], h, -1).transpose(1, 2).reshape(b * h, t.shape[1], -1) for t in (q, k, v))
        nk, nv = (t.view(1, 1, -1).expand(b * h, 1, -1) for t in self.null_kv.unbind(dim=-2))
        k = torch.cat((nk, k), dim=-2)
        v = torch.cat((nv, v), dim=-2)
        q = q.softmax(dim=-1)
        k = k.softmax(dim=-2)
        q = q * self.scale
        ctx = torch.einsum("b n d, b n e -> b d e", k, v)
        out = torch.einsum("b n d, b d e -> b n e", q, ctx)
        out = out.view(b, h, n, -1).transpose(1, 2).reshape(b, n, -1)
        return self.to_out(out)


class Attention(nn.Module):
    """Multi-query self attention: one shared K/V head."""

    def __init__(self, dim, *, dim_head=64, heads=8, context_dim=None):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        inner_dim = dim_head * heads
        self.norm = LayerNorm(dim)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, dim_head * 2, bias=False)
        self.to_context = (
            nn.Sequential(nn.LayerNorm(context_dim), nn.Linear(context_dim, dim_head * 2)) if exists(context_dim) else None
        )
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim, bias=False), LayerNorm(dim))

    def forward(self, x, context=None):
        b, n = x.shape[:2]
        x = self.norm(x)
        q = self.to_q(x)
        k, v = self.to_kv(x).chunk(2, dim=-1)
        q = q.view(b, n, self.heads, -1).transpose(1, 2) * self.scale
        nk, nv = (t.view(1, 1, -1).expand(b, 1, -1) for t in self.null_kv.unbind(dim=-2))
        k = torch.cat((nk, k), dim=-2)
        v = torch.cat((nv, v), dim=-2)
        if exists(context):
            assert exists(self.to_context)
            ck, cv = self.to_context(context).chunk(2, dim=-1)
            k = torch.cat((ck, k), dim=-2)
            v = torch.cat((cv, v), dim=-2)
        sim = torch.einsum("b h i d, b j d -> b h i j", q, k)
        attn = sim.softmax(dim=-1, dtype=torch.float32)
        out = torch.einsum("b h i j, b j d -> b h i d", attn, v)
        out = out.transpose(1, 2).reshape(b, n, -1)
        return self.to_out(out)


def FeedForward(dim, mult=2):
    hidden_dim = int(dim * mult)
    return nn.Sequential(
        LayerNorm(dim), nn.Linear(dim, hidden_dim, bias=False), nn.GELU(), LayerNorm(hidden_dim),
        nn.Linear(hidden_dim, dim, bias=False),
    )


def ChanFeedForward(dim, mult=2):
    hidden_dim = int(dim * mult)
    return nn.Sequential(
        ChanLayerNorm(dim), nn.Conv2d(dim, hidden_dim, 1, bias=False), nn.GELU(), ChanLayerNorm(hidden_dim),
        nn.Conv2d(hidden_dim, dim, 1, bias=False),
    )


class EinopsToAndFrom(nn.Module):
    """'b c h w' <-> 'b (h w) c' wrapper; keeps the ``.fn`` key prefix of 1.18.x."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, **kwargs):
        b, c, h, w = x.shape
        y = self.fn(x.flatten(2).transpose(1, 2), **kwargs)
        return y.transpose(1, 2).reshape(b, c, h, w)


class TransformerBlock(nn.Module):
    def __init__(self, dim, *, depth=1, heads=8, dim_head=32, ff_mult=2, context_dim=None):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(
                nn.ModuleList(
                    [
                        EinopsToAndFrom(Attention(dim=dim, heads=heads, dim_head=dim_head, context_dim=context_dim)),
                        ChanFeedForward(dim=dim, mult=ff_mult),
                    ]
                )
            )

    def forward(self, x, context=None):
        for attn, ff in self.layers:
            x = attn(x, context=context) + x
            x = ff(x) + x
        return x


class LinearAttention(nn.Module):
    """imagen_pytorch.LinearAttention (1.18.x): q / k / v = 1x1 conv + depthwise 3x3 conv of the channel-normalised map; softmax of
    q over the head dimension and of k over positions (pixels + optional context tokens), context = k^T v, out = q context,
    SiLU, 1x1 conv + ChanLayerNorm.  Dropout(0.05) of the to_q / to_k / to_v stacks is inactive at sampling time."""

    def __init__(self, dim, dim_head=32, heads=8, dropout=0.05, context_dim=None, **kwargs):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        inner_dim = dim_head * heads
        self.norm = ChanLayerNorm(dim)
        self.nonlin = nn.SiLU()

        def qkv():
            return nn.Sequential(nn.Dropout(dropout), nn.Conv2d(dim, inner_dim, 1, bias=False),
                                 nn.Conv2d(inner_dim, inner_dim, 3, bias=False, padding=1, groups=inner_dim))

        self.to_q, self.to_k, self.to_v = qkv(), qkv(), qkv()
        self.to_context = nn.Sequential(nn.LayerNorm(context_dim), nn.Linear(context_dim, inner_dim * 2, bias=False)) if exists(context_dim) else None
        self.to_out = nn.Sequential(nn.Conv2d(inner_dim, dim, 1, bias=False), ChanLayerNorm(dim))

    def forward(self, fmap, context=None):
        h, x, y = self.heads, *fmap.shape[-2:]
        b = fmap.shape[0]
        fmap = self.norm(fmap)
        q, k, v = (fn(fmap) for fn in (self.to_q, self.to_k, self.to_v))
        q, k, v = (t.reshape(b, h, -1, x * y).transpose(-1, -2).reshape(b * h, x * y, -1) for t in (q, k, v))  # 'b (h c) x y -> (b h) (x y) c'
        if exists(context):
            assert exists(self.to_context)
            ck, cv = self.to_context(context).chunk(2, dim=-1)
            ck, cv = (t.reshape(b, t.shape[1], h, -1).transpose(1, 2).reshape(b * h, t.shape[1], -1) for t in (ck, cv))  # 'b n (h d) -> (b h) n d'
            k = torch.cat((k, ck), dim=-2)
            v = torch.cat((v, cv), dim=-2)
        q = q.softmax(dim=-1)
        k = k.softmax(dim=-2)
        q = q * self.scale
        context = torch.einsum("b n d, b n e -> b d e", k, v)
        out = torch.einsum("b n d, b d e -> b n e", q, context)
        out = out.reshape(b, h, x, y, -1).permute(0, 1, 4, 2, 3).reshape(b, -1, x, y)  # '(b h) (x y) d -> b (h d) x y'
        out = self.nonlin(out)
        return self.to_out(out)


class LinearAttentionTransformerBlock(nn.Module):
    def __init__(self, dim, *, depth=1, heads=8, dim_head=32, ff_mult=2, context_dim=None, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([LinearAttention(dim=dim, heads=heads, dim_head=dim_head, context_dim=context_dim),
                                              ChanFeedForward(dim=dim, mult=ff_mult)]))

    def forward(self, x, context=None):
        for attn, ff in self.layers:
            x = attn(x, context=context) + x
            x = ff(x) + x
        return x


class ResnetBlock(nn.Module):
    def __init__(self, dim, dim_out, *, cond_dim=None, time_cond_dim=None, groups=8, linear_attn=False, use_gca=False, **attn_kwargs):
        super().__init__()
        self.time_mlp = None
        if exists(time_cond_dim):
            self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_cond_dim, dim_out * 2))
        self.cross_attn = None
        if exists(cond_dim):
            attn_klass = CrossAttention if not linear_attn else LinearCrossAttention
            self.cross_attn = EinopsToAndFrom(attn_klass(dim=dim_out, context_dim=cond_dim, **attn_kwargs))
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.gca = GlobalContext(dim_in=dim_out, dim_out=dim_out) if use_gca else Always(1)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else Identity()

    def forward(self, x, time_emb=None, cond=None):
        scale_shift = None
        if exists(self.time_mlp) and exists(time_emb):
            time_emb = self.time_mlp(time_emb)[:, :, None, None]
            scale_shift = time_emb.chunk(2, dim=1)
        h = self.block1(x)
        if exists(self.cross_attn):
            assert exists(cond)
            h = self.cross_attn(h, context=cond) + h
        h = self.block2(h, scale_shift=scale_shift)
        h = h * self.gca(h)
        return h + self.res_conv(x)


class PerceiverAttention(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        inner_dim = dim_head * heads
        self.norm = nn.LayerNorm(dim)
        self.norm_latents = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim, bias=False), nn.LayerNorm(dim))

    def forward(self, x, latents):
        x = self.norm(x)
        latents = self.norm_latents(latents)
        b, h = x.shape[0], self.heads
        q = self.to_q(latents)
        kv_input = torch.cat((x, latents), dim=-2)
        k, v = self.to_kv(kv_input).chunk(2, dim=-1)
        q, k, v = (t.view(b, t.shape[1], h, -1).transpose(1, 2) for t in (q, k, v))
        q = q * self.scale
        sim = torch.einsum("... i d, ... j d  -> ... i j", q, k)
        attn = sim.softmax(dim=-1, dtype=torch.float32)
        out = torch.einsum("... i j, ... j d -> ... i d", attn, v)
        out = out.transpose(1, 2).reshape(b, latents.shape[1], -1)
        return self.to_out(out)


class _MeanPooledLatents(nn.Module):
    """Sequential(LayerNorm(dim), Linear(dim, dim*n), Rearrange('b (n d) -> b n d'))."""

    def __init__(self, dim, n):
        super().__init__()
        self.n = n
        self.add_module("0", LayerNorm(dim))
        self.add_module("1", nn.Linear(dim, dim * n))

    def forward(self, x):
        x = getattr(self, "1")(getattr(self, "0")(x))
        return x.view(x.shape[0], self.n, -1)


class PerceiverResampler(nn.Module):
    def __init__(self, *, dim, depth, dim_head=64, heads=8, num_latents=64, num_latents_mean_pooled=4,
                 max_seq_len=512, ff_mult=4):
        super().__init__()
        self.pos_emb = nn.Embedding(max_seq_len, dim)
        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.to_latents_from_mean_pooled_seq = (
            _MeanPooledLatents(dim, num_latents_mean_pooled) if num_latents_mean_pooled > 0 else None
        )
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(
                nn.ModuleList([PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads), FeedForward(dim=dim, mult=ff_mult)])
            )

    def forward(self, x):
        n = x.shape[1]
        pos_emb = self.pos_emb(torch.arange(n, device=x.device))
        x_with_pos = x + pos_emb
        latents = self.latents[None].expand(x.shape[0], -1, -1)
        if exists(self.to_latents_from_mean_pooled_seq):
            meanpooled_seq = x.mean(dim=1)
            latents = torch.cat((self.to_latents_from_mean_pooled_seq(meanpooled_seq), latents), dim=-2)
        for attn, ff in self.layers:
            latents = attn(x_with_pos, latents) + latents
            latents = ff(latents) + latents
        return latents


class _ToTokens(nn.Module):
    """Sequential(Linear(Tc, cond_dim*r), Rearrange('b (r d) -> b r d'))."""

    def __init__(self, time_cond_dim, cond_dim, r):
        super().__init__()
        self.r = r
        self.add_module("0", nn.Linear(time_cond_dim, cond_dim * r))

    def forward(self, x):
        x = getattr(self, "0")(x)
        return x.view(x.shape[0], self.r, -1)


# --------------------------------------------------------------------------- Unet (A.4, A.6, A.8)
class Unet(nn.Module):
    def __init__(
        self, *, dim, text_embed_dim=768, num_resnet_blocks=1, cond_dim=None, num_image_tokens=4, num_time_tokens=2,
        learned_sinu_pos_emb_dim=16, dim_mults=(1, 2, 4, 8), cond_images_channels=0, channels=3, channels_out=None,
        attn_dim_head=64, attn_heads=8, ff_mult=2.0, lowres_cond=False, layer_attns=True, layer_attns_depth=1,
        layer_mid_attns_depth=1, attend_at_middle=True, layer_cross_attns=True, cond_on_text=True, max_text_len=256,
        init_dim=None, resnet_groups=8, init_cross_embed_kernel_sizes=(3, 7, 15), attn_pool_text=True,
        attn_pool_num_latents=32, memory_efficient=False, init_conv_to_final_conv_residual=False,
        use_global_context_attn=True, scale_skip_connection=True, final_resnet_block=True, final_conv_kernel_size=3,
        use_linear_attn=False, use_linear_cross_attn=False,
    ):
        super().__init__()
        self._locals = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        self.channels = channels
        self.channels_out = default(channels_out, channels)
        init_channels = channels * (1 + int(lowres_cond))
        init_dim = default(init_dim, dim)
        self.has_cond_image = cond_images_channels > 0
        self.cond_images_channels = cond_images_channels
        init_channels += cond_images_channels
        self.init_conv = CrossEmbedLayer(init_channels, dim_out=init_dim, kernel_sizes=init_cross_embed_kernel_sizes, stride=1)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        cond_dim = default(cond_dim, dim)
        time_cond_dim = dim * 4 * (2 if lowres_cond else 1)

        self.to_time_hiddens = nn.Sequential(
            LearnedSinusoidalPosEmb(learned_sinu_pos_emb_dim), nn.Linear(learned_sinu_pos_emb_dim + 1, time_cond_dim), nn.SiLU()
        )
        self.to_time_cond = nn.Sequential(nn.Linear(time_cond_dim, time_cond_dim))
        self.to_time_tokens = _ToTokens(time_cond_dim, cond_dim, num_time_tokens)
        self.lowres_cond = lowres_cond
        if lowres_cond:
            self.to_lowres_time_hiddens = nn.Sequential(
                LearnedSinusoidalPosEmb(learned_sinu_pos_emb_dim), nn.Linear(learned_sinu_pos_emb_dim + 1, time_cond_dim), nn.SiLU()
            )
            self.to_lowres_time_cond = nn.Sequential(nn.Linear(time_cond_dim, time_cond_dim))
            self.to_lowres_time_tokens = _ToTokens(time_cond_dim, cond_dim, num_time_tokens)
        self.norm_cond = nn.LayerNorm(cond_dim)

        self.text_to_cond = None
        if cond_on_text:
            self.text_to_cond = nn.Linear(text_embed_dim, cond_dim)
        self.cond_on_text = cond_on_text
        self.attn_pool = (
            PerceiverResampler(dim=cond_dim, depth=2, dim_head=attn_dim_head, heads=attn_heads, num_latents=attn_pool_num_latents)
            if attn_pool_text else None
        )
        self.max_text_len = max_text_len
        self.null_text_embed = nn.Parameter(torch.randn(1, max_text_len, cond_dim))
        self.null_text_hidden = nn.Parameter(torch.randn(1, time_cond_dim))
        self.to_text_non_attn_cond = None
        if cond_on_text:
            self.to_text_non_attn_cond = nn.Sequential(
                nn.LayerNorm(cond_dim), nn.Linear(cond_dim, time_cond_dim), nn.SiLU(), nn.Linear(time_cond_dim, time_cond_dim)
            )

        attn_kwargs = dict(heads=attn_heads, dim_head=attn_dim_head)
        num_layers = len(in_out)
        num_resnet_blocks = cast_tuple(num_resnet_blocks, num_layers)
        resnet_groups = cast_tuple(resnet_groups, num_layers)
        layer_attns = cast_tuple(layer_attns, num_layers)
        layer_attns_depth = cast_tuple(layer_attns_depth, num_layers)
        layer_cross_attns = cast_tuple(layer_cross_attns, num_layers)

        self.init_resnet_block = (
            ResnetBlock(init_dim, init_dim, time_cond_dim=time_cond_dim, groups=resnet_groups[0], use_gca=use_global_context_attn, **attn_kwargs)
            if memory_efficient else None
        )
        self.skip_connect_scale = 1.0 if not scale_skip_connection else (2 ** -0.5)

        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        use_linear_attn = cast_tuple(use_linear_attn, num_layers)
        use_linear_cross_attn = cast_tuple(use_linear_cross_attn, num_layers)
        layer_params = [num_resnet_blocks, resnet_groups, layer_attns, layer_attns_depth, layer_cross_attns, use_linear_attn, use_linear_cross_attn]
        reversed_layer_params = [tuple(reversed(p)) for p in layer_params]
        skip_connect_dims = []

        for ind, ((dim_in, dim_out), n_blocks, groups, layer_attn, attn_depth, layer_cross_attn, layer_lin, layer_lin_cross) in enumerate(zip(in_out, *layer_params)):
            is_last = ind >= (num_layers - 1)
            layer_cond_dim = cond_dim if layer_cross_attn else None
            current_dim = dim_in
            pre_downsample = None
            if memory_efficient:
                pre_downsample = Downsample(dim_in, dim_out)
                current_dim = dim_out
            skip_connect_dims.append(current_dim)
            post_downsample = None
            if not memory_efficient:
                post_downsample = (
                    Downsample(current_dim, dim_out) if not is_last
                    else Parallel(nn.Conv2d(dim_in, dim_out, 3, padding=1), nn.Conv2d(dim_in, dim_out, 1))
                )
            self.downs.append(
                nn.ModuleList(
                    [
                        pre_downsample,
                        ResnetBlock(current_dim, current_dim, cond_dim=layer_cond_dim, linear_attn=layer_lin_cross, time_cond_dim=time_cond_dim, groups=groups, **attn_kwargs),
                        nn.ModuleList(
                            [ResnetBlock(current_dim, current_dim, time_cond_dim=time_cond_dim, groups=groups, use_gca=use_global_context_attn)
                             for _ in range(n_blocks)]
                        ),
                        TransformerBlock(dim=current_dim, depth=attn_depth, ff_mult=ff_mult, context_dim=cond_dim, **attn_kwargs)
                        if layer_attn else (LinearAttentionTransformerBlock(dim=current_dim, depth=attn_depth, ff_mult=ff_mult, context_dim=cond_dim, **attn_kwargs)
                                            if layer_lin else Identity()),
                        post_downsample,
                    ]
                )
            )

        mid_dim = dims[-1]
        self.mid_block1 = ResnetBlock(mid_dim, mid_dim, cond_dim=cond_dim, time_cond_dim=time_cond_dim, groups=resnet_groups[-1], **attn_kwargs)
        self.mid_attn = TransformerBlock(mid_dim, depth=layer_mid_attns_depth, **attn_kwargs) if attend_at_middle else None
        self.mid_block2 = ResnetBlock(mid_dim, mid_dim, cond_dim=cond_dim, time_cond_dim=time_cond_dim, groups=resnet_groups[-1], **attn_kwargs)

        for ind, ((dim_in, dim_out), n_blocks, groups, layer_attn, attn_depth, layer_cross_attn, layer_lin, layer_lin_cross) in enumerate(
            zip(reversed(in_out), *reversed_layer_params)
        ):
            is_last = ind == (len(in_out) - 1)
            layer_cond_dim = cond_dim if layer_cross_attn else None
            skip_connect_dim = skip_connect_dims.pop()
            self.ups.append(
                nn.ModuleList(
                    [
                        ResnetBlock(dim_out + skip_connect_dim, dim_out, cond_dim=layer_cond_dim, linear_attn=layer_lin_cross, time_cond_dim=time_cond_dim, groups=groups, **attn_kwargs),
                        nn.ModuleList(
                            [ResnetBlock(dim_out + skip_connect_dim, dim_out, time_cond_dim=time_cond_dim, groups=groups, use_gca=use_global_context_attn)
                             for _ in range(n_blocks)]
                        ),
                        TransformerBlock(dim=dim_out, depth=attn_depth, ff_mult=ff_mult, context_dim=cond_dim, **attn_kwargs)
                        if layer_attn else (LinearAttentionTransformerBlock(dim=dim_out, depth=attn_depth, ff_mult=ff_mult, context_dim=cond_dim, **attn_kwargs)
                                            if layer_lin else Identity()),
                        PixelShuffleUpsample(dim_out, dim_in) if not is_last or memory_efficient else Identity(),
                    ]
                )
            )

        self.init_conv_to_final_conv_residual = init_conv_to_final_conv_residual
        final_conv_dim = dim + (dim if init_conv_to_final_conv_residual else 0)
        self.final_res_block = (
            ResnetBlock(final_conv_dim, dim, time_cond_dim=time_cond_dim, groups=resnet_groups[0], use_gca=True)
            if final_resnet_block else None
        )
        final_conv_dim_in = dim if final_resnet_block else final_conv_dim
        final_conv_dim_in += channels if lowres_cond else 0
        self.final_conv = nn.Conv2d(final_conv_dim_in, self.channels_out, final_conv_kernel_size, padding=final_conv_kernel_size // 2)
        nn.init.zeros_(self.final_conv.weight)
        nn.init.zeros_(self.final_conv.bias)

    def cast_model_parameters(self, *, lowres_cond, text_embed_dim, channels, channels_out, cond_on_text):
        if (
            lowres_cond == self.lowres_cond and channels == self.channels and cond_on_text == self.cond_on_text
            and text_embed_dim == self._locals["text_embed_dim"] and channels_out == self.channels_out
        ):
            return self
        updated = dict(lowres_cond=lowres_cond, text_embed_dim=text_embed_dim, channels=channels, channels_out=channels_out,
                       cond_on_text=cond_on_text)
        return self.__class__(**{**self._locals, **updated})

    def forward_with_cond_scale(self, *args, cond_scale=1.0, **kwargs):
        logits = self.forward(*args, **kwargs)
        if cond_scale == 1:
            return logits
        null_logits = self.forward(*args, cond_drop_prob=1.0, **kwargs)
        return null_logits + (logits - null_logits) * cond_scale

    def forward(self, x, time, *, lowres_cond_img=None, lowres_noise_times=None, text_embeds=None, text_mask=None,
                cond_images=None, cond_drop_prob=0.0, taps=None):
        batch_size = x.shape[0]
        assert not (self.lowres_cond and not exists(lowres_cond_img))
        assert not (self.lowres_cond and not exists(lowres_noise_times))
        if exists(lowres_cond_img):
            x = torch.cat((x, lowres_cond_img), dim=1)
        assert not (self.has_cond_image ^ exists(cond_images))
        if exists(cond_images):
            assert cond_images.shape[1] == self.cond_images_channels
            cond_images = resize_image_to(cond_images, x.shape[-1], mode="nearest")
            x = torch.cat((cond_images, x), dim=1)

        x = self.init_conv(x)
        if exists(taps):
            taps["init_conv"] = x
        if self.init_conv_to_final_conv_residual:
            init_conv_residual = x.clone()

        time_hiddens = self.to_time_hiddens(time)
        time_tokens = self.to_time_tokens(time_hiddens)
        t = self.to_time_cond(time_hiddens)
        if self.lowres_cond:
            lowres_time_hiddens = self.to_lowres_time_hiddens(lowres_noise_times)
            lowres_time_tokens = self.to_lowres_time_tokens(lowres_time_hiddens)
            lowres_t = self.to_lowres_time_cond(lowres_time_hiddens)
            t = t + lowres_t
            time_tokens = torch.cat((time_tokens, lowres_time_tokens), dim=-2)

        text_tokens = None
        if exists(text_embeds) and self.cond_on_text:
            keep = torch.full((batch_size,), cond_drop_prob < 1.0, dtype=torch.bool, device=x.device)
            if 0.0 < cond_drop_prob < 1.0:
                keep = torch.zeros((batch_size,), device=x.device).uniform_(0, 1) < (1 - cond_drop_prob)
            text_keep_mask_embed = keep[:, None, None]
            text_keep_mask_hidden = keep[:, None]
            text_tokens = self.text_to_cond(text_embeds)[:, : self.max_text_len]
            if exists(text_mask):
                text_mask = text_mask[:, : self.max_text_len]
            remainder = self.max_text_len - text_tokens.shape[1]
            if remainder > 0:
                text_tokens = F.pad(text_tokens, (0, 0, 0, remainder))
            if exists(text_mask):
                if remainder > 0:
                    text_mask = F.pad(text_mask, (0, remainder), value=False)
                text_keep_mask_embed = text_mask[:, :, None] & text_keep_mask_embed
            text_tokens = torch.where(text_keep_mask_embed, text_tokens, self.null_text_embed)
            if exists(self.attn_pool):
                text_tokens = self.attn_pool(text_tokens)
            mean_pooled_text_tokens = text_tokens.mean(dim=-2)
            text_hiddens = self.to_text_non_attn_cond(mean_pooled_text_tokens)
            text_hiddens = torch.where(text_keep_mask_hidden, text_hiddens, self.null_text_hidden)
            t = t + text_hiddens

        c = time_tokens if not exists(text_tokens) else torch.cat((time_tokens, text_tokens), dim=-2)
        c = self.norm_cond(c)
        if exists(taps):
            taps["t"], taps["c"] = t, c

        if exists(self.init_resnet_block):
            x = self.init_resnet_block(x, t)
            if exists(taps):
                taps["init_resnet_block"] = x

        hiddens = []
        for li, (pre_downsample, init_block, resnet_blocks, attn_block, post_downsample) in enumerate(self.downs):
            if exists(pre_downsample):
                x = pre_downsample(x)
            x = init_block(x, t, c)
            for resnet_block in resnet_blocks:
                x = resnet_block(x, t)
                hiddens.append(x)
            x = attn_block(x, c) if not isinstance(attn_block, Identity) else x
            hiddens.append(x)
            if exists(post_downsample):
                x = post_downsample(x)
            if exists(taps):
                taps[f"down{li}"] = x

        x = self.mid_block1(x, t, c)
        if exists(taps):
            taps["mid_block1"] = x
        if exists(self.mid_attn):
            x = self.mid_attn(x)
            if exists(taps):
                taps["mid_attn"] = x
        x = self.mid_block2(x, t, c)
        if exists(taps):
            taps["mid_block2"] = x

        def add_skip_connection(x):
            return torch.cat((x, hiddens.pop() * self.skip_connect_scale), dim=1)

        for li, (init_block, resnet_blocks, attn_block, upsample) in enumerate(self.ups):
            x = add_skip_connection(x)
            x = init_block(x, t, c)
            for resnet_block in resnet_blocks:
                x = add_skip_connection(x)
                x = resnet_block(x, t)
            x = attn_block(x, c) if not isinstance(attn_block, Identity) else x
            x = upsample(x)
            if exists(taps):
                taps[f"up{li}"] = x

        if self.init_conv_to_final_conv_residual:
            x = torch.cat((x, init_conv_residual), dim=1)
        if exists(self.final_res_block):
            x = self.final_res_block(x, t)
            if exists(taps):
                taps["final_res_block"] = x
        if exists(lowres_cond_img):
            x = torch.cat((x, lowres_cond_img), dim=1)
        return self.final_conv(x)


class NullUnet(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.lowres_cond = False
        self.dummy_parameter = nn.Parameter(torch.tensor([0.0]))

    def cast_model_parameters(self, *args, **kwargs):
        return self

    def forward(self, x, *args, **kwargs):
        return x


# --------------------------------------------------------------------------- Imagen (A.2, A.3)
def default_noise_fn(site, shape, **key):
    return torch.randn(shape)


class Imagen(nn.Module):
    def __init__(
        self, unets, *, image_sizes, text_embed_dim=None, channels=3, timesteps=1000, cond_drop_prob=0.1,
        noise_schedules="cosine", pred_objectives="noise", random_crop_sizes=None, lowres_noise_schedule="linear",
        lowres_sample_noise_level=0.2, condition_on_text=True, auto_normalize_img=True, dynamic_thresholding=True,
        dynamic_thresholding_percentile=0.95,
    ):
        super().__init__()
        self.condition_on_text = condition_on_text
        self.unconditional = not condition_on_text
        self.channels = channels
        unets = cast_tuple(unets)
        num_unets = len(unets)
        timesteps = cast_tuple(timesteps, num_unets)
        noise_schedules = cast_tuple(noise_schedules)
        noise_schedules = pad_tuple_to_length(noise_schedules, 2, "cosine")
        noise_schedules = pad_tuple_to_length(noise_schedules, num_unets, "linear")
        self.noise_schedulers = nn.ModuleList(
            [GaussianDiffusionContinuousTimes(noise_schedule=s, timesteps=t) for t, s in zip(timesteps, noise_schedules)]
        )
        self.lowres_noise_schedule = GaussianDiffusionContinuousTimes(noise_schedule=lowres_noise_schedule)
        self.pred_objectives = cast_tuple(pred_objectives, num_unets)
        self.text_embed_dim = default(text_embed_dim, 768)
        self.unets = nn.ModuleList([])
        for ind, one_unet in enumerate(unets):
            is_first = ind == 0
            one_unet = one_unet.cast_model_parameters(
                lowres_cond=not is_first, cond_on_text=self.condition_on_text, text_embed_dim=self.text_embed_dim if self.condition_on_text else None,
                channels=self.channels, channels_out=self.channels,
            )
            self.unets.append(one_unet)
        self.image_sizes = cast_tuple(image_sizes)
        assert num_unets == len(self.image_sizes)
        self.sample_channels = cast_tuple(self.channels, num_unets)
        lowres_conditions = tuple(u.lowres_cond for u in self.unets)
        assert lowres_conditions == (False, *((True,) * (num_unets - 1)))
        self.lowres_sample_noise_level = lowres_sample_noise_level
        self.cond_drop_prob = cond_drop_prob
        self.can_classifier_guidance = cond_drop_prob > 0.0
        self.normalize_img = (lambda img: img * 2 - 1) if auto_normalize_img else (lambda img: img)
        self.unnormalize_img = (lambda img: (img + 1) * 0.5) if auto_normalize_img else (lambda img: img)
        self.dynamic_thresholding = cast_tuple(dynamic_thresholding, num_unets)
        self.dynamic_thresholding_percentile = dynamic_thresholding_percentile

    def p_mean_variance(self, unet, x, t, *, noise_scheduler, text_embeds=None, text_mask=None, cond_images=None,
                        lowres_cond_img=None, lowres_noise_times=None, cond_scale=1.0, t_next=None, pred_objective="noise",
                        dynamic_threshold=True, taps=None):
        pred = unet.forward_with_cond_scale(
            x, noise_scheduler.get_condition(t), text_embeds=text_embeds, text_mask=text_mask, cond_images=cond_images,
            cond_scale=cond_scale, lowres_cond_img=lowres_cond_img,
            lowres_noise_times=self.lowres_noise_schedule.get_condition(lowres_noise_times),
        )
        if exists(taps):
            taps["pred"] = pred
        if pred_objective == "noise":
            x_start = noise_scheduler.predict_start_from_noise(x, t=t, noise=pred)
        elif pred_objective == "x_start":
            x_start = pred
        elif pred_objective == "v":
            x_start = noise_scheduler.predict_start_from_v(x, t=t, v=pred)
        else:
            raise ValueError(pred_objective)
        if dynamic_threshold:
            s = torch.quantile(x_start.flatten(1).abs(), self.dynamic_thresholding_percentile, dim=-1)
            s.clamp_(min=1.0)
            s = right_pad_dims_to(x_start, s)
            x_start = x_start.clamp(-s, s) / s
        else:
            x_start = x_start.clamp(-1.0, 1.0)
        if exists(taps):
            taps["x_start"] = x_start
        return noise_scheduler.q_posterior(x_start=x_start, x_t=x, t=t, t_next=t_next), x_start

    def p_sample(self, unet, x, t, *, noise, noise_scheduler, t_next=None, taps=None, **kwargs):
        b = x.shape[0]
        (model_mean, _, model_log_variance), x_start = self.p_mean_variance(
            unet, x=x, t=t, t_next=t_next, noise_scheduler=noise_scheduler, taps=taps, **kwargs
        )
        is_last_sampling_timestep = t_next == 0
        nonzero_mask = (1 - is_last_sampling_timestep.float()).reshape(b, *((1,) * (x.ndim - 1)))
        pred = model_mean + nonzero_mask * (0.5 * model_log_variance).exp() * noise
        return pred, x_start

    def p_sample_loop(self, unet, shape, *, noise_scheduler, noise_fn, unet_number, lowres_cond_img=None, lowres_noise_times=None,
                      text_embeds=None, text_mask=None, cond_images=None, inpaint_images=None, inpaint_masks=None,
                      inpaint_resample_times=5, cond_scale=1.0, pred_objective="noise", dynamic_threshold=True, step_taps=None):
        batch = shape[0]
        img = noise_fn("init", shape, unet=unet_number)
        has_inpainting = exists(inpaint_images) and exists(inpaint_masks)
        resample_times = inpaint_resample_times if has_inpainting else 1
        if has_inpainting:
            inpaint_images = self.normalize_img(inpaint_images)
            inpaint_images = resize_image_to(inpaint_images, shape[-1])
            inpaint_masks = resize_image_to(inpaint_masks[:, None].float(), shape[-1]).bool()
        timesteps = noise_scheduler.get_sampling_timesteps(batch, device=img.device)
        for step, (times, times_next) in enumerate(timesteps):
            is_last_timestep = times_next == 0
            for r in reversed(range(resample_times)):
                is_last_resample_step = r == 0
                if has_inpainting:
                    noised, *_ = noise_scheduler.q_sample(
                        inpaint_images, t=times, noise=noise_fn("inpaint", shape, unet=unet_number, step=step, r=r))
                    img = img * ~inpaint_masks + noised * inpaint_masks
                taps = {} if exists(step_taps) else None
                if exists(taps):
                    taps["x_in"] = img
                img, x_start = self.p_sample(
                    unet, img, times, t_next=times_next, noise=noise_fn("p_sample", shape, unet=unet_number, step=step, r=r),
                    text_embeds=text_embeds, text_mask=text_mask, cond_images=cond_images, cond_scale=cond_scale,
                    lowres_cond_img=lowres_cond_img, lowres_noise_times=lowres_noise_times, noise_scheduler=noise_scheduler,
                    pred_objective=pred_objective, dynamic_threshold=dynamic_threshold, taps=taps,
                )
                if has_inpainting and not (is_last_resample_step or bool(torch.all(is_last_timestep))):
                    renoised = noise_scheduler.q_sample_from_to(
                        img, times_next, times, noise=noise_fn("renoise", shape, unet=unet_number, step=step, r=r))
                    img = torch.where(right_pad_dims_to(img, is_last_timestep), img, renoised)
                if exists(taps):
                    taps["img"] = img
                    step_taps.append(taps)
        img = img.clamp(-1.0, 1.0)
        if has_inpainting:
            img = img * ~inpaint_masks + inpaint_images * inpaint_masks
        return self.unnormalize_img(img)

    @torch.no_grad()
    def sample(self, text_embeds=None, text_masks=None, cond_images=None, inpaint_images=None, inpaint_masks=None,
               inpaint_resample_times=5, batch_size=1, cond_scale=1.0, lowres_sample_noise_level=None, start_at_unet_number=1,
               start_image_or_video=None, stop_at_unet_number=None, return_all_unet_outputs=False, return_pil_images=False,
               device=None, use_tqdm=True, noise_fn=None, step_taps=None):
        self.eval()
        noise_fn = noise_fn if exists(noise_fn) else default_noise_fn
        if exists(cond_images) and cond_images.dtype == torch.uint8:
            cond_images = cond_images.float() / 255
        if not self.unconditional:
            assert exists(text_embeds), "text must be passed in if the network was not trained without text"
            text_masks = default(text_masks, lambda: torch.any(text_embeds != 0.0, dim=-1))
            batch_size = text_embeds.shape[0]
        if exists(inpaint_images):
            if self.unconditional and batch_size == 1:
                batch_size = inpaint_images.shape[0]
            assert inpaint_images.shape[0] == batch_size
        assert not (self.condition_on_text and not exists(text_embeds))
        assert not (not self.condition_on_text and exists(text_embeds))
        assert not (exists(text_embeds) and text_embeds.shape[-1] != self.text_embed_dim)
        assert not (exists(inpaint_images) ^ exists(inpaint_masks))
        outputs = []
        lowres_sample_noise_level = default(lowres_sample_noise_level, self.lowres_sample_noise_level)
        num_unets = len(self.unets)
        cond_scale = cast_tuple(cond_scale, num_unets)
        img = None
        if start_at_unet_number > 1:
            assert start_at_unet_number <= num_unets
            assert not exists(stop_at_unet_number) or start_at_unet_number <= stop_at_unet_number
            assert exists(start_image_or_video)
            img = resize_image_to(start_image_or_video, self.image_sizes[start_at_unet_number - 2])
        for unet_number, unet, channel, image_size, noise_scheduler, pred_objective, dynamic_threshold, unet_cond_scale in zip(
            range(1, num_unets + 1), self.unets, self.sample_channels, self.image_sizes, self.noise_schedulers,
            self.pred_objectives, self.dynamic_thresholding, cond_scale,
        ):
            if unet_number < start_at_unet_number:
                continue
            assert not isinstance(unet, NullUnet), "one cannot sample from null / placeholder unets"
            lowres_cond_img = lowres_noise_times = None
            if unet.lowres_cond:
                lowres_noise_times = self.lowres_noise_schedule.get_times(batch_size, lowres_sample_noise_level, device=img.device)
                lowres_cond_img = resize_image_to(img, image_size)
                lowres_cond_img = self.normalize_img(lowres_cond_img)
                lowres_cond_img, *_ = self.lowres_noise_schedule.q_sample(
                    x_start=lowres_cond_img, t=lowres_noise_times,
                    noise=noise_fn("lowres_aug", lowres_cond_img.shape, unet=unet_number))
            shape = (batch_size, self.channels, image_size, image_size)
            img = self.p_sample_loop(
                unet, shape, text_embeds=text_embeds, text_mask=text_masks, cond_images=cond_images, inpaint_images=inpaint_images,
                inpaint_masks=inpaint_masks, inpaint_resample_times=inpaint_resample_times, cond_scale=unet_cond_scale,
                lowres_cond_img=lowres_cond_img, lowres_noise_times=lowres_noise_times, noise_scheduler=noise_scheduler,
                pred_objective=pred_objective, dynamic_threshold=dynamic_threshold, noise_fn=noise_fn, unet_number=unet_number,
                step_taps=step_taps,
            )
            outputs.append(img)
            if exists(stop_at_unet_number) and stop_at_unet_number == unet_number:
                break
        out = outputs[-1] if not return_all_unet_outputs else outputs
        if not return_pil_images:
            return out
        from torchvision.transforms import ToPILImage  # pragma: no cover - only for API parity

        return [ToPILImage()(i) for i in out]


def restore_parts(state_dict_target, state_dict_from):
    """imagen_pytorch.trainer.restore_parts: copy tensors with matching name and shape."""
    for name, param in state_dict_from.items():
        if name not in state_dict_target:
            continue
        if param.size() == state_dict_target[name].size():
            state_dict_target[name].copy_(param)
    return state_dict_target


def randomize_zero_init_(module: nn.Module, std: float = 0.02, seed: int = 1):
    """Zero-initialised tensors (final_conv, PixelShuffle biases ...) make parity vacuous: redraw them."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            if p.numel() > 1 and bool((p == 0).all()):
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return module
