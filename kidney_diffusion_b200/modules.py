"""Parameter containers that reproduce the module tree (and therefore the state_dict key names and shapes) of
imagen-pytorch 1.18.5's ``Unet`` as the reference constructs it (train_ultra_res_v_param.py:27-62, train.py:28-67,
train_uncond.py:28-63).  They hold weights only: none of these modules is ever *called* on the product path -- the
forward pass is executed by ``unet_exec.UnetExecutor`` with hand-written CUDA kernels.  torch.nn layers are used purely
as parameter holders with the reference's default initialisation.
"""
from __future__ import annotations

import torch
from torch import nn


def exists(v):
    return v is not None


def default(v, d):
    return v if exists(v) else (d() if callable(d) else d)


def cast_tuple(val, length=None):
    if isinstance(val, list):
        val = tuple(val)
    out = val if isinstance(val, tuple) else ((val,) * default(length, 1))
    if exists(length):
        assert len(out) == length
    return out


class _NoCall(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container; the forward pass runs through the CUDA executor "
            "(kidney_diffusion_b200 has no PyTorch fallback path)"
        )


class GainLayerNorm(_NoCall):
    """imagen-pytorch LayerNorm / ChanLayerNorm: gain ``g`` only (shape [C] or [C,1,1])."""

    def __init__(self, feats, dim=-1):
        super().__init__()
        self.g = nn.Parameter(torch.ones(feats, *((1,) * (-dim - 1))))


class LearnedSinusoidalPosEmb(_NoCall):
    def __init__(self, dim):
        super().__init__()
        self.weights = nn.Parameter(torch.randn(dim // 2))


class CrossEmbedLayer(_NoCall):
    def __init__(self, dim_in, kernel_sizes, dim_out=None, stride=2):
        super().__init__()
        dim_out = default(dim_out, dim_in)
        kernel_sizes = sorted(kernel_sizes)
        n = len(kernel_sizes)
        dim_scales = [int(dim_out / (2 ** i)) for i in range(1, n)]
        dim_scales = [*dim_scales, dim_out - sum(dim_scales)]
        self.kernel_sizes = kernel_sizes
        self.convs = nn.ModuleList(
            [nn.Conv2d(dim_in, d, k, stride=stride, padding=(k - stride) // 2) for k, d in zip(kernel_sizes, dim_scales)]
        )


def Downsample(dim, dim_out=None):
    # Sequential(Rearrange, Conv2d): the conv sits at index 1 like in the reference
    return nn.Sequential(nn.Identity(), nn.Conv2d(dim * 4, default(dim_out, dim), 1))


class PixelShuffleUpsample(_NoCall):
    def __init__(self, dim, dim_out=None):
        super().__init__()
        dim_out = default(dim_out, dim)
        conv = nn.Conv2d(dim, dim_out * 4, 1)
        self.net = nn.Sequential(conv, nn.SiLU(), nn.PixelShuffle(2))
        o, i, h, w = conv.weight.shape
        cw = torch.empty(o // 4, i, h, w)
        nn.init.kaiming_uniform_(cw)
        conv.weight.data.copy_(cw.repeat_interleave(4, dim=0))
        nn.init.zeros_(conv.bias.data)


class Parallel(_NoCall):
    def __init__(self, *fns):
        super().__init__()
        self.fns = nn.ModuleList(fns)


class Block(_NoCall):
    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        self.groupnorm = nn.GroupNorm(groups, dim)
        self.activation = nn.SiLU()
        self.project = nn.Conv2d(dim, dim_out, 3, padding=1)


class GlobalContext(_NoCall):
    def __init__(self, *, dim_in, dim_out):
        super().__init__()
        self.to_k = nn.Conv2d(dim_in, 1, 1)
        hidden = max(3, dim_out // 2)
        self.net = nn.Sequential(nn.Conv2d(dim_in, hidden, 1), nn.SiLU(), nn.Conv2d(hidden, dim_out, 1), nn.Sigmoid())


class CrossAttention(_NoCall):
    def __init__(self, dim, *, context_dim=None, dim_head=64, heads=8):
        super().__init__()
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        inner = dim_head * heads
        context_dim = default(context_dim, dim)
        self.norm = GainLayerNorm(dim)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), GainLayerNorm(dim))


class Attention(_NoCall):
    def __init__(self, dim, *, dim_head=64, heads=8, context_dim=None):
        super().__init__()
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        inner = dim_head * heads
        self.norm = GainLayerNorm(dim)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, dim_head * 2, bias=False)
        self.to_context = (
            nn.Sequential(nn.LayerNorm(context_dim), nn.Linear(context_dim, dim_head * 2)) if exists(context_dim) else None
        )
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), GainLayerNorm(dim))


def ChanFeedForward(dim, mult=2):
    hidden = int(dim * mult)
    return nn.Sequential(
        GainLayerNorm(dim, dim=-3), nn.Conv2d(dim, hidden, 1, bias=False), nn.GELU(), GainLayerNorm(hidden, dim=-3),
        nn.Conv2d(hidden, dim, 1, bias=False),
    )


def FeedForward(dim, mult=2):
    hidden = int(dim * mult)
    return nn.Sequential(
        GainLayerNorm(dim), nn.Linear(dim, hidden, bias=False), nn.GELU(), GainLayerNorm(hidden), nn.Linear(hidden, dim, bias=False)
    )


class EinopsToAndFrom(_NoCall):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


class TransformerBlock(_NoCall):
    def __init__(self, dim, *, depth=1, heads=8, dim_head=32, ff_mult=2, context_dim=None):
        super().__init__()
        self.layers = nn.ModuleList(
            [
                nn.ModuleList(
                    [
                        EinopsToAndFrom(Attention(dim=dim, heads=heads, dim_head=dim_head, context_dim=context_dim)),
                        ChanFeedForward(dim=dim, mult=ff_mult),
                    ]
                )
                for _ in range(depth)
            ]
        )


class LinearAttention(_NoCall):
    """imagen-pytorch LinearAttention: to_q / to_k / to_v = Sequential(Dropout, Conv2d 1x1, depthwise Conv2d 3x3), no biases."""

    def __init__(self, dim, dim_head=32, heads=8, dropout=0.05, context_dim=None, **kwargs):
        super().__init__()
        self.scale, self.heads, self.dim_head = dim_head ** -0.5, heads, dim_head
        inner = dim_head * heads
        self.norm = GainLayerNorm(dim, dim=-3)
        self.nonlin = nn.SiLU()

        def qkv():
            return nn.Sequential(nn.Dropout(dropout), nn.Conv2d(dim, inner, 1, bias=False),
                                 nn.Conv2d(inner, inner, 3, bias=False, padding=1, groups=inner))

        self.to_q, self.to_k, self.to_v = qkv(), qkv(), qkv()
        self.to_context = nn.Sequential(nn.LayerNorm(context_dim), nn.Linear(context_dim, inner * 2, bias=False)) if exists(context_dim) else None
        self.to_out = nn.Sequential(nn.Conv2d(inner, dim, 1, bias=False), GainLayerNorm(dim, dim=-3))


class LinearAttentionTransformerBlock(_NoCall):
    def __init__(self, dim, *, depth=1, heads=8, dim_head=32, ff_mult=2, context_dim=None, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList(
            [nn.ModuleList([LinearAttention(dim=dim, heads=heads, dim_head=dim_head, context_dim=context_dim), ChanFeedForward(dim=dim, mult=ff_mult)])
             for _ in range(depth)]
        )


class ResnetBlock(_NoCall):
    def __init__(self, dim, dim_out, *, cond_dim=None, time_cond_dim=None, groups=8, linear_attn=False, use_gca=False, **attn_kwargs):
        super().__init__()
        self.dim, self.dim_out, self.groups = dim, dim_out, groups
        self.linear_cross_attn = bool(linear_attn) and exists(cond_dim)  # LinearCrossAttention: same parameters, different forward
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_cond_dim, dim_out * 2)) if exists(time_cond_dim) else None
        self.cross_attn = (
            EinopsToAndFrom(CrossAttention(dim=dim_out, context_dim=cond_dim, **attn_kwargs)) if exists(cond_dim) else None
        )
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.gca = GlobalContext(dim_in=dim_out, dim_out=dim_out) if use_gca else None
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else None


class PerceiverAttention(_NoCall):
    def __init__(self, *, dim, dim_head=64, heads=8):
        super().__init__()
        self.scale, self.heads = dim_head ** -0.5, heads
        inner = dim_head * heads
        self.norm = nn.LayerNorm(dim)
        self.norm_latents = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), nn.LayerNorm(dim))


class PerceiverResampler(_NoCall):
    def __init__(self, *, dim, depth, dim_head=64, heads=8, num_latents=64, num_latents_mean_pooled=4, max_seq_len=512, ff_mult=4):
        super().__init__()
        self.pos_emb = nn.Embedding(max_seq_len, dim)
        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.num_latents_mean_pooled = num_latents_mean_pooled
        self.to_latents_from_mean_pooled_seq = (
            nn.Sequential(GainLayerNorm(dim), nn.Linear(dim, dim * num_latents_mean_pooled)) if num_latents_mean_pooled > 0 else None
        )
        self.layers = nn.ModuleList(
            [
                nn.ModuleList([PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads), FeedForward(dim=dim, mult=ff_mult)])
                for _ in range(depth)
            ]
        )
