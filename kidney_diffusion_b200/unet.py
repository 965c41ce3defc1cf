"""Drop-in ``Unet`` / ``NullUnet`` with the constructor contract of imagen-pytorch 1.18.5 as the reference uses it
(train_ultra_res_v_param.py:27-62, train.py:28-67, train_uncond.py:28-63; defaults per SURVEY.md appendix A.4).

The module owns ordinary ``nn.Parameter``s under the reference's key names (so ``load_state_dict`` of a reference
checkpoint works); ``forward`` hands them, packed once into TMA-friendly fp16 layouts, to ``UnetExecutor`` which runs
the whole pass with the hand-written CUDA kernels of libkidney_b200.  There is no PyTorch fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from .modules import (
    CrossEmbedLayer, Downsample, LearnedSinusoidalPosEmb, LinearAttentionTransformerBlock, Parallel, PerceiverResampler, PixelShuffleUpsample, ResnetBlock,
    TransformerBlock, cast_tuple, default, exists,
)


class Unet(nn.Module):
    # "fp16": tcgen05 tensor-core path (fp16 storage, fp32 accumulation; per-step rel-L2 <= 1e-2 vs the fp32 reference).
    # "fp32": precise path (csrc/kd_precise.cu: fp32 storage and CUDA-core accumulation; per-step rel-L2 <= 1e-4), 20-50x slower.
    precision = "fp16"

    def __init__(
        self, *, dim, image_embed_dim=1024, text_embed_dim=768, num_resnet_blocks=1, cond_dim=None, num_image_tokens=4,
        num_time_tokens=2, learned_sinu_pos_emb_dim=16, out_dim=None, dim_mults=(1, 2, 4, 8), cond_images_channels=0, channels=3,
        channels_out=None, attn_dim_head=64, attn_heads=8, ff_mult=2.0, lowres_cond=False, layer_attns=True, layer_attns_depth=1,
        layer_mid_attns_depth=1, layer_attns_add_text_cond=True, attend_at_middle=True, layer_cross_attns=True,
        use_linear_attn=False, use_linear_cross_attn=False, cond_on_text=True, max_text_len=256, init_dim=None, resnet_groups=8,
        init_conv_kernel_size=7, init_cross_embed=True, init_cross_embed_kernel_sizes=(3, 7, 15), cross_embed_downsample=False,
        cross_embed_downsample_kernel_sizes=(2, 4), attn_pool_text=True, attn_pool_num_latents=32, dropout=0.0,
        memory_efficient=False, init_conv_to_final_conv_residual=False, use_global_context_attn=True, scale_skip_connection=True,
        final_resnet_block=True, final_conv_kernel_size=3, self_cond=False, resize_mode="nearest", combine_upsample_fmaps=False,
        pixel_shuffle_upsample=True,
    ):
        super().__init__()
        self._locals = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        unsupported = dict(
            cross_embed_downsample=cross_embed_downsample,
            self_cond=self_cond, combine_upsample_fmaps=combine_upsample_fmaps,
        )
        for k, v in unsupported.items():
            if v not in (False, (False,), None) and any(cast_tuple(v)):
                raise NotImplementedError(f"Unet({k}={v!r}) is not used by any reference script and is not built on the CUDA path")
        if not (init_cross_embed and pixel_shuffle_upsample and final_resnet_block and final_conv_kernel_size == 3 and dropout == 0.0):
            raise NotImplementedError("only the init_cross_embed / pixel_shuffle_upsample / final_resnet_block configuration is built")
        assert attn_heads > 1
        assert dim % 64 == 0, "the tcgen05 implicit-GEMM path needs channel counts that are multiples of 64"

        self.channels = channels
        self.channels_out = default(channels_out, channels)
        init_channels = channels * (1 + int(lowres_cond))
        init_dim = default(init_dim, dim)
        self.has_cond_image = cond_images_channels > 0
        self.cond_images_channels = cond_images_channels
        init_channels += cond_images_channels
        self.init_channels = init_channels
        self.init_conv = CrossEmbedLayer(init_channels, dim_out=init_dim, kernel_sizes=init_cross_embed_kernel_sizes, stride=1)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        cond_dim = default(cond_dim, dim)
        self.cond_dim = cond_dim
        time_cond_dim = dim * 4 * (2 if lowres_cond else 1)
        self.time_cond_dim = time_cond_dim
        self.num_time_tokens = num_time_tokens

        self.to_time_hiddens = nn.Sequential(
            LearnedSinusoidalPosEmb(learned_sinu_pos_emb_dim), nn.Linear(learned_sinu_pos_emb_dim + 1, time_cond_dim), nn.SiLU()
        )
        self.to_time_cond = nn.Sequential(nn.Linear(time_cond_dim, time_cond_dim))
        self.to_time_tokens = nn.Sequential(nn.Linear(time_cond_dim, cond_dim * num_time_tokens))
        self.lowres_cond = lowres_cond
        if lowres_cond:
            self.to_lowres_time_hiddens = nn.Sequential(
                LearnedSinusoidalPosEmb(learned_sinu_pos_emb_dim), nn.Linear(learned_sinu_pos_emb_dim + 1, time_cond_dim), nn.SiLU()
            )
            self.to_lowres_time_cond = nn.Sequential(nn.Linear(time_cond_dim, time_cond_dim))
            self.to_lowres_time_tokens = nn.Sequential(nn.Linear(time_cond_dim, cond_dim * num_time_tokens))
        self.norm_cond = nn.LayerNorm(cond_dim)

        self.text_to_cond = nn.Linear(text_embed_dim, cond_dim) if cond_on_text else None
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
        assert attn_dim_head == 64, "attention kernels are built for dim_head = 64 (the reference's value)"
        num_layers = len(in_out)
        num_resnet_blocks = cast_tuple(num_resnet_blocks, num_layers)
        resnet_groups = cast_tuple(resnet_groups, num_layers)
        layer_attns = cast_tuple(layer_attns, num_layers)
        layer_attns_depth = cast_tuple(layer_attns_depth, num_layers)
        layer_cross_attns = cast_tuple(layer_cross_attns, num_layers)
        self.memory_efficient = memory_efficient

        self.init_resnet_block = (
            ResnetBlock(init_dim, init_dim, time_cond_dim=time_cond_dim, groups=resnet_groups[0], use_gca=use_global_context_attn)
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
                                            if layer_lin else nn.Identity()),
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
                                            if layer_lin else nn.Identity()),
                        PixelShuffleUpsample(dim_out, dim_in) if not is_last or memory_efficient else nn.Identity(),
                    ]
                )
            )
        self.init_conv_to_final_conv_residual = init_conv_to_final_conv_residual
        final_conv_dim = dim + (dim if init_conv_to_final_conv_residual else 0)
        self.final_res_block = ResnetBlock(final_conv_dim, dim, time_cond_dim=time_cond_dim, groups=resnet_groups[0], use_gca=True)
        final_conv_dim_in = dim + (channels if lowres_cond else 0)
        self.final_conv = nn.Conv2d(final_conv_dim_in, self.channels_out, final_conv_kernel_size, padding=final_conv_kernel_size // 2)
        nn.init.zeros_(self.final_conv.weight)
        nn.init.zeros_(self.final_conv.bias)
        self._executor = None

    # ------------------------------------------------------------------ reference API
    def cast_model_parameters(self, *, lowres_cond, text_embed_dim, channels, channels_out, cond_on_text):
        if (
            lowres_cond == self.lowres_cond and channels == self.channels and cond_on_text == self.cond_on_text
            and text_embed_dim == self._locals["text_embed_dim"] and channels_out == self.channels_out
        ):
            return self
        updated = dict(lowres_cond=lowres_cond, text_embed_dim=text_embed_dim, channels=channels, channels_out=channels_out,
                       cond_on_text=cond_on_text)
        return self.__class__(**{**self._locals, **updated})

    def executor(self):
        from .unet_exec import UnetExecutor

        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("kidney_diffusion_b200.Unet runs on a B200 only: move the module to a CUDA device (no CPU fallback)")
        stamp = (dev, sum(p._version for p in self.parameters()), self.precision)
        if self._executor is None or self._executor.stamp != stamp:
            self._executor = UnetExecutor(self, dev, stamp, precision=self.precision)
        return self._executor

    def _apply(self, fn, *args, **kwargs):
        self._executor = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._executor = None
        return super().load_state_dict(*args, **kwargs)

    def forward_with_cond_scale(self, *args, cond_scale=1.0, **kwargs):
        logits = self.forward(*args, **kwargs)
        if cond_scale == 1:
            return logits
        from . import ops

        null_logits = self.forward(*args, cond_drop_prob=1.0, **kwargs)
        return ops.axpby(logits, null_logits, cond_scale, 1.0 - cond_scale)  # null + (logits - null) * cond_scale

    @torch.no_grad()
    def forward(self, x, time, *, lowres_cond_img=None, lowres_noise_times=None, text_embeds=None, text_mask=None, cond_images=None,
                self_cond=None, cond_drop_prob=0.0):
        """x: NCHW fp32 noisy image; time: log-SNR per sample (B,).  Returns the eps / v prediction, NCHW fp32."""
        assert not (self.lowres_cond and not exists(lowres_cond_img)), "low resolution conditioning image must be present"
        assert not (self.lowres_cond and not exists(lowres_noise_times)), "low resolution conditioning noise time must be present"
        assert not (self.has_cond_image ^ exists(cond_images)), (
            "you either requested to condition on an image on the unet, but the conditioning image is not supplied, or vice versa")
        if exists(cond_images):
            assert cond_images.shape[1] == self.cond_images_channels, (
                "the number of channels on the conditioning image you are passing in does not match what you specified on "
                "initialiation of the unet")
        ex = self.executor()
        ex.set_conditioning(cond_images=cond_images, lowres_cond_img=lowres_cond_img, text_embeds=text_embeds, text_mask=text_mask,
                            cond_drop_prob=cond_drop_prob, image_size=x.shape[-1])
        return ex.forward(x, time, lowres_noise_times)


class NullUnet(nn.Module):
    """Placeholder stage (imagen-pytorch NullUnet); subclassable the way the reference's FixedNullUnet does
    (train_ultra_res_v_param.py:65-75)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.lowres_cond = False
        self.dummy_parameter = nn.Parameter(torch.tensor([0.0]))

    def cast_model_parameters(self, *args, **kwargs):
        return self

    def forward(self, x, *args, **kwargs):
        return x
