"""Drop-in ``Imagen`` for the sampling path (imagen-pytorch 1.18.5 ``Imagen.sample -> p_sample_loop -> p_sample ->
p_mean_variance``; reference call sites sample_ultra_res.py:183-195, sample_cond.py:40-48, sample_uncond.py:49-55,
construction train_ultra_res_v_param.py:78-92 / train.py:83-95 / train_uncond.py:79-93).

Host side only orchestrates: per step it launches the UNet (a captured CUDA graph of hand-written kernels), the exact
dynamic-threshold select (K7) and the fused p_sample update (K6).  Training (``forward`` / loss) is out of scope.
"""
from __future__ import annotations

import zlib

import torch
import torch.nn.functional as F
from torch import nn

from . import ops, schedule, torch_ops  # noqa: F401  (torch_ops registers torch.ops.kidney_b200.*)
from .modules import cast_tuple, default, exists
from .unet import NullUnet, Unet

SITES = {"lowres_aug": 1, "init": 2, "inpaint": 3, "p_sample": 4, "renoise": 5}
K = torch.ops.kidney_b200  # the sampler-level kernels are called as PyTorch custom ops (CUDA dispatch key only: no CPU kernel exists)


def pad_tuple_to_length(t, length, fillvalue=None):
    remain = length - len(t)
    return t if remain <= 0 else (*t, *((fillvalue,) * remain))


def resize_image_to(image, target_image_size, mode="nearest"):
    if image.shape[-1] == target_image_size:
        return image
    return F.interpolate(image, target_image_size, mode=mode)


class NoiseSchedulerSpec(nn.Module):
    """Parameter-free stand-in for GaussianDiffusionContinuousTimes (keeps Imagen's module tree shape)."""

    def __init__(self, *, noise_schedule, timesteps=1000):
        super().__init__()
        assert noise_schedule in schedule.LOG_SNR, f"invalid noise schedule {noise_schedule}"
        self.noise_schedule = noise_schedule
        self.num_timesteps = timesteps


class CounterNoise:
    """Counter-based noise: every randn site of the sampler is keyed by (seed, stream key, unet, step, r, site), so
    results are identical for any GPU count / patch-to-rank assignment (SURVEY.md section 8e invariance requirement).
    ``stream_key`` is one int for the whole batch or a list with one key per batch element (the patch-grid sampler keys
    each patch by its index so that batching patches never changes a patch's noise).
    The reference draws from the unseeded global torch generator; tests inject identical tensors into both paths."""

    def __init__(self, seed=0, stream_key=0):
        self.seed, self.stream_key = seed, stream_key

    @staticmethod
    def _key(stream_key, unet, step, r, site):
        return zlib.crc32(f"{stream_key}/{unet}/{step}/{r}/{SITES[site]}".encode()) | (SITES[site] << 40) | (unet << 48)

    def __call__(self, site, shape, device, unet=0, step=0, r=0):
        if isinstance(self.stream_key, (list, tuple)):
            assert len(self.stream_key) == shape[0]
            out = torch.empty(tuple(shape), device=device, dtype=torch.float32)
            for b, k in enumerate(self.stream_key):
                ops.randn_into(out[b], self.seed, self._key(k, unet, step, r, site))
            return out
        return ops.randn(shape, self.seed, self._key(self.stream_key, unet, step, r, site), device)


class _StepGraph:
    """One UNet forward for fixed (B, S) captured as a CUDA graph (~800 kernel launches per step otherwise)."""

    def __init__(self, ex, B, S, channels, lowres_t, device, drop=0.0, pool=None):
        self.drop = drop
        self.x = torch.zeros((B, channels, S, S), device=device, dtype=torch.float32)
        self.time = torch.zeros((B,), device=device, dtype=torch.float32)
        self.lowres_t = lowres_t.clone() if lowres_t is not None else None
        self.ex = ex
        stream = torch.cuda.Stream(device=device)
        stream.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(stream):
            for _ in range(2):  # warm-up: function attributes, allocator pools
                ex.forward(self.x, self.time, self.lowres_t, drop=drop)
        torch.cuda.current_stream(device).wait_stream(stream)
        before = ops.launch_count
        self.graph = torch.cuda.CUDAGraph()
        # all step graphs of one conditioning kind share a memory pool: they are replayed one at a time and `pred` is consumed
        # (dynamic threshold + update) before the next replay, so the pool is as large as the biggest graph, not their sum --
        # the patch-grid sampler meets every batch size 1..16 per stage
        with torch.cuda.graph(self.graph, pool=pool):
            self.pred = ex.forward(self.x, self.time, self.lowres_t, drop=drop)
        self.launches = ops.launch_count - before

    def __call__(self, img, time_row, lowres_t):
        self.x.copy_(img)
        self.time.copy_(time_row)
        if self.lowres_t is not None:
            self.lowres_t.copy_(lowres_t)
        self.graph.replay()
        ops.launch_count += self.launches
        return self.pred


class StageRun:
    """One stage of the cascade: p_sample_loop state resident in HBM.  ``step(k, r)`` is one *patch-step* (the unit of the
    headline metric): [RePaint blend] -> UNet forward -> exact dynamic threshold -> fused p_sample update [+ re-noise]."""

    def __init__(self, imagen, unet_number, shape, *, noise, lowres_cond_img, lowres_noise_level, text_embeds, text_mask, cond_images,
                 inpaint_images, inpaint_masks, inpaint_resample_times, cond_scale):
        assert not (cond_scale != 1.0 and not imagen.can_classifier_guidance), (
            "imagen was not trained with conditional dropout, and thus one cannot use classifier free guidance (cond_scale anything other than 1)")
        self.cond_scale = float(cond_scale)
        self.im, self.unet_number, self.shape, self.noise = imagen, unet_number, tuple(shape), noise
        self.unet = unet = imagen.unets[unet_number - 1]
        spec = imagen.noise_schedulers[unet_number - 1]
        self.sched = spec.noise_schedule
        self.objective = imagen.pred_objectives[unet_number - 1]
        self.dynamic_threshold = imagen.dynamic_thresholding[unet_number - 1]
        self.device = device = imagen.device
        B, _, S, _ = shape
        self.B, self.S = B, S
        self.img = noise("init", self.shape, device, unet=unet_number).contiguous()
        self.has_inpainting = exists(inpaint_images) and exists(inpaint_masks)
        self.resample_times = inpaint_resample_times if self.has_inpainting else 1
        self.inpaint, self.mask_u8 = None, None
        if self.has_inpainting:
            self.inpaint = resize_image_to(imagen.normalize_img(inpaint_images.float()), S).contiguous()
            self.mask_u8 = resize_image_to(inpaint_masks[:, None].float(), S).bool()[:, 0].to(torch.uint8).contiguous()
        self.ex = unet.executor()
        if self.cond_scale != 1.0:  # classifier-free guidance (sample.py:59): also prepare the null-conditioning context
            self.ex.set_conditioning(cond_images=cond_images, lowres_cond_img=lowres_cond_img, text_embeds=text_embeds, text_mask=text_mask,
                                     cond_drop_prob=1.0, image_size=S)
        self.ex.set_conditioning(cond_images=cond_images, lowres_cond_img=lowres_cond_img, text_embeds=text_embeds, text_mask=text_mask,
                                 cond_drop_prob=0.0, image_size=S)
        self.lowres_t = None
        if unet.lowres_cond:
            lt = schedule.log_snr(imagen.lowres_noise_schedule.noise_schedule, lowres_noise_level)
            self.lowres_t = torch.full((B,), float(lt), device=device, dtype=torch.float32)
        self.times = schedule.sampling_times(spec.num_timesteps)
        self.num_steps = len(self.times)
        self.scal = [schedule.step_scalars(self.sched, t, tn) for t, tn in self.times]
        # pinned + non_blocking: a pageable host-to-device copy would make the host wait for everything queued on the stream, i.e.
        # for the previous batch of a patch-grid run, and the GPU would idle while the host prepares this one
        tt = torch.tensor([[s["log_snr"]] * B for s in self.scal], dtype=torch.float32)
        self.time_table = tt.pin_memory().to(device, non_blocking=True) if torch.device(device).type == "cuda" else tt.to(device)
        self.ws = torch.empty(ops.lib().kd_dynthresh_workspace_bytes(B), device=device, dtype=torch.uint8)

    def step(self, step, r=0):
        im, sc, n = self.im, self.scal[step], self.unet_number
        t, t_next = self.times[step]
        if self.has_inpainting:
            K.inpaint_blend(self.img, self.inpaint, self.mask_u8, self.noise("inpaint", self.shape, self.device, unet=n, step=step, r=r),
                            sc["alpha"], sc["sigma"])
        pred = im._unet_step(self.unet, self.img, self.time_table[step], self.lowres_t, self.B, self.S, ex=self.ex)
        if self.cond_scale != 1.0:  # Unet.forward_with_cond_scale: null + (cond - null) * scale
            null = im._unet_step(self.unet, self.img, self.time_table[step], self.lowres_t, self.B, self.S, drop=1.0, ex=self.ex)
            pred = K.axpby(pred, null, self.cond_scale, 1.0 - self.cond_scale)
        s = None
        if self.dynamic_threshold:
            s = K.dynthresh(self.img, pred, self.objective, sc["alpha"], sc["sigma"], im.dynamic_thresholding_percentile, self.ws)
        renoise, rn = None, (0.0, 0.0, 1.0)
        if self.has_inpainting and not (r == 0 or bool(t_next == 0)):
            renoise = self.noise("renoise", self.shape, self.device, unet=n, step=step, r=r)
            rn = schedule.renoise_scalars(self.sched, t_next, t)
        x_in = self.img
        self.img = K.ddpm_step(x_in, pred, self.noise("p_sample", self.shape, self.device, unet=n, step=step, r=r), s, self.objective,
                               sc["alpha"], sc["sigma"], sc["one_minus_c"], sc["c"], sc["alpha_next"], sc["std"], renoise, rn[0], rn[1], rn[2])
        if exists(im.step_hook):
            im.step_hook(dict(unet=n, step=step, r=r, x_in=x_in, pred=pred, img=self.img))
        return self.img

    def finish(self):
        K.finalize_image(self.img, self.inpaint if self.has_inpainting else None, self.mask_u8)
        img = self.img
        if not self.im.auto_normalize_img:  # finalize_image un-normalises; undo for the (unused by the reference) raw mode
            img = img * 2 - 1
        return img


class Imagen(nn.Module):
    def __init__(
        self, unets, *, image_sizes, text_encoder_name=None, text_embed_dim=None, channels=3, timesteps=1000, cond_drop_prob=0.1,
        loss_type="l2", noise_schedules="cosine", pred_objectives="noise", random_crop_sizes=None, lowres_noise_schedule="linear",
        lowres_sample_noise_level=0.2, per_sample_random_aug_noise_level=False, condition_on_text=True, auto_normalize_img=True,
        dynamic_thresholding=True, dynamic_thresholding_percentile=0.95, only_train_unet_number=None, temporal_downsample_factor=1,
        resize_cond_video_frames=True, resize_mode="nearest", min_snr_loss_weight=True, min_snr_gamma=5, p2_loss_weight_gamma=0.5,
        p2_loss_weight_k=1,
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
            [NoiseSchedulerSpec(noise_schedule=s, timesteps=t) for t, s in zip(timesteps, noise_schedules)])
        self.lowres_noise_schedule = NoiseSchedulerSpec(noise_schedule=lowres_noise_schedule)
        self.pred_objectives = cast_tuple(pred_objectives, num_unets)
        self.text_embed_dim = default(text_embed_dim, 768)
        self.unets = nn.ModuleList([])
        for ind, one_unet in enumerate(unets):
            assert isinstance(one_unet, (Unet, NullUnet))
            one_unet = one_unet.cast_model_parameters(
                lowres_cond=not ind == 0, cond_on_text=self.condition_on_text,
                text_embed_dim=self.text_embed_dim if self.condition_on_text else None, channels=self.channels, channels_out=self.channels)
            self.unets.append(one_unet)
        self.image_sizes = cast_tuple(image_sizes)
        assert num_unets == len(self.image_sizes), (
            f"you did not supply the correct number of u-nets ({num_unets}) for resolutions {self.image_sizes}")
        self.sample_channels = cast_tuple(self.channels, num_unets)
        lowres_conditions = tuple(u.lowres_cond for u in self.unets)
        assert lowres_conditions == (False, *((True,) * (num_unets - 1))), (
            "the first unet must be unconditioned (by low resolution image), and the rest of the unets must have `lowres_cond` set to True")
        self.random_crop_sizes = cast_tuple(random_crop_sizes, num_unets)
        self.lowres_sample_noise_level = lowres_sample_noise_level
        self.cond_drop_prob = cond_drop_prob
        self.can_classifier_guidance = cond_drop_prob > 0.0
        self.auto_normalize_img = auto_normalize_img
        self.dynamic_thresholding = cast_tuple(dynamic_thresholding, num_unets)
        self.dynamic_thresholding_percentile = dynamic_thresholding_percentile
        self.register_buffer("_temp", torch.tensor([0.0]), persistent=False)
        self.use_cuda_graph = True
        self.noise_fn = None  # tests inject (site, shape, **key) -> tensor; default CounterNoise
        self.noise_seed = 0
        self._graphs = {}
        self._pools = {}
        self.step_hook = None  # optional callable(dict) per inner iteration (parity taps)
        # overflow guard of the fp16 activation path: when True, sample() runs without CUDA graphs, counts every stored activation
        # value that was clipped at +-65504 (kd_count_saturated) and leaves the total in `last_saturation_count`
        self.check_saturation = False
        self.last_saturation_count = None

    @property
    def device(self):
        return self._temp.device

    def set_precision(self, precision):
        """"fp16" (default): tcgen05 tensor-core path.  "fp32": the precise CUDA-core path of every stage's UNet (csrc/kd_precise.cu;
        per-step parity 1e-4 against the fp32 reference, 20-50x slower).  The sampler update is fp32 in both."""
        assert precision in ("fp16", "fp32"), precision
        changed = False
        for u in self.unets:
            if hasattr(u, "precision") and u.precision != precision:
                u.precision = precision
                changed = True
        if changed:  # captured step graphs belong to the other executor
            self._graphs = {}
        return self

    def normalize_img(self, img):
        return img * 2 - 1 if self.auto_normalize_img else img

    def _apply(self, fn, *a, **k):
        self._graphs = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._graphs = {}
        return super().load_state_dict(*a, **k)

    def forward(self, *a, **k):  # pragma: no cover
        raise NotImplementedError("training (Imagen.forward / loss) is outside the sampling hot path built here")

    # ------------------------------------------------------------------ one stage
    def _unet_step(self, unet, img, time_row, lowres_t, B, S, drop=0.0, ex=None):
        # `ex`: the executor a StageRun looked up once.  unet.executor() re-validates the packed weights against every parameter's
        # version counter -- 4.7 ms of host time for the 930 parameters of the 1024^2 UNet, as long as a whole step of the 64^2 stage.
        if ex is None:
            ex = unet.executor()
        if not self.use_cuda_graph:
            return ex.forward(img, time_row, lowres_t, drop=drop)
        has_text = float(drop) in ex.text_by_drop
        key = (id(ex), B, S, exists(ex.init_base), exists(ex.lowres_img), float(drop) if has_text else None)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) > 96:
                self._graphs.clear()
            # conditional and null-conditioning graphs (classifier-free guidance) keep separate pools: both predictions are
            # alive when they are mixed
            pkey = (str(img.device), float(drop))
            if pkey not in self._pools:
                self._pools[pkey] = torch.cuda.graph_pool_handle()
            g = self._graphs[key] = _StepGraph(ex, B, S, self.channels, lowres_t, img.device, drop=drop, pool=self._pools[pkey])
        return g(img, time_row, lowres_t)

    def stage_run(self, unet_number, shape, *, noise, lowres_cond_img=None, lowres_noise_level=None, text_embeds=None, text_mask=None,
                  cond_images=None, inpaint_images=None, inpaint_masks=None, inpaint_resample_times=5, cond_scale=1.0):
        """Prepare one cascade stage (x-independent conditioning, schedule tables, static buffers) and return its runner."""
        return StageRun(self, unet_number, shape, noise=noise, lowres_cond_img=lowres_cond_img, lowres_noise_level=lowres_noise_level,
                        text_embeds=text_embeds, text_mask=text_mask, cond_images=cond_images, inpaint_images=inpaint_images,
                        inpaint_masks=inpaint_masks, inpaint_resample_times=inpaint_resample_times, cond_scale=cond_scale)

    def p_sample_loop(self, unet_number, shape, **kwargs):
        run = self.stage_run(unet_number, shape, **kwargs)
        for step in range(run.num_steps):
            for r in reversed(range(run.resample_times)):
                run.step(step, r)
        return run.finish()

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def sample(self, texts=None, text_masks=None, text_embeds=None, video_frames=None, cond_images=None, cond_video_frames=None,
               post_cond_video_frames=None, inpaint_videos=None, inpaint_images=None, inpaint_masks=None, inpaint_resample_times=5,
               init_images=None, skip_steps=None, batch_size=1, cond_scale=1.0, lowres_sample_noise_level=None,
               start_at_unet_number=1, start_image_or_video=None, stop_at_unet_number=None, return_all_unet_outputs=False,
               return_pil_images=False, device=None, use_tqdm=True, use_one_unet_in_gpu=True, noise_key=None):
        if self.training:  # nn.Module.eval() walks all 11 000 sub-modules (7 ms): only when there is something to switch
            self.eval()
        device = torch.device(default(device, self.device))
        if device.type != "cuda":
            raise RuntimeError("kidney_diffusion_b200.Imagen.sample runs on a B200 only (no CPU fallback); call .cuda() / .to('cuda') first")
        if next(self.parameters()).device != device:
            self.to(device)
        assert not exists(texts), "raw-text encoding (T5) is not part of the reference's sampling path; pass text_embeds"
        assert not exists(init_images) and not exists(skip_steps), "init_images / skip_steps are unused by the reference"
        to_dev = lambda t: t.to(device) if exists(t) else None
        cond_images, text_embeds, text_masks, inpaint_images, inpaint_masks, start_image_or_video = map(
            to_dev, (cond_images, text_embeds, text_masks, inpaint_images, inpaint_masks, start_image_or_video))
        if exists(cond_images) and cond_images.dtype == torch.uint8:
            cond_images = cond_images.float() / 255
        if not self.unconditional:
            assert exists(text_embeds), (
                "text must be passed in if the network was not trained without text `condition_on_text` must be set to `False` when training")
            text_masks = default(text_masks, lambda: torch.any(text_embeds != 0.0, dim=-1))
            batch_size = text_embeds.shape[0]
        inpaint_images = default(inpaint_videos, inpaint_images)
        if exists(inpaint_images):
            if self.unconditional and batch_size == 1:
                batch_size = inpaint_images.shape[0]
            assert inpaint_images.shape[0] == batch_size, (
                "number of inpainting images must be equal to the specified batch size on sample `sample(batch_size=<int>)``")
        assert not (self.condition_on_text and not exists(text_embeds)), "text or text encodings must be passed into imagen if specified"
        assert not (not self.condition_on_text and exists(text_embeds)), "imagen specified not to be conditioned on text, yet it is presented"
        assert not (exists(text_embeds) and text_embeds.shape[-1] != self.text_embed_dim), (
            f"invalid text embedding dimension being passed in (should be {self.text_embed_dim})")
        assert not (exists(inpaint_images) ^ exists(inpaint_masks)), "inpaint images and masks must be both passed in to do inpainting"

        guard = None
        if self.check_saturation:
            guard = dict(graph=self.use_cuda_graph, counter=torch.zeros(1, dtype=torch.int64, device=device))
            self.use_cuda_graph, ops.sat_counter = False, guard["counter"]
        try:
            return self._sample(device, texts, text_masks, text_embeds, cond_images, inpaint_videos, inpaint_images, inpaint_masks,
                                inpaint_resample_times, batch_size, cond_scale, lowres_sample_noise_level, start_at_unet_number,
                                start_image_or_video, stop_at_unet_number, return_all_unet_outputs, return_pil_images, noise_key)
        finally:
            if guard is not None:
                ops.sat_counter, self.use_cuda_graph = None, guard["graph"]
                self.last_saturation_count = int(guard["counter"].item())
                if self.last_saturation_count:
                    import warnings

                    warnings.warn(f"{self.last_saturation_count} activation values were clipped at the fp16 range (+-65504) during sample(): "
                                  "these weights overflow the 16-bit activation path")

    def _sample(self, device, texts, text_masks, text_embeds, cond_images, inpaint_videos, inpaint_images, inpaint_masks, inpaint_resample_times,
                batch_size, cond_scale, lowres_sample_noise_level, start_at_unet_number, start_image_or_video, stop_at_unet_number,
                return_all_unet_outputs, return_pil_images, noise_key):
        noise = self.noise_fn
        if noise is None:
            if noise_key is None:
                # the reference draws from torch's global generator: every call gets fresh noise, reproducible under
                # torch.manual_seed.  One draw from that generator seeds this call's counter-based streams.
                noise = CounterNoise(int(torch.randint(0, 2 ** 62, (1,)).item()), 0)
            else:  # explicit stream key(s) (patch-grid sampler): noise is a pure function of (noise_seed, key, site)
                noise = CounterNoise(self.noise_seed, noise_key)
        outputs = []
        lowres_sample_noise_level = default(lowres_sample_noise_level, self.lowres_sample_noise_level)
        num_unets = len(self.unets)
        cond_scale = cast_tuple(cond_scale, num_unets)
        img = None
        if start_at_unet_number > 1:
            assert start_at_unet_number <= num_unets, "must start a unet that is less than the total number of unets"
            assert not exists(stop_at_unet_number) or start_at_unet_number <= stop_at_unet_number
            assert exists(start_image_or_video), "starting image or video must be supplied if only doing upscaling"
            img = resize_image_to(start_image_or_video.float(), self.image_sizes[start_at_unet_number - 2])
        for unet_number, unet, image_size, spec, pred_objective, dynamic_threshold, unet_cond_scale in zip(
                range(1, num_unets + 1), self.unets, self.image_sizes, self.noise_schedulers, self.pred_objectives,
                self.dynamic_thresholding, cond_scale):
            if unet_number < start_at_unet_number:
                continue
            assert not isinstance(unet, NullUnet), "one cannot sample from null / placeholder unets"
            lowres_cond_img = None
            if unet.lowres_cond:
                lowres_cond_img = self.normalize_img(resize_image_to(img, image_size)).contiguous()
                la, lsig = schedule.alpha_sigma(self.lowres_noise_schedule.noise_schedule, lowres_sample_noise_level)
                lowres_cond_img = K.q_sample(lowres_cond_img, noise("lowres_aug", tuple(lowres_cond_img.shape), device, unet=unet_number),
                                             float(la), float(lsig))
            shape = (batch_size, self.channels, image_size, image_size)
            img = self.p_sample_loop(
                unet_number, shape, noise=noise, lowres_cond_img=lowres_cond_img, lowres_noise_level=lowres_sample_noise_level,
                text_embeds=text_embeds, text_mask=text_masks, cond_images=cond_images, inpaint_images=inpaint_images,
                inpaint_masks=inpaint_masks, inpaint_resample_times=inpaint_resample_times, cond_scale=unet_cond_scale)
            outputs.append(img)
            if exists(stop_at_unet_number) and stop_at_unet_number == unet_number:
                break
        out = outputs[-1] if not return_all_unet_outputs else outputs
        if not return_pil_images:
            return out
        from torchvision.transforms import ToPILImage

        pil = [[ToPILImage()(i) for i in o.cpu().unbind(0)] for o in (out if return_all_unet_outputs else [out])]
        return pil if return_all_unet_outputs else pil[0]
