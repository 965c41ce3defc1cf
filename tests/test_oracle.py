"""Closed-form anchors that pin the oracle restatement of imagen-pytorch 1.18.5 (the dependency itself is not installable
here and the reference ships no tests: "parity unpinned") and the product's host-side schedule code."""
import math

import pytest
import torch

from kidney_diffusion_b200 import schedule
from oracle import imagen_oracle as O


def test_schedule_values_survey_8c():
    t = lambda v: torch.tensor([v])
    cos, lin = O.alpha_cosine_log_snr, O.beta_linear_log_snr
    assert abs(float(cos(t(0.0))) - 8.7692) < 2e-3 and abs(float(cos(t(0.2))) - 2.1814) < 1e-3
    assert abs(float(cos(t(0.5))) + 0.0249) < 1e-3 and abs(float(cos(t(1.0))) + 33.891) < 2e-2
    assert abs(float(lin(t(0.0))) - 9.2103) < 1e-3 and abs(float(lin(t(0.2))) - 0.7093) < 1e-3 and abs(float(lin(t(1.0))) + 10.0001) < 1e-3
    a, s = O.log_snr_to_alpha_sigma(lin(t(0.2)))
    assert abs(float(a) - 0.81869) < 1e-4 and abs(float(s) - 0.57424) < 1e-4
    ts = torch.linspace(0, 1, 101)
    for f in (cos, lin):
        ls = f(ts)
        assert bool((ls[1:] < ls[:-1]).all())  # monotone decreasing
        a, s = O.log_snr_to_alpha_sigma(ls)
        assert float((a ** 2 + s ** 2 - 1).abs().max()) < 1e-6


def test_q_posterior_identities_and_v_inversion():
    sched = O.GaussianDiffusionContinuousTimes(noise_schedule="cosine", timesteps=100)
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(2, 3, 8, 8, generator=g), torch.randn(2, 3, 8, 8, generator=g)
    t = torch.full((2,), 0.6)
    xt, _, alpha, sigma = sched.q_sample(x0, t, eps)
    mean, var, _ = sched.q_posterior(x0, xt, t, t_next=t)  # t_next == t -> c = 0 -> mean = x_t, var = 0
    assert torch.allclose(mean, xt, atol=1e-6) and float(var.abs().max()) < 1e-12
    mean0, _, _ = sched.q_posterior(x0, xt, t, t_next=torch.zeros(2))  # t_next -> 0: x0 dominated
    assert float((mean0 - x0).abs().max()) < 0.05
    v = alpha * eps - sigma * x0
    assert torch.allclose(sched.predict_start_from_v(xt, t, v), x0, atol=1e-5)
    assert torch.allclose(sched.predict_start_from_noise(xt, t, eps), x0, atol=1e-4)


def test_product_schedule_matches_oracle_bitwise():
    for name in ("cosine", "linear"):
        sched = O.GaussianDiffusionContinuousTimes(noise_schedule=name, timesteps=50)
        for (t, tn), (ot, otn) in zip(schedule.sampling_times(50), sched.get_sampling_timesteps(1, device="cpu")):
            assert float(t) == float(ot) and float(tn) == float(otn)
            sc = schedule.step_scalars(name, t, tn)
            ls = sched.log_snr(ot)
            a, s = O.log_snr_to_alpha_sigma(ls)
            assert sc["log_snr"] == float(ls) and sc["alpha"] == float(a) and sc["sigma"] == float(s)
            _, var, logvar = sched.q_posterior(torch.zeros(1, 1), torch.zeros(1, 1), ot, t_next=otn)
            std = float((1 - (otn == 0).float()) * (0.5 * logvar).exp())
            assert sc["std"] == std


def test_fresh_unet_outputs_zero_and_key_inventory():
    torch.manual_seed(0)
    kw = dict(dim=32, dim_mults=(1, 2), num_resnet_blocks=1, layer_attns=(False, True), layer_cross_attns=(False, True))
    u = O.Unet(**kw, cond_on_text=False, text_embed_dim=None)
    out = u(torch.randn(2, 3, 16, 16), torch.tensor([1.0, -2.0]))
    assert out.shape == (2, 3, 16, 16) and float(out.abs().max()) == 0.0  # final_conv is zero-initialised
    keys = list(u.state_dict().keys())
    for expect in ("init_conv.convs.2.weight", "to_time_hiddens.0.weights", "downs.1.1.cross_attn.fn.null_kv", "downs.0.2.0.gca.to_k.weight",
                   "downs.1.3.layers.0.0.fn.to_context.1.weight", "downs.1.3.layers.0.1.4.weight", "mid_block1.time_mlp.1.weight",
                   "ups.0.3.net.0.weight", "final_conv.bias", "downs.1.4.fns.0.weight"):
        assert expect in keys, expect


def test_sample_range_shape_and_inpaint_round_trip():
    torch.manual_seed(1)
    kw = dict(dim=32, dim_mults=(1, 2), num_resnet_blocks=1, layer_attns=False, layer_cross_attns=(False, True))
    im = O.Imagen(unets=(O.Unet(**kw),), image_sizes=(16,), timesteps=(3,), condition_on_text=False)
    O.randomize_zero_init_(im)
    inp, mask = torch.rand(2, 3, 16, 16), torch.zeros(2, 16, 16)
    mask[:, :4] = 1
    out = im.sample(batch_size=2, inpaint_images=inp, inpaint_masks=mask, inpaint_resample_times=2)
    assert out.shape == (2, 3, 16, 16) and float(out.min()) >= 0 and float(out.max()) <= 1
    m = mask.bool()[:, None].expand_as(inp)
    assert float((out[m] - inp[m]).abs().max()) < 1e-6


def test_state_dict_parity_between_oracle_and_product():
    from kidney_diffusion_b200 import Unet

    torch.manual_seed(2)
    for kw in (dict(dim=64, dim_mults=(1, 2, 3, 4), num_resnet_blocks=2, layer_attns=(False, True, True, True), layer_cross_attns=(False, True, True, True)),
               dict(dim=64, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(1, 2, 2, 2), memory_efficient=True, layer_attns=False,
                    layer_cross_attns=(False, False, False, True), init_conv_to_final_conv_residual=True, cond_images_channels=3, lowres_cond=True),
               dict(dim=64, dim_mults=(1, 2), cond_dim=128, text_embed_dim=3, num_resnet_blocks=1, cond_images_channels=4)):
        a, b = O.Unet(**kw).state_dict(), Unet(**kw).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)


def test_restore_parts():
    from kidney_diffusion_b200.trainer import restore_parts

    tgt = {"a": torch.zeros(2), "b": torch.zeros(3)}
    out = restore_parts(tgt, {"a": torch.ones(2), "b": torch.ones(4), "c": torch.ones(1)})
    assert out["a"].tolist() == [1, 1] and out["b"].tolist() == [0, 0, 0]
