"""Host-side decisions of the C-ABI library that need no GPU: kernel / tiling choice, statistics-buffer geometry, reduction
block counts (which must not depend on the batch size -- that is what keeps 1/2/4/8-GPU grid runs bit-identical), weight
panel sizes.  Pure host functions of libkidney_b200.so; no kernel is launched."""
import ctypes

import pytest


@pytest.fixture(scope="module")
def lib():
    from kidney_diffusion_b200 import _lib
    from kidney_diffusion_b200.build import build_library

    build_library()
    return _lib.load()


def _layout(lib, **kw):
    from kidney_diffusion_b200._lib import KdConvDesc

    d = dict(mode=0, B=1, H=64, W=64, Ca=128, Cb=0, Cout=128, ksize=3, act=0, out_mode=0, out_f32=0, addend_f32=0)
    d.update(kw)
    desc = KdConvDesc(*[d[k] for k in ("mode", "B", "H", "W", "Ca", "Cb", "Cout", "ksize", "act", "out_mode", "out_f32", "addend_f32")])
    lay = (ctypes.c_int * 4)()
    assert lib.kd_conv_stats_layout(ctypes.byref(desc), lay) == 0
    return list(lay)


def test_conv_kernel_choice_and_stats_geometry(lib):
    # 3x3 on >= 16 x 8 images, Cout >= 128: halo kernel (fused pre-activation supported), 16 x 8 pixel tiles, one image per tile
    rows, tiles, tb, pre = _layout(lib, B=3, H=64, W=64)
    assert pre == 1 and tb == 1 and tiles == (64 // 16) * (64 // 8) and rows == 3 * tiles * 4
    # partial tiles round up
    rows, tiles, tb, pre = _layout(lib, B=1, H=40, W=24)
    assert pre == 1 and tiles == 3 * 3 and rows == tiles * 4
    # 1x1 conv and tiny images: CTA-pair tap-loop kernel (statistics yes, pre-activation no)
    assert _layout(lib, ksize=1)[3] == 0 and _layout(lib, ksize=1)[0] > 0
    rows, tiles, tb, pre = _layout(lib, B=4, H=8, W=8)
    assert pre == 0 and tb == 2 and rows == 2 * tiles * 4  # 128-pixel tile = two 8x8 images
    # narrow layers (Cout < 128) run the single-CTA kernel: no fused side outputs
    assert _layout(lib, Cout=64)[0] == 0 and _layout(lib, Cout=64)[3] == 0
    # fp32 output and pixel-shuffle output do not emit statistics
    assert _layout(lib, out_f32=1)[0] == 0
    assert _layout(lib, ksize=1, Cout=512, out_mode=1)[0] == 0
    # the 2x2 stride-2 downsample taps do
    assert _layout(lib, mode=1, H=32, W=32, Cout=256)[0] > 0


def test_reduction_block_counts_are_batch_independent_and_bounded(lib):
    for hw, c in ((1024 * 1024, 128), (512 * 512, 128), (64 * 64, 1024), (8 * 8, 1024), (16, 64)):
        n = lib.kd_elementwise_blocks(hw, c)
        assert 1 <= n <= 148 * 4
        oct_ = c // 8
        lanes = max(1, (256 // oct_ if oct_ < 256 else 1))
        assert n * 16 * lanes >= hw or n == 148 * 4 or n == 1 or n * 16 * lanes + 16 * lanes > hw  # >= 16 pixels per lane unless capped
    assert lib.kd_elementwise_blocks(0, 128) == 0 and lib.kd_elementwise_blocks(64, 12) == 0
    # octet-reduce splits: a function of the partial-row count of ONE image only
    assert lib.kd_oct_reduce_splits(4, 8192, 1) == 144          # 1024^2 conv output: 32768 rows -> capped at 144
    assert lib.kd_oct_reduce_splits(4, 32, 1) == 2              # 64 x 64: 128 rows -> 2 splits
    assert lib.kd_oct_reduce_splits(4, 1, 2) == 1
    assert lib.kd_oct_reduce_splits(0, 1, 1) == 0


def test_init_conv_panel_width_and_final_conv_pack(lib):
    # K ordered (ky, c, kx padded to 16), rounded up to 64-wide chunks of four (ky, c) rows
    assert lib.kd_init_conv_kp(3, 15) == 768
    assert lib.kd_init_conv_kp(1, 15) == 256
    assert lib.kd_init_conv_kp(3, 3) == 192
    assert lib.kd_init_conv_kp(2, 7) == 256
    # hi + lo fp16 filter split: per 32-channel chunk 2 x 9 taps x 8 rows x 40 (32 + padding)
    assert lib.kd_final_conv_pack_elems(128) == 4 * 2 * 9 * 8 * 40


def test_argument_validation_reports_through_kd_last_error(lib):
    from kidney_diffusion_b200._lib import KdConvDesc

    bad = KdConvDesc(0, 1, 16, 16, 100, 0, 128, 3, 0, 0, 0, 0)  # Ca not a multiple of 64
    rc = lib.kd_conv_gemm(ctypes.byref(bad), ctypes.c_void_p(16), None, ctypes.c_void_p(16), None, None, None, ctypes.c_void_p(16), None)
    assert rc != 0
    assert b"multiples of 64" in lib.kd_last_error()
    assert lib.kd_set_conv_impl(3) != 0 and lib.kd_set_conv_impl(7) != 0 and lib.kd_set_conv_impl(0) == 0


def test_precise_namespace_covers_every_op_the_executor_uses():
    """UnetExecutor runs the same forward code over `ops` (fp16 tensor-core path) or `ops_f32` (precise path): every op it
    reaches through `self.K` must exist in both namespaces, except the tensor-core-only fast paths it guards with `self.precise`."""
    import inspect
    import re

    from kidney_diffusion_b200 import ops, ops_f32, unet_exec

    src = inspect.getsource(unet_exec.UnetExecutor)
    used = set(re.findall(r"self\.K\.([A-Za-z_0-9]+)", src))
    fast_only = {"init_conv", "init_conv_kp", "im2col_nchw", "gemm_rows", "gn_finalize_oct", "gn_apply"}   # behind `if self.precise: ... return`
    precise_only = {"groupnorm", "init_conv_nchw"}
    assert fast_only <= used and precise_only <= used
    for name in sorted(used - precise_only):
        assert hasattr(ops, name), f"ops.{name} missing"
    for name in sorted(used - fast_only):
        assert hasattr(ops_f32, name), f"ops_f32.{name} missing"
    assert ops_f32.ACT_DTYPE.is_floating_point and ops_f32.ACT_DTYPE.itemsize == 4 and ops.ACT_DTYPE.itemsize == 2
    assert not ops_f32.conv_pre_supported(1, 64, 64, 128, 0, 128)


def test_precision_switch_on_the_host_side():
    import torch

    from kidney_diffusion_b200 import Imagen, Unet

    u = Unet(dim=64, dim_mults=(1, 2), num_resnet_blocks=1, layer_attns=False, layer_cross_attns=False, cond_on_text=False, text_embed_dim=None)
    assert Unet.precision == "fp16" and u.precision == "fp16"
    im = Imagen(unets=(u,), image_sizes=(32,), timesteps=(2,), condition_on_text=False)
    u = im.unets[0]
    im._graphs["stale"] = object()
    assert im.set_precision("fp32") is im
    assert u.precision == "fp32" and im._graphs == {} and Unet.precision == "fp16"   # per instance, captured graphs dropped
    with pytest.raises(AssertionError):
        im.set_precision("bf16")
    with pytest.raises(RuntimeError, match="no CPU fallback"):   # either precision needs the GPU: nothing falls back to torch
        u(torch.zeros(1, 3, 32, 32), torch.zeros(1))
