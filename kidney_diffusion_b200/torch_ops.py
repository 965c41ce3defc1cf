"""PyTorch custom-op registration of the C-ABI kernels (`torch.ops.kidney_b200.*`).

BASELINE.json north_star: "Python host code calls hand-written CUDA built for sm_100a through a thin C-ABI layer exposed as
PyTorch custom ops".  Each operator below is defined with a schema in the `kidney_b200` namespace and implemented for the CUDA
dispatch key ONLY by the ctypes wrapper of `ops.py` (which enqueues the kernel of libkidney_b200.so on the current stream).
There is deliberately no CPU / Meta implementation: calling an op with CPU tensors fails in the dispatcher ("no kernel for CPU")
-- the path has no fallback.  The sampler (`imagen.StageRun`) and the patch-grid code call these ops; the UNet executor calls
`ops.py` directly (its conv wrapper attaches fused side outputs to the result, which a schema cannot express).
"""
from __future__ import annotations

import torch

from . import ops

_LIB = torch.library.Library("kidney_b200", "DEF")
REGISTERED = []


def _reg(schema, fn):
    name = schema.split("(")[0]
    _LIB.define(schema)
    _LIB.impl(name, fn, "CUDA")
    REGISTERED.append(name)


def _u64(v):
    return int(v) & (2 ** 64 - 1)


def _strip(t):
    """strided [3, rows, cols] view -> (tensor, channel stride, row stride) as ops.border_pack expects."""
    if t is None:
        return None
    assert t.dim() == 3 and t.stride(2) == 1
    return (t, t.stride(0), t.stride(1))


def _ddpm_step(x_t, pred, noise, s, objective, alpha, sigma, one_minus_c, c, alpha_next, std, renoise, rn_k1, rn_num, rn_alpha):
    sc = dict(alpha=alpha, sigma=sigma, one_minus_c=one_minus_c, c=c, alpha_next=alpha_next, std=std)
    return ops.ddpm_step(x_t, pred, noise, s, objective, sc, renoise=renoise, rn=(rn_k1, rn_num, rn_alpha))


def _inpaint_blend(img, inpaint, mask, noise, alpha, sigma):
    ops.inpaint_blend(img, inpaint, mask, noise, alpha, sigma)


def _finalize_image(img, inpaint, mask):
    ops.finalize_image(img, inpaint, mask)


def _border_pack(S, overlap_pos, orientation, above, side, corner, like):
    return ops.border_pack(S, overlap_pos, orientation, _strip(above), _strip(side), _strip(corner), like.device)


def _conv2d_nhwc(xa, w, bias, xb, mode, ksize, act, out_mode, out_f32, addend, addend_scale):
    out = ops.conv_gemm(xa, w, bias, xb=xb, mode=mode, ksize=ksize, act=act, out_mode=out_mode, out_f32=out_f32, addend=addend,
                        addend_scale=addend_scale)
    return out


_reg("conv2d_nhwc(Tensor xa, Tensor w, Tensor? bias, Tensor? xb, int mode, int ksize, int act, int out_mode, bool out_f32, Tensor? addend, "
     "Tensor? addend_scale) -> Tensor", _conv2d_nhwc)
_reg("linear_small(Tensor x, Tensor w, Tensor? bias, int pre_act, int post_act) -> Tensor",
     lambda x, w, bias, pre_act, post_act: ops.linear_small(x, w, bias, pre_act=pre_act, post_act=post_act))
_reg("layernorm_h16(Tensor x, Tensor g, Tensor? bias, Tensor? residual, float eps) -> Tensor",
     lambda x, g, bias, residual, eps: ops.layernorm_h16(x, g, bias, residual, eps))
_reg("attn_mqa(Tensor q, Tensor kv, int heads, float scale) -> Tensor", lambda q, kv, heads, scale: ops.attn_mqa(q, kv, heads, scale))
_reg("attn_cross(Tensor q, Tensor kv, Tensor null_kv, int heads, float scale) -> Tensor",
     lambda q, kv, null_kv, heads, scale: ops.attn_cross(q, kv, null_kv, heads, scale))
_reg("final_conv(Tensor xa, Tensor? xb, Tensor w, Tensor bias) -> Tensor", lambda xa, xb, w, bias: ops.final_conv(xa, xb, w, bias))
_reg("dynthresh(Tensor x_t, Tensor pred, str objective, float alpha, float sigma, float q, Tensor? workspace) -> Tensor",
     lambda x_t, pred, objective, alpha, sigma, q, workspace: ops.dynthresh(x_t, pred, objective, alpha, sigma, q, workspace))
_reg("ddpm_step(Tensor x_t, Tensor pred, Tensor noise, Tensor? s, str objective, float alpha, float sigma, float one_minus_c, float c, "
     "float alpha_next, float std, Tensor? renoise, float rn_k1, float rn_num, float rn_alpha) -> Tensor", _ddpm_step)
_reg("inpaint_blend(Tensor(a!) img, Tensor inpaint, Tensor mask, Tensor? noise, float alpha, float sigma) -> ()", _inpaint_blend)
_reg("finalize_image(Tensor(a!) img, Tensor? inpaint, Tensor? mask) -> ()", _finalize_image)
_reg("q_sample(Tensor x0, Tensor noise, float alpha, float sigma) -> Tensor", lambda x0, noise, alpha, sigma: ops.q_sample(x0, noise, alpha, sigma))
_reg("randn_like(Tensor like, int seed, int key) -> Tensor",
     lambda like, seed, key: ops.randn(tuple(like.shape), _u64(seed), _u64(key), like.device))
_reg("axpby(Tensor x, Tensor y, float a, float b) -> Tensor", lambda x, y, a, b: ops.axpby(x, y, a, b))
_reg("border_pack(int S, int overlap_pos, int orientation, Tensor? above, Tensor? side, Tensor? corner, Tensor like) -> (Tensor, Tensor)", _border_pack)
_reg("cond_gather(Tensor zoomed, Tensor(a!) out, int off, int shift_y, int shift_x, float fill, int patch_width, int center_top) -> ()",
     lambda zoomed, out, off, shift_y, shift_x, fill, patch_width, center_top: ops.cond_gather(zoomed, out, off, shift_y, shift_x, fill, patch_width,
                                                                                                center_top) and None)
