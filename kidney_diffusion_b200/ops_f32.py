"""fp32 ("precise") counterparts of the activation ops of ``ops`` (same names and argument meaning, fp32 NHWC tensors), backed
by ``csrc/kd_precise.cu``.  ``UnetExecutor`` swaps this namespace in when ``Unet.precision == "fp32"``: the forward code is
the same, only the kernels differ (fp32 weights, fp32 storage, fp32 FFMA accumulation on the CUDA cores; no fused GroupNorm
prologue / statistics / logits -- the executor falls back to its explicit passes).  This is the "fp32 path" of
BASELINE.json's north_star (per-step parity 1e-4 against the fp32 reference); it is 20-50x slower than the tensor-core path
and is never selected implicitly.  The conditioning towers are fp32 in both modes and are shared (``linear_small``,
``sinu_emb``, ``layernorm_f32``, ``axpby``).  torch is used for memory and for layout glue (cat / permute), not for compute.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import check
from .ops import (ACT_GELU, ACT_NONE, ACT_SIGMOID, ACT_SILU, _count, _ptr, _stream, _timed, axpby, layernorm_f32, lib, linear_small,  # noqa: F401
                  sinu_emb)

ACT_DTYPE = torch.float32
F32 = torch.float32


def _chk(t, name):
    assert t.dtype == F32 and t.is_cuda and t.is_contiguous(), f"{name}: expected a contiguous fp32 CUDA tensor"


def conv_pre_supported(*args, **kwargs):
    return False  # the GroupNorm + SiLU prologue is a tensor-core-path fusion; here the executor runs groupnorm() first


def stats_of(x):
    return None


@_timed
def conv_gemm(xa, w, bias=None, xb=None, *, mode=0, ksize=3, act=ACT_NONE, out_mode=0, out_f32=True, addend=None, addend_scale=None,
              out=None, want_stats=False, logit_w=None, pre_coef=None, **_):
    """Same contract as ops.conv_gemm on fp32 tensors.  mode 1 = Downsample (pixel-unshuffle + 1x1 conv = 2x2 conv, stride 2)."""
    assert pre_coef is None
    _chk(xa, "xa")
    _chk(w, "w")
    B, Hin, Win, Ca = xa.shape
    Cb = 0
    if xb is not None:
        _chk(xb, "xb")
        assert xb.shape[:3] == xa.shape[:3]
        Cb = xb.shape[3]
    if mode == 1:
        ks, stride, pad = 2, 2, 0
    else:
        ks, stride, pad = ksize, 1, ksize // 2
    Cout = w.shape[0]
    assert w.shape[1] == ks * ks * (Ca + Cb), f"weight K {w.shape[1]} != {ks * ks}*({Ca}+{Cb})"
    H, W = (Hin // 2, Win // 2) if mode == 1 else (Hin, Win)
    oshape = (B, 2 * H, 2 * W, Cout // 4) if out_mode == 1 else (B, H, W, Cout)
    if out is None:
        out = torch.empty(oshape, device=xa.device, dtype=F32)
    else:
        _chk(out, "out")
        assert tuple(out.shape) == oshape
    if addend is not None:
        _chk(addend, "addend")
        assert addend.shape == out.shape
    if addend_scale is not None:
        _chk(addend_scale, "addend_scale")
        assert addend_scale.shape == (B, Cout)
    check(lib().kd_conv_f32(_ptr(xa), Ca, _ptr(xb), Cb, _ptr(w), _ptr(bias), _ptr(addend), _ptr(addend_scale), _ptr(out), B, Hin, Win, Cout,
                            ks, stride, pad, act, out_mode, _stream()), "kd_conv_f32")
    _count()
    return out


@_timed
def groupnorm(xa, xb, b_scale, G, gamma, beta, scale_shift, act=ACT_SILU, eps=1e-5):
    """act(GroupNorm([xa | b_scale * xb]) * (scale + 1) + shift) -> (ya, yb), the two channel ranges of the result."""
    _chk(xa, "xa")
    B, H, W, Ca = xa.shape
    Cb = xb.shape[3] if xb is not None else 0
    C, HW = Ca + Cb, H * W
    gs = C // G
    nch = lib().kd_gn_chunks_f32(HW)
    partial = torch.empty((B, G, nch, 2), device=xa.device, dtype=torch.float64)
    check(lib().kd_gn_stats_f32(_ptr(xa), Ca, _ptr(xb), Cb, float(b_scale), B, HW, G, _ptr(partial), _stream()), "kd_gn_stats_f32")
    mr = torch.empty((B, G, 2), device=xa.device, dtype=F32)
    check(lib().kd_gn_finalize_f32(_ptr(partial), B, HW, G, gs, eps, _ptr(mr), _stream()), "kd_gn_finalize_f32")
    ss_stride = 0
    if scale_shift is not None:
        assert scale_shift.dtype == F32 and scale_shift.stride(-1) == 1 and scale_shift.shape == (B, 2 * C)
        ss_stride = scale_shift.stride(0)
    outs = []
    for x, off, sc in ((xa, 0, 1.0), (xb, Ca, float(b_scale))):
        if x is None:
            outs.append(None)
            continue
        y = torch.empty_like(x)
        check(lib().kd_gn_apply_f32(_ptr(x), _ptr(y), B, HW, x.shape[3], off, gs, G, sc, _ptr(mr), _ptr(gamma), _ptr(beta), _ptr(scale_shift),
                                    ss_stride, C, act, _stream()), "kd_gn_apply_f32")
        outs.append(y)
    _count(2 + (2 if xb is not None else 1))
    return outs[0], outs[1]


@_timed
def rowdot(x, w, bias):
    _chk(x, "x")
    B, H, W, C = x.shape
    out = torch.empty((B, H * W), device=x.device, dtype=F32)
    check(lib().kd_rowdot_f32(_ptr(x), _ptr(w), _ptr(bias), _ptr(out), B * H * W, C, _stream()), "kd_rowdot_f32")
    _count()
    return out


@_timed
def gca_gate(x, logits, w0, b0, w1, b1):
    """GlobalContext: softmax of the to_k logits over pixels, pooled = softmax @ x, gate = sigmoid(W1 silu(W0 pooled + b0) + b1)."""
    _chk(x, "x")
    B, H, W, C = x.shape
    HW = H * W
    nch = lib().kd_gn_chunks_f32(HW)
    ml = torch.empty((B, 2), device=x.device, dtype=F32)
    part = torch.empty((B, nch, C), device=x.device, dtype=F32)
    pooled = torch.empty((B, C), device=x.device, dtype=F32)
    check(lib().kd_softmax_pool_f32(_ptr(x), _ptr(logits), B, HW, C, _ptr(ml), _ptr(part), _ptr(pooled), _stream()), "kd_softmax_pool_f32")
    _count(3)
    return linear_small(linear_small(pooled, w0, b0, post_act=ACT_SILU), w1, b1, post_act=ACT_SIGMOID)


@_timed
def gate_residual(h, gate, res, want_stats=False):
    _chk(h, "h")
    _chk(res, "res")
    B, H, W, C = h.shape
    out = torch.empty_like(h)
    check(lib().kd_gate_residual_f32(_ptr(h), _ptr(gate), _ptr(res), _ptr(out), B, H * W, C, _stream()), "kd_gate_residual_f32")
    _count()
    return out


def layernorm_h16(x, g, bias=None, residual=None, eps=1e-5):
    """imagen-pytorch LayerNorm over the channel axis (gain only) of an NHWC fp32 map, then `+ residual` (name kept from ops)."""
    y = layernorm_f32(x, g, bias, eps)
    return y if residual is None else axpby(y, residual, 1.0, 1.0)


def _attn(q, ldq, k, v, B, N, J, heads, scale, per_head):
    out = torch.empty((B, N, heads * 64), device=q.device, dtype=F32)
    assert k.stride(-1) == 1 and v.stride(-1) == 1
    check(lib().kd_attn_f32(_ptr(q), ldq, _ptr(k), k.stride(1), k.stride(0), 64 if per_head else 0, _ptr(v), v.stride(1), v.stride(0),
                            64 if per_head else 0, _ptr(out), B, N, J, heads, float(scale), _stream()), "kd_attn_f32")
    _count()
    return out


def kv_assemble(qkv, kv_col, ctx_kv, null_kv):
    """[context kv | null kv | the map's own kv] along the key axis (Attention.forward: null prepended, then context), fp32 [B, J, 128]."""
    B = qkv.shape[0]
    parts = []
    if ctx_kv is not None:
        parts.append(ctx_kv)
    parts += [null_kv.reshape(1, 1, 128).expand(B, 1, 128), qkv[:, :, kv_col:kv_col + 128]]
    return torch.cat(parts, 1).contiguous()


@_timed
def attn_mqa(q, kv, heads, scale):
    """q: fp32 [B,N,ld] (first heads*64 columns are the queries); kv: fp32 [B,J,128] (k | v, one head shared by all query heads)."""
    B, N, ld = q.shape
    return _attn(q, ld, kv[:, :, :64], kv[:, :, 64:], B, N, kv.shape[1], heads, scale, per_head=False)


@_timed
def attn_cross(q, kv, null_kv, heads, scale):
    """q: fp32 [B,N,heads*64]; kv: fp32 [B,Jc,2*heads*64] (k | v per head); null_kv fp32 [2,64] prepended for every head."""
    B, N, ld = q.shape
    inner = heads * 64
    null_row = torch.cat((null_kv[0].repeat(heads), null_kv[1].repeat(heads))).view(1, 1, 2 * inner).expand(B, 1, 2 * inner)
    tokens = torch.cat((null_row, kv), 1).contiguous()
    return _attn(q, ld, tokens[:, :, :inner], tokens[:, :, inner:], B, N, tokens.shape[1], heads, scale, per_head=True)


def final_conv(xa, xb, w, bias):
    """xa: NHWC fp32; xb: NCHW fp32 or None; w: fp32 [Cout,3,3,Ca+Cb] -> NCHW fp32."""
    xb_h = None if xb is None else xb.permute(0, 2, 3, 1).contiguous()
    y = conv_gemm(xa, w.reshape(w.shape[0], -1), bias, xb=xb_h, ksize=3)
    return y.permute(0, 3, 1, 2).contiguous()


def init_conv_nchw(img, ksize, w, bias, addend, out):
    """CrossEmbedLayer on an NCHW fp32 image with the merged ksize x ksize filter w [dim, ksize*ksize*C]."""
    return conv_gemm(img.permute(0, 2, 3, 1).contiguous(), w, bias, ksize=ksize, addend=addend, out=out)


@_timed
def dwconv3x3(x, w):
    """Depthwise 3x3 convolution (zero padding, no bias): x NHWC fp32 [B,H,W,C], w fp32 [C,3,3]."""
    _chk(x, "x")
    _chk(w, "w")
    B, H, W, C = x.shape
    assert w.shape == (C, 3, 3)
    y = torch.empty_like(x)
    check(lib().kd_dwconv3x3_f32(_ptr(x), _ptr(w), _ptr(y), B, H, W, C, _stream()), "kd_dwconv3x3_f32")
    _count()
    return y


@_timed
def linear_attention(qkv, heads, scale, ctx_kv=None, act=ACT_SILU, pixels_kv=True):
    """Same contract as ops.linear_attention on fp32 tensors: qkv [B, N, 3*heads*64] (q | k | v), or only q with pixels_kv=False;
    ctx_kv fp32 [B, J, 2*heads*64] (k | v of the context tokens) -> [B, N, heads*64]."""
    _chk(qkv, "qkv")
    B, N, ld = qkv.shape
    inner = heads * 64
    assert ld == (3 * inner if pixels_kv else inner)
    parts = []
    if pixels_kv:
        parts.append(qkv[:, :, inner:])
    if ctx_kv is not None:
        assert ctx_kv.shape[0] == B and ctx_kv.shape[2] == 2 * inner
        parts.append(ctx_kv)
    assert parts, "linear attention needs keys: pixels, context tokens or both"
    kv = parts[0] if len(parts) == 1 and parts[0].is_contiguous() else torch.cat(parts, 1).contiguous()  # [B, J, 2*inner]: k | v, layout glue
    J = kv.shape[1]
    ctx = torch.empty((B, heads, 64, 64), device=qkv.device, dtype=F32)
    out = torch.empty((B, N, inner), device=qkv.device, dtype=F32)
    k, v = kv[:, :, :inner], kv[:, :, inner:]
    check(lib().kd_linattn_f32(_ptr(qkv), ld, _ptr(k), _ptr(v), kv.stride(1), kv.stride(0), B, N, J, heads, float(scale), act, _ptr(ctx), _ptr(out),
                               _stream()), "kd_linattn_f32")
    _count(2)
    return out
