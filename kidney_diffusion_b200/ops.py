"""Thin tensor-level wrappers over the C ABI (one Python function per entry point of include/kidney_b200.h).

PyTorch is used only for device memory and streams: every function checks dtype / layout, allocates the output with
``torch.empty`` and enqueues the hand-written CUDA kernel on the current stream.  Nothing here computes with torch ops.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib
from ._lib import KdConvDesc, KdConvFusion, check

ACT_NONE, ACT_SILU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3
ACT_DTYPE = torch.float16  # 16-bit activation / weight dtype of the CUDA path (kd_common.cuh: h16), fp32 accumulation
PRED = {"noise": 0, "v": 1, "x_start": 2}

launch_count = 0  # kernels launched through this module (bench.py reports it as gpu_launches)
conv_profile = None  # bench.py sets this to a list: every tensor-core GEMM launch appends (flops, start_event, end_event)


op_profile = None  # bench / profiling: list of (op name, start_event, end_event, tensor bytes in + out) for EVERY op when set
sat_counter = None  # overflow guard: int64 device tensor [1]; while set, every op that stores an fp16 activation tensor counts clipped values


def _timed(fn):
    """CUDA-event bracket around an op while `op_profile` is a list (no cost otherwise)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        if sat_counter is not None:
            out = fn(*a, **k)
            for t in (out if isinstance(out, (tuple, list)) else [out]):
                if isinstance(t, torch.Tensor) and t.dtype == ACT_DTYPE and t.is_cuda and t.is_contiguous() and t.numel():
                    check(lib().kd_count_saturated(_ptr(t), t.numel(), _ptr(sat_counter), _stream()), "kd_count_saturated")
            return out
        if op_profile is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        nbytes = 0  # bytes of the distinct tensors the op touches once each: the HBM-traffic floor of the launch
        seen = set()
        for t in list(a) + list(k.values()) + (list(out) if isinstance(out, (tuple, list)) else [out]):
            if isinstance(t, torch.Tensor) and t.data_ptr() not in seen:
                seen.add(t.data_ptr())
                nbytes += t.numel() * t.element_size()
        op_profile.append((fn.__name__, e0, e1, nbytes))
        return out

    return wrapper


class _ConvTimer:
    """CUDA-event bracket around one kd_conv_gemm launch on the launching stream (only active while profiling)."""

    def __init__(self, flops, label=None):
        self.flops, self.label = flops, label

    def __enter__(self):
        if conv_profile is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if conv_profile is not None:
            self.e1.record()
            conv_profile.append((self.flops, self.e0, self.e1, self.label))
        return False


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise _lib.KdError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.KdError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.KdError(f"{name}: expected a contiguous tensor")


def _count(n=1):
    global launch_count
    launch_count += n


def lib():
    l = _lib.load()
    _lib.require_b200()
    return l


# ------------------------------------------------------------------------------------------------ K1 conv / GEMM
def set_conv_impl(impl):
    """0 = automatic (default), 1 = single-CTA 128x128 kernel, 2 = CTA-pair cta_group::2 kernel, 4 = pair kernel without halo,
    8 = automatic without split-K (tests / profiling)."""
    check(lib().kd_set_conv_impl(int(impl)), "kd_set_conv_impl")


class OctStats:
    """Per-(row group, channel octet) {sum, sumsq} partials of a tensor + their layout; `reduced()` -> [B, C/8, 2] fp32."""

    def __init__(self, partial, rpt, tiles, TB, B, n_oct):
        self.partial, self.rpt, self.tiles, self.TB, self.B, self.n_oct = partial, rpt, tiles, TB, B, n_oct
        self._reduced = None

    def reduced(self):
        """[B, NS, C/8, 2] split sums (NS <= 144 row ranges per image; kd_gn_finalize_oct adds them in fixed order)."""
        if self._reduced is None:
            self.ns = lib().kd_oct_reduce_splits(self.rpt, self.tiles, self.TB)
            out = torch.empty((self.B, self.ns, self.n_oct, 2), device=self.partial.device, dtype=torch.float32)
            check(lib().kd_oct_reduce(_ptr(self.partial), self.rpt, self.tiles, self.TB, self.B, self.n_oct, _ptr(out), _stream()),
                  "kd_oct_reduce")
            _count()
            self._reduced = out
        return self._reduced


@_timed
def oct_stats(x):
    """Standalone octet statistics of an NHWC fp16 tensor (used when the producer kernel could not fuse them)."""
    _chk(x, ACT_DTYPE, "x")
    B, H, W, C = x.shape
    nblk = _nblk(H * W, C, B)
    partial = torch.empty((B, nblk, C // 8, 2), device=x.device, dtype=torch.float32)
    check(lib().kd_oct_stats(_ptr(x), B, H * W, C, _ptr(partial), nblk, _stream()), "kd_oct_stats")
    _count()
    return OctStats(partial, 1, nblk, 1, B, C // 8)


FUSED_STATS = True  # debugging switch: False forces the standalone statistics pass
FUSED_REDUCE = True  # debugging switch: False runs kd_oct_reduce and kd_gn_finalize_oct as separate launches
FUSED_PRE = True    # debugging switch: False keeps the separate GroupNorm-apply pass in front of the 3x3 convolutions


def stats_of(x):
    """Octet statistics attached to `x` by its producer (conv epilogue / gate_residual), else computed standalone; cached."""
    st = getattr(x, "_kd_stats", None) if FUSED_STATS else getattr(x, "_kd_stats_alone", None)
    if st is None:
        st = oct_stats(x)
        if FUSED_STATS:
            x._kd_stats = st
        else:
            x._kd_stats_alone = st
    return st


@_timed
def gn_finalize_oct(stats_a, scale_a, stats_b, scale_b, group_size, num_groups, count, eps=1e-5, *, gamma=None, beta=None,
                    scale_shift=None, want_coef=False):
    """mean / rstd per (b, group); with want_coef also the per-channel affine [B, C, 2] = {A, B} of GroupNorm (+ time
    scale/shift) on the raw sources (input of conv_gemm(pre_coef=...)).  Returns mean_rstd or (mean_rstd, coef)."""
    B = stats_a.B
    dev = stats_a.partial.device
    mean_rstd = torch.empty((B, num_groups, 2), device=dev, dtype=torch.float32)
    coef = None
    ss_stride = 0
    if want_coef:
        C = 8 * (stats_a.n_oct + (0 if stats_b is None else stats_b.n_oct))
        coef = torch.empty((B, C, 2), device=dev, dtype=torch.float32)
        _chk(gamma, torch.float32, "gamma")
        _chk(beta, torch.float32, "beta")
        if scale_shift is not None:
            assert scale_shift.dtype == torch.float32 and scale_shift.stride(1) == 1 and scale_shift.shape[1] == 2 * C
            ss_stride = scale_shift.stride(0)
    extra = (_ptr(gamma) if want_coef else None, _ptr(beta) if want_coef else None, _ptr(scale_shift) if want_coef else None, ss_stride,
             _ptr(coef), _stream())
    fused = FUSED_REDUCE and stats_a._reduced is None and (stats_b is None or stats_b._reduced is None)
    if fused:  # reduce + finalize in one launch
        L = lib()
        n_a = L.kd_oct_reduce_splits(stats_a.rpt, stats_a.tiles, stats_a.TB) * stats_a.n_oct
        n_b = 0 if stats_b is None else L.kd_oct_reduce_splits(stats_b.rpt, stats_b.tiles, stats_b.TB) * stats_b.n_oct
        scratch = torch.empty((B * (n_a + n_b) * 2,), device=dev, dtype=torch.float32)
        sb_args = (None, 0, 0, 0, 0) if stats_b is None else (_ptr(stats_b.partial), stats_b.rpt, stats_b.tiles, stats_b.TB, stats_b.n_oct)
        check(L.kd_gn_reduce_finalize(_ptr(stats_a.partial), stats_a.rpt, stats_a.tiles, stats_a.TB, stats_a.n_oct, scale_a, *sb_args, scale_b,
                                      B, num_groups, group_size, float(count), eps, _ptr(scratch), _ptr(_arrival_counters(dev, B)),
                                      _ptr(mean_rstd), *extra), "kd_gn_reduce_finalize")
    else:
        sa = stats_a.reduced()
        sb = stats_b.reduced() if stats_b is not None else None
        check(lib().kd_gn_finalize_oct(_ptr(sa), stats_a.n_oct, stats_a.ns, scale_a, _ptr(sb), 0 if stats_b is None else stats_b.n_oct,
                                       0 if stats_b is None else stats_b.ns, scale_b, B, num_groups, group_size, float(count), eps,
                                       _ptr(mean_rstd), *extra), "kd_gn_finalize_oct")
    _count()
    return (mean_rstd, coef) if want_coef else mean_rstd


_COUNTERS = {}


def _arrival_counters(device, n):
    """Zeroed, self-resetting arrival counters of the 'last block finalizes' kernels (one per batch image)."""
    key = str(device)
    t = _COUNTERS.get(key)
    if t is None or t.numel() < n:
        t = _COUNTERS[key] = torch.zeros((max(4096, n),), device=device, dtype=torch.int32)
    return t


def conv_pre_supported(B, H, W, Ca, Cb, Cout, ksize=3):
    """True when kd_conv_gemm_fused can apply the GroupNorm affine + SiLU to its input itself (3x3 halo kernel)."""
    d = KdConvDesc(0, B, H, W, Ca, Cb, Cout, ksize, ACT_NONE, 0, 0, 0)
    lay = (ctypes.c_int * 4)()
    check(lib().kd_conv_stats_layout(ctypes.byref(d), lay), "kd_conv_stats_layout")
    return FUSED_PRE and lay[3] == 1


@_timed
def conv_gemm(xa, w, bias=None, xb=None, *, mode=0, ksize=3, out_hw=None, act=ACT_NONE, out_mode=0, out_f32=False,
              addend=None, addend_scale=None, out=None, want_stats=False, logit_w=None, pre_coef=None):
    """xa / xb: NHWC fp16 [B,H,W,C]; w: packed fp16 [Cout, taps*(Ca+Cb)]; returns NHWC (fp16 or fp32).
    want_stats / logit_w: ask the epilogue for fused GroupNorm statistics / GlobalContext logits of the output; they are
    attached to the result as `_kd_stats` / `_kd_logits` when this shape's kernel can emit them (else the consumer runs
    its standalone kernel)."""
    _chk(xa, ACT_DTYPE, "xa")
    _chk(w, ACT_DTYPE, "w")
    B, Hin, Win, Ca = xa.shape
    Cb = 0
    if xb is not None:
        _chk(xb, ACT_DTYPE, "xb")
        assert xb.shape[:3] == xa.shape[:3]
        Cb = xb.shape[3]
    H, W = (Hin // 2, Win // 2) if mode == 1 else (Hin, Win)
    Cout = w.shape[0]
    taps = 4 if mode == 1 else (ksize * ksize if mode == 0 else 1)
    assert w.shape[1] == taps * (Ca + Cb), f"weight K {w.shape[1]} != {taps}*({Ca}+{Cb})"
    if out is None:
        oshape = (B, 2 * H, 2 * W, Cout // 4) if out_mode == 1 else (B, H, W, Cout)
        out = torch.empty(oshape, device=xa.device, dtype=torch.float32 if out_f32 else ACT_DTYPE)
    addend_f32 = 0
    if addend is not None:
        assert addend.is_contiguous() and addend.shape == out.shape
        addend_f32 = 1 if addend.dtype == torch.float32 else 0
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    if addend_scale is not None:
        _chk(addend_scale, torch.float32, "addend_scale")
        assert addend_scale.shape == (B, Cout)
    d = KdConvDesc(mode, B, H, W, Ca, Cb, Cout, ksize, act, out_mode, 1 if out_f32 else 0, addend_f32)
    stats = None
    lay = None
    if (want_stats or logit_w is not None) and FUSED_STATS:
        lay = (ctypes.c_int * 4)()
        check(lib().kd_conv_stats_layout(ctypes.byref(d), lay), "kd_conv_stats_layout")
        if lay[0] > 0 and want_stats:
            stats = OctStats(torch.empty((lay[0], Cout // 8, 2), device=xa.device, dtype=torch.float32), 4, lay[1], lay[2], B, Cout // 8)
    logit_parts = None
    if logit_w is not None and lay is not None and lay[0] > 0 and out_mode == 0 and Cout % 64 == 0 and addend_scale is None:
        _chk(logit_w, torch.float32, "logit_w")
        logit_parts = torch.empty((Cout // 64, B, H * W), device=xa.device, dtype=torch.float32)
    with _ConvTimer(2.0 * B * H * W * Cout * taps * (Ca + Cb), (mode, B, H, W, Ca + Cb, Cout, ksize if mode == 0 else 2)):
        if pre_coef is not None:
            _chk(pre_coef, torch.float32, "pre_coef")
            assert pre_coef.shape == (B, Ca + Cb, 2)
        ws_bytes = lib().kd_conv_splitk_workspace_bytes(ctypes.byref(d))
        ws = torch.empty((ws_bytes,), device=xa.device, dtype=torch.uint8) if ws_bytes else None
        fz = KdConvFusion(None if stats is None else _ptr(stats.partial), None if logit_parts is None else _ptr(logit_w),
                          _ptr(logit_parts), _ptr(pre_coef), _ptr(ws), ws_bytes)
        check(lib().kd_conv_gemm_fused(ctypes.byref(d), _ptr(xa), _ptr(xb), _ptr(w), _ptr(bias), _ptr(addend), _ptr(addend_scale),
                                       _ptr(out), ctypes.byref(fz), _stream()), "kd_conv_gemm")
    if logit_parts is not None:
        out._kd_logits = logit_parts
    _count()
    if stats is not None:
        out._kd_stats = stats
    return out


@_timed
def gemm_rows(x, w, bias=None, *, act=ACT_NONE, out_f32=False, addend=None, out=None, algo_k=None):
    """Plain GEMM y[M,N] = act(x[M,K] @ w[N,K]^T + bias) + addend on tensor cores (mode 2)."""
    _chk(x, ACT_DTYPE, "x")
    _chk(w, ACT_DTYPE, "w")
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32 if out_f32 else ACT_DTYPE)
    else:
        assert out.is_contiguous() and out.shape == (M, N)
        out_f32 = out.dtype == torch.float32
    addend_f32 = 0
    if addend is not None:
        assert addend.is_contiguous() and addend.shape == (M, N)
        addend_f32 = 1 if addend.dtype == torch.float32 else 0
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    d = KdConvDesc(2, 1, 1, M, K, 0, N, 1, act, 0, 1 if out_f32 else 0, addend_f32)
    with _ConvTimer(2.0 * M * N * (algo_k if algo_k is not None else K), (2, 1, 1, M, K, N, 1)):  # algo_k: un-padded K (algorithmic FLOPs)
        check(lib().kd_conv_gemm(ctypes.byref(d), _ptr(x), None, _ptr(w), _ptr(bias), _ptr(addend), None, _ptr(out), _stream()),
              "kd_conv_gemm(mode 2)")
    _count()
    return out


# ------------------------------------------------------------------------------------------------ small linear / time embedding
@_timed
def linear_small(x, w, bias=None, *, pre_act=ACT_NONE, post_act=ACT_NONE, out=None, ldy=None):
    """fp32 y[M,N] = post(pre(x[M,K]) @ w[N,K]^T + bias); `out` may be a strided row view (ldy)."""
    _chk(w, torch.float32, "w")
    assert x.dtype == torch.float32 and x.is_cuda
    M, K = x.shape
    ldx = x.stride(0)
    assert x.stride(1) == 1
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32)
        ldy = N
    elif ldy is None:
        ldy = out.stride(0)
    check(lib().kd_linear_small(_ptr(x), M, K, ldx, _ptr(w), _ptr(bias), _ptr(out), N, ldy, pre_act, post_act, _stream()),
          "kd_linear_small")
    _count()
    return out


@_timed
def sinu_emb(t, weights):
    _chk(t, torch.float32, "t")
    _chk(weights, torch.float32, "weights")
    B, half = t.shape[0], weights.shape[0]
    out = torch.empty((B, 2 * half + 1), device=t.device, dtype=torch.float32)
    check(lib().kd_sinu_emb(_ptr(t), _ptr(weights), B, half, _ptr(out), _stream()), "kd_sinu_emb")
    _count()
    return out


# ------------------------------------------------------------------------------------------------ GroupNorm
def _nblk(HW, C, B):
    # blocks per image of the pixel-chunked reduction kernels; independent of the batch size on purpose: a sample's
    # reduction order (hence its bits) must not depend on which other patches share its batch (patch-grid invariance
    # across GPU counts)
    return int(lib().kd_elementwise_blocks(HW, C))


@_timed
def gn_stats(x, c_offset, group_size, num_groups):
    _chk(x, ACT_DTYPE, "x")
    B, H, W, C = x.shape
    HW = H * W
    nblk = _nblk(HW, C, B)
    partial = torch.empty((B, nblk, num_groups, 2), device=x.device, dtype=torch.float32)
    check(lib().kd_gn_stats(_ptr(x), B, HW, C, c_offset, group_size, num_groups, _ptr(partial), nblk, _stream()), "kd_gn_stats")
    _count()
    return partial


@_timed
def gn_finalize(partial_a, scale_a, partial_b, scale_b, count, eps=1e-5):
    B, nblk_a, G, _ = partial_a.shape
    nblk_b = 0 if partial_b is None else partial_b.shape[1]
    mean_rstd = torch.empty((B, G, 2), device=partial_a.device, dtype=torch.float32)
    check(lib().kd_gn_finalize(_ptr(partial_a), nblk_a, scale_a, _ptr(partial_b), nblk_b, scale_b, B, G, float(count), eps,
                               _ptr(mean_rstd), _stream()), "kd_gn_finalize")
    _count()
    return mean_rstd


@_timed
def gn_apply(x, mean_rstd, gamma, beta, *, c_offset, group_size, num_groups, src_scale=1.0, scale_shift=None, ctot=None,
             act=ACT_SILU):
    _chk(x, ACT_DTYPE, "x")
    B, H, W, C = x.shape
    y = torch.empty_like(x)
    ctot = ctot if ctot is not None else C
    ss_stride = 0
    if scale_shift is not None:
        # may be a column slice of the [B, sum(2C)] table produced by one kd_linear_small launch for all blocks
        assert scale_shift.dtype == torch.float32 and scale_shift.stride(-1) == 1 and scale_shift.shape == (B, 2 * ctot)
        ss_stride = scale_shift.stride(0)
    check(lib().kd_gn_apply(_ptr(x), _ptr(y), B, H * W, C, c_offset, group_size, num_groups, src_scale, _ptr(mean_rstd),
                            _ptr(gamma), _ptr(beta), _ptr(scale_shift), ss_stride, ctot, act, _stream()), "kd_gn_apply")
    _count()
    return y


# ------------------------------------------------------------------------------------------------ GlobalContext
@_timed
def rowdot(x, w, bias):
    _chk(x, ACT_DTYPE, "x")
    B, H, W, C = x.shape
    out = torch.empty((B, H * W), device=x.device, dtype=torch.float32)
    check(lib().kd_rowdot(_ptr(x), _ptr(w), _ptr(bias), _ptr(out), B, H * W, C, _stream()), "kd_rowdot")
    _count()
    return out


@_timed
def gca_pool(x, logits):
    """logits: [B, HW] or [n_parts, B, HW] partial logits (summed per pixel in fixed order inside the kernel)."""
    B, H, W, C = x.shape
    HW = H * W
    n_parts = logits.shape[0] if logits.dim() == 3 else 1
    nblk = _nblk(HW, C, B)
    part = torch.empty((B, nblk, C), device=x.device, dtype=torch.float32)
    ml = torch.empty((B, nblk, 2), device=x.device, dtype=torch.float32)
    check(lib().kd_gca_pool(_ptr(x), _ptr(logits), n_parts, B, HW, C, nblk, _ptr(part), _ptr(ml), _stream()), "kd_gca_pool")
    pooled = torch.empty((B, C), device=x.device, dtype=torch.float32)
    check(lib().kd_gca_finalize(_ptr(part), _ptr(ml), B, nblk, C, _ptr(pooled), _stream()), "kd_gca_finalize")
    _count(2)
    return pooled


FUSED_GCA_GATE = False  # opt-in: kd_gca_gate (finalize + MLP + sigmoid as one 8-CTA cluster launch) instead of kd_gca_finalize + two kd_linear_small; measured neutral at B = 16 (193.0 vs 194.9 ms) and 3 % slower at B = 1 (8 CTAs stream the MLP weights instead of 64-128), see profiles/README.md


@_timed
def gca_gate(x, logits, w0, b0, w1, b1):
    """GlobalContext gate [B, C] of NHWC fp16 `x`: softmax-pool over pixels (kd_gca_pool) then finalize + MLP + sigmoid in one
    cluster launch (kd_gca_gate)."""
    B, H, W, C = x.shape
    HW = H * W
    hid = w0.shape[0]
    nblk = _nblk(HW, C, B)
    if not (FUSED_GCA_GATE and C % 4 == 0 and hid % 4 == 0 and 4 * (C + hid + hid // 8 + 16 + nblk) <= 48 * 1024):
        pooled = gca_pool(x, logits)
        return linear_small(linear_small(pooled, w0, b0, post_act=ACT_SILU), w1, b1, post_act=ACT_SIGMOID)
    n_parts = logits.shape[0] if logits.dim() == 3 else 1
    part = torch.empty((B, nblk, C), device=x.device, dtype=torch.float32)
    ml = torch.empty((B, nblk, 2), device=x.device, dtype=torch.float32)
    check(lib().kd_gca_pool(_ptr(x), _ptr(logits), n_parts, B, HW, C, nblk, _ptr(part), _ptr(ml), _stream()), "kd_gca_pool")
    gate = torch.empty((B, C), device=x.device, dtype=torch.float32)
    for t, name in ((w0, "w0"), (b0, "b0"), (w1, "w1"), (b1, "b1")):
        _chk(t, torch.float32, name)
    check(lib().kd_gca_gate(_ptr(part), _ptr(ml), B, nblk, C, hid, _ptr(w0), _ptr(b0), _ptr(w1), _ptr(b1), _ptr(gate), _stream()), "kd_gca_gate")
    _count(2)
    return gate


@_timed
def gate_residual(h, gate, res, want_stats=False):
    _chk(h, ACT_DTYPE, "h")
    B, H, W, C = h.shape
    out = torch.empty_like(h)
    stats = None
    if want_stats:
        nblk = lib().kd_elementwise_blocks(H * W, C)
        stats = OctStats(torch.empty((B, nblk, C // 8, 2), device=h.device, dtype=torch.float32), 1, nblk, 1, B, C // 8)
    check(lib().kd_gate_residual(_ptr(h), _ptr(gate), _ptr(res), _ptr(out), None if stats is None else _ptr(stats.partial), B, H * W, C,
                                 _stream()), "kd_gate_residual")
    _count()
    if stats is not None:
        out._kd_stats = stats
    return out


# ------------------------------------------------------------------------------------------------ LayerNorm
@_timed
def layernorm_h16(x, g, bias=None, residual=None, eps=1e-5):
    _chk(x, ACT_DTYPE, "x")
    C = x.shape[-1]
    M = x.numel() // C
    y = torch.empty_like(x)
    check(lib().kd_layernorm_h16(_ptr(x), _ptr(g), _ptr(bias), _ptr(residual), _ptr(y), M, C, eps, _stream()), "kd_layernorm_h16")
    _count()
    return y


@_timed
def layernorm_f32(x, g, bias=None, eps=1e-5):
    _chk(x, torch.float32, "x")
    C = x.shape[-1]
    M = x.numel() // C
    y = torch.empty_like(x)
    check(lib().kd_layernorm_f32(_ptr(x), _ptr(g), _ptr(bias), _ptr(y), M, C, eps, _stream()), "kd_layernorm_f32")
    _count()
    return y


# ------------------------------------------------------------------------------------------------ attention
@_timed
def kv_assemble(qkv, kv_col, ctx_kv, null_kv):
    """qkv: fp16 [B,N,ld]; ctx_kv: fp32 [B,Jc,128] or None; null_kv fp32 [2,64] -> fp16 [B, Jc+1+N, 128]."""
    B, N, ld = qkv.shape
    Jc = 0 if ctx_kv is None else ctx_kv.shape[1]
    out = torch.empty((B, Jc + 1 + N, 128), device=qkv.device, dtype=ACT_DTYPE)
    check(lib().kd_kv_assemble(_ptr(qkv), ld, kv_col, _ptr(ctx_kv), Jc, _ptr(null_kv), _ptr(out), B, N, _stream()), "kd_kv_assemble")
    _count()
    return out


ATTN_TC_MIN_TOKENS = 256  # kd_attn_mqa_tc (tcgen05) from this many query tokens; below, the mma.sync kernel (tests may raise it)


@_timed
def attn_mqa(q, kv, heads, scale):
    """q: fp16 [B,N,ld] (first heads*64 columns are the queries); kv: fp16 [B,J,128]."""
    B, N, ld = q.shape
    J = kv.shape[1]
    out = torch.empty((B, N, heads * 64), device=q.device, dtype=ACT_DTYPE)
    if N >= ATTN_TC_MIN_TOKENS:  # tcgen05 kernel (choice by the per-sample token count only: batch-invariant)
        vt = torch.empty((lib().kd_attn_vt_elems(B, J),), device=q.device, dtype=ACT_DTYPE)
        check(lib().kd_attn_mqa_tc(_ptr(q), ld, _ptr(kv), _ptr(vt), _ptr(out), B, N, J, heads, scale, _stream()), "kd_attn_mqa_tc")
        _count(2)
        return out
    check(lib().kd_attn_mqa(_ptr(q), ld, _ptr(kv), _ptr(out), B, N, J, heads, scale, _stream()), "kd_attn_mqa")
    _count()
    return out


@_timed
def attn_cross(q, kv, null_kv, heads, scale):
    """q: fp16 [B,N,heads*64]; kv: fp32 [B,Jc,2*heads*64]; null_kv fp32 [2,64]."""
    B, N, ld = q.shape
    Jc = kv.shape[1]
    out = torch.empty((B, N, heads * 64), device=q.device, dtype=ACT_DTYPE)
    check(lib().kd_attn_cross(_ptr(q), ld, _ptr(kv), _ptr(null_kv), _ptr(out), B, N, Jc, heads, scale, _stream()), "kd_attn_cross")
    _count()
    return out


@_timed
def attn_small_f32(q, kv, heads, scale):
    """q: fp32 [B,Nq,heads*64]; kv: fp32 [B,J,2*heads*64] (k | v) -> fp32 [B,Nq,heads*64]."""
    _chk(q, torch.float32, "q")
    _chk(kv, torch.float32, "kv")
    B, Nq, _ = q.shape
    out = torch.empty_like(q)
    check(lib().kd_attn_small_f32(_ptr(q), _ptr(kv), _ptr(out), B, Nq, kv.shape[1], heads, scale, _stream()), "kd_attn_small_f32")
    _count()
    return out


@_timed
def dwconv3x3(x, w):
    """Depthwise 3x3 convolution (zero padding, no bias): x NHWC fp16 [B,H,W,C], w fp32 [C,3,3]."""
    _chk(x, ACT_DTYPE, "x")
    _chk(w, torch.float32, "w")
    B, H, W, C = x.shape
    assert w.shape == (C, 3, 3)
    y = torch.empty_like(x)
    check(lib().kd_dwconv3x3(_ptr(x), _ptr(w), _ptr(y), B, H, W, C, _stream()), "kd_dwconv3x3")
    _count()
    return y


@_timed
def linear_attention(qkv, heads, scale, ctx_kv=None, act=ACT_SILU, pixels_kv=True):
    """qkv: fp16 [B, N, 3*heads*64] (q | k | v after the depthwise convs), or with pixels_kv=False only the queries [B, N, heads*64];
    ctx_kv: fp32 [B, J, 2*heads*64] (k | v of the context tokens) or None -> fp16 [B, N, heads*64] = act(scale * softmax_d(q) @
    (softmax_n(k)^T v)), n running over the pixels (pixels_kv) and the tokens."""
    _chk(qkv, ACT_DTYPE, "qkv")
    B, N, ld = qkv.shape
    inner = heads * 64
    assert ld == (3 * inner if pixels_kv else inner)
    J = 0
    if ctx_kv is not None:
        _chk(ctx_kv, torch.float32, "ctx_kv")
        J = ctx_kv.shape[1]
        assert ctx_kv.shape == (B, J, 2 * inner)
    n_kv = N if pixels_kv else 0
    assert n_kv + J > 0
    nbytes = lib().kd_linattn_workspace_bytes(B, n_kv, heads)
    ws = torch.empty((nbytes,), device=qkv.device, dtype=torch.uint8)
    ctx = torch.empty((B, heads, 64, 64), device=qkv.device, dtype=torch.float32)
    check(lib().kd_linattn_context(_ptr(qkv), ld, inner, 2 * inner, B, n_kv, heads, _ptr(ctx_kv), J, _ptr(ws), nbytes, _ptr(ctx), _stream()),
          "kd_linattn_context")
    out = torch.empty((B, N, inner), device=qkv.device, dtype=ACT_DTYPE)
    check(lib().kd_linattn_apply(_ptr(qkv), ld, 0, _ptr(ctx), _ptr(out), B, N, heads, scale, act, _stream()), "kd_linattn_apply")
    _count(4)
    return out


@_timed
def axpby(x, y, a, b):
    _chk(x, torch.float32, "x")
    _chk(y, torch.float32, "y")
    assert x.shape == y.shape
    out = torch.empty_like(x)
    check(lib().kd_axpby(_ptr(x), _ptr(y), float(a), float(b), _ptr(out), x.numel(), _stream()), "kd_axpby")
    _count()
    return out


# ------------------------------------------------------------------------------------------------ edge convs
@_timed
def im2col_nchw(x, ksize, Kp):
    _chk(x, torch.float32, "x")
    B, C, H, W = x.shape
    out = torch.empty((B * H * W, Kp), device=x.device, dtype=ACT_DTYPE)
    check(lib().kd_im2col_nchw(_ptr(x), B, C, H, W, ksize, _ptr(out), Kp, _stream()), "kd_im2col_nchw")
    _count()
    return out


def init_conv_kp(C, ksize):
    return lib().kd_init_conv_kp(C, ksize)


@_timed
def init_conv(x, ksize, w_packed, bias, addend, out, algo_taps=None):
    """CrossEmbedLayer slice for <= 3 channels: out (NHWC fp16) = conv_ks(x NCHW fp32) + bias + addend.
    algo_taps: un-merged taps per (input channel, output channel) for the algorithmic FLOP count."""
    _chk(x, torch.float32, "x")
    _chk(w_packed, ACT_DTYPE, "w_packed")
    _chk(out, ACT_DTYPE, "out")
    B, C, H, W = x.shape
    Cout = out.shape[-1]
    assert w_packed.shape == (Cout, init_conv_kp(C, ksize)), (w_packed.shape, Cout, C, ksize)
    if addend is not None:
        _chk(addend, ACT_DTYPE, "addend")
        assert addend.shape == out.shape
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    with _ConvTimer(2.0 * B * H * W * Cout * C * (algo_taps if algo_taps is not None else ksize * ksize), (3, B, H, W, C, Cout, ksize)):
        check(lib().kd_init_conv(_ptr(x), B, C, H, W, ksize, _ptr(w_packed), _ptr(bias), _ptr(addend), _ptr(out), Cout, _stream()),
              "kd_init_conv")
    _count()
    return out


@_timed
def final_conv(xa, xb, w, bias):
    """xa: NHWC fp16; xb: NCHW fp32 or None; w: fp32 [Cout,3,3,Ca+Cb] -> NCHW fp32 [B,Cout,H,W]."""
    _chk(xa, ACT_DTYPE, "xa")
    B, H, W, Ca = xa.shape
    Cb = 0 if xb is None else xb.shape[1]
    Cout = w.shape[0]
    out = torch.empty((B, Cout, H, W), device=xa.device, dtype=torch.float32)
    ws = getattr(w, "_kd_split", None)
    if ws is None:  # one-time hi + lo fp16 split of the filter (cached on the weight tensor)
        ws = torch.empty((lib().kd_final_conv_pack_elems(Ca),), device=xa.device, dtype=ACT_DTYPE)
        check(lib().kd_final_conv_pack(_ptr(w), Cout, Ca, Cb, _ptr(ws), _stream()), "kd_final_conv_pack")
        w._kd_split = ws
    check(lib().kd_final_conv(_ptr(xa), Ca, _ptr(xb), Cb, _ptr(w), _ptr(ws), _ptr(bias), _ptr(out), B, H, W, Cout, _stream()), "kd_final_conv")
    _count()
    return out


# ------------------------------------------------------------------------------------------------ sampler update
def quantile_ranks(n, q=0.95):
    """Mirror of ATen quantile_compute (linear interpolation): ranks = q * (n - 1) evaluated in float32."""
    rank = torch.tensor(q, dtype=torch.float32) * (n - 1)
    lo = rank.floor()
    weight = float(rank - lo)
    hi = int(rank.ceil().item())
    return int(lo.item()), hi, weight


@_timed
def dynthresh(x_t, pred, objective, alpha, sigma, q=0.95, workspace=None):
    _chk(x_t, torch.float32, "x_t")
    _chk(pred, torch.float32, "pred")
    B = x_t.shape[0]
    n_per = x_t[0].numel()
    lo, hi, weight = quantile_ranks(n_per, q)
    nbytes = lib().kd_dynthresh_workspace_bytes(B)
    if workspace is None:
        workspace = torch.empty(nbytes, device=x_t.device, dtype=torch.uint8)
    s = torch.empty((B,), device=x_t.device, dtype=torch.float32)
    check(lib().kd_dynthresh(_ptr(x_t), _ptr(pred), B, n_per, PRED[objective], alpha, sigma, lo, hi, weight, _ptr(workspace), nbytes,
                             _ptr(s), _stream()), "kd_dynthresh")
    _count(10)
    return s


@_timed
def ddpm_step(x_t, pred, noise, s, objective, sc, *, renoise=None, rn=(0.0, 0.0, 1.0), out=None, x0_out=None):
    """sc: dict of fp32 python floats alpha, sigma, one_minus_c, c, alpha_next, std."""
    _chk(x_t, torch.float32, "x_t")
    B = x_t.shape[0]
    n_per = x_t[0].numel()
    if out is None:
        out = torch.empty_like(x_t)
    check(lib().kd_ddpm_step(_ptr(x_t), _ptr(pred), _ptr(noise), _ptr(s), _ptr(out), _ptr(x0_out), B, n_per, PRED[objective],
                             sc["alpha"], sc["sigma"], sc["one_minus_c"], sc["c"], sc["alpha_next"], sc["std"], _ptr(renoise),
                             rn[0], rn[1], rn[2], _stream()), "kd_ddpm_step")
    _count()
    return out


@_timed
def inpaint_blend(img, inpaint, mask, noise, alpha, sigma):
    B, C, H, W = img.shape
    _chk(mask, torch.uint8, "mask")
    check(lib().kd_inpaint_blend(_ptr(img), _ptr(inpaint), _ptr(mask), _ptr(noise), alpha, sigma, B, C, H * W, _stream()),
          "kd_inpaint_blend")
    _count()
    return img


@_timed
def finalize_image(img, inpaint=None, mask=None):
    B, C, H, W = img.shape
    check(lib().kd_finalize_image(_ptr(img), _ptr(inpaint), _ptr(mask), B, C, H * W, _stream()), "kd_finalize_image")
    _count()
    return img


@_timed
def q_sample(x0, noise, alpha, sigma):
    out = torch.empty_like(x0)
    check(lib().kd_q_sample(_ptr(x0), _ptr(noise), alpha, sigma, _ptr(out), x0.numel(), _stream()), "kd_q_sample")
    _count()
    return out


@_timed
def randn(shape, seed, key, device):
    out = torch.empty(shape, device=device, dtype=torch.float32)
    check(lib().kd_randn(_ptr(out), out.numel(), seed & (2**64 - 1), key & (2**64 - 1), _stream()), "kd_randn")
    _count()
    return out


def neighbour_strips(S, ov, orientation, above=None, side=None, corner=None):
    """Strip views (tensor, channel stride, row stride) into FULL neighbour patches [3,S,S] (sample_ultra_res.py:156-170)."""
    def view(t, y0, x0):
        if t is None:
            return None
        assert t.shape[-3:] == (3, S, S) and t.is_contiguous()
        t3 = t.reshape(3, S, S)
        return (t3[:, y0:, x0:], S * S, S)
    side_x0 = S - ov if orientation == -1 else 0  # o=-1: neighbour's right columns face us; o=+1: its left columns
    return view(above, S - ov, 0), view(side, 0, side_x0), view(corner, S - ov, side_x0)


@_timed
def randn_into(out, seed, key):
    assert out.is_contiguous() and out.dtype == torch.float32 and out.is_cuda
    check(lib().kd_randn(_ptr(out), out.numel(), seed & (2**64 - 1), key & (2**64 - 1), _stream()), "kd_randn")
    _count()
    return out


@_timed
def border_pack(S, overlap_pos, orientation, above, side, corner, device):
    """above / side / corner: None or (tensor_view, channel_stride, row_stride) strips (see neighbour_strips) -> (inpaint, mask)."""
    inpaint = torch.empty((3, S, S), device=device, dtype=torch.float32)
    mask = torch.empty((S, S), device=device, dtype=torch.uint8)
    args = []
    for st in (above, side, corner):
        if st is None:
            args += [None, 0, 0]
        else:
            t, cs, rs = st
            assert t.dtype == torch.float32 and t.is_cuda
            args += [_ptr(t), cs, rs]
    check(lib().kd_border_pack(_ptr(inpaint), _ptr(mask), *args, S, overlap_pos, orientation, _stream()), "kd_border_pack")
    _count()
    return inpaint, mask


# ------------------------------------------------------------------------------------------------ K9 peer mailbox / N2 / N3
def strip_push(view, cs, rs, dst_ptr, flag_ptr, value):
    """view: fp32 strided strip [C, rows, cols] (element (c,y,x) at c*cs + y*rs + x); dst_ptr / flag_ptr: raw (peer) addresses."""
    assert view.dtype == torch.float32 and view.is_cuda and view.stride(-1) == 1
    C, rows, cols = view.shape
    check(lib().kd_strip_push(_ptr(view), cs, rs, C, rows, cols, ctypes.c_void_p(dst_ptr), ctypes.c_void_p(flag_ptr), value, _stream()),
          "kd_strip_push")
    _count(2)


def flag_wait(flag_ptr, value, timeout_s, status):
    check(lib().kd_flag_wait(ctypes.c_void_p(flag_ptr), value, float(timeout_s), _ptr(status), _stream()), "kd_flag_wait")
    _count()


@_timed
def cond_gather(zoomed, out, off, shift_y, shift_x, fill, patch_width=0, center_top=0):
    """zoomed: fp32 [3,W,W]; out: fp32 [3|6,P,P] (see kd_cond_gather)."""
    _chk(zoomed, torch.float32, "zoomed")
    _chk(out, torch.float32, "out")
    W, P = zoomed.shape[-1], out.shape[-1]
    assert zoomed.shape == (3, W, W) and out.shape[0] in (3, 6) and out.shape[1] == P
    check(lib().kd_cond_gather(_ptr(zoomed), W, _ptr(out), out.shape[0], P, off, shift_y, shift_x, fill, patch_width, center_top, _stream()),
          "kd_cond_gather")
    _count()
    return out


@_timed
def canvas_fill(zoomed, canvas_ptr, Wc, cell_index, n, patch_dist, P):
    """Background of the stitched canvas (bilinear upsample of `zoomed`, or zeros when None) where no patch covers it."""
    if zoomed is not None:
        _chk(zoomed, torch.float32, "zoomed")
    _chk(cell_index, torch.int32, "cell_index")
    check(lib().kd_canvas_fill(_ptr(zoomed), 0 if zoomed is None else zoomed.shape[-1], ctypes.c_void_p(canvas_ptr), Wc, _ptr(cell_index), n,
                               patch_dist, P, _stream()), "kd_canvas_fill")
    _count()


@_timed
def patch_paste(patch, canvas_ptr, Wc, cell_index, n, patch_dist, k, i, j):
    _chk(patch, torch.float32, "patch")
    P = patch.shape[-1]
    assert patch.numel() == 3 * P * P
    check(lib().kd_patch_paste(_ptr(patch), ctypes.c_void_p(canvas_ptr), Wc, _ptr(cell_index), n, patch_dist, P, k, i, j, _stream()),
          "kd_patch_paste")
    _count()
