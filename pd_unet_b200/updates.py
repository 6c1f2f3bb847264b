"""The elementwise steps of the unrolled primal-dual iteration, fused (csrc/updates.cu):

    concat(a, b[, c])                -- torch.cat(..., dim=1) feeding a primal / dual block
    residual_slice(state, delta, k)  -- state + delta, plus the channel the next operator consumes
    axpby(alpha, x, beta, y)
    angular_upsample(sino, factor, mode) -- PD-UNet's sparse-view -> full-view linear interpolation

All are differentiable; float32 CUDA tensors only (no CPU path).
"""
from __future__ import annotations

from typing import Optional, Tuple

import os

import torch

from ._lib import LAYOUT_NCHW, LAYOUT_NHWC, WRAP_MODES, PduError, check, lib, require_cuda, stream_ptr


def _plane(t: torch.Tensor) -> int:
    n = 1
    for s in t.shape[2:]:
        n *= s
    return n


def _is_channels_last(t: torch.Tensor) -> bool:
    """4-D tensor whose memory is [B, H, W, C] (torch.channels_last) and not also plain contiguous."""
    return t.dim() == 4 and t.shape[1] > 1 and not t.is_contiguous() and \
        t.is_contiguous(memory_format=torch.channels_last)


def _in_layout(t: torch.Tensor, nhwc: bool, name: str) -> torch.Tensor:
    """float32 CUDA tensor whose bytes are in the requested layout (copying only if they are not)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise PduError(f"{name} is on {t.device}: the pd_unet_b200 operators run on CUDA only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be torch.float32, got {t.dtype}")
    if nhwc:
        return t.contiguous(memory_format=torch.channels_last)
    return t.contiguous()


def _empty(shape, like: torch.Tensor, nhwc: bool) -> torch.Tensor:
    fmt = torch.channels_last if nhwc else torch.contiguous_format
    return torch.empty(shape, dtype=torch.float32, device=like.device, memory_format=fmt)


class _Concat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c, scale_b, pad_to):
        nhwc = _is_channels_last(a)          # the state tensor decides; one-channel inputs fit either layout
        a = _in_layout(a, nhwc, "a")
        # channels-last state with multi-channel b / c that are still planar (operator outputs): the layout change is
        # done inside the concatenation (pdu_concat_mixed_f32) instead of by a copy pass per operand
        planar = [False, False]
        if nhwc:
            for j, t in enumerate((b, c)):
                planar[j] = (t is not None and t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] > 1
                             and t.is_contiguous())
        if not planar[0]:
            b = _in_layout(b, nhwc, "b")
        if c is not None and not planar[1]:
            c = _in_layout(c, nhwc, "c")
        for t in (b, c):
            if t is not None and (t.shape[0] != a.shape[0] or t.shape[2:] != a.shape[2:]):
                raise ValueError("concat: tensors must agree in every axis but the channel axis")
        ca, cb, cc = a.shape[1], b.shape[1], (c.shape[1] if c is not None else 0)
        ctx.split, ctx.scale_b = (ca, cb, cc), scale_b
        c_out = ca + cb + cc
        if pad_to > 1:
            c_out = (c_out + pad_to - 1) // pad_to * pad_to
        out = _empty((a.shape[0], c_out) + tuple(a.shape[2:]), a, nhwc)
        if out.numel():
            with torch.cuda.device(a.device):
                if planar[0] or planar[1]:
                    check(lib().pdu_concat_mixed_f32(out.data_ptr(), a.data_ptr(), b.data_ptr(),
                                                     c.data_ptr() if c is not None else None, a.shape[0], ca, cb, cc, c_out,
                                                     _plane(a), scale_b, int(planar[0]), int(planar[1]), stream_ptr()),
                          "pdu_concat_mixed_f32")
                else:
                    check(lib().pdu_concat_f32(out.data_ptr(), a.data_ptr(), b.data_ptr(),
                                               c.data_ptr() if c is not None else None, a.shape[0], ca, cb, cc,
                                               c_out, _plane(a), scale_b, LAYOUT_NHWC if nhwc else LAYOUT_NCHW, stream_ptr()),
                          "pdu_concat_f32")
        return out

    @staticmethod
    def backward(ctx, g):
        ca, cb, cc = ctx.split
        gb = g[:, ca:ca + cb]
        if ctx.scale_b != 1.0:
            gb = gb * ctx.scale_b
        return g[:, :ca], gb, (g[:, ca + cb:ca + cb + cc] if cc else None), None, None


def concat(a: torch.Tensor, b: torch.Tensor, c: Optional[torch.Tensor] = None, scale_b: float = 1.0,
           pad_to: int = 0) -> torch.Tensor:
    """cat([a, scale_b * b(, c)], dim=1) for [B, c_i, ...] tensors in one pass.  If `a` is a
    channels_last tensor the result is channels_last too (what cuDNN's tensor-core convolutions want).
    pad_to > 1 appends zero channels up to the next multiple of pad_to (for a convolution whose
    weights are zero-padded the same way: 6 input channels make cuDNN fall back to a CUDA-core kernel)."""
    return _Concat.apply(a, b, c, float(scale_b), int(pad_to))


class _ConcatUpsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, sparse, factor, mode, scale_b, scale_c, pad_to):
        nhwc = _is_channels_last(a)
        a = _in_layout(a, nhwc, "a")
        b = _in_layout(b, nhwc, "b")
        sparse = require_cuda(sparse, torch.float32, "sparse")
        B, ca = a.shape[:2]
        As, D = sparse.shape[-2], sparse.shape[-1]
        if b.shape[1] != 1 or b.shape[0] != B or tuple(b.shape[2:]) != tuple(a.shape[2:]):
            raise ValueError("concat_upsampled: b must be [B, 1, views, det] like a")
        if tuple(a.shape[2:]) != (As * factor, D) or sparse.numel() != B * As * D:
            raise ValueError(f"concat_upsampled: a is {tuple(a.shape)}, sparse {tuple(sparse.shape)} x{factor} views")
        ctx.ca, ctx.scale_b = ca, scale_b
        c_out = ca + 2
        if pad_to > 1:
            c_out = (c_out + pad_to - 1) // pad_to * pad_to
        out = _empty((B, c_out) + tuple(a.shape[2:]), a, nhwc)
        if out.numel():
            with torch.cuda.device(a.device):
                check(lib().pdu_concat_upsample_f32(out.data_ptr(), a.data_ptr(), b.data_ptr(), sparse.data_ptr(), B, ca, c_out, As,
                                                    factor, D, WRAP_MODES[mode], scale_b, scale_c,
                                                    LAYOUT_NHWC if nhwc else LAYOUT_NCHW, stream_ptr()), "pdu_concat_upsample_f32")
        return out

    @staticmethod
    def backward(ctx, g):
        ca = ctx.ca
        gb = g[:, ca:ca + 1]
        if ctx.scale_b != 1.0:
            gb = gb * ctx.scale_b
        return g[:, :ca], gb, None, None, None, None, None, None      # the measured data takes no gradient


def concat_upsampled(a: torch.Tensor, b: torch.Tensor, sparse: torch.Tensor, factor: int, mode: str = "flip",
                     scale_b: float = 1.0, scale_c: float = 1.0, pad_to: int = 0) -> torch.Tensor:
    """cat([a, scale_b * b, scale_c * angular_upsample(sparse, factor, mode)], dim=1) in one pass: the dual update's input
    with the measured sparse-view sinogram [B, (1,) A_sparse, D] interpolated on the fly."""
    if mode not in WRAP_MODES:
        raise ValueError(f"mode must be one of {tuple(WRAP_MODES)}")
    return _ConcatUpsample.apply(a, b, sparse, int(factor), mode, float(scale_b), float(scale_c), int(pad_to))


class _ResidualSlice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, delta, k, kn):
        nhwc = _is_channels_last(delta) or _is_channels_last(state)
        state = _in_layout(state, nhwc, "state")
        delta = _in_layout(delta, nhwc, "delta")
        if state.shape != delta.shape or state.dim() < 3:
            raise ValueError("residual_slice: state and delta must be the same [B, C, ...] shape")
        B, Cn = state.shape[:2]
        if not (0 <= k and kn >= 1 and k + kn <= Cn):
            raise ValueError(f"residual_slice: channels [{k}, {k + kn}) out of range for {Cn} channels")
        ctx.k, ctx.kn = k, kn
        out = _empty(tuple(state.shape), state, nhwc)
        sl = torch.empty((B, kn) + tuple(state.shape[2:]), dtype=torch.float32, device=state.device)
        if out.numel():
            with torch.cuda.device(state.device):
                check(lib().pdu_residual_slice_f32(out.data_ptr(), sl.data_ptr(), state.data_ptr(), delta.data_ptr(),
                                                   B, Cn, _plane(state), k, kn,
                                                   LAYOUT_NHWC if nhwc else LAYOUT_NCHW, stream_ptr()),
                      "pdu_residual_slice_f32")
        return out, sl

    @staticmethod
    def backward(ctx, g_out, g_slice):
        g = g_out
        if g_slice is not None:
            g = g_out.clone()
            g[:, ctx.k:ctx.k + ctx.kn] += g_slice
        return g, g, None, None


def residual_slice(state: torch.Tensor, delta: torch.Tensor, k: int = 0, kn: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (state + delta, its channels k:k+kn as a contiguous planar [B, kn, ...] tensor), one pass over
    the data.  Follows the layout of its inputs (planar or channels_last)."""
    return _ResidualSlice.apply(state, delta, int(k), int(kn))


class _Axpby(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, alpha, beta):
        x = require_cuda(x, torch.float32, "x")
        y = require_cuda(y, torch.float32, "y")
        if x.shape != y.shape:
            raise ValueError("axpby: shapes differ")
        ctx.ab = (alpha, beta)
        out = torch.empty_like(x)
        if out.numel():
            with torch.cuda.device(x.device):
                check(lib().pdu_axpby_f32(out.data_ptr(), alpha, x.data_ptr(), beta, y.data_ptr(), x.numel(),
                                          stream_ptr()), "pdu_axpby_f32")
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.ab
        return g * a, g * b, None, None


def axpby(alpha: float, x: torch.Tensor, beta: float, y: torch.Tensor) -> torch.Tensor:
    return _Axpby.apply(x, y, float(alpha), float(beta))


def _upsample_call(fn_name: str, src: torch.Tensor, a_sparse: int, factor: int, mode: str, out_views: int):
    src = require_cuda(src, torch.float32, "sinogram")
    lead, D = src.shape[:-2], src.shape[-1]
    flat = src.reshape(-1, src.shape[-2], D)
    out = torch.empty((flat.shape[0], out_views, D), dtype=torch.float32, device=src.device)
    if out.numel():
        with torch.cuda.device(src.device):
            check(getattr(lib(), fn_name)(flat.data_ptr(), out.data_ptr(), flat.shape[0], a_sparse, factor, D,
                                          WRAP_MODES[mode], stream_ptr()), fn_name)
    return out.reshape(*lead, out_views, D)


class _AngularUpsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sino, factor, mode):
        ctx.cfg = (sino.shape[-2], factor, mode)
        return _upsample_call("pdu_angular_upsample_f32", sino, sino.shape[-2], factor, mode, sino.shape[-2] * factor)

    @staticmethod
    def backward(ctx, g):
        a_sparse, factor, mode = ctx.cfg
        return _upsample_call("pdu_angular_upsample_adj_f32", g, a_sparse, factor, mode, a_sparse), None, None


class _AngularUpsampleAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, full, factor, mode):
        if full.shape[-2] % factor:
            raise ValueError("angular_upsample_adjoint: view count must be a multiple of the factor")
        ctx.cfg = (full.shape[-2] // factor, factor, mode)
        return _upsample_call("pdu_angular_upsample_adj_f32", full, ctx.cfg[0], factor, mode, ctx.cfg[0])

    @staticmethod
    def backward(ctx, g):
        a_sparse, factor, mode = ctx.cfg
        return _upsample_call("pdu_angular_upsample_f32", g, a_sparse, factor, mode, a_sparse * factor), None, None


def angular_upsample(sino: torch.Tensor, factor: int, mode: str = "flip") -> torch.Tensor:
    """[..., A_sparse, D] -> [..., A_sparse * factor, D]; view i f + r = (1 - r/f) s[i] + (r/f) s[i+1].
    mode: 'flip' (parallel beam over pi: the view after the last is the first, detector reversed),
    'periodic' (fan beam over 2 pi) or 'clamp'."""
    if mode not in WRAP_MODES:
        raise ValueError(f"mode must be one of {tuple(WRAP_MODES)}")
    if factor < 1:
        raise ValueError("factor must be >= 1")
    return _AngularUpsample.apply(sino, int(factor), mode)


def angular_upsample_adjoint(full: torch.Tensor, factor: int, mode: str = "flip") -> torch.Tensor:
    if mode not in WRAP_MODES:
        raise ValueError(f"mode must be one of {tuple(WRAP_MODES)}")
    return _AngularUpsampleAdjoint.apply(full, int(factor), mode)


def bias_prelu_(y: torch.Tensor, bias: torch.Tensor, slope: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In place y <- prelu(y + bias[c], slope[c]) for a [B, C, ...] float32 CUDA tensor in planar or
    channels_last layout; slope None means bias only.  Inference epilogue of Conv2d + PReLU as one pass.
    Not differentiable (the modules in pd_unet_b200.model fall back to the ATen ops when gradients are on)."""
    if not y.is_cuda:
        raise PduError(f"y is on {y.device}: the pd_unet_b200 operators run on CUDA only (no CPU fallback)")
    if y.dtype != torch.float32 or y.dim() < 3:
        raise TypeError("bias_prelu_: y must be a float32 [B, C, ...] tensor")
    nhwc = _is_channels_last(y)
    if not nhwc and not y.is_contiguous():
        raise ValueError("bias_prelu_: y must be contiguous (planar or channels_last) to be updated in place")
    Cn = y.shape[1]
    bias = require_cuda(bias.detach(), torch.float32, "bias")
    if bias.numel() != Cn:
        raise ValueError(f"bias_prelu_: bias has {bias.numel()} values for {Cn} channels")
    n_slope = 0
    if slope is not None:
        slope = require_cuda(slope.detach(), torch.float32, "slope")
        n_slope = slope.numel()
        if n_slope not in (1, Cn):
            raise ValueError(f"bias_prelu_: slope has {n_slope} values for {Cn} channels")
    if y.numel():
        with torch.cuda.device(y.device):
            check(lib().pdu_bias_prelu_f32(y.data_ptr(), bias.data_ptr(), slope.data_ptr() if slope is not None else None,
                                           max(n_slope, 1), y.shape[0], Cn, _plane(y),
                                           LAYOUT_NHWC if nhwc else LAYOUT_NCHW, stream_ptr()), "pdu_bias_prelu_f32")
    return y


# Below this size a training step is bound by the host (a Python autograd.Function + ctypes calls cost more than
# ATen's C++ nodes; an eager PD-UNet step is ~1000 launches) and the fused epilogue does not pay.  Measured on B200:
#   CT 256^2 x 8 slices (16.8 M-element maps), 1 GPU, gate 0 / 1 M / 4 M / 8 M: 24.6 / 24.6 / 25.2 / 26.9 ms
#                                                                    (ATen ops only: 34.8; under 2-GPU DDP 25.5 vs 35.9)
#   MRI 320^2 x 2 slices (6.5 M-element maps): 1 GPU neutral (17.0 .. 17.7 ms either way), 2-GPU DDP 20.5 fused vs 18.7
#   128^2 x 2 slices (1 M): 7.2 ms fused vs 6.4
# so the gate sits between 6.5 M and 16.8 M elements.
FUSED_TRAIN_MIN_ELEMS = int(os.environ.get("PDU_FUSED_TRAIN_MIN_ELEMS", 8 << 20))


def fused_train_min_elems() -> int:
    """The size gate of the fused training epilogues; 0 while a CUDA graph is being captured (a replayed step has no
    host cost per launch, so the fused kernels win at every size: MRI 320^2 x 2 slices 15.3 -> 11.0 ms)."""
    return 0 if torch.cuda.is_current_stream_capturing() else FUSED_TRAIN_MIN_ELEMS


def _bias_prelu_train_ok(y: torch.Tensor, bias: torch.Tensor, slope: Optional[torch.Tensor]) -> bool:
    """Shapes the fused training epilogue serves: float32 CUDA channels-last [B, C, H, W], C in {4, 8, ..., 256}."""
    if slope is None or not y.is_cuda or y.dtype != torch.float32 or y.dim() != 4 or y.numel() == 0:
        return False
    Cn = y.shape[1]
    return (_is_channels_last(y) and Cn % 4 == 0 and Cn // 4 <= 64 and 256 % (Cn // 4) == 0 and bias.numel() == Cn
            and slope.numel() in (1, Cn) and y.data_ptr() % 16 == 0)


class _on_device:
    """torch.cuda.device(dev) without the context-manager cost when dev is already current (the usual case)."""

    def __init__(self, dev):
        self.ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


class _BiasPReLU(torch.autograd.Function):
    """out = prelu(y + bias, slope) with a one-pass backward (pdu_bias_prelu_fwd_f32 / pdu_bias_prelu_bwd_f32)."""

    @staticmethod
    def forward(ctx, y, bias, slope):
        y = y.contiguous(memory_format=torch.channels_last)
        bias_c, slope_c = bias.detach().contiguous(), slope.detach().contiguous()
        out = torch.empty_like(y, memory_format=torch.channels_last)
        with _on_device(y.device):
            check(lib().pdu_bias_prelu_fwd_f32(y.data_ptr(), out.data_ptr(), bias_c.data_ptr(), slope_c.data_ptr(),
                                               slope_c.numel(), y.shape[0], y.shape[1], _plane(y), LAYOUT_NHWC,
                                               stream_ptr()), "pdu_bias_prelu_fwd_f32")
        ctx.save_for_backward(y, bias_c, slope_c)
        return out

    @staticmethod
    def backward(ctx, g):
        y, bias, slope = ctx.saved_tensors
        g = g.contiguous(memory_format=torch.channels_last)
        gz = torch.empty_like(y, memory_format=torch.channels_last)
        gb, ga = torch.empty_like(bias), torch.empty_like(slope)
        with _on_device(y.device):
            L = lib()
            ws = torch.empty(L.pdu_bias_prelu_bwd_workspace_bytes(y.shape[1]), dtype=torch.uint8, device=y.device)
            check(L.pdu_bias_prelu_bwd_f32(g.data_ptr(), y.data_ptr(), bias.data_ptr(), slope.data_ptr(), slope.numel(),
                                           gz.data_ptr(), gb.data_ptr(), ga.data_ptr(), ws.data_ptr(), ws.numel(),
                                           y.shape[0], y.shape[1], _plane(y), LAYOUT_NHWC, stream_ptr()),
                  "pdu_bias_prelu_bwd_f32")
        return gz, gb, ga


class _BiasAdd(torch.autograd.Function):
    """out = y + bias[c] in place (y is the fresh output of a bias-free convolution); backward passes the gradient
    through untouched and sums it per channel with pdu_channel_sum_f32 (two stages, fixed order)."""

    @staticmethod
    def forward(ctx, y, bias):
        ctx.mark_dirty(y)
        ctx.n_channels = y.shape[1]
        ctx.bias_dtype = bias.dtype
        return bias_prelu_(y, bias, None)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous(memory_format=torch.channels_last)
        gb = torch.empty(ctx.n_channels, dtype=torch.float32, device=g.device)
        with _on_device(g.device):
            L = lib()
            ws = torch.empty(L.pdu_bias_prelu_bwd_workspace_bytes(g.shape[1]), dtype=torch.uint8, device=g.device)
            check(L.pdu_channel_sum_f32(g.data_ptr(), gb.data_ptr(), ws.data_ptr(), ws.numel(), g.shape[0], g.shape[1],
                                        _plane(g), LAYOUT_NHWC, stream_ptr()), "pdu_channel_sum_f32")
        return g, gb


def bias_add(y: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """Differentiable y + bias[c] for the output of a bias-free (transposed) convolution without activation; large
    channels-last maps take the in-place add and the two-stage channel sum, everything else the ATen ops."""
    if not y.is_cuda:
        raise PduError(f"y is on {y.device}: the pd_unet_b200 operators run on CUDA only (no CPU fallback)")
    Cn = y.shape[1] if y.dim() == 4 else 0
    if (y.is_cuda and y.dtype == torch.float32 and y.dim() == 4 and y.numel() >= fused_train_min_elems() and _is_channels_last(y)
            and Cn % 4 == 0 and Cn // 4 <= 64 and 256 % (Cn // 4) == 0 and bias.numel() == Cn and y.data_ptr() % 16 == 0
            and not y.is_leaf):        # in place: y must be the (unsaved) output of the convolution, never a user tensor
        return _BiasAdd.apply(y, bias)
    return y + bias.view((1, -1) + (1,) * (y.dim() - 2))


def bias_prelu(y: torch.Tensor, bias: torch.Tensor, slope: torch.Tensor) -> torch.Tensor:
    """Differentiable prelu(y + bias[c], slope) for the output y of a bias-free convolution: the training form of
    `bias_prelu_`.  One forward pass and one backward pass (input gradient + bias and slope gradients, reproducible)
    instead of ATen's bias add, PReLU, PReLU backward and two full-size reductions.  Shapes the fused kernels do
    not serve (planar layout, odd channel counts) go through the equivalent ATen ops on the GPU."""
    if not y.is_cuda:
        raise PduError(f"y is on {y.device}: the pd_unet_b200 operators run on CUDA only (no CPU fallback)")
    if _bias_prelu_train_ok(y, bias, slope) and y.numel() >= fused_train_min_elems():
        return _BiasPReLU.apply(y, bias, slope)
    return torch.nn.functional.prelu(y + bias.view((1, -1) + (1,) * (y.dim() - 2)), slope)


def bias_prelu_place_(y: torch.Tensor, bias: torch.Tensor, slope: Optional[torch.Tensor], dst: torch.Tensor,
                      pooled: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """One pass over a channels_last convolution output y [B, C, H, W]: v = prelu(y + bias, slope) is written
    into `dst`, a channel slice (view) of a channels_last buffer -- e.g. the decoder's concatenation buffer
    -- and, if `pooled` [B, C, H/2, W/2] (channels_last) is given, its 2x2 max into pooled.  Replaces the
    bias add, PReLU, max_pool2d and torch.cat passes of a UNet encoder block.  Inference only."""
    if not (y.is_cuda and y.dtype == torch.float32 and y.dim() == 4):
        raise PduError("bias_prelu_place_: y must be a float32 CUDA [B, C, H, W] tensor (no CPU fallback)")
    B, Cn, H, W = y.shape
    if not y.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("bias_prelu_place_: y must be channels_last")
    pix = dst.stride(3)
    if tuple(dst.shape) != (B, Cn, H, W) or dst.dtype != torch.float32 or dst.device != y.device or \
            dst.stride(1) != 1 or dst.stride(2) != W * pix or dst.stride(0) != H * W * pix:
        raise ValueError("bias_prelu_place_: dst must be a channel slice of a channels_last buffer shaped like y")
    bias = require_cuda(bias.detach(), torch.float32, "bias")
    n_slope = 1
    if slope is not None:
        slope = require_cuda(slope.detach(), torch.float32, "slope")
        n_slope = slope.numel()
    if pooled is not None and (tuple(pooled.shape) != (B, Cn, H // 2, W // 2) or
                               not pooled.is_contiguous(memory_format=torch.channels_last) or pooled.dtype != torch.float32):
        raise ValueError("bias_prelu_place_: pooled must be a channels_last [B, C, H/2, W/2] float32 tensor")
    if y.numel():
        with torch.cuda.device(y.device):
            check(lib().pdu_bias_prelu_place_f32(y.data_ptr(), bias.data_ptr(), slope.data_ptr() if slope is not None else None,
                                                 n_slope, dst.data_ptr(), pix, pooled.data_ptr() if pooled is not None else None,
                                                 B, Cn, H, W, stream_ptr()), "pdu_bias_prelu_place_f32")
    return pooled
