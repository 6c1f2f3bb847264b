"""Primal-Dual UNet assembled around the B200 operators.

The reference model lives in unmounted branches (/root/reference/README.md:5), so this is a
[RECALL]/[DESIGN] assembly of what the paper the README cites (arXiv 2112.13443) describes: Adler &
Oktem's learned primal-dual unrolling in which the image-space (primal) block is a UNet, the
data-space (dual) block stays a small CNN, and the measured sparse-view sinogram is first upsampled
to the full angular grid.  Parameter names are this repo's own; they are NOT claimed to match the
reference checkpoints (unverifiable -- see DESIGN.md).

    h <- h + Dual_i (cat(h, K f[:, :kc] / s, g))       data space
    f <- f + UNet_i (cat(f, K* h[:, :kc] / s))         image space

K / K* are the operators of pd_unet_b200.radon (CT) or pd_unet_b200.nufft (MRI), kc = 1 real
channel for CT and 2 (re, im) for MRI; cat and the residual + slice steps are the fused kernels of
pd_unet_b200.updates.  The convolutions are stock PyTorch / cuDNN as BASELINE.json prescribes.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import nn

from . import updates
from .nufft import KbNufft, KbNufftAdjoint
from .radon import _BaseRadon


def _fused_epilogue_ok(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled()


class ConvAct(nn.Module):
    """3x3 convolution (cuDNN) + optional PReLU.  In inference the bias add and the activation run as
    one in-place pass of pdu_bias_prelu_f32 instead of ATen's two; with gradients on, the convolution runs
    without its bias and `updates.bias_prelu` (one forward pass, one backward pass that also yields the bias and
    slope gradients) replaces bias add + PReLU and their backward kernels.  Parameters are those of the wrapped
    modules, so state_dict keys do not change."""

    def __init__(self, cin: int, cout: int, act: bool = True, kernel: int = 3):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel, padding=kernel // 2)
        self.act = nn.PReLU(cout) if act else None
        self._wpad = None          # (key, weight zero-padded along the input channels); inference-only cache

    def _weight_for(self, cin: int) -> torch.Tensor:
        """The convolution weight, zero-padded to `cin` input channels when the fused concat delivered a
        channel count rounded up for the tensor cores (the extra inputs are zeros, so the result is the same)."""
        w = self.conv.weight
        if cin == w.shape[1]:
            return w
        key = (w.data_ptr(), w._version, cin, w.device)
        if self._wpad is None or self._wpad[0] != key:
            wp = torch.zeros((w.shape[0], cin) + tuple(w.shape[2:]), dtype=w.dtype, device=w.device)
            wp[:, :w.shape[1]] = w.detach()
            self._wpad = (key, wp.contiguous(memory_format=torch.channels_last))
        return self._wpad[1]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if _fused_epilogue_ok(x):
            y = nn.functional.conv2d(x, self._weight_for(x.shape[1]), None, self.conv.stride, self.conv.padding)
            return updates.bias_prelu_(y, self.conv.bias, self.act.weight if self.act is not None else None)
        if (self.act is not None and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
                and x.is_contiguous(memory_format=torch.channels_last)
                and x.shape[0] * self.conv.out_channels * x.shape[2] * x.shape[3] >= updates.fused_train_min_elems()):
            # training: bias-free cuDNN convolution (autograd's own dgrad / wgrad) + the fused differentiable epilogue
            y = nn.functional.conv2d(x, self.conv.weight, None, self.conv.stride, self.conv.padding)
            return updates.bias_prelu(y, self.conv.bias, self.act.weight)
        if (self.act is None and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and torch.is_grad_enabled()
                and x.is_contiguous(memory_format=torch.channels_last)
                and x.shape[0] * self.conv.out_channels * x.shape[2] * x.shape[3] >= updates.fused_train_min_elems()):
            return updates.bias_add(nn.functional.conv2d(x, self.conv.weight, None, self.conv.stride, self.conv.padding),
                                    self.conv.bias)
        y = self.conv(x)
        return self.act(y) if self.act is not None else y


class UpConv(nn.Module):
    """2x2 stride-2 transposed convolution with the same fused bias epilogue."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = nn.ConvTranspose2d(cin, cout, 2, stride=2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if _fused_epilogue_ok(x):
            y = nn.functional.conv_transpose2d(x, self.conv.weight, None, 2)
            return updates.bias_prelu_(y, self.conv.bias, None)
        if (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and torch.is_grad_enabled()
                and x.is_contiguous(memory_format=torch.channels_last)
                and 4 * x.shape[0] * self.conv.out_channels * x.shape[2] * x.shape[3] >= updates.fused_train_min_elems()):
            return updates.bias_add(nn.functional.conv_transpose2d(x, self.conv.weight, None, 2), self.conv.bias)
        return self.conv(x)


def _conv_block(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(ConvAct(cin, cout), ConvAct(cout, cout))


class UNet(nn.Module):
    """Plain 2-D UNet: `depth` poolings, `base` features doubling per level, 2x2 transposed-convolution
    up-sampling (the original UNet's "up-conv"; every layer is a cuDNN convolution)."""

    def __init__(self, cin: int, cout: int, base: int = 32, depth: int = 3):
        super().__init__()
        self.down = nn.ModuleList()
        ch = cin
        for d in range(depth + 1):
            self.down.append(_conv_block(ch, base << d))
            ch = base << d
        self.upconv = nn.ModuleList()
        self.up = nn.ModuleList()
        for d in reversed(range(depth)):
            self.upconv.append(UpConv(ch, base << d))
            self.up.append(_conv_block(2 * (base << d), base << d))
            ch = base << d
        self.head = ConvAct(ch, cout, act=False, kernel=1)
        self.pool = nn.MaxPool2d(2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self._fused_ok(x):
            return self._forward_fused(x)
        skips = []
        for i, blk in enumerate(self.down):
            x = blk(x)
            if i + 1 < len(self.down):
                skips.append(x)
                x = self.pool(x)
        for upc, blk in zip(self.upconv, self.up):
            s = skips.pop()
            x = upc(x)
            if x.shape[-2:] != s.shape[-2:]:           # odd sizes: pad the up-sampled map to the skip
                x = nn.functional.pad(x, (0, s.shape[-1] - x.shape[-1], 0, s.shape[-2] - x.shape[-2]))
            x = blk(torch.cat([x, s], dim=1))
        return self.head(x)

    def _fused_ok(self, x: torch.Tensor) -> bool:
        depth = len(self.upconv)
        return (_fused_epilogue_ok(x) and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)
                and x.shape[-1] % (1 << depth) == 0 and x.shape[-2] % (1 << depth) == 0
                and self.down[0][0].conv.out_channels % 4 == 0)

    def _forward_fused(self, x: torch.Tensor) -> torch.Tensor:
        """Inference path: every encoder block ends in ONE pass (pdu_bias_prelu_place_f32) that applies bias + PReLU,
        drops the result into the skip half of the decoder's concatenation buffer and emits the 2x2 max-pooled map;
        the up-convolution's bias epilogue fills the other half -- no separate max_pool2d or torch.cat passes."""
        cats = []
        for i, blk in enumerate(self.down):
            x = blk[0](x)
            if i + 1 == len(self.down):
                x = blk[1](x)
                break
            last = blk[1]
            y = nn.functional.conv2d(x, last._weight_for(x.shape[1]), None, 1, last.conv.padding)
            B, Cn, H, W = y.shape
            cat = torch.empty((B, 2 * Cn, H, W), dtype=y.dtype, device=y.device, memory_format=torch.channels_last)
            x = torch.empty((B, Cn, H // 2, W // 2), dtype=y.dtype, device=y.device, memory_format=torch.channels_last)
            updates.bias_prelu_place_(y, last.conv.bias, last.act.weight, cat[:, Cn:], x)
            cats.append(cat)
        for upc, blk in zip(self.upconv, self.up):
            cat = cats.pop()
            Cn = cat.shape[1] // 2
            y = nn.functional.conv_transpose2d(x, upc.conv.weight, None, 2)
            updates.bias_prelu_place_(y, upc.conv.bias, None, cat[:, :Cn], None)
            x = blk(cat)
        return self.head(x)


class DualBlock(nn.Module):
    """Three 3x3 convolutions with PReLU, as in the learned primal-dual data-space block."""

    def __init__(self, cin: int, cout: int, features: int = 32):
        super().__init__()
        self.net = nn.Sequential(ConvAct(cin, features), ConvAct(features, features), ConvAct(features, cout, act=False))

    def forward(self, x):
        return self.net(x)


class PrimalDualUNet(nn.Module):
    """Generic unrolling over a pair of callables.

    op_forward:  [B, kc, *image] float32 -> [B, kc_data, *data] float32
    op_adjoint:  [B, kc_data, *data]     -> [B, kc, *image]
    """

    def __init__(self, op_forward: Callable, op_adjoint: Callable, image_channels: int = 1, data_channels: int = 1,
                 n_iter: int = 4, n_primal: int = 4, n_dual: int = 4, unet_base: int = 32, unet_depth: int = 3,
                 dual_features: int = 32, op_scale: float = 1.0, channels_last: bool = True):
        super().__init__()
        if n_primal < image_channels or n_dual < data_channels:
            raise ValueError("n_primal / n_dual must hold at least one operator-sized slice")
        self.op_forward, self.op_adjoint = op_forward, op_adjoint
        self.kc, self.kd = image_channels, data_channels
        self.n_iter, self.n_primal, self.n_dual = n_iter, n_primal, n_dual
        self.op_scale = float(op_scale)
        self.dual = nn.ModuleList(DualBlock(n_dual + 2 * data_channels, n_dual, dual_features) for _ in range(n_iter))
        self.primal = nn.ModuleList(UNet(n_primal + image_channels, n_primal, unet_base, unet_depth)
                                    for _ in range(n_iter))
        # channels-last activations and weights: cuDNN's tensor-core convolutions run without the
        # NCHW<->NHWC conversion kernels, and the fused cat / residual kernels follow the same layout
        self.channels_last = bool(channels_last)
        if self.channels_last:
            self.to(memory_format=torch.channels_last)

    def forward(self, g: Optional[torch.Tensor], image_shape, dual_cat: Optional[Callable] = None,
                data_like: Optional[torch.Tensor] = None, data_shape=None) -> torch.Tensor:
        """g: measured data on the full grid [B, kd, *data].  Returns the reconstruction [B, kc, *image].
        dual_cat (with g None): callable (h, kf, scale_b, pad_to) -> the dual block's input; the CT flavour passes the
        concatenation that interpolates the sparse-view sinogram on the fly, so the full-view g never exists
        (data_like gives dtype / device, data_shape the [*data] axes)."""
        ref = g if g is not None else data_like
        dshape = tuple(g.shape[2:]) if g is not None else tuple(data_shape)
        B = ref.shape[0]
        fmt = torch.channels_last if (self.channels_last and len(dshape) == 2) else torch.contiguous_format
        h = torch.empty((B, self.n_dual) + dshape, dtype=ref.dtype, device=ref.device, memory_format=fmt).zero_()
        f = torch.empty((B, self.n_primal) + tuple(image_shape), dtype=ref.dtype, device=ref.device, memory_format=fmt).zero_()
        f_op = ref.new_zeros((B, self.kc) + tuple(image_shape))
        inv = 1.0 / self.op_scale
        # inference: round the concatenated channel count up to 8 (zero channels; the first convolution pads
        # its weights to match) -- with 6 or 5 input channels cuDNN falls back to a CUDA-core kernel
        pad = 8 if (ref.is_cuda and not torch.is_grad_enabled()) else 0
        if dual_cat is None:
            dual_cat = lambda hh, kk, scale_b, pad_to: updates.concat(hh, kk, g, scale_b=scale_b, pad_to=pad_to)
        for i in range(self.n_iter):
            # the 1/op_scale normalisation of each operator output rides in the concat kernel
            # the primal state starts at zero and the operator is linear: K 0 = 0, so the first projection is skipped
            kf = self.op_forward(f_op) if i > 0 else ref.new_zeros((B, self.kd) + dshape)
            h, h_op = updates.residual_slice(h, self.dual[i](dual_cat(h, kf, inv, pad)), 0, self.kd)
            kth = self.op_adjoint(h_op)
            f, f_op = updates.residual_slice(f, self.primal[i](updates.concat(f, kth, scale_b=inv, pad_to=pad)), 0, self.kc)
        return f_op


class PrimalDualUNetCT(PrimalDualUNet):
    """CT flavour: sparse-view sinogram [B, 1, A_sparse, D] -> image [B, 1, N, N].

    radon: a pd_unet_b200.radon.Radon / RadonFanbeam over the FULL view set (A_sparse * upsample views).
    adjoint: 'fbp' (ramp filter + backprojection) or 'backprojection'."""

    def __init__(self, radon: _BaseRadon, upsample: int = 8, adjoint: str = "fbp", **kw):
        if adjoint not in ("fbp", "backprojection"):
            raise ValueError("adjoint must be 'fbp' or 'backprojection'")
        if radon.n_angles % upsample:
            raise ValueError("the full view count must be a multiple of the upsampling factor")
        self_radon = radon
        back = (lambda s: self_radon.fbp(s)) if adjoint == "fbp" else (lambda s: self_radon.backprojection(s))
        kw.setdefault("op_scale", float(radon.resolution))
        super().__init__(lambda x: self_radon.forward(x), back, 1, 1, **kw)
        self.radon = radon                      # plain object: contributes nothing to state_dict
        self.upsample = int(upsample)
        self.wrap = "flip" if radon.geom.geom == 0 else "periodic"

    def forward(self, sparse_sino: torch.Tensor) -> torch.Tensor:
        """sparse_sino [B, 1, A_sparse, D].  The angular upsampling to the full view set is fused with the dual
        update: every iteration's concatenation interpolates the sparse views on the fly (1 / op_scale included)."""
        n = self.radon.resolution
        inv = 1.0 / self.op_scale
        sparse = sparse_sino.contiguous()
        if not sparse.is_cuda:                   # host tensors (the CPU plumbing tests): the plain two-step form
            g = updates.angular_upsample(sparse, self.upsample, self.wrap) * inv
            return super().forward(g, (n, n))
        cat = lambda h, kf, scale_b, pad_to: updates.concat_upsampled(h, kf, sparse, self.upsample, self.wrap, scale_b=scale_b,
                                                                       scale_c=inv, pad_to=pad_to)
        dshape = (sparse.shape[-2] * self.upsample, sparse.shape[-1])
        return super().forward(None, (n, n), dual_cat=cat, data_like=sparse, data_shape=dshape)


class PrimalDualUNetMRI(PrimalDualUNet):
    """Radial-MRI flavour: k-space samples [B, coils, spokes * readout] complex64 (+ trajectory, coil maps, density
    compensation) -> complex image [B, 1, N, N].

    Complex tensors travel through the CNNs as (re, im) channel pairs: data-space tensors are
    [B, 2 coils, spokes, readout] float32 and the image is [B, 2, N, N] -- exactly the `split` layouts of
    pd_unet_b200.nufft, so the operators read and write them directly: no permute / contiguous / view_as_complex pass
    and no separate density-compensation multiply surround a NUFFT call (the adjoint takes the weights as `kweight`).

    Spoke upsampling (BASELINE.json configs[0]: "32 spokes upsampled to 256"): when `forward` is given the sparse
    acquisition (`omega_sparse`, optionally `dcf_sparse`) the measured samples are first regridded onto the model's
    full trajectory, g = A_full A_sparse^H (dcf_sparse y) -- the step the reference does with torchkbnufft on the CPU --
    and the unrolled iterations then run on the full set of spokes, as the CT flavour does after its angular
    upsampling."""

    def __init__(self, im_size, n_spokes: int, n_readout: int, coils: int = 1, **kw):
        self.im_size = tuple(im_size)
        self.n_spokes, self.n_readout, self.coils = n_spokes, n_readout, coils
        self._omega: Optional[torch.Tensor] = None
        self._smaps: Optional[torch.Tensor] = None
        self._dcf: Optional[torch.Tensor] = None
        kw.setdefault("op_scale", float(im_size[0]))
        nn.Module.__init__(self)
        fwd, adj = KbNufft(im_size), KbNufftAdjoint(im_size)

        def op_forward(x):               # [B, 2, N, N] -> [B, 2 coils, spokes, readout]
            y = fwd(x, self._omega, smaps=self._smaps, norm="ortho", split=True)
            return y.view(x.shape[0], 2 * self.coils, n_spokes, n_readout)

        def op_adjoint(y):               # [B, 2 coils, spokes, readout] -> [B, 2, N, N], density compensated
            z = y.reshape(y.shape[0], 2 * self.coils, n_spokes * n_readout)
            return adj(z, self._omega, smaps=self._smaps, norm="ortho", split=True, kweight=self._dcf)

        PrimalDualUNet.__init__(self, op_forward, op_adjoint, 2, 2 * coils, **kw)
        self.nufft, self.nufft_adjoint = fwd, adj

    @staticmethod
    def _real_weights(dcf: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """Density compensation as float32 [M] (calc_density_compensation_function returns complex [1, 1, M])."""
        if dcf is None:
            return None
        w = dcf.real if dcf.is_complex() else dcf
        return w.reshape(-1).to(torch.float32).contiguous()

    def upsample_spokes(self, kdata: torch.Tensor, omega_sparse: torch.Tensor, omega_full: torch.Tensor,
                        smaps: Optional[torch.Tensor] = None, dcf_sparse: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Sparse acquisition [B, coils, M_sparse] complex64 -> the same object sampled on the full trajectory, in the
        model's data layout [B, 2 coils, M_full] float32: A_full A_sparse^H (dcf_sparse y)."""
        B = kdata.shape[0]
        ys = torch.view_as_real(kdata).permute(0, 1, 3, 2).reshape(B, 2 * self.coils, -1).contiguous()
        x = self.nufft_adjoint(ys, omega_sparse, smaps=smaps, norm="ortho", split=True, kweight=self._real_weights(dcf_sparse))
        return self.nufft(x, omega_full, smaps=smaps, norm="ortho", split=True)

    def forward(self, kdata: torch.Tensor, omega: torch.Tensor, smaps: Optional[torch.Tensor] = None,
                dcf: Optional[torch.Tensor] = None, omega_sparse: Optional[torch.Tensor] = None,
                dcf_sparse: Optional[torch.Tensor] = None) -> torch.Tensor:
        if kdata.shape[1] != self.coils:
            raise ValueError(f"kdata has {kdata.shape[1]} coils, the model was built for {self.coils}")
        if self.coils > 1 and smaps is None:
            raise ValueError("multi-coil data needs smaps")
        self._omega, self._smaps, self._dcf = omega, smaps, self._real_weights(dcf)
        B = kdata.shape[0]
        if omega_sparse is not None:
            g = self.upsample_spokes(kdata, omega_sparse, omega, smaps, dcf_sparse)
        else:
            g = torch.view_as_real(kdata).permute(0, 1, 3, 2).reshape(B, 2 * self.coils, -1)
        g = (g * (1.0 / self.op_scale)).reshape(B, 2 * self.coils, self.n_spokes, self.n_readout).contiguous()
        out = PrimalDualUNet.forward(self, g, self.im_size)          # [B, 2, N, N]
        return torch.complex(out[:, 0], out[:, 1])[:, None]
