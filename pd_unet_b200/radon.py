"""CT operators with the call shapes of torch_radon v1 ([RECALL]; the
reference names the library at /root/reference/README.md:3-5 only through its unmounted branches):

    Radon(resolution, angles, det_count=-1, det_spacing=1.0, clip_to_circle=False)
    RadonFanbeam(resolution, angles, source_distance, det_distance=-1, det_count=-1,
                 det_spacing=-1, clip_to_circle=False)
    .forward(x)  .backprojection(sino)  .backward(sino)  .filter_sinogram(sino, filter_name="ramp")

Inputs are float32 CUDA tensors with any leading batch dimensions; the last two are the image
[N, N] or the sinogram [n_angles, det_count].  Both directions are differentiable, each using the
other as its gradient exactly as the library does.  Like torch_radon's, these operators are plain
objects (not nn.Modules): they hold no parameters and add nothing to a model's state_dict.

Everything numeric happens in libpdu_b200.so (pd_unet_b200/csrc/radon_fwd.cu, radon_adj.cu,
filter.cu).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import numpy as np
import torch

from ._lib import PDU_GEOM_FAN, PDU_GEOM_PARALLEL, RadonGeomC, check, lib, require_cuda, stream_ptr

# bench.py sets this to time operator calls with CUDA events inside a larger step:
# EVENT_HOOK(kind) -> (start_event, end_event) or None, kind in {"fwd", "adj", "filter"}.
EVENT_HOOK = None


class _Timed:
    def __init__(self, kind):
        self.pair = EVENT_HOOK(kind) if EVENT_HOOK is not None else None

    def __enter__(self):
        if self.pair:
            self.pair[0].record()

    def __exit__(self, *exc):
        if self.pair:
            self.pair[1].record()


FILTERS = ("ramp", "ram-lak", "shepp-logan", "cosine", "hamming", "hann")


def _fourier_filter(size: int, name: str) -> np.ndarray:
    """Frequency response of the band-limited ramp (Kak & Slaney eq. 61, the construction
    scikit-image and [RECALL] torch_radon's FourierFilters use), optionally windowed."""
    n = np.concatenate((np.arange(1, size / 2 + 1, 2, dtype=np.int64), np.arange(size / 2 - 1, 0, -2, dtype=np.int64)))
    f = np.zeros(size)
    f[0] = 0.25
    f[1::2] = -1.0 / (np.pi * n) ** 2
    resp = 2.0 * np.real(np.fft.fft(f))
    name = name.lower()
    if name in ("ramp", "ram-lak"):
        return resp
    if name == "shepp-logan":
        omega = np.pi * np.fft.fftfreq(size)[1:]
        resp[1:] *= np.sin(omega) / omega
    elif name == "cosine":
        resp *= np.fft.fftshift(np.sin(np.linspace(0, np.pi, size, endpoint=False)))
    elif name == "hamming":
        resp *= np.fft.fftshift(np.hamming(size))
    elif name == "hann":
        resp *= np.fft.fftshift(np.hanning(size))
    else:
        raise ValueError(f"unknown filter {name!r}; choose from {FILTERS}")
    return resp


def filter_taps(det_count: int, n_angles: int, filter_name: str = "ramp") -> np.ndarray:
    """float64 [2 D - 1] spatial taps h[-(D-1) .. D-1], already scaled by pi / (2 n_angles): the
    zero-padded circular FFT product of filter_sinogram restricted to the D kept samples."""
    padded = max(64, int(2 ** math.ceil(math.log2(2 * det_count))))
    h = np.real(np.fft.ifft(_fourier_filter(padded, filter_name)))
    k = np.arange(-(det_count - 1), det_count)
    return h[k % padded] * (math.pi / (2.0 * n_angles))


class _Forward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, op):
        ctx.op = op
        return op._project(x)

    @staticmethod
    def backward(ctx, grad):
        return ctx.op._backproject(grad), None


class _Backprojection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sino, op):
        ctx.op = op
        return op._backproject(sino)

    @staticmethod
    def backward(ctx, grad):
        return ctx.op._project(grad), None


class _FanFbp(torch.autograd.Function):
    """Fan-beam FBP (cosine pre-weight fused into the filter, 1 / U^2 weighting in the backprojector).  Inference only:
    the transpose of the distance-weighted backprojector is a pixel-driven scatter this library does not have, and
    unlike for the plain pair there is no partner operator to stand in as its gradient."""

    @staticmethod
    def forward(ctx, sino, op, name):
        return op._backproject(op._filter(sino, name, weighted=True), fbp_weight=True)

    @staticmethod
    def backward(ctx, grad):
        raise NotImplementedError("fan-beam FBP with its distance weighting is not differentiable here; train through "
                                  "RadonFanbeam.backprojection (adjoint='backprojection') or fbp(..., fan_weights=False)")


class _Filter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sino, op, name):
        ctx.op, ctx.name = op, name
        return op._filter(sino, name)

    @staticmethod
    def backward(ctx, grad):
        # every supported response is real and even, so the Toeplitz matrix is symmetric
        return ctx.op._filter(grad, ctx.name), None, None


class _BaseRadon:
    def __init__(self, resolution: int, angles, det_count: int, det_spacing: float, clip_to_circle: bool,
                 geom: int, s_dist: float = 0.0, d_dist: float = 0.0):
        if isinstance(angles, torch.Tensor):
            angles = angles.detach().cpu().numpy()
        angles = np.asarray(angles, dtype=np.float64).reshape(-1)
        if angles.size == 0:
            raise ValueError("angles is empty")
        if resolution <= 0:
            raise ValueError("resolution must be positive")
        self.resolution = int(resolution)
        self.angles = angles                      # as the user gave them
        self._internal = -angles                  # [RECALL] torch_radon BaseRadon negates once
        self.det_count = int(det_count) if det_count > 0 else self.resolution
        self.det_spacing = float(det_spacing)
        self.clip_to_circle = bool(clip_to_circle)
        self.geom = RadonGeomC(geom, self.resolution, int(angles.size), self.det_count, self.det_spacing,
                               float(s_dist), float(d_dist), int(self.clip_to_circle))
        # (cos, sin) in float64, rounded once: the kernels never evaluate a trigonometric function
        self._trig_host = np.stack([np.cos(self._internal), np.sin(self._internal)], axis=-1).astype(np.float32)
        self._trig: Dict[torch.device, torch.Tensor] = {}
        self._taps: Dict[Tuple[torch.device, str], torch.Tensor] = {}
        self._cosw: Dict[torch.device, torch.Tensor] = {}

    # ------------------------------------------------------------------ public API
    @property
    def n_angles(self) -> int:
        return int(self.angles.size)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [..., N, N] -> sinogram [..., n_angles, det_count]."""
        return _Forward.apply(x, self)

    def backprojection(self, sinogram: torch.Tensor) -> torch.Tensor:
        """sinogram [..., n_angles, det_count] -> image [..., N, N] (unfiltered adjoint-type sum)."""
        return _Backprojection.apply(sinogram, self)

    def backward(self, sinogram: torch.Tensor) -> torch.Tensor:
        return self.backprojection(sinogram)

    def filter_sinogram(self, sinogram: torch.Tensor, filter_name: str = "ramp") -> torch.Tensor:
        return _Filter.apply(sinogram, self, filter_name)

    def fbp(self, sinogram: torch.Tensor, filter_name: str = "ramp", fan_weights: bool = True) -> torch.Tensor:
        """Filtered backprojection.  Parallel beam: backprojection(filter_sinogram(s)).  Fan beam (views over 2 pi, flat
        equispaced detector), fan_weights=True: Kak & Slaney section 3.4.2 -- the projections are multiplied by the cosine
        of the fan angle (inside the filter kernel), ramp filtered, and backprojected with the 1 / U^2 distance weight;
        fan_weights=False gives the unweighted composition (differentiable, like the parallel-beam form)."""
        if self.geom.geom == PDU_GEOM_FAN and fan_weights:
            return _FanFbp.apply(sinogram, self, filter_name)
        return self.backprojection(self.filter_sinogram(sinogram, filter_name))

    __call__ = forward

    # ------------------------------------------------------------------ C-ABI calls
    def _trig_on(self, device: torch.device) -> torch.Tensor:
        t = self._trig.get(device)
        if t is None:
            t = torch.from_numpy(self._trig_host).to(device)
            self._trig[device] = t
        return t

    def _project(self, x: torch.Tensor) -> torch.Tensor:
        x = require_cuda(x, torch.float32, "image")
        n, A, D = self.resolution, self.n_angles, self.det_count
        if x.dim() < 2 or x.shape[-1] != n or x.shape[-2] != n:
            raise ValueError(f"image must end in [{n}, {n}], got {tuple(x.shape)}")
        lead = x.shape[:-2]
        flat = x.reshape(-1, n, n)
        out = torch.empty((flat.shape[0], A, D), dtype=torch.float32, device=x.device)
        if flat.shape[0] == 0:
            return out.reshape(*lead, A, D)
        with torch.cuda.device(x.device):
            L = lib()
            ws_bytes = L.pdu_radon_workspace_bytes(C.byref(self.geom), 1)
            trig = self._trig_on(x.device)
            timer = _Timed("fwd")
            timer.__enter__()
            for b0 in range(0, flat.shape[0], 65535):
                part = flat[b0:b0 + 65535]
                ws = torch.empty(ws_bytes * part.shape[0], dtype=torch.uint8, device=x.device)
                check(L.pdu_radon_fwd_f32(part.data_ptr(), out[b0:].data_ptr(), trig.data_ptr(), part.shape[0],
                                          C.byref(self.geom), ws.data_ptr(), ws.numel(), stream_ptr()),
                      "pdu_radon_fwd_f32")
            timer.__exit__()
        return out.reshape(*lead, A, D)

    def _backproject(self, s: torch.Tensor, fbp_weight: bool = False) -> torch.Tensor:
        s = require_cuda(s, torch.float32, "sinogram")
        n, A, D = self.resolution, self.n_angles, self.det_count
        if s.dim() < 2 or s.shape[-1] != D or s.shape[-2] != A:
            raise ValueError(f"sinogram must end in [{A}, {D}], got {tuple(s.shape)}")
        lead = s.shape[:-2]
        flat = s.reshape(-1, A, D)
        out = torch.empty((flat.shape[0], n, n), dtype=torch.float32, device=s.device)
        if flat.shape[0] == 0:
            return out.reshape(*lead, n, n)
        with torch.cuda.device(s.device):
            L = lib()
            trig = self._trig_on(s.device)
            timer = _Timed("adj")
            timer.__enter__()
            for b0 in range(0, flat.shape[0], 65535):
                part = flat[b0:b0 + 65535]
                check(L.pdu_radon_adj_weighted_f32(part.data_ptr(), out[b0:].data_ptr(), trig.data_ptr(), part.shape[0],
                                                   C.byref(self.geom), 1 if fbp_weight else 0, None, 0, stream_ptr()),
                      "pdu_radon_adj_f32")
            timer.__exit__()
        return out.reshape(*lead, n, n)

    def fan_cosine_weights(self) -> np.ndarray:
        """float64 [D]: cos of the fan angle of every detector bin, (s + d) / sqrt((s + d)^2 + u^2)."""
        u = (np.arange(self.det_count, dtype=np.float64) + 0.5 - self.det_count / 2.0) * self.det_spacing
        k = float(self.geom.s_dist) + float(self.geom.d_dist)
        return k / np.sqrt(k * k + u * u)

    def _filter(self, s: torch.Tensor, name: str, weighted: bool = False) -> torch.Tensor:
        s = require_cuda(s, torch.float32, "sinogram")
        A, D = self.n_angles, self.det_count
        if s.dim() < 2 or s.shape[-1] != D or s.shape[-2] != A:
            raise ValueError(f"sinogram must end in [{A}, {D}], got {tuple(s.shape)}")
        out = torch.empty_like(s)
        rows = s.numel() // D
        if rows == 0:
            return out
        key = (s.device, name.lower())
        with torch.cuda.device(s.device):
            L = lib()
            entry = self._taps.get(key)
            if entry is None:
                taps = torch.from_numpy(filter_taps(D, A, name).astype(np.float32)).to(s.device)
                ws = torch.empty(max(1, L.pdu_filter_workspace_bytes(D)), dtype=torch.uint8, device=s.device)
                check(L.pdu_filter_prepare_f32(taps.data_ptr(), ws.data_ptr(), ws.numel(), D, stream_ptr()),
                      "pdu_filter_prepare_f32")
                entry = (taps, ws)
                self._taps[key] = entry
            taps, ws = entry
            cw = None
            if weighted:
                cw = self._cosw.get(s.device)
                if cw is None:
                    cw = torch.from_numpy(self.fan_cosine_weights().astype(np.float32)).to(s.device)
                    self._cosw[s.device] = cw
            with _Timed("filter"):
                check(L.pdu_filter_sinogram_weighted_f32(s.data_ptr(), out.data_ptr(), taps.data_ptr(),
                                                         cw.data_ptr() if cw is not None else None, ws.data_ptr(),
                                                         ws.numel(), rows, D, stream_ptr()), "pdu_filter_sinogram_f32")
        return out


class Radon(_BaseRadon):
    """Parallel-beam projector.  [RECALL] torch_radon.Radon."""

    def __init__(self, resolution: int, angles, det_count: int = -1, det_spacing: float = 1.0,
                 clip_to_circle: bool = False):
        super().__init__(resolution, angles, det_count, det_spacing, clip_to_circle, PDU_GEOM_PARALLEL)


class RadonFanbeam(_BaseRadon):
    """Fan-beam projector with a flat equispaced detector.  [RECALL] torch_radon.RadonFanbeam:
    det_distance < 0 means "same as source_distance"; det_spacing < 0 means "the magnification
    (source_distance + det_distance) / source_distance", so the detector covers the object."""

    def __init__(self, resolution: int, angles, source_distance: float, det_distance: float = -1,
                 det_count: int = -1, det_spacing: float = -1, clip_to_circle: bool = False):
        if source_distance <= 0:
            raise ValueError("source_distance must be positive")
        if det_distance < 0:
            det_distance = source_distance
        if det_spacing < 0:
            det_spacing = (source_distance + det_distance) / source_distance
        self.source_distance = float(source_distance)
        self.det_distance = float(det_distance)
        super().__init__(resolution, angles, det_count, det_spacing, clip_to_circle, PDU_GEOM_FAN,
                         source_distance, det_distance)
