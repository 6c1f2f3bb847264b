"""Radial-MRI operators with the call shapes of torchkbnufft >= 1.0 ([RECALL], SURVEY.md section 8b;
the reference reaches the library through its unmounted MRI branch, /root/reference/README.md:3-5):

    KbNufft(im_size, grid_size=None, numpoints=6, n_shift=None, table_oversamp=2**10,
            kbwidth=2.34, order=0.0)(image, omega, interp_mats=None, smaps=None, norm=None)
    KbNufftAdjoint(...)(data, omega, interp_mats=None, smaps=None, norm=None)
    KbInterp(...)(image, omega) / KbInterpAdjoint(...)(data, omega)
    calc_density_compensation_function(ktraj, im_size, num_iterations=10, ...)

image: complex64 [B, C, N0, N1]; omega: float32 [2, M] radians (or [B, 2, M]); data: complex64
[B, C, M].  With smaps [1 or B, coils, N0, N1] the forward expands a one-channel image to the
coils and the adjoint combines them.  Both directions are differentiable (each is the other's
conjugate transpose).  The modules register no parameters and no persistent buffers, so a model
containing them has the same state_dict as one without.

Everything numeric happens in libpdu_b200.so (pd_unet_b200/csrc/nufft.cu).  No CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from ._lib import NUFFT_IMAGE_SPLIT, NUFFT_KDATA_SPLIT, PduError, check, lib, require_cuda, stream_ptr


# ----------------------------------------------------------------------------- tables (host, float64)
def kaiser_bessel_table(n: int, k: int, numpoints: int, table_oversamp: int, kbwidth: float) -> np.ndarray:
    """complex128 [J L + 1]: kb(u) exp(-i (2 pi / K) ((N - 1) / 2) u) at u = q / L - J / 2
    (Fessler's NUFFT table, order 0)."""
    J, L = numpoints, table_oversamp
    u = np.arange(J * L + 1, dtype=np.float64) / L - J / 2.0
    inside = np.abs(u) < J / 2.0
    arg = np.sqrt(np.where(inside, 1.0 - (u / (J / 2.0)) ** 2, 0.0))
    alpha = kbwidth * J
    kb = np.where(inside, np.i0(alpha * arg) / np.i0(alpha), 0.0)
    return kb * np.exp(-1j * (2.0 * np.pi / k) * ((n - 1) / 2.0) * u)


def kaiser_bessel_scaling(n: int, k: int, numpoints: int, kbwidth: float) -> np.ndarray:
    """float64 [N]: reciprocal of the Kaiser-Bessel kernel's Fourier transform (apodisation correction)."""
    J = numpoints
    alpha = kbwidth * J
    om = (np.arange(n, dtype=np.float64) - (n - 1) / 2.0) / k
    w2 = alpha ** 2 - (np.pi * J * om) ** 2
    w = np.sqrt(np.abs(w2))
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(w2 > 0, np.sinh(w) / w, np.sin(w) / w)
    ratio = np.where(w == 0, 1.0, ratio)
    return np.i0(alpha) / (J * ratio)


class _Plan:
    """Owns one pdu_nufft_plan_t per device."""

    def __init__(self, im_size, grid_size, numpoints, n_shift, table_oversamp, kbwidth, order):
        if len(im_size) != 2:
            raise NotImplementedError("only 2-D transforms are implemented (radial MRI slices)")
        if float(order) != 0.0:
            raise NotImplementedError("only order-0 Kaiser-Bessel kernels are implemented")
        self.im_size = tuple(int(v) for v in im_size)
        self.grid_size = tuple(int(v) for v in (grid_size or [2 * n for n in self.im_size]))
        self.n_shift = tuple(int(v) for v in (n_shift or [n // 2 for n in self.im_size]))
        self.numpoints = int(numpoints if np.isscalar(numpoints) else numpoints[0])
        self.table_oversamp = int(table_oversamp if np.isscalar(table_oversamp) else table_oversamp[0])
        self.kbwidth = float(kbwidth)
        self.tables = [kaiser_bessel_table(n, k, self.numpoints, self.table_oversamp, self.kbwidth).astype(np.complex64)
                       for n, k in zip(self.im_size, self.grid_size)]
        self.scaling = [kaiser_bessel_scaling(n, k, self.numpoints, self.kbwidth).astype(np.float32)
                        for n, k in zip(self.im_size, self.grid_size)]
        self._handles: Dict[int, C.c_void_p] = {}
        # per-trajectory precomputation (row bins of the fused path, CSR form of the adjoint interpolator), keyed by
        # the trajectory tensor's identity; a fixed trajectory is reused by every unrolled iteration / training step
        self._traj: "OrderedDict[tuple, dict]" = OrderedDict()
        self.traj_cache_bytes = 512 << 20
        # "auto": the gather where it is faster (measured on B200: >= 8 planes per call, see _csr_for; with fewer
        # planes the atomic scatter wins -- 67 vs 180 us at 256 spokes x 1 plane); True: always (bit-reproducible
        # adjoint); False: never.  Only the generic path (grids without a fused path, KbInterpAdjoint) looks at it.
        self.use_csr = "auto"
        # the fused path (csrc/nufft_fused.cu): True -- wherever the library has one for the grid; False -- never;
        # "auto" -- where it measured faster on B200 (tools/prof_nufft.py, profiles/r02_nufft_timings.txt): from 8 planes
        # per call up (below that a call is ~700 short-lived CTAs and the generic path's wide gathers win: 24 vs 36 us
        # at 256^2 x 1 plane), and for the adjoint only while the plane count is moderate (_bins_for: 64 planes of
        # 640^2 with coil maps: fused 294 us, generic sorted gather + register FFT 253 us; forward 206 vs 261 us)
        self.use_fused = "auto"
        self.fused_max_row = 4096     # "auto": no fused path for a trajectory whose heaviest grid row has more entries

    # -------------------------------------------------------------- per-trajectory cache
    def _entry(self, omega: torch.Tensor) -> dict:
        key = (omega.data_ptr(), omega._version, tuple(omega.shape), omega.device)
        ent = self._traj.get(key)
        if ent is None:
            ent = {"omega": omega, "bins": None, "csr": None, "events": []}   # omega kept alive: its address cannot be reused
            self._traj[key] = ent
        else:
            self._traj.move_to_end(key)
        return ent

    def _trim(self) -> None:
        def nbytes(e):
            return sum(t.numel() for t in (e["bins"], e["csr"]) if t is not None)
        while len(self._traj) > 1 and (len(self._traj) > 16 or sum(nbytes(e) for e in self._traj.values()) > self.traj_cache_bytes):
            self._traj.popitem(last=False)

    @staticmethod
    def _built(ent: dict) -> None:
        """Remember where the build ran so that a later use on another stream can wait for it."""
        ev = torch.cuda.Event()
        ev.record()
        ent["events"].append((torch.cuda.current_stream(), ev))

    @staticmethod
    def _wait(ent: dict) -> None:
        cur = torch.cuda.current_stream()
        if torch.cuda.is_current_stream_capturing():
            return                      # graph capture follows eager warm-up calls and a device synchronisation
        for st, ev in ent["events"]:
            if st != cur:
                cur.wait_event(ev)

    def _keep_prefix(self, buf: torch.Tensor, persist: int) -> torch.Tensor:
        """Only the head of a build buffer is needed afterwards; the sort scratch behind it (about two thirds) goes
        back to the allocator (stream-ordered, so the pending build kernels are safe)."""
        return buf[:persist].clone() if persist < buf.numel() else buf

    def _bins_for(self, omega: torch.Tensor, planes: int = 1 << 30, adjoint: bool = False, split: bool = False):
        """Row bins of the fused path for this trajectory (None when the generic path is to be used)."""
        L, h = lib(), self.handle(omega.device)
        if not self.use_fused or not L.pdu_nufft_has_fused_path(h):
            return None
        if self.use_fused == "auto":
            # measured policy (profiles/r02_nufft_timings.txt, profiles/r02_nufft_policy.txt): the row-binned kernels pay
            # per (sample, row) entry, so they win for sparse trajectories -- samples per grid cell rho <= 0.15
            # (configs[3]: 0.075) -- and lose by 2 - 8 x for dense ones (rho >= 0.5).  (While the generic adjoint still
            # scattered with atomics below 16 planes the fused one was also used up to rho = 0.3 there; against the
            # sorted gather from 8 planes it no longer pays: rho = 0.25, 8 / 12 planes, generic / fused us: 128^2
            # 48 / 60, 48 / 61; 256^2 91 / 110, 118 / 118; 512^2 294 / 253, 339 / 521.)
            rho = omega.shape[1] / float(self.grid_size[0] * self.grid_size[1])
            sparse = rho <= 0.15
            # adjoint: from a grid-dependent plane count on, the sorted gather (4 lanes per cell x 16 planes, non-empty cells
            # only, since r02) + FFT passes beat the row-binned kernel (tools/prof_nufft_adj_policy.py, REPS=21, generic /
            # fused us: 128^2 x 32 spokes 16 / 48 / 64 planes 44 / 65 / 71 against 38 / 63 / 77; 256^2 x 48 16 / 24 / 32 / 64
            # 71 / 85 / 101 / 173 against 65 / 91 / 114 / 208; 320^2 x 48 16 / 24 / 32 / 64 89 / 114 / 138 / 245 against
            # 85 / 116 / 146 / 273; 512^2 x 96 8 / 16 / 32 124 / 208 / 392 against 124 / 226 / 484; 1024^2 x 128 8 / 16
            # 419 / 744 against 484 / 1111)
            many = adjoint and planes >= {256: 64, 512: 24, 640: 32, 1024: 16}.get(self.grid_size[0], 8)
            if planes < 8 or not sparse or many:
                return None
        ent = self._entry(omega)
        if ent["bins"] is None:
            persist = C.c_size_t(0)
            nbytes = L.pdu_nufft_bins_bytes(h, omega.shape[1], C.byref(persist))
            buf = torch.empty(nbytes, dtype=torch.uint8, device=omega.device)
            check(L.pdu_nufft_bins_build(h, omega.data_ptr(), omega.shape[1], buf.data_ptr(), buf.numel(), stream_ptr()),
                  "pdu_nufft_bins_build")
            ent["bins"] = self._keep_prefix(buf, persist.value)
            self._built(ent)
            self._trim()
            # entries of the heaviest grid row (header word 7): one small read-back per trajectory, at build time
            ent["max_row"] = (int(ent["bins"][:64].cpu().view(torch.int32)[7])
                              if not torch.cuda.is_current_stream_capturing() else 0)
        else:
            self._wait(ent)
        if self.use_fused == "auto" and ent.get("max_row", 0) > self.fused_max_row:
            return None       # a few very heavy rows (dense or concentrated trajectory) serialise the row kernels
        return ent["bins"]

    def _csr_for(self, omega: torch.Tensor, planes: int = 1 << 30):
        """The sorted-gather form of the adjoint interpolator for this trajectory, built on first use."""
        # from 8 planes (16 before the gather went to 4 lanes per cell; tools/prof_nufft_csr_policy.py, atomics / sorted us:
        # 8 planes of 512^2 x 256 spokes 406 / 282, of 320^2 x 48 83 / 73, of 256^2 x 256 251 / 266; 4 planes 134 / 175)
        # -- where the plane-interleaved copy of the samples fits the call's scratch (M <= about three quarters of the grid cells); a denser
        # trajectory runs the planar form of the gather, which only pays from 16 planes (256^2 x 512 spokes, 8 planes:
        # 396 against 374 us for the scatter; 16 planes: 527 against 722)
        if self.use_csr == "auto":
            n0, n1 = self.im_size
            k0, k1 = self.grid_size
            fits = omega.shape[1] * ((planes + 3) & ~3) <= planes * (max(n0 * k1, k0 * n1) + n0 * n1)
            if planes < (8 if fits else 16):
                return None
        if self.use_csr is False:
            return None
        ent = self._entry(omega)
        if ent["csr"] is None:
            L, h = lib(), self.handle(omega.device)
            persist = C.c_size_t(0)
            nbytes = L.pdu_nufft_csr_bytes2(h, omega.shape[1], C.byref(persist))
            buf = torch.empty(nbytes, dtype=torch.uint8, device=omega.device)
            check(L.pdu_nufft_csr_build(h, omega.data_ptr(), omega.shape[1], buf.data_ptr(), buf.numel(), stream_ptr()),
                  "pdu_nufft_csr_build")
            ent["csr"] = self._keep_prefix(buf, persist.value)
            self._built(ent)
            self._trim()
        else:
            self._wait(ent)
        return ent["csr"]

    def handle(self, device: torch.device) -> C.c_void_p:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            h = C.c_void_p()
            t0, t1 = (np.ascontiguousarray(t) for t in self.tables)
            s0, s1 = (np.ascontiguousarray(s) for s in self.scaling)
            with torch.cuda.device(idx):
                check(lib().pdu_nufft_plan_create(C.byref(h), self.im_size[0], self.im_size[1], self.grid_size[0],
                                                  self.grid_size[1], self.numpoints, self.table_oversamp,
                                                  self.n_shift[0], self.n_shift[1], t0.ctypes.data, t1.ctypes.data,
                                                  s0.ctypes.data, s1.ctypes.data), "pdu_nufft_plan_create")
            self._handles[idx] = h
        return h

    def __deepcopy__(self, memo):  # noqa: D105
        # device handles are per-process resources: a copy starts without any and re-creates them lazily
        return _Plan(self.im_size, self.grid_size, self.numpoints, self.n_shift, self.table_oversamp, self.kbwidth, 0.0)

    def __getstate__(self):
        raise TypeError("a NUFFT plan holds device handles and cannot be pickled; rebuild the module instead")

    def __del__(self):
        try:
            for h in self._handles.values():
                lib().pdu_nufft_plan_destroy(h)
        except Exception:
            pass

    def scale(self, norm: Optional[str]) -> float:
        if norm is None:
            return 1.0
        if norm == "ortho":
            return 1.0 / math.sqrt(self.grid_size[0] * self.grid_size[1])
        raise ValueError("norm must be None or 'ortho'")

    # -------------------------------------------------------------- C-ABI calls (single trajectory)
    def _omega(self, omega: torch.Tensor) -> torch.Tensor:
        omega = require_cuda(omega, torch.float32, "omega")
        if omega.dim() != 2 or omega.shape[0] != 2:
            raise ValueError(f"omega must be [2, M], got {tuple(omega.shape)}")
        return omega

    def _smaps(self, smaps, batch):
        if smaps is None:
            return None, 0, 1
        smaps = require_cuda(smaps, torch.complex64, "smaps")
        if smaps.dim() == 3:
            smaps = smaps[None]
        if smaps.dim() != 4 or tuple(smaps.shape[-2:]) != self.im_size or smaps.shape[0] not in (1, batch):
            raise ValueError(f"smaps must be [1 or {batch}, coils, {self.im_size[0]}, {self.im_size[1]}]")
        return smaps, smaps.shape[1], smaps.shape[0]

    # Layouts.  split=False: complex64 tensors, image [B, C, N0, N1], data [B, C, M] (torchkbnufft's).
    # split=True: float32 tensors with the real and imaginary parts as neighbouring channels, image [B, 2 C, N0, N1],
    # data [B, 2 C, M] -- what PD-UNet's CNN blocks carry, so the model needs no permute / view_as_complex passes.
    @staticmethod
    def _split_to_complex(x: torch.Tensor, weight: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B, 2 C, *s] float32 -> [B, C, *s] complex64 (x weight[*s]), one pass of the library's layout kernel."""
        B, C2 = x.shape[:2]
        out = torch.empty((B, C2 // 2) + tuple(x.shape[2:]), dtype=torch.complex64, device=x.device)
        n = out[0, 0].numel()
        if out.numel():
            check(lib().pdu_complex_from_split_f32(x.data_ptr(), out.data_ptr(), weight.data_ptr() if weight is not None else None,
                                                   B * (C2 // 2), n, stream_ptr()), "pdu_complex_from_split_f32")
        return out

    @staticmethod
    def _complex_to_split(z: torch.Tensor) -> torch.Tensor:
        B, Cc = z.shape[:2]
        out = torch.empty((B, 2 * Cc) + tuple(z.shape[2:]), dtype=torch.float32, device=z.device)
        if out.numel():
            check(lib().pdu_split_from_complex_f32(z.data_ptr(), out.data_ptr(), B * Cc, z[0, 0].numel(), stream_ptr()),
                  "pdu_split_from_complex_f32")
        return out

    def _chunks(self, L, h, B: int, coils: int, M: int):
        """Batch chunks that keep the fused path's scratch under 1 GiB."""
        per = max(1, L.pdu_nufft_binned_workspace_bytes(h, coils, M))
        cb = max(1, min(B, (1 << 30) // per))
        return [(b0, min(B, b0 + cb)) for b0 in range(0, B, cb)]

    def forward(self, image, omega, smaps, norm, split: bool = False):
        image = require_cuda(image, torch.float32 if split else torch.complex64, "image")
        if image.dim() != 4 or tuple(image.shape[-2:]) != self.im_size or (split and image.shape[1] % 2):
            raise ValueError(f"image must be [B, C, {self.im_size[0]}, {self.im_size[1]}], got {tuple(image.shape)}")
        omega = self._omega(omega)
        B, M = image.shape[0], omega.shape[1]
        ci = image.shape[1] // 2 if split else image.shape[1]
        smaps, coils, sb = self._smaps(smaps, B)
        if smaps is None:
            coils = ci
        elif ci != 1:
            raise ValueError("with smaps the image must have one channel")
        out = (torch.empty((B, 2 * coils, M), dtype=torch.float32, device=image.device) if split else
               torch.empty((B, coils, M), dtype=torch.complex64, device=image.device))
        if B == 0 or M == 0:
            return out
        with torch.cuda.device(image.device):
            L, h = lib(), self.handle(image.device)
            bins = self._bins_for(omega, B * coils, split=split)
            if bins is not None:
                flags = (NUFFT_IMAGE_SPLIT | NUFFT_KDATA_SPLIT) if split else 0
                for b0, b1 in self._chunks(L, h, B, coils, M):
                    nb = b1 - b0
                    ws = torch.empty(L.pdu_nufft_binned_workspace_bytes(h, nb * coils, M), dtype=torch.uint8, device=image.device)
                    sm = None if smaps is None else (smaps if sb == 1 else smaps[b0:b1])
                    check(L.pdu_nufft_fwd_binned_c64(h, image[b0:b1].data_ptr(), out[b0:b1].data_ptr(),
                                                     sm.data_ptr() if sm is not None else None, nb, coils, 1 if sb == 1 else nb, M,
                                                     self.scale(norm), bins.data_ptr(), flags, ws.data_ptr(), ws.numel(),
                                                     stream_ptr()), "pdu_nufft_fwd_binned_c64")
                return out
            if split:
                return self._complex_to_split(self.forward(self._split_to_complex(image), omega, smaps, norm))
            ws = torch.empty(L.pdu_nufft_workspace_bytes(h, B * coils), dtype=torch.uint8, device=image.device)
            check(L.pdu_nufft_fwd_c64(h, image.data_ptr(), out.data_ptr(), omega.data_ptr(),
                                      smaps.data_ptr() if smaps is not None else None, B, coils, sb, M,
                                      self.scale(norm), ws.data_ptr(), ws.numel(), stream_ptr()), "pdu_nufft_fwd_c64")
        return out

    def adjoint(self, data, omega, smaps, norm, split: bool = False, kweight: Optional[torch.Tensor] = None):
        """kweight: float32 [M] multiplied into the samples first (density compensation)."""
        data = require_cuda(data, torch.float32 if split else torch.complex64, "data")
        omega = self._omega(omega)
        if data.dim() != 3 or data.shape[-1] != omega.shape[1] or (split and data.shape[1] % 2):
            raise ValueError(f"data must be [B, C, {omega.shape[1]}], got {tuple(data.shape)}")
        B, M = data.shape[0], data.shape[2]
        coils = data.shape[1] // 2 if split else data.shape[1]
        smaps, sc, sb = self._smaps(smaps, B)
        if smaps is not None and sc != coils:
            raise ValueError(f"data has {coils} coils, smaps {sc}")
        if kweight is not None:
            kweight = require_cuda(kweight, torch.float32, "kweight").reshape(-1)
            if kweight.numel() != M:
                raise ValueError(f"kweight must have {M} entries")
        co = 1 if smaps is not None else coils
        out = (torch.empty((B, 2 * co) + self.im_size, dtype=torch.float32, device=data.device) if split else
               torch.empty((B, co) + self.im_size, dtype=torch.complex64, device=data.device))
        if B == 0:
            return out
        if M == 0:
            return out.zero_()
        with torch.cuda.device(data.device):
            L, h = lib(), self.handle(data.device)
            bins = self._bins_for(omega, B * coils, adjoint=True, split=split)
            if bins is not None:
                flags = (NUFFT_IMAGE_SPLIT | NUFFT_KDATA_SPLIT) if split else 0
                for b0, b1 in self._chunks(L, h, B, coils, M):
                    nb = b1 - b0
                    ws = torch.empty(L.pdu_nufft_binned_workspace_bytes(h, nb * coils, M), dtype=torch.uint8, device=data.device)
                    sm = None if smaps is None else (smaps if sb == 1 else smaps[b0:b1])
                    check(L.pdu_nufft_adj_binned_c64(h, data[b0:b1].data_ptr(), out[b0:b1].data_ptr(),
                                                     sm.data_ptr() if sm is not None else None,
                                                     kweight.data_ptr() if kweight is not None else None, nb, coils,
                                                     1 if sb == 1 else nb, M, self.scale(norm), bins.data_ptr(), flags,
                                                     ws.data_ptr(), ws.numel(), stream_ptr()), "pdu_nufft_adj_binned_c64")
                return out
            if split:            # generic path: the library's layout kernels around the complex64 entry points
                return self._complex_to_split(self.adjoint(self._split_to_complex(data, kweight), omega, smaps, norm))
            if kweight is not None:
                return self.adjoint(data * kweight, omega, smaps, norm)
            ws = torch.empty(L.pdu_nufft_workspace_bytes(h, B * coils), dtype=torch.uint8, device=data.device)
            csr = self._csr_for(omega, B * coils)
            check(L.pdu_nufft_adj_csr_c64(h, data.data_ptr(), out.data_ptr(), omega.data_ptr(),
                                          smaps.data_ptr() if smaps is not None else None, B, coils, sb, M,
                                          self.scale(norm), csr.data_ptr() if csr is not None else None, ws.data_ptr(),
                                          ws.numel(), stream_ptr()), "pdu_nufft_adj_c64")
        return out

    def interp(self, grid, omega):
        grid = require_cuda(grid, torch.complex64, "grid")
        if grid.dim() != 4 or tuple(grid.shape[-2:]) != self.grid_size:
            raise ValueError(f"grid must be [B, C, {self.grid_size[0]}, {self.grid_size[1]}]")
        omega = self._omega(omega)
        B, Cc, M = grid.shape[0], grid.shape[1], omega.shape[1]
        out = torch.empty((B, Cc, M), dtype=torch.complex64, device=grid.device)
        if B * Cc == 0 or M == 0:
            return out
        with torch.cuda.device(grid.device):
            check(lib().pdu_nufft_interp_fwd_c64(self.handle(grid.device), grid.data_ptr(), out.data_ptr(),
                                                 omega.data_ptr(), B * Cc, M, stream_ptr()), "pdu_nufft_interp_fwd_c64")
        return out

    def interp_adjoint(self, data, omega):
        data = require_cuda(data, torch.complex64, "data")
        omega = self._omega(omega)
        if data.dim() != 3 or data.shape[-1] != omega.shape[1]:
            raise ValueError(f"data must be [B, C, {omega.shape[1]}]")
        B, Cc, M = data.shape
        if B * Cc == 0 or M == 0:
            return torch.zeros((B, Cc) + self.grid_size, dtype=torch.complex64, device=data.device)
        with torch.cuda.device(data.device):
            csr = self._csr_for(omega, B * Cc)
            if csr is not None:      # the gather writes every cell
                out = torch.empty((B, Cc) + self.grid_size, dtype=torch.complex64, device=data.device)
                check(lib().pdu_nufft_interp_adj_csr_c64(self.handle(data.device), data.data_ptr(), out.data_ptr(),
                                                         csr.data_ptr(), B * Cc, M, stream_ptr()),
                      "pdu_nufft_interp_adj_csr_c64")
            else:                    # the scatter accumulates
                out = torch.zeros((B, Cc) + self.grid_size, dtype=torch.complex64, device=data.device)
                check(lib().pdu_nufft_interp_adj_c64(self.handle(data.device), data.data_ptr(), out.data_ptr(),
                                                     omega.data_ptr(), B * Cc, M, stream_ptr()), "pdu_nufft_interp_adj_c64")
        return out


def _per_trajectory(fn, x, omega, smaps, *rest):
    """torchkbnufft lets omega carry a batch axis ([B, 2, M]); each trajectory is its own call."""
    if omega.dim() == 2:
        return fn(x, omega, smaps, *rest)
    if omega.dim() != 3 or omega.shape[0] != x.shape[0]:
        raise ValueError("batched omega must be [B, 2, M] with the batch of the data")
    outs = []
    for b in range(x.shape[0]):
        sm = None if smaps is None else (smaps if smaps.shape[0] == 1 else smaps[b:b + 1])
        outs.append(fn(x[b:b + 1], omega[b], sm, *rest))
    return torch.cat(outs, dim=0)


class _NufftFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, omega, smaps, plan, norm, adjoint, split=False, kweight=None):
        ctx.plan, ctx.norm, ctx.adjoint, ctx.split, ctx.kweight = plan, norm, adjoint, split, kweight
        ctx.save_for_backward(omega, smaps) if smaps is not None else ctx.save_for_backward(omega)
        if adjoint:
            return _per_trajectory(plan.adjoint, x, omega, smaps, norm, split, kweight)
        return _per_trajectory(plan.forward, x, omega, smaps, norm, split)

    @staticmethod
    def backward(ctx, grad):
        saved = ctx.saved_tensors
        omega, smaps = saved[0], (saved[1] if len(saved) > 1 else None)
        if ctx.adjoint:       # x = A^H (w y)  =>  dy = w A dx
            g = _per_trajectory(ctx.plan.forward, grad.contiguous(), omega, smaps, ctx.norm, ctx.split)
            if ctx.kweight is not None:
                g = g * ctx.kweight.reshape(-1)
        else:
            g = _per_trajectory(ctx.plan.adjoint, grad.contiguous(), omega, smaps, ctx.norm, ctx.split)
        return g, None, None, None, None, None, None, None


class _InterpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, omega, plan, adjoint):
        ctx.plan, ctx.adjoint = plan, adjoint
        ctx.save_for_backward(omega)
        fn = plan.interp_adjoint if adjoint else plan.interp
        return _per_trajectory(lambda a, o, s: fn(a, o), x, omega, None)

    @staticmethod
    def backward(ctx, grad):
        (omega,) = ctx.saved_tensors
        fn = ctx.plan.interp if ctx.adjoint else ctx.plan.interp_adjoint
        return _per_trajectory(lambda a, o, s: fn(a, o), grad.contiguous(), omega, None), None, None, None


class _KbModule(nn.Module):
    def __init__(self, im_size: Sequence[int], grid_size: Optional[Sequence[int]] = None, numpoints=6,
                 n_shift: Optional[Sequence[int]] = None, table_oversamp=2 ** 10, kbwidth: float = 2.34,
                 order=0.0, dtype=None, device=None):
        super().__init__()
        order0 = float(order if np.isscalar(order) else order[0])
        self._plan = _Plan(im_size, grid_size, numpoints, n_shift, table_oversamp, kbwidth, order0)
        self.im_size, self.grid_size, self.n_shift = self._plan.im_size, self._plan.grid_size, self._plan.n_shift
        self.numpoints, self.table_oversamp = self._plan.numpoints, self._plan.table_oversamp

    @staticmethod
    def _no_interp_mats(interp_mats):
        if interp_mats is not None:
            raise NotImplementedError("sparse-matrix interpolation is not implemented; pass interp_mats=None "
                                      "(table interpolation, torchkbnufft's default)")


class KbNufft(_KbModule):
    """image [B, C, N0, N1] -> k-space samples [B, C (or coils), M].  [RECALL] torchkbnufft.KbNufft."""

    def forward(self, image, omega, interp_mats=None, smaps=None, norm: Optional[str] = None, split: bool = False):
        """split=True (an extension): image float32 [B, 2 C, N0, N1] with (re, im) as neighbouring channels ->
        float32 [B, 2 coils, M]; no complex tensors, no layout passes around the operator."""
        self._no_interp_mats(interp_mats)
        return _NufftFn.apply(image, omega, smaps, self._plan, norm, False, split)


class KbNufftAdjoint(_KbModule):
    """samples [B, C, M] -> image [B, C (or 1), N0, N1].  [RECALL] torchkbnufft.KbNufftAdjoint."""

    def forward(self, data, omega, interp_mats=None, smaps=None, norm: Optional[str] = None, split: bool = False,
                kweight: Optional[torch.Tensor] = None):
        """Extensions: split=True as in KbNufft; kweight float32 [M] is multiplied into the samples on load (the
        density compensation that otherwise is a separate pass over the data)."""
        self._no_interp_mats(interp_mats)
        return _NufftFn.apply(data, omega, smaps, self._plan, norm, True, split, kweight)


class KbInterp(_KbModule):
    """oversampled Cartesian k-space [B, C, K0, K1] -> samples [B, C, M].  [RECALL] torchkbnufft.KbInterp."""

    def forward(self, image, omega, interp_mats=None):
        self._no_interp_mats(interp_mats)
        return _InterpFn.apply(image, omega, self._plan, False)


class KbInterpAdjoint(_KbModule):
    """samples [B, C, M] -> gridded k-space [B, C, K0, K1].  [RECALL] torchkbnufft.KbInterpAdjoint."""

    def forward(self, data, omega, interp_mats=None):
        self._no_interp_mats(interp_mats)
        return _InterpFn.apply(data, omega, self._plan, True)


def calc_density_compensation_function(ktraj: torch.Tensor, im_size: Sequence[int], num_iterations: int = 10,
                                       grid_size: Optional[Sequence[int]] = None, numpoints=6,
                                       n_shift: Optional[Sequence[int]] = None, table_oversamp=2 ** 10,
                                       kbwidth: float = 2.34, order=0.0) -> torch.Tensor:
    """Pipe & Menon's iteration w <- w / |G G^H w| with the table interpolator only.
    [RECALL] torchkbnufft.calc_density_compensation_function; returns complex64 [1, 1, M]."""
    plan = _Plan(im_size, grid_size, numpoints, n_shift, table_oversamp, kbwidth, order)
    omega = require_cuda(ktraj, torch.float32, "ktraj")
    w = torch.ones((1, 1, omega.shape[1]), dtype=torch.complex64, device=omega.device)
    for _ in range(num_iterations):
        new = plan.interp(plan.interp_adjoint(w, omega), omega)
        w = w / new.abs()
    return w
