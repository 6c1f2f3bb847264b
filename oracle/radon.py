"""Float64 CPU restatement of the CT operators (TEST INFRASTRUCTURE, parity unpinned).

What is restated, and from where.  The reference repository names torch_radon
as its CT operator library (/root/reference/README.md:3-5 only points at
branches; BASELINE.json north_star names the library).  torch_radon is not
mounted, so the conventions below are [RECALL] of torch_radon v1.0
(`Radon`, `RadonFanbeam`, `filter_sinogram`), each kept as an explicit
parameter:

* image f[y, x], N x N; pixel (y, x) has its centre at (x + .5, y + .5) in
  texture coordinates and at (x + .5 - N/2, y + .5 - N/2) in world units;
  samples are bilinear with a zero border (CUDA texture `linear` + `border`);
* the user-facing angle theta is negated once by the host wrapper
  ([RECALL] `self.angles = -angles`); everything here takes the *internal*
  angle and only ever sees its cosine and sine (`trig_table`);
* forward projection is ray driven: a ray is clipped against the image square
  (or the inscribed circle), cut into n = ceil(length) equal steps of length
  <= 1, sampled at both end points and every step between, and the plain sum
  is scaled by the step length;
* backprojection is pixel driven with linear interpolation along the detector
  and a final 1/det_spacing; the fan-beam flavour weights each view by the
  magnification (s+d)/(s - t);
* filter_sinogram is the scikit-image ramp construction in a zero-padded
  power-of-two FFT, scaled by pi / (2 * n_angles).

Discrete decisions (clip interval, step count) are taken in IEEE float32 by
`ray_setup_f32`, op for op what pd_unet_b200/csrc/radon_common.cuh does with
__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn, so oracle and kernel sample the
same points.  Everything else is float64.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

PARALLEL = 0
FAN = 1


@dataclass(frozen=True)
class RadonGeom:
    """Scan geometry.  Mirrors pdu_radon_geom_t in include/pdu.h."""
    n: int                      # image is n x n
    n_angles: int
    det_count: int
    det_spacing: float = 1.0
    geom: int = PARALLEL
    s_dist: float = 0.0         # source -> rotation centre (fan only)
    d_dist: float = 0.0         # rotation centre -> detector (fan only)
    clip_to_circle: bool = False


def trig_table(angles) -> np.ndarray:
    """[A, 2] float32 (cos, sin) of the INTERNAL angles, rounded once from float64."""
    a = np.asarray(angles, dtype=np.float64)
    return np.stack([np.cos(a), np.sin(a)], axis=-1).astype(np.float32)


# ----------------------------------------------------------------------------
# ray setup, float32 op-for-op (mirror of radon_common.cuh :: ray_setup)
# ----------------------------------------------------------------------------
def ray_setup_f32(g: RadonGeom, trig: np.ndarray):
    """Returns dict of [A, D] arrays: xc0, yc0, vx, vy, step (float32), n_steps (int32).

    (xc0, yc0) is the first sample in pixel-centre coordinates (texture
    coordinate minus one half), so the bilinear cell is floor(xc), floor(yc).
    n_steps == -1 marks a ray that misses the volume.
    """
    f = np.float32
    A, D = g.n_angles, g.det_count
    cs = trig[:, 0].astype(f)[:, None]
    sn = trig[:, 1].astype(f)[:, None]
    d = np.arange(D, dtype=f)[None, :]
    v = f(g.n) * f(0.5)
    u = ((d - f(D) * f(0.5)) + f(0.5)) * f(g.det_spacing)
    if g.geom == PARALLEL:
        sx, sy = u, f(g.n)
        ex, ey = u, -f(g.n)
    else:
        sx, sy = np.zeros_like(u), f(g.s_dist)
        ex, ey = u, -f(g.d_dist)
    sx = np.broadcast_to(sx, (1, D)).astype(f)
    ex = np.broadcast_to(ex, (1, D)).astype(f)
    with np.errstate(all="ignore"):
        rsx = sx * cs - f(sy) * sn
        rsy = sx * sn + f(sy) * cs
        rex = ex * cs - f(ey) * sn
        rey = ex * sn + f(ey) * cs
        dx = rex - rsx
        dy = rey - rsy
        eps = f(1e-6)
        dx = np.where(dx >= 0, np.maximum(dx, eps), np.minimum(dx, -eps)).astype(f)
        dy = np.where(dy >= 0, np.maximum(dy, eps), np.minimum(dy, -eps)).astype(f)
        if not g.clip_to_circle:
            ax0 = (-v - rsx) / dx
            ax1 = (v - rsx) / dx
            ay0 = (-v - rsy) / dy
            ay1 = (v - rsy) / dy
            a_s = np.maximum(np.minimum(ax0, ax1), np.minimum(ay0, ay1))
            a_e = np.minimum(np.maximum(ax0, ax1), np.maximum(ay0, ay1))
            hit = np.ones_like(a_s, dtype=bool)
        else:
            a = dx * dx + dy * dy
            b = rsx * dx + rsy * dy
            c = (rsx * rsx + rsy * rsy) - v * v
            delta = b * b - a * c
            hit = delta > 0
            sq = np.sqrt(np.where(hit, delta, f(0))).astype(f)
            a_s = (-b - sq) / a
            a_e = (-b + sq) / a
        a_s = np.maximum(a_s, f(0)).astype(f)
        a_e = np.minimum(a_e, f(1)).astype(f)
        hit &= a_s < a_e
        x0 = (rsx + dx * a_s) + v
        y0 = (rsy + dy * a_s) + v
        x1 = (rsx + dx * a_e) + v
        y1 = (rsy + dy * a_e) + v
        lx = x1 - x0
        ly = y1 - y0
        length = np.sqrt(lx * lx + ly * ly).astype(f)
        n_steps = np.ceil(length).astype(np.int32)
        hit &= n_steps > 0
        nf = np.where(hit, n_steps, 1).astype(f)
        vx = (lx / nf).astype(f)
        vy = (ly / nf).astype(f)
        step = np.sqrt(vx * vx + vy * vy).astype(f)
        xc0 = (x0 - f(0.5)).astype(f)
        yc0 = (y0 - f(0.5)).astype(f)
    n_steps = np.where(hit, n_steps, -1).astype(np.int32)
    z = f(0)
    return dict(xc0=np.where(hit, xc0, z), yc0=np.where(hit, yc0, z),
                vx=np.where(hit, vx, z), vy=np.where(hit, vy, z),
                step=np.where(hit, step, z), n_steps=n_steps)


def _texq(f: torch.Tensor) -> torch.Tensor:
    """An interpolation fraction as the texture unit holds it: 9-bit fixed point with 8 fractional bits (CUDA C
    programming guide, linear filtering) -- round to nearest (even) multiple of 2^-8."""
    return torch.round(f * 256.0) / 256.0


def _bilinear_zero(img_flat: torch.Tensor, n: int, xc: torch.Tensor, yc: torch.Tensor, tex_weights: bool = False):
    """img_flat [B, n*n] float64; xc, yc [...] float64 pixel-centre coords -> [B, ...]."""
    ix = torch.floor(xc)
    iy = torch.floor(yc)
    fx = xc - ix
    fy = yc - iy
    if tex_weights:
        fx, fy = _texq(fx), _texq(fy)
    ix = ix.long()
    iy = iy.long()
    out = None
    for dy_, wy in ((0, 1.0 - fy), (1, fy)):
        for dx_, wx in ((0, 1.0 - fx), (1, fx)):
            xx = ix + dx_
            yy = iy + dy_
            ok = (xx >= 0) & (xx < n) & (yy >= 0) & (yy < n)
            idx = (yy.clamp(0, n - 1) * n + xx.clamp(0, n - 1)).reshape(-1)
            val = img_flat[:, idx].reshape((img_flat.shape[0],) + tuple(xc.shape))
            term = val * (wy * wx * ok)
            out = term if out is None else out + term
    return out


def radon_forward(img, trig: np.ndarray, g: RadonGeom, angle_chunk: int = 8, tex_weights: bool = False) -> torch.Tensor:
    """img [B, n, n] -> sinogram [B, A, D] float64.  Restates [RECALL] torch_radon
    radon_forward_kernel (ray driven, clipped, unit-ish step, texture bilinear).
    tex_weights: emulate the texture unit's 8-bit interpolation weights (the twin of the library's "tex_weights"
    option).  The sample coordinates are then rounded to float32 like the kernel's single FMA, so that both sides
    quantise the same fractions."""
    img = torch.as_tensor(img, dtype=torch.float64)
    B = img.shape[0]
    n, A, D = g.n, g.n_angles, g.det_count
    assert img.shape[1:] == (n, n)
    rs = ray_setup_f32(g, trig)
    flat = img.reshape(B, n * n)
    out = torch.zeros(B, A, D, dtype=torch.float64)
    for a0 in range(0, A, angle_chunk):
        a1 = min(A, a0 + angle_chunk)
        ns = torch.from_numpy(rs["n_steps"][a0:a1])
        smax = int(ns.max())
        if smax < 0:
            continue
        j = torch.arange(smax + 1, dtype=torch.float64)[None, None, :]
        x0 = torch.from_numpy(rs["xc0"][a0:a1]).double()[..., None]
        y0 = torch.from_numpy(rs["yc0"][a0:a1]).double()[..., None]
        vx = torch.from_numpy(rs["vx"][a0:a1]).double()[..., None]
        vy = torch.from_numpy(rs["vy"][a0:a1]).double()[..., None]
        live = (j <= ns[..., None].double())
        xs, ys = x0 + j * vx, y0 + j * vy
        if tex_weights:
            xs, ys = xs.float().double(), ys.float().double()
        vals = _bilinear_zero(flat, n, xs, ys, tex_weights)
        s = (vals * live).sum(-1)
        out[:, a0:a1] = s * torch.from_numpy(rs["step"][a0:a1]).double()
    return out


def radon_backprojection(sino, trig: np.ndarray, g: RadonGeom, angle_chunk: int = 16, fbp_weight: bool = False,
                         tex_weights: bool = False) -> torch.Tensor:
    """sinogram [B, A, D] -> image [B, n, n] float64.  Restates [RECALL] torch_radon
    radon_backward_kernel (pixel driven, linear along the detector, zero border).
    fbp_weight (fan beam): every tap once more times s / (s - t) -- with the magnification weight this is the
    1 / U^2 of fan-beam FBP."""
    sino = torch.as_tensor(sino, dtype=torch.float64)
    B = sino.shape[0]
    n, A, D = g.n, g.n_angles, g.det_count
    assert sino.shape[1:] == (A, D)
    ids = 1.0 / float(np.float32(g.det_spacing))
    c = torch.arange(n, dtype=torch.float64) - n / 2.0 + 0.5
    dx = c[None, :].expand(n, n)          # x = column
    dy = c[:, None].expand(n, n)          # y = row
    cr = D / 2.0
    acc = torch.zeros(B, n, n, dtype=torch.float64)
    cs_all = torch.from_numpy(trig[:, 0].astype(np.float64))
    sn_all = torch.from_numpy(trig[:, 1].astype(np.float64))
    for a0 in range(0, A, angle_chunk):
        a1 = min(A, a0 + angle_chunk)
        cs = cs_all[a0:a1, None, None]
        sn = sn_all[a0:a1, None, None]
        if g.geom == PARALLEL:
            jc = (cs * dx + sn * dy) * ids + cr
            w = torch.ones_like(jc)
        else:
            k = float(np.float32(g.s_dist)) + float(np.float32(g.d_dist))
            den = float(np.float32(g.s_dist)) + sn * dx - cs * dy
            iden = k / den
            jc = (cs * dx + sn * dy) * ids * iden + cr
            w = iden * (float(np.float32(g.s_dist)) / den) if fbp_weight else iden
        jb = jc - 0.5
        i0 = torch.floor(jb)
        fr = jb - i0
        if tex_weights:
            fr = _texq(fr)
        i0 = i0.long()
        rows = sino[:, a0:a1, :]                                   # [B, a, D]
        tot = torch.zeros(B, a1 - a0, n, n, dtype=torch.float64)
        for off, wt in ((0, 1.0 - fr), (1, fr)):
            ii = i0 + off
            ok = (ii >= 0) & (ii < D)
            idx = ii.clamp(0, D - 1).reshape(a1 - a0, n * n)
            val = torch.gather(rows, 2, idx[None].expand(B, -1, -1)).reshape(B, a1 - a0, n, n)
            tot = tot + val * (wt * ok)
        acc += (tot * w).sum(1)
    if g.clip_to_circle:
        acc = acc * ((dx * dx + dy * dy) <= (n / 2.0) ** 2)
    return acc * ids


# ----------------------------------------------------------------------------
# sinogram filtering ([RECALL] torch_radon filter_sinogram + scikit-image filters)
# ----------------------------------------------------------------------------
# Written independently of pd_unet_b200/radon.py on purpose: the product builds the response
# the scikit-image way (np.fft of a constant table, np.hamming / np.hanning, np.fft.fftshift); here every quantity
# comes from its closed form and plain cosine sums, so a test that compares the two compares two constructions.
def _ramp_kernel(size: int) -> np.ndarray:
    """Kak & Slaney eq. 61: the band-limited ramp sampled at the integers, laid out circularly on `size` points:
    f[0] = 1/4, f[k] = -1 / (pi d)^2 for odd distance d = min(k, size - k), 0 for even d."""
    k = np.arange(size)
    d = np.minimum(k, size - k).astype(np.float64)
    f = np.zeros(size)
    odd = (d % 2) == 1
    f[odd] = -1.0 / (np.pi * d[odd]) ** 2
    f[0] = 0.25
    return f


def _cos_sum(values: np.ndarray, chunk: int = 256) -> np.ndarray:
    """out[i] = sum_k values[k] cos(2 pi i k / P): a real DFT of an even sequence without an FFT.  The angle is
    reduced with integer arithmetic ((i k) mod P) before the cosine, so the sum is accurate to ~1e-16 P."""
    P = values.shape[0]
    k = np.arange(P, dtype=np.int64)
    out = np.empty(P)
    for i0 in range(0, P, chunk):
        i = np.arange(i0, min(P, i0 + chunk), dtype=np.int64)
        out[i0:i0 + i.size] = np.cos((2.0 * np.pi / P) * ((i[:, None] * k[None, :]) % P)) @ values
    return out


_RESP_CACHE = {}


def _fourier_filter(size: int, name: str) -> np.ndarray:
    """Frequency response on `size` (even) points: 2 DFT(ramp kernel), times the named window.  Windows are written
    out as functions of the centred frequency index c = (i + size/2) mod size (what np.fft.fftshift of a symmetric
    window amounts to)."""
    name = name.lower()
    key = (size, name)
    if key in _RESP_CACHE:
        return _RESP_CACHE[key]
    ff = 2.0 * _cos_sum(_ramp_kernel(size))
    i = np.arange(size)
    c = (i + size // 2) % size
    if name in ("ramp", "ram-lak"):
        pass
    elif name == "shepp-logan":
        freq = np.where(i < size - size // 2, i, i - size) / size          # cycles per sample, as np.fft.fftfreq
        om = np.pi * freq[1:]
        ff[1:] = ff[1:] * np.sin(om) / om
    elif name == "cosine":
        ff = ff * np.sin(np.pi * c / size)
    elif name == "hamming":
        ff = ff * (0.54 - 0.46 * np.cos(2.0 * np.pi * c / (size - 1)))
    elif name == "hann":
        ff = ff * (0.5 - 0.5 * np.cos(2.0 * np.pi * c / (size - 1)))
    else:
        raise ValueError(f"unknown filter {name!r}")
    _RESP_CACHE[key] = ff
    return ff


def padded_size(det_count: int) -> int:
    return max(64, int(2 ** math.ceil(math.log2(2 * det_count))))


_TAPS_CACHE = {}


def filter_taps(det_count: int, name: str = "ramp") -> np.ndarray:
    key = (det_count, name.lower())
    if key not in _TAPS_CACHE:
        _TAPS_CACHE[key] = _filter_taps(det_count, name)
    return _TAPS_CACHE[key].copy()


def _filter_taps(det_count: int, name: str = "ramp") -> np.ndarray:
    """Spatial taps h[-(D-1) .. D-1] (length 2D-1, float64) such that the padded
    circular FFT filter equals the linear convolution q[i] = sum_j p[j] h[i-j].
    ramp: straight from the closed form (h = 2 x ramp kernel); windowed: inverse cosine sum of the response."""
    P = padded_size(det_count)
    if name.lower() in ("ramp", "ram-lak"):
        h = 2.0 * _ramp_kernel(P)
    else:
        h = _cos_sum(_fourier_filter(P, name)) / P                          # response is real and even
    k = np.arange(-(det_count - 1), det_count)
    return h[k % P]


def filter_matrix(det_count: int, n_angles: int, name: str = "ramp") -> np.ndarray:
    """[D, D] float64 Toeplitz H with out = sino @ H, including pi/(2A)."""
    t = filter_taps(det_count, name)
    j = np.arange(det_count)[:, None]
    i = np.arange(det_count)[None, :]
    return t[(i - j) + det_count - 1] * (np.pi / (2.0 * n_angles))


def filter_sinogram_fft(sino, name: str = "ramp") -> torch.Tensor:
    """[..., A, D] float64, the FFT route as [RECALL] torch_radon does it: zero-pad to a power of two >= 2 D,
    multiply the spectrum by the response, inverse, crop, scale by pi / (2 A)."""
    sino = torch.as_tensor(sino, dtype=torch.float64)
    D = sino.shape[-1]
    A = sino.shape[-2]
    P = padded_size(D)
    ff = torch.from_numpy(_fourier_filter(P, name))
    pad = torch.nn.functional.pad(sino, (0, P - D))
    q = torch.fft.ifft(torch.fft.fft(pad, dim=-1) * ff, dim=-1).real
    return q[..., :D] * (math.pi / (2.0 * A))


def filter_sinogram(sino, name: str = "ramp") -> torch.Tensor:
    """[..., A, D] float64: the same filter as the linear convolution it is, out = sino @ H (no FFT anywhere on this
    route; tests/test_oracle_radon.py checks it against filter_sinogram_fft)."""
    sino = torch.as_tensor(sino, dtype=torch.float64)
    D, A = sino.shape[-1], sino.shape[-2]
    return sino @ torch.from_numpy(filter_matrix(D, A, name))


def fan_cosine_weights(g: RadonGeom) -> np.ndarray:
    """float64 [D]: the pre-weight of fan-beam FBP for a flat equispaced detector (Kak & Slaney section 3.4.2): with the
    detector scaled back to the rotation centre, s' = u s / (s + d), the projection is multiplied by
    s / sqrt(s^2 + s'^2) = (s + d) / sqrt((s + d)^2 + u^2), the cosine of the ray's fan angle."""
    u = (np.arange(g.det_count, dtype=np.float64) + 0.5 - g.det_count / 2.0) * g.det_spacing
    k = float(g.s_dist) + float(g.d_dist)
    return k / np.sqrt(k * k + u * u)


def fbp(sino, trig, g: RadonGeom, name: str = "ramp", fan_weights: bool = True) -> torch.Tensor:
    """Parallel beam: backprojection(filter(s)).  Fan beam with fan_weights (Kak & Slaney section 3.4.2, flat equispaced
    detector, views over 2 pi): cosine pre-weight, ramp filter, backprojection weighted by 1 / U^2."""
    if g.geom == FAN and fan_weights:
        s = torch.as_tensor(sino, dtype=torch.float64) * torch.from_numpy(fan_cosine_weights(g))
        return radon_backprojection(filter_sinogram(s, name), trig, g, fbp_weight=True)
    return radon_backprojection(filter_sinogram(sino, name), trig, g)
