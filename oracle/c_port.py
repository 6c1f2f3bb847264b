"""ctypes front end of oracle/radon_c.c (TEST INFRASTRUCTURE, parity unpinned -- see __init__.py).

build() compiles the C restatement with gcc into oracle/_build/ (git-ignored); the functions take
and return float64 torch tensors with the signatures of oracle.radon.  Used by the full-size parity
tests and by bench.py's cpu_baseline / --impl reference legs, never by pd_unet_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import torch

from .radon import RadonGeom, filter_taps

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "radon_c.c")
_OUT = os.path.join(_HERE, "_build", "libpdu_oracle.so")
_lib = None


class _Geom(C.Structure):
    _fields_ = [("geom", C.c_int32), ("n", C.c_int32), ("n_angles", C.c_int32), ("det_count", C.c_int32),
                ("det_spacing", C.c_float), ("s_dist", C.c_float), ("d_dist", C.c_float), ("clip", C.c_int32)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", _OUT, _SRC, "-lm"],
                       check=True)
    return _OUT


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def n_threads() -> int:
    return int(_load().pduo_get_threads())


def set_threads(n: int) -> None:
    """Override OMP_NUM_THREADS (torchrun pins it to 1 for its workers)."""
    _load().pduo_set_threads(C.c_int(int(n)))


def _g(g: RadonGeom) -> _Geom:
    return _Geom(g.geom, g.n, g.n_angles, g.det_count, g.det_spacing, g.s_dist, g.d_dist, int(g.clip_to_circle))


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def ray_setup(g: RadonGeom, trig: np.ndarray) -> dict:
    A, D = g.n_angles, g.det_count
    out = {k: np.empty((A, D), np.float32) for k in ("xc0", "yc0", "vx", "vy", "step")}
    out["n_steps"] = np.empty((A, D), np.int32)
    t = np.ascontiguousarray(trig, np.float32)
    gg = _g(g)
    _load().pduo_ray_setup(C.byref(gg), _p(t), *(_p(out[k]) for k in ("xc0", "yc0", "vx", "vy", "step", "n_steps")))
    return out


def radon_forward(img, trig: np.ndarray, g: RadonGeom) -> torch.Tensor:
    x = np.ascontiguousarray(torch.as_tensor(img, dtype=torch.float64).numpy())
    B = x.shape[0]
    out = np.empty((B, g.n_angles, g.det_count), np.float64)
    t = np.ascontiguousarray(trig, np.float32)
    gg = _g(g)
    _load().pduo_radon_forward(_p(x), _p(out), _p(t), C.c_int(B), C.byref(gg))
    return torch.from_numpy(out)


def radon_backprojection(sino, trig: np.ndarray, g: RadonGeom) -> torch.Tensor:
    s = np.ascontiguousarray(torch.as_tensor(sino, dtype=torch.float64).numpy())
    B = s.shape[0]
    out = np.empty((B, g.n, g.n), np.float64)
    t = np.ascontiguousarray(trig, np.float32)
    gg = _g(g)
    _load().pduo_radon_backproj(_p(s), _p(out), _p(t), C.c_int(B), C.byref(gg))
    return torch.from_numpy(out)


def filter_sinogram(sino, name: str = "ramp") -> torch.Tensor:
    s = np.ascontiguousarray(torch.as_tensor(sino, dtype=torch.float64).numpy())
    D, A = s.shape[-1], s.shape[-2]
    taps = np.ascontiguousarray(filter_taps(D, name) * (np.pi / (2.0 * A)))
    out = np.empty_like(s)
    _load().pduo_filter(_p(s), _p(out), _p(taps), C.c_long(s.size // D), C.c_int(D))
    return torch.from_numpy(out)


def fbp(sino, trig, g: RadonGeom, name: str = "ramp") -> torch.Tensor:
    return radon_backprojection(filter_sinogram(sino, name), trig, g)
