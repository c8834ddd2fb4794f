"""Mirror of the hot-path part of utils/utils.py: ``logmeanexp`` (:9-11) and the distribution containers (:48-96)."""
from __future__ import annotations

from collections import namedtuple
from typing import Any, NamedTuple, Optional

import torch

from . import functional as F

__all__ = ["logmeanexp", "Dist", "DistributionTuple", "normalize", "fill_canvas", "sample_grid", "write_ppm"]


def logmeanexp(log_w: torch.Tensor, axis: int) -> torch.Tensor:
    """``log(mean(exp(log_w - max), axis)) + max`` (utils/utils.py:9-11) on the sm_100a kernel; differentiable
    (gradient = softmax over ``axis``, as the reference's un-stopped max gives)."""
    return F.logmeanexp(log_w, axis)


class Dist(namedtuple("Dist", "dist sample axes")):
    """utils/utils.py:48-71."""

    @property
    def z(self):
        return self.sample

    @property
    def x(self):
        return self.sample

    @property
    def p(self):
        return self.dist

    @property
    def q(self):
        return self.dist


class DistributionTuple(NamedTuple):
    """utils/utils.py:83-96."""

    dist: Any
    sample: Optional[torch.Tensor] = None
    axes: tuple = (-1, -2, -3)

    @property
    def z(self):
        return self.sample

    @property
    def x(self):
        return self.sample


# ---- the byte side of the path: uint8 images in, quantised sample grids out --------------------------------------------
def normalize(img: torch.Tensor) -> torch.Tensor:
    """``uint8 -> float32 / 255`` (utils/data.py:15-16).  The kernels accept the uint8 tensor directly and form exactly
    this value on the fly, so calling it is only needed for code that wants the float image itself."""
    return img.to(torch.float32) / 255.0


def fill_canvas(img: torch.Tensor, n: int, h: int, w: int, c: int) -> torch.Tensor:
    """``[>= n*n, h, w, c] -> [n*h, n*w, c]`` image grid, row-major (utils/utils.py:74-80), any dtype, on the device:
    a view + one copy instead of the reference's n*n slice assignments."""
    if img.shape[0] < n * n or tuple(img.shape[1:]) != (h, w, c):
        raise ValueError(f"need at least {n * n} images of shape {(h, w, c)}, got {tuple(img.shape)}")
    return img[: n * n].reshape(n, n, h, w, c).permute(0, 2, 1, 3, 4).reshape(n * h, n * w, c)


def sample_grid(pxz, n: int = 8, **sample_kwargs) -> torch.Tensor:
    """The sample canvas of ``_plot_samples`` (models/model05.py:200-216) as bytes: draws one sample per image from
    ``pxz`` (parameters ``[>= n*n, h, w, 10*n_mix]``), quantised in the sampling kernel (``rint(255 * clip(x, 0, 1))``)
    -> uint8 ``[n*h, n*w, 3]``."""
    out = pxz.sample(return_quantised=True, **sample_kwargs)
    xq = out[1]
    h, w, c = xq.shape[-3:]
    return fill_canvas(xq.reshape(-1, h, w, c), n, h, w, c)


def write_ppm(path: str, canvas_u8: torch.Tensor) -> None:
    """Binary PPM (P6) of a uint8 ``[H, W, 3]`` canvas -- a dependency-free stand-in for the reference's
    ``tf.summary.image`` / PNG output (models/model05.py:186-190)."""
    if canvas_u8.dtype != torch.uint8 or canvas_u8.dim() != 3 or canvas_u8.shape[2] != 3:
        raise ValueError("canvas must be uint8 [H, W, 3]")
    data = canvas_u8.contiguous().cpu().numpy().tobytes()
    with open(path, "wb") as f:
        f.write(f"P6\n{canvas_u8.shape[1]} {canvas_u8.shape[0]}\n255\n".encode())
        f.write(data)
