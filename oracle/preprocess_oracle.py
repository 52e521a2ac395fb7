"""CPU oracle of the input transform (TEST INFRASTRUCTURE ONLY - imported by tests/ and smoke checks, never by msclip_b200/).

Restates, in plain Python / numpy integer arithmetic, what tools/zero_shot.py:202-207 of the reference computes for one
decoded RGB image:  transforms.Resize(S, BICUBIC) -> CenterCrop(S) -> ToTensor() -> Normalize(mean, std).

The arithmetic lives in third-party code that is NOT under /root/reference: torchvision (0.26 in this image; reference pins
0.7.0, INSTALL.md:20-23) forwards PIL images to Pillow (12.2 here) `Image.resize`, i.e. ImagingResample in Pillow's
src/libImaging/Resample.c.  Published algorithm restated below: per axis a window of half-width 2 * max(scale, 1) around
(xx + 0.5) * scale, bicubic weights (a = -0.5) normalised in double, rounded half away from zero to 22-bit fixed point;
horizontal pass into uint8 (int32 accumulation from 1 << 21, >> 22, clip to 0..255), then the vertical pass on the rounded
result.  Pinned in tests/test_preprocess.py against the live torchvision + Pillow of this image on random images of many
sizes (bit-exact), and against the committed fixture tests/golden/preprocess_cases.npz (generated with the same stack by
`python oracle/preprocess_oracle.py`).
"""
from __future__ import annotations

import math
import os

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc: ([out][ksize] int coefficients, [out] xmin, [out] count)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    xmins = np.zeros(out_size, dtype=np.int64)
    counts = np.zeros(out_size, dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            v = v * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if v < 0 else int(0.5 + v)
        xmins[xx], counts[xx] = xmin, xmax
    return kk, xmins, counts


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """Resample along axis 0 of a uint8 array [n, ...] with Pillow's fixed-point arithmetic."""
    kk, xmins, counts = precompute_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        n, x0 = int(counts[xx]), int(xmins[xx])
        acc = np.tensordot(kk[xx, :n], src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resized_size(h: int, w: int, size: int):
    """torchvision _compute_resized_output_size for an int size: the shorter edge becomes `size`."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def resize_center_crop_u8(img: np.ndarray, size: int) -> np.ndarray:
    """uint8 [H, W, 3] -> uint8 [size, size, 3]: horizontal pass first, then vertical (Pillow's order), then the crop."""
    h, w = img.shape[:2]
    new_h, new_w = resized_size(h, w, size)
    tmp = _resample_axis0(np.ascontiguousarray(img.transpose(1, 0, 2)), new_w).transpose(1, 0, 2)   # [H, new_w, 3]
    res = _resample_axis0(np.ascontiguousarray(tmp), new_h)                                          # [new_h, new_w, 3]
    top, left = int(round((new_h - size) / 2.0)), int(round((new_w - size) / 2.0))
    return res[top:top + size, left:left + size]


def transform(img: np.ndarray, size: int, mean, std) -> np.ndarray:
    """The whole transform: float32 [3, size, size]."""
    u8 = resize_center_crop_u8(img, size)
    x = u8.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return ((x - m) / s).astype(np.float32)


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)      # lib/config/default.py:84-85
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
FIXTURE_SIZES = [(224, 224), (300, 225), (225, 300), (500, 375), (97, 640), (1024, 768), (231, 229), (64, 80)]


def fixture_image(h: int, w: int, seed: int) -> np.ndarray:
    r = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127 + 80 * np.sin(xx / 9.0 + seed)[..., None] * np.cos(yy / 7.0)[..., None] * np.array([1.0, 0.7, -0.8])
    return np.clip(base + r.randint(-60, 60, size=(h, w, 3)), 0, 255).astype(np.uint8)


def make_fixture(path: str) -> None:
    """Golden vectors from the live torchvision + Pillow stack: for every fixture size the CRC-free raw bytes of the
    resized + cropped image (S = 224) and of the float tensor."""
    import torch  # noqa: F401
    from PIL import Image
    from torchvision import transforms
    t_u8 = transforms.Compose([transforms.Resize(224, interpolation=Image.BICUBIC), transforms.CenterCrop((224, 224))])
    t_all = transforms.Compose([t_u8, transforms.ToTensor(), transforms.Normalize(mean=CLIP_MEAN, std=CLIP_STD)])
    out = {}
    for i, (h, w) in enumerate(FIXTURE_SIZES):
        pil = Image.fromarray(fixture_image(h, w, i))
        out[f"u8_{h}x{w}"] = np.asarray(t_u8(pil))
        out[f"f32_{h}x{w}"] = t_all(pil).numpy()[:, ::7, ::5]
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    make_fixture(os.path.join(here, "tests", "golden", "preprocess_cases.npz"))
    print("wrote tests/golden/preprocess_cases.npz")
