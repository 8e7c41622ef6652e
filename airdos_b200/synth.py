"""Seeded synthetic inputs shaped like the TartanAir-Shibuya stream AirDOS runs on.

numpy only (no cv2) so the same generator runs in the tests, in ``bench.py`` and on the GPU
box.  Shapes follow BASELINE.json's configs (640x480 u8 stereo pairs; BA windows of K key-frames /
P points / 6 observations per point); distributions follow SURVEY.md Appendix E.
"""
from __future__ import annotations

import numpy as np

# Examples/Stereo/config/tartanair.yaml:20-25 (reference camera); cy moved to 240 for 480 rows.
FX = 772.548
FY = 772.548
CX = 320.0
CY = 240.0
BF = 193.137


def _upsample4(a: np.ndarray) -> np.ndarray:
    """Bilinear x4 up-sampling of a 2-D float array (edge clamped)."""
    h, w = a.shape
    ys = (np.arange(h * 4) + 0.5) / 4 - 0.5
    xs = (np.arange(w * 4) + 0.5) / 4 - 0.5
    y0 = np.clip(np.floor(ys).astype(int), 0, h - 1)
    x0 = np.clip(np.floor(xs).astype(int), 0, w - 1)
    y1 = np.clip(y0 + 1, 0, h - 1)
    x1 = np.clip(x0 + 1, 0, w - 1)
    fy = np.clip(ys - y0, 0, 1)[:, None]
    fx = np.clip(xs - x0, 0, 1)[None, :]
    top = a[y0][:, x0] * (1 - fx) + a[y0][:, x1] * fx
    bot = a[y1][:, x0] * (1 - fx) + a[y1][:, x1] * fx
    return top * (1 - fy) + bot * fy


def _smooth(a: np.ndarray) -> np.ndarray:
    """Separable [1 2 1]/4 binomial smoothing, edge replicated."""
    p = np.pad(a, 1, mode="edge")
    a = (p[1:-1, :-2] + 2 * p[1:-1, 1:-1] + p[1:-1, 2:]) * 0.25
    p = np.pad(a, 1, mode="edge")
    return (p[:-2, 1:-1] + 2 * p[1:-1, 1:-1] + p[2:, 1:-1]) * 0.25


def make_image(seed: int, width: int = 640, height: int = 480, n_shapes: int = 200) -> np.ndarray:
    """One textured u8 image with plenty of corners (float64 scene, before camera noise)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(40, 216, size=((height + 3) // 4, (width + 3) // 4)).astype(np.float64)
    img = _upsample4(base)[:height, :width]
    for _ in range(n_shapes):
        w = int(rng.integers(6, 60))
        h = int(rng.integers(6, 60))
        x = int(rng.integers(0, width - 6))
        y = int(rng.integers(0, height - 6))
        img[y:y + h, x:x + w] = float(rng.integers(0, 256))
    return _smooth(img)


def _finish(scene: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    noisy = scene + rng.normal(0.0, 2.0, size=scene.shape)
    return np.clip(np.rint(noisy), 0, 255).astype(np.uint8)


def make_stereo_pair(frame: int, width: int = 640, height: int = 480):
    """(left, right) u8 images.  The right image is the left scene displaced by d = bf / Z with a
    piece-wise constant depth Z in [3, 40] m per 40-row band (linear interpolation in x)."""
    scene = make_image(1000 + frame, width, height)
    rng = np.random.default_rng(2000 + frame)
    bands = (height + 39) // 40
    depth = rng.uniform(3.0, 40.0, size=bands)
    disp = np.repeat(BF / depth, 40)[:height]
    xs = np.arange(width)[None, :] + disp[:, None]          # right(x) = left(x + d)
    x0 = np.floor(xs).astype(int)
    fx = xs - x0
    x0c = np.clip(x0, 0, width - 1)
    x1c = np.clip(x0 + 1, 0, width - 1)
    rows = np.arange(height)[:, None]
    right_scene = scene[rows, x0c] * (1 - fx) + scene[rows, x1c] * fx
    return _finish(scene, rng), _finish(right_scene, rng)


def make_stereo_batch(n_pairs: int, width: int = 640, height: int = 480, start: int = 0) -> np.ndarray:
    """u8 array [n_pairs, 2, height, width] (index 0 = left, 1 = right)."""
    out = np.empty((n_pairs, 2, height, width), np.uint8)
    for f in range(n_pairs):
        out[f, 0], out[f, 1] = make_stereo_pair(start + f, width, height)
    return out


def make_mask(seed: int, width: int = 640, height: int = 480, n_rect: int = 3) -> np.ndarray:
    """Extractor mask as Frame::ExtractORB builds it (src/Frame.cc:553-560): 255 = keep, 0 = human."""
    rng = np.random.default_rng(seed)
    m = np.full((height, width), 255, np.uint8)
    for _ in range(n_rect):
        w = int(rng.integers(40, 200))
        h = int(rng.integers(60, 300))
        x = int(rng.integers(0, width - 40))
        y = int(rng.integers(0, height - 60))
        m[y:y + h, x:x + w] = 0
    return m
