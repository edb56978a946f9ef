"""Deterministic synthetic test images.

Image ``i`` is seeded ``0xB2000000 + i`` through a counter-based Philox
generator, so the oracle and the CUDA path are always fed identical bytes.
Content is a blend of four layers so that every AC-strategy class and every
homogeneity-metric branch of the proposals is exercised:

1. smooth low-frequency sinusoid gradients with random phase,
2. band-limited 1/f noise (sigma ~ 12 LSB),
3. hard-edged axis-aligned rectangles and 1-px lines (text-like),
4. flat patches, including pure black (the 0/0 -> NaN path of
   ``CalculateHomogeneitySimilarityIndices``,
   proposals/homogeneity-partitioning.diff:183-211) and saturated colours.

The reference harness reads user-supplied PNGs from ``./test_images``
(benchmark-jpegxl/src/benchmark.rs:406-433); there is no dataset here, so
this generator stands in for it.
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0xB2000000


def _rng(index: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=SEED_BASE + int(index)))


def synth_image(width: int, height: int, index: int = 0) -> np.ndarray:
    """Return a ``(height, width, 3)`` uint8 sRGB image, C-contiguous."""
    rng = _rng(index)
    h, w = int(height), int(width)
    yy = np.arange(h, dtype=np.float32)[:, None]
    xx = np.arange(w, dtype=np.float32)[None, :]

    # layer 1: smooth gradients ------------------------------------------------
    img = np.empty((h, w, 3), dtype=np.float32)
    for c in range(3):
        acc = np.full((h, w), 128.0, dtype=np.float32)
        for _ in range(3):
            fx = rng.uniform(0.2, 2.5) / max(w, 1)
            fy = rng.uniform(0.2, 2.5) / max(h, 1)
            ph = rng.uniform(0, 2 * np.pi)
            amp = rng.uniform(15.0, 45.0)
            acc += amp * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph).astype(np.float32)
        img[:, :, c] = acc

    # layer 2: 1/f band-limited noise (separable box-filter pyramid; cheap, no FFT)
    noise = np.zeros((h, w), dtype=np.float32)
    for octave, sigma in ((1, 5.0), (4, 7.0), (16, 8.0)):
        hh, ww = (h + octave - 1) // octave + 1, (w + octave - 1) // octave + 1
        coarse = rng.standard_normal((hh, ww), dtype=np.float32) * sigma
        up = np.repeat(np.repeat(coarse, octave, axis=0), octave, axis=1)[:h, :w]
        noise += up
    # modulate noise strength spatially so smooth and busy regions coexist
    gate = 0.5 + 0.5 * np.sin(2 * np.pi * (xx / max(w, 1) * 1.5 + yy / max(h, 1) * 0.75))
    gate = (gate.astype(np.float32)) ** 2
    for c in range(3):
        img[:, :, c] += noise * gate * (1.0 if c == 1 else 0.8)

    # layer 3: hard rectangles and lines ----------------------------------------
    n_rect = max(4, (w * h) // (256 * 256) * 3)
    for _ in range(int(n_rect)):
        rw = int(rng.integers(2, max(3, w // 6)))
        rh = int(rng.integers(2, max(3, h // 6)))
        x0 = int(rng.integers(0, max(1, w - 1)))
        y0 = int(rng.integers(0, max(1, h - 1)))
        col = rng.integers(0, 256, size=3).astype(np.float32)
        if rng.random() < 0.4:  # thin line
            if rng.random() < 0.5:
                rh = 1
            else:
                rw = 1
        img[y0:y0 + rh, x0:x0 + rw, :] = col

    # layer 4: flat patches (black, white, saturated) ---------------------------
    flats = [(0, 0, 0), (255, 255, 255), (255, 0, 0), (0, 0, 255), (0, 255, 0), (17, 17, 17)]
    n_flat = max(3, (w * h) // (512 * 512) * 2)
    for k in range(int(n_flat)):
        pw = int(rng.integers(8, max(9, w // 8)))
        ph_ = int(rng.integers(8, max(9, h // 8)))
        # snap to the 8-px block grid so whole blocks are flat
        x0 = int(rng.integers(0, max(1, w - 1))) // 8 * 8
        y0 = int(rng.integers(0, max(1, h - 1))) // 8 * 8
        img[y0:y0 + ph_ // 8 * 8, x0:x0 + pw // 8 * 8, :] = np.array(
            flats[k % len(flats)], dtype=np.float32)

    out = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(out)


def synth_batch(width: int, height: int, count: int, first_index: int = 0) -> np.ndarray:
    """``(count, height, width, 3)`` uint8 batch of distinct images."""
    return np.stack([synth_image(width, height, first_index + i) for i in range(count)])
