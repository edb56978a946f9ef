"""Generates tests/golden/*.json.

1. homogeneity_cases.json — known-answer vectors for the thesis' homogeneity metric (H2-H8,
   proposals/homogeneity-partitioning.diff:17-235).  The reference ships no vectors, so the expected
   values come from an INDEPENDENT restatement of the diff written here in numpy float32 (not from the
   C++ oracle), plus closed-form cases (constant block -> all ratios 1 -> DCT; all-zero block -> 0/0 ->
   NaN -> DCT; half flat / half busy -> DCT8X4 / DCT4X8; one busy quadrant -> DCT4X4).
2. codestream_pins.json — SHA-256 of oracle codestreams on seeded synthetic images (drift detector;
   regenerate deliberately when the bitstream design changes).

Run from the repo root:  python tools/make_golden.py
"""
import base64
import hashlib
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
f32 = np.float32


def laplacian(Y, x, y, xs, ys, bx, by, stride, ysize):
    mask = [[0, -1, 0], [-1, -4, -1], [0, -1, 0]]
    out = np.zeros((ys, xs), dtype=f32)
    for i in range(by, by + ys):
        for j in range(bx, bx + xs):
            s = f32(0)
            for k in (-1, 0, 1):
                for l in (-1, 0, 1):
                    xx, yy = x + j + l, y + i + k
                    if 0 <= xx < stride and 0 <= yy < ysize:      # size_t wrap at -1 fails the '<' test
                        s = f32(s + f32(Y[yy, xx] * f32(mask[k + 1][l + 1])))
            out[i - by, j - bx] = s
    return out


def zero_crossings(lap, thr):
    ys, xs = lap.shape
    cnt = 0
    for i in range(ys):
        in_edge = False
        for j in range(xs):
            if lap[i, j] > thr:
                if not in_edge:
                    cnt += 1
                in_edge = True
            else:
                in_edge = False
    avg_h = f32(cnt) / f32(ys)
    cnt = 0
    for j in range(xs):
        in_edge = False
        for i in range(ys):
            if lap[i, j] > thr:
                if not in_edge:
                    cnt += 1
                in_edge = True
            else:
                in_edge = False
    avg_v = f32(cnt) / f32(xs)
    return int(f32(avg_h + avg_v))            # size_t return: truncation


def sml(Y, x, y, xs, ys, bx, by, stride, ysize):
    s = f32(0)
    for i in range(by, by + ys):
        for j in range(bx, bx + xs):
            xx, yy = x + j, y + i
            if xx + 1 >= stride or yy + 1 >= ysize:
                continue
            if xx < 1 or yy < 1:                # defined behaviour for the diff's out-of-bounds read
                continue
            p = Y[yy, xx]
            a = abs(f32(f32(f32(2) * p - Y[yy, xx - 1]) - Y[yy, xx + 1]))
            b = abs(f32(f32(f32(2) * p - Y[yy - 1, xx]) - Y[yy + 1, xx]))
            s = f32(s + f32(a + b))
    return s


def colorfulness(X, B, x, y, xs, ys, bx, by):
    n = f32(xs * ys)
    def mean(P):
        s = f32(0)
        for i in range(by, by + ys):
            for j in range(bx, bx + xs):
                s = f32(s + P[y + i, x + j])
        return f32(s / n)
    def var(P, m):
        s = f32(0)
        for i in range(by, by + ys):
            for j in range(bx, bx + xs):
                d = f32(P[y + i, x + j] - m)
                s = f32(s + f32(d * d))
        return f32(s / n)
    mx, mb = mean(X), mean(B)
    vx, vb = var(X, mx), var(B, mb)
    a = f32(np.sqrt(f32(vx + vb)))
    b = f32(np.sqrt(f32(f32(mx * mx) + f32(mb * mb))))
    return f32(np.float64(a) + 0.3 * np.float64(b))


def homogeneity(X, Y, B, x, y, xs, ys, bx, by, d):
    stride, ysize = Y.shape[1], Y.shape[0]
    thr = 0.25
    if d > 10:
        thr = 0.40
    elif d <= 2:
        thr = 0.15
    lap = laplacian(Y, x, y, xs, ys, bx, by, stride, ysize)
    zc = zero_crossings(lap, f32(thr))
    return f32(f32(f32(zc) + sml(Y, x, y, xs, ys, bx, by, stride, ysize)) + colorfulness(X, B, x, y, xs, ys, bx, by))


def indices(X, Y, B, x, y, d):
    H = lambda xs, ys, bx, by: homogeneity(X, Y, B, x, y, xs, ys, bx, by, d)
    with np.errstate(all="ignore"):
        h1, h2 = H(8, 4, 0, 0), H(8, 4, 0, 4)
        v1, v2 = H(4, 8, 0, 0), H(4, 8, 4, 0)
        d1 = f32(H(4, 4, 0, 0) + f32(H(4, 4, 4, 4) / f32(2)))
        d2 = f32(H(4, 4, 0, 4) + f32(H(4, 4, 4, 0) / f32(2)))
        r = lambda a, b: f32(max(a, b)) / f32(min(a, b)) if not (np.isnan(a) or np.isnan(b)) else f32(np.nan)
        return [float(r(h1, h2)), float(r(v1, v2)), float(r(d1, d2))]


def partition(r_h, r_v, r_d, d):
    thr = 1.60
    if d > 10:
        thr = 1.80
    elif d <= 3:
        thr = 1.50
    if r_d > thr:
        return 3       # DCT4X4
    if r_h > r_v and r_h > thr:
        return 13      # DCT8X4
    if r_v > r_h and r_v > thr:
        return 12      # DCT4X8
    return 0


def make_cases():
    rng = np.random.default_rng(20240318)
    cases = []

    def add(name, X, Y, B, d, px=8, py=8, expect_partition=None):
        r = indices(X, Y, B, px, py, d)
        p = partition(*[f32(v) for v in r], d)
        if expect_partition is not None:
            assert p == expect_partition, (name, r, p)
        enc = lambda a: base64.b64encode(np.ascontiguousarray(a, dtype=f32).tobytes()).decode()
        cases.append({"name": name, "d": d, "px": px, "py": py, "shape": list(Y.shape), "X": enc(X), "Y": enc(Y), "B": enc(B),
                      "r": [None if np.isnan(v) else (("inf" if v > 0 else "-inf") if np.isinf(v) else v) for v in r],
                      "partition": p})

    shape = (24, 24)
    const = lambda v: np.full(shape, v, dtype=f32)
    add("constant_block_ratios_are_one", const(0.01), const(0.4), const(0.3), 1.0, expect_partition=0)
    add("all_zero_block_nan_to_dct", const(0.0), const(0.0), const(0.0), 1.0, expect_partition=0)
    base_x, base_b = const(0.002), const(0.05)
    Y = const(0.3); Y[12:, :] += (rng.random((12, 24)) * 0.2).astype(f32)           # bottom half busy
    add("top_flat_bottom_busy_dct8x4", base_x, Y, base_b, 1.0, expect_partition=None)
    Y = const(0.3); Y[:, 12:] += (rng.random((24, 12)) * 0.2).astype(f32)           # right half busy
    add("left_flat_right_busy_dct4x8", base_x, Y, base_b, 1.0, expect_partition=None)
    Y = const(0.3); Y[8:12, 8:12] += (rng.random((4, 4)) * 0.3).astype(f32)         # one busy quadrant
    add("one_busy_quadrant", base_x, Y, base_b, 1.0)
    Y = (rng.random(shape) * 0.5).astype(f32)
    Xn = ((rng.random(shape) - 0.5) * 0.02).astype(f32); Bn = (rng.random(shape) * 0.4).astype(f32)
    for d in (1.0, 2.0, 2.5, 3.0, 3.5, 10.0, 10.5):                                 # threshold switches at d = 2, 3, 10
        add(f"noise_d{d}", Xn, Y, Bn, d)
    add("first_row_first_col_block", Xn, Y, Bn, 1.0, px=0, py=0)                    # UB case: defined as skip
    add("last_block_bounds", Xn, Y, Bn, 1.0, px=16, py=16)
    Yneg = (Y - f32(0.3)).astype(f32)                                               # negative Y: Laplacian can exceed thr
    add("negative_luma_zero_crossings", Xn, Yneg, Bn, 1.0)
    return cases


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "homogeneity_cases.json"), "w") as f:
        json.dump(make_cases(), f, indent=0)
    import oracle_lib
    pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
    ora = oracle_lib.load()
    pins = []
    for (w, h, idx, d, effort, proposal, flags) in [(64, 64, 1, 1.0, 7, 0, 1), (264, 136, 2, 1.0, 7, 0, 1), (256, 256, 3, 2.0, 7, 0, 0),
                                                    (200, 120, 4, 1.0, 7, 1, 0), (200, 120, 4, 1.0, 7, 2, 0), (200, 120, 4, 1.0, 7, 3, 0),
                                                    (96, 72, 5, 8.0, 5, 3, 0), (200, 120, 4, 1.0, 7, 3, 16), (200, 120, 4, 1.0, 7, 3, 32)]:
        fr = ora.encode(pkg.synth_image(w, h, idx), d, effort, proposal, flags)
        cs = fr.dump("codestream").tobytes()
        pins.append({"w": w, "h": h, "index": idx, "distance": d, "effort": effort, "proposal": proposal, "flags": flags,
                     "bytes": len(cs), "sha256": hashlib.sha256(cs).hexdigest(),
                     "acs_sha256": hashlib.sha256(fr.dump("acs").tobytes()).hexdigest()})
    with open(os.path.join(ROOT, "tests", "golden", "codestream_pins.json"), "w") as f:
        json.dump(pins, f, indent=1)
    print("wrote", len(pins), "pins")


if __name__ == "__main__":
    main()
