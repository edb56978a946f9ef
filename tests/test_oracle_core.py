"""CPU tests of the oracle's building blocks (no GPU)."""
import numpy as np
import pytest


def dct_ref_1d(x):
    n = len(x)
    k = np.arange(n)[:, None]
    i = np.arange(n)[None, :]
    m = np.cos(np.pi * (2 * i + 1) * k / (2 * n)) * np.where(k == 0, 1.0, np.sqrt(2.0)) / n
    return m @ x


@pytest.mark.parametrize("rows,cols", [(8, 8), (16, 16), (32, 32), (16, 8), (8, 16), (32, 16), (16, 32), (4, 4), (8, 4), (4, 8),
                                       (64, 64), (64, 32), (32, 64), (32, 8), (8, 32)])
def test_dct2d_matches_definition(oracle, rows, cols):
    import ctypes
    rng = np.random.default_rng(rows * 100 + cols)
    px = rng.random((rows, cols), dtype=np.float32)
    out = np.zeros(rows * cols, dtype=np.float32)
    oracle.lib.jxo_dct2d(px.ctypes.data, cols, rows, cols, out.ctypes.data)
    ref = np.apply_along_axis(dct_ref_1d, 1, px.astype(np.float64))   # horizontal
    ref = np.apply_along_axis(dct_ref_1d, 0, ref)                      # vertical -> ref[vf][hf]
    got = out.reshape(cols, rows).T if rows >= cols else out.reshape(rows, cols)
    assert np.abs(got - ref).max() < (4e-6 if max(rows, cols) == 64 else 2e-6)
    # DC = mean
    assert abs(out[0] - px.mean()) < 1e-6
    back = np.zeros((rows, cols), dtype=np.float32)
    oracle.lib.jxo_idct2d(out.ctypes.data, rows, cols, back.ctypes.data, cols)
    assert np.abs(back - px).max() < (1e-5 if max(rows, cols) == 64 else 5e-6)


@pytest.mark.parametrize("strategy,rows,cols", [(0, 8, 8), (3, 8, 8), (12, 8, 8), (13, 8, 8), (4, 16, 16), (5, 32, 32),
                                                (6, 16, 8), (7, 8, 16), (10, 32, 16), (11, 16, 32), (1, 8, 8), (2, 8, 8),
                                                (8, 32, 8), (9, 8, 32), (18, 64, 64), (19, 64, 32), (20, 32, 64)])
def test_transform_roundtrip(oracle, strategy, rows, cols):
    rng = np.random.default_rng(strategy)
    px = rng.random((rows, cols), dtype=np.float32)
    coef = oracle.transform(strategy, px)
    back = oracle.inverse_transform(strategy, coef, rows, cols)
    assert np.abs(back - px).max() < 2e-5
    assert abs(coef[0] - px.mean()) < 1e-6   # coefficient 0 carries the mean for every strategy


def test_dct2x2_and_identity_definitions(oracle):
    """DCT2X2 = three levels of 2x2 Hadamard averages (libjxl DCT2TopBlock<8>, <4>, <2>); IDENTITY = per 4x4 quadrant
    the mean, and every pixel minus pixel (1, 1) of its quadrant, with the four means Hadamard-combined."""
    rng = np.random.default_rng(7)
    px = rng.random((8, 8), dtype=np.float32)
    c = oracle.transform(2, px).reshape(8, 8).astype(np.float64)
    p = px.astype(np.float64)
    lvl1 = {k: np.zeros((4, 4)) for k in ("s", "h", "v", "d")}
    for y in range(4):
        for x in range(4):
            a, b, cc, d = p[2 * y, 2 * x], p[2 * y, 2 * x + 1], p[2 * y + 1, 2 * x], p[2 * y + 1, 2 * x + 1]
            lvl1["s"][y, x] = (a + b + cc + d) / 4; lvl1["h"][y, x] = (a + b - cc - d) / 4
            lvl1["v"][y, x] = (a - b + cc - d) / 4; lvl1["d"][y, x] = (a - b - cc + d) / 4
    assert np.abs(c[0:4, 4:8] - lvl1["h"]).max() < 1e-6 and np.abs(c[4:8, 0:4] - lvl1["v"]).max() < 1e-6
    assert np.abs(c[4:8, 4:8] - lvl1["d"]).max() < 1e-6
    assert abs(c[0, 0] - p.mean()) < 1e-6
    ci = oracle.transform(1, px).reshape(8, 8).astype(np.float64)
    for qy in range(2):
        for qx in range(2):
            quad = p[4 * qy:4 * qy + 4, 4 * qx:4 * qx + 4]
            for iy in range(4):
                for ix in range(4):
                    if (iy, ix) in ((0, 0), (1, 1)):
                        continue
                    assert abs(ci[qy + 2 * iy, qx + 2 * ix] - (quad[iy, ix] - quad[1, 1])) < 1e-6
            assert abs(ci[qy + 2, qx + 2] - (quad[0, 0] - quad[1, 1])) < 1e-6
    means = np.array([[p[0:4, 0:4].mean(), p[0:4, 4:8].mean()], [p[4:8, 0:4].mean(), p[4:8, 4:8].mean()]])
    assert abs(ci[0, 0] - means.mean()) < 1e-6
    assert abs(ci[0, 1] - (means[0].sum() - means[1].sum()) / 4) < 1e-6


def test_quant_weights_new_kinds(oracle):
    """Tables of the transforms added in round 2: IDENTITY (kind 1), DCT2X2 (2), DCT64X64 (11), DCT32X64 (12)."""
    w1 = oracle.quant_weights(1)
    assert w1.shape == (3, 64) and w1[0, 0] == 280.0 and w1[0, 1] == 3160.0 and w1[0, 8] == 3160.0 and w1[1, 9] == 864.0
    w2 = oracle.quant_weights(2)
    assert w2[0, 1] == 3840.0 and w2[0, 9] == 2560.0 and w2[0, 2] == 1280.0 and w2[0, 8 * 7 + 7] == 300.0
    w64 = oracle.quant_weights(11)
    assert w64.shape == (3, 4096) and np.all(w64 > 0) and np.all(np.diff(w64[1].reshape(64, 64)[0]) <= 1e-3)
    w3264 = oracle.quant_weights(12)
    assert w3264.shape == (3, 2048) and np.all(w3264 > 0)


def test_natural_order_dct8_is_zigzag(oracle):
    order = oracle.natural_order(0)
    assert list(order[:16]) == [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5]
    assert sorted(order) == list(range(64))
    for s in (4, 5, 6, 7, 10, 11):
        o = oracle.natural_order(s)
        assert sorted(o) == list(range(len(o)))


def test_quant_weights_dct8(oracle):
    w = oracle.quant_weights(0)
    assert w.shape == (3, 64)
    assert abs(w[0, 0] - 3150.0) < 1e-3 and abs(w[1, 0] - 560.0) < 1e-3 and abs(w[2, 0] - 512.0) < 1e-3
    assert np.all(w > 0)
    # symmetric in (x, y) for a square transform, non-increasing along the first row
    for c in range(3):
        m = w[c].reshape(8, 8)
        assert np.allclose(m, m.T, rtol=1e-6)
    assert np.all(np.diff(w[1].reshape(8, 8)[0]) <= 1e-3)


def test_srgb_lut_and_cbrt(oracle):
    lut = np.zeros(256, dtype=np.float32)
    oracle.lib.jxo_srgb_lut(lut.ctypes.data)
    v = np.arange(256) / 255.0
    exact = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)
    assert np.abs(lut - exact).max() < 2e-6       # within 1e-5 of the exact curve
    for x in (0.0, 1e-6, 0.0037930732, 0.05, 0.5, 1.0, 1.3):
        assert abs(oracle.lib.jxo_cbrt(x) - np.cbrt(x)) <= 1e-6 * max(1.0, np.cbrt(x))


def test_xyb_known_values(oracle, pkg):
    img = np.zeros((8, 8, 3), dtype=np.uint8)
    img[:, 4:, :] = 255
    f = oracle.encode(img, 1.0, 7, 0, 3)
    d = oracle.dims(8, 8)
    xyb = f.dump("xyb").reshape(3, d["ys_pad"], d["pitch"])
    # black: X = 0, Y ~ 0, B ~ 0 ; white: X = 0 (L == M), Y = B = cbrt(1 + bias) - cbrt(bias)
    assert abs(xyb[0, 0, 0]) < 1e-6 and abs(xyb[1, 0, 0]) < 1e-6 and abs(xyb[2, 0, 0]) < 1e-6
    white = np.cbrt(1.0 + 0.0037930732552754493) - np.cbrt(0.0037930732552754493)
    assert abs(xyb[1, 0, 7] - white) < 2e-6 and abs(xyb[2, 0, 7] - white) < 2e-6 and abs(xyb[0, 0, 7]) < 1e-6
    assert np.all(xyb[:, :, 8:] == 0)             # pitch padding is zero-filled


def test_oracle_frame_smoke(oracle, pkg):
    img = pkg.synth_image(100, 60, 1)            # ragged: not a multiple of 8
    f = oracle.encode(img, 1.0, 7, 0, 1)
    assert f.error == ""
    d = oracle.dims(100, 60)
    assert (d["xs_pad"], d["ys_pad"], d["bxs"], d["bys"]) == (104, 64, 13, 8)
    qp = f.dump("quant_params")
    assert qp[0] >= 1 and qp[1] >= 1
    raw = f.dump("raw_qf")
    assert raw.min() >= 1 and raw.max() <= 256
    coeffs = f.dump("coeffs").reshape(d["num_groups"], 1024, 3, 64)
    assert np.all(coeffs[:, :, :, 0] == 0)       # DC position is carried by the DC image
    nz = f.dump("nzeros").reshape(3, d["bys"], d["bxs"])
    # nzeros must equal the count of non-zero AC coefficients of each block (Y is slot 0)
    got = (coeffs[0, :, 0, :] != 0).sum(axis=1).reshape(32, 32)[: d["bys"], : d["bxs"]]
    assert np.array_equal(got, nz[1])


def test_oracle_rejects_bad_params(oracle, pkg):
    img = pkg.synth_image(16, 16, 0)
    assert oracle.encode(img, 0.0, 7).error != ""
    assert oracle.encode(img, 1.0, 0).error != ""
    assert oracle.encode(img, 1.0, 7, 9).error != ""


def test_xyb_matches_the_float64_definition(oracle, pkg):
    """Stage U1 against the published definition evaluated in float64 (exact sRGB EOTF, opsin matrix and bias of
    SURVEY 8a row U1, exact cube roots): the oracle's fp32 path with its rational-polynomial EOTF and Newton cube root
    stays within the north star's 1e-5 of it on a whole synthetic image (absolute, against values of order 1)."""
    img = pkg.synth_image(160, 96, 4)
    f = oracle.encode(img, 1.0, 7, 0, 3)
    d = oracle.dims(160, 96)
    xyb = f.dump("xyb").reshape(3, d["ys_pad"], d["pitch"])[:, :96, :160].astype(np.float64)
    f.close()
    v = img.astype(np.float64) / 255.0
    lin = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)
    m = np.array([[0.30, 0.622, 0.078], [0.23, 0.692, 0.078],
                  [0.24342268924547819, 0.20476744424496821, 0.55180986650955360]])
    bias = 0.0037930732552754493
    mix = np.maximum(lin @ m.T + bias, 0.0)
    lms = np.cbrt(mix) - np.cbrt(bias)
    want = np.stack([(lms[..., 0] - lms[..., 1]) / 2, (lms[..., 0] + lms[..., 1]) / 2, lms[..., 2]])
    assert np.abs(xyb - want).max() < 1e-5


def test_gaborish_round_trip_and_kernel(pkg, oracle):
    """Row U3 (opt-in): the encoder's 5x5 kernel is the least-squares inverse of the decoder's default 3x3 Gaborish kernel
    (tools/gen_gab_inverse.py): sharpening then blurring a smooth image returns it to within 1 %; the flag sets the loop
    filter bit, the self-decoder applies the blur, and the decoded pixels equal the encoder-side reconstruction."""
    import subprocess, sys, os, re
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "gen_gab_inverse.py")],
                         capture_output=True, text=True, check=True).stdout
    weights = [float(v) for v in re.findall(r"([-0-9.e+]+)f", out)]
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "jxo_xyb.cc")).read()
    for wv in weights:
        assert f"{np.float32(wv):.9e}f" in src
    assert abs(weights[0] + 4 * (weights[1] + weights[2] + weights[3] + weights[5]) + 8 * weights[4] - 1.0) < 1e-6
    img = pkg.synth_image(264, 200, 21)
    a = oracle.encode(img, 1.0, 7, 0, 0)
    b = oracle.encode(img, 1.0, 7, 0, 16)
    assert b.error == "" and a.dump("codestream").tobytes() != b.dump("codestream").tobytes()
    assert np.array_equal(a.dump("mask1x1"), b.dump("mask1x1"))            # the quant field stage uses the pre-sharpening planes
    dec = oracle.decode_pixels(b.dump("codestream").tobytes(), 264, 200)
    sse = ((dec.astype(np.int64) - img.astype(np.int64)) ** 2).reshape(-1, 3).sum(0)
    assert [int(v) for v in sse] == [int(v) for v in b.sse(img)]
    mse = float(sum(int(v) for v in sse)) / img.size
    assert 10 * np.log10(255.0 ** 2 / mse) > 36.0
