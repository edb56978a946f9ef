"""GPU parity: every intermediate of the CUDA path against the CPU oracle on the same seeded
inputs, through the C ABI.  Bar: bit-exact (integers AND floats: both sides follow the same
operation order, fmaf only where written)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLOAT_STAGES = ("xyb", "mask1x1", "qf_float", "mask", "homog")
INT_STAGES = ("quant_params", "raw_qf", "dc_quant", "nzeros")


def _bits(a):
    return a.view(np.uint8)


def _valid_coeffs(c, d):
    """keeps only the block slots that exist in the frame (edge groups have unused slots)"""
    c = c.reshape(d["gys"], d["gxs"], 32, 32, 3, 64)
    full = c.transpose(0, 2, 1, 3, 4, 5).reshape(d["gys"] * 32, d["gxs"] * 32, 3, 64)
    return full[: d["bys"], : d["bxs"]]


def compare_all(pkg, oracle, enc, img, distance, effort, proposal, flags, stages):
    enc.encode(img, distance, effort, proposal, flags)
    ora = oracle.encode(img, distance, effort, proposal, flags)
    assert ora.error == ""
    d = pkg.frame_dims(img.shape[1], img.shape[0])
    for st in stages:
        a, b = enc.dump(st), ora.dump(st)
        assert a.shape == b.shape, st
        if st == "coeffs":
            a, b = _valid_coeffs(a, d), _valid_coeffs(b, d)
        if st in FLOAT_STAGES:
            # north star tolerance for XYB / DCT outputs is 1e-5 relative; we require bit equality
            same = _bits(np.ascontiguousarray(a)) == _bits(np.ascontiguousarray(b))
            if not same.all():
                bad = np.flatnonzero(~(a == b) & ~(np.isnan(a) & np.isnan(b)))
                assert bad.size == 0, f"{st}: {bad.size} mismatches, first at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}"
        else:
            assert np.array_equal(a, b), f"{st}: {np.count_nonzero(a != b)} mismatches"


@pytest.mark.parametrize("w,h", [(512, 512), (100, 60), (8, 8), (264, 40), (1, 1), (257, 9)])
@pytest.mark.parametrize("flags", [1, 3])
def test_stage_parity_dct8(pkg, oracle, encoder, w, h, flags):
    img = pkg.synth_image(w, h, w * 7 + h)
    compare_all(pkg, oracle, encoder, img, 1.0, 7, 3, flags, FLOAT_STAGES + INT_STAGES + ("coeffs",))


def test_stage_parity_very_small_distance(pkg, oracle, encoder):
    """distance 0.02 on hard edges: quantised values beyond the DCT8 kernel's 1024-entry dequantisation-bias table
    (its formula path) and quants at the 255 / 256 clamps."""
    rng = np.random.default_rng(5)
    img = pkg.synth_image(264, 136, 9)
    img[::3, ::5] = 255
    img[40:80, 100:160] = rng.integers(0, 2, (40, 60, 3), dtype=np.uint8) * 255
    for d in (0.02, 0.1):
        compare_all(pkg, oracle, encoder, img, d, 7, 0, 1, INT_STAGES + ("coeffs",))
        assert np.abs(encoder.dump("coeffs").astype(np.int32)).max() >= 1024 or d > 0.02


@pytest.mark.parametrize("distance", [0.5, 1.0, 1.5, 3.0, 8.0, 14.0])
def test_stage_parity_distances(pkg, oracle, encoder, distance):
    img = pkg.synth_image(320, 256, 11)
    compare_all(pkg, oracle, encoder, img, distance, 7, 3, 1, FLOAT_STAGES + INT_STAGES + ("coeffs",))


@pytest.mark.parametrize("effort", [3, 5, 9])
def test_stage_parity_efforts(pkg, oracle, encoder, effort):
    img = pkg.synth_image(200, 136, 5)
    compare_all(pkg, oracle, encoder, img, 1.0, effort, 0, 1, INT_STAGES + ("coeffs",))


def test_extreme_images(pkg, oracle, encoder):
    for fill in (0, 255):
        img = np.full((64, 72, 3), fill, dtype=np.uint8)
        compare_all(pkg, oracle, encoder, img, 1.0, 7, 3, 1, FLOAT_STAGES + INT_STAGES + ("coeffs",))
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(72, 96, 3), dtype=np.uint8)
    compare_all(pkg, oracle, encoder, img, 1.0, 7, 3, 1, FLOAT_STAGES + INT_STAGES + ("coeffs",))


def test_error_behaviour(pkg, encoder):
    img = pkg.synth_image(16, 16, 0)
    for kwargs in (dict(distance=0.0), dict(effort=0), dict(effort=10), dict(proposal=7)):
        with pytest.raises(pkg.EncodeError):
            encoder.encode(img, **kwargs)
    with pytest.raises(pkg.EncodeError):
        encoder.encode(np.zeros((4, 4), dtype=np.uint8))
    for shape in ((0, 0, 3), (0, 8, 3), (8, 0, 3)):          # empty inputs are refused, not crashed on
        with pytest.raises(pkg.EncodeError):
            encoder.encode(np.zeros(shape, dtype=np.uint8))
    encoder.encode(img)   # the context stays usable after an error (skip-and-continue)


def test_strided_input(pkg, oracle, encoder):
    big = pkg.synth_image(300, 80, 2)
    view = big[:, 10:210, :]          # row stride larger than 3*width
    enc_bytes = encoder.encode(view, 1.0, 7, 0, 1)
    a = encoder.dump("coeffs")
    ora = oracle.encode(np.ascontiguousarray(view), 1.0, 7, 0, 1)
    d = pkg.frame_dims(200, 80)
    assert np.array_equal(_valid_coeffs(a, d), _valid_coeffs(ora.dump("coeffs"), d))


# ---------------------------------------------------------------------------------------------
# entropy stage (U6-U9): tokens, histograms, clustering, ANS group streams, modular DC / metadata,
# frame assembly — all bit-exact against the oracle, and the GPU codestream must decode.
ENTROPY_STAGES = ("token_offsets", "tokens", "histograms", "num_clusters", "context_map", "codestream")


def compare_entropy(pkg, oracle, enc, img, distance=1.0, effort=7, proposal=0, flags=1):
    data, st = enc.encode(img, distance, effort, proposal, flags)
    ora = oracle.encode(img, distance, effort, proposal, flags)
    assert ora.error == ""
    for stg in ENTROPY_STAGES:
        a, b = enc.dump(stg), ora.dump(stg)
        assert a.shape == b.shape, f"{stg}: shape {a.shape} vs {b.shape}"
        if not np.array_equal(a, b):
            bad = np.flatnonzero(a != b)
            raise AssertionError(f"{stg}: {bad.size} mismatches, first at {bad[:8]}: {a[bad[:8]]} vs {b[bad[:8]]}")
    ref = ora.dump("codestream")
    assert np.array_equal(np.frombuffer(data, dtype=np.uint8), ref)
    assert st.codestream_bytes == ref.size and abs(st.bpp - 8.0 * ref.size / (img.shape[0] * img.shape[1])) < 1e-9
    dec = oracle.decode(data)
    assert dec.error == "", dec.error
    for stg in ("dc_quant", "acs", "raw_qf", "nzeros"):
        assert np.array_equal(dec.dump(stg), enc.dump(stg)), stg
    d = pkg.frame_dims(img.shape[1], img.shape[0])
    assert np.array_equal(_valid_coeffs(dec.dump("coeffs"), d), _valid_coeffs(enc.dump("coeffs"), d))
    return st


@pytest.mark.parametrize("w,h", [(64, 64), (256, 256), (8, 8), (1, 1), (257, 9), (264, 300), (520, 260), (1000, 700)])
def test_entropy_parity_sizes(pkg, oracle, encoder, w, h):
    compare_entropy(pkg, oracle, encoder, pkg.synth_image(w, h, w + 3 * h))


@pytest.mark.parametrize("distance", [0.5, 1.5, 3.0, 8.0, 14.0])
def test_entropy_parity_distances(pkg, oracle, encoder, distance):
    compare_entropy(pkg, oracle, encoder, pkg.synth_image(320, 200, 7), distance)


def test_entropy_parity_extremes(pkg, oracle, encoder):
    for fill in (0, 255):
        compare_entropy(pkg, oracle, encoder, np.full((40, 72, 3), fill, dtype=np.uint8))
    rng = np.random.default_rng(3)
    compare_entropy(pkg, oracle, encoder, rng.integers(0, 256, size=(72, 96, 3), dtype=np.uint8))
    compare_entropy(pkg, oracle, encoder, rng.integers(0, 256, size=(300, 280, 3), dtype=np.uint8), distance=0.1)


def test_entropy_parity_multi_dc_group(pkg, oracle, encoder):
    # 2304 x 2100 px: 2 x 2 DC groups, 9 x 9 AC groups
    st = compare_entropy(pkg, oracle, encoder, pkg.synth_image(2304, 2100, 13))
    assert st.num_dc_groups == 4 and st.num_groups == 81


def test_full_size_properties(pkg, oracle, encoder):
    """BASELINE config 2 size (3840x2160): too slow to diff every stage in a test, so check size-independent
    properties: the codestream decodes, and decodes to exactly the integers the device holds."""
    img = pkg.synth_image(3840, 2160, 1)
    data, st = encoder.encode(img, 1.0, 7, 0, 1)
    dec = oracle.decode(data)
    assert dec.error == "", dec.error
    d = pkg.frame_dims(3840, 2160)
    assert np.array_equal(dec.dump("dc_quant"), encoder.dump("dc_quant"))
    assert np.array_equal(_valid_coeffs(dec.dump("coeffs"), d), _valid_coeffs(encoder.dump("coeffs"), d))
    assert st.num_groups == 135 and st.num_dc_groups == 4 and 0.2 < st.bpp < 4.0


def test_batch_matches_single(pkg, oracle, encoder):
    """jxlb200_encode_batch keeps several images in flight on separate streams; every codestream must equal the
    one-at-a-time result (and so the oracle's)."""
    imgs = [pkg.synth_image(200 + 24 * i, 136 + 8 * i, 40 + i) for i in range(7)]
    dists = [0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 1.0]
    encoder.set_pipelines(3)
    datas, sts = encoder.encode_batch(imgs, dists, 7, 0, 1)
    for im, d, data, st in zip(imgs, dists, datas, sts):
        single, _ = encoder.encode(im, d, 7, 0, 1)
        assert data == single
        assert np.array_equal(np.frombuffer(data, dtype=np.uint8), oracle.encode(im, d, 7, 0, 1).dump("codestream"))
        assert st.codestream_bytes == len(data)
    import torch
    d_imgs = [torch.from_numpy(imgs[0]).cuda() for _ in range(5)]
    sts, ms = encoder.encode_batch_device([t.data_ptr() for t in d_imgs], imgs[0].shape[1], imgs[0].shape[0],
                                          3 * imgs[0].shape[1], 0.5, 7, 0, 1)
    assert ms > 0 and all(s.codestream_bytes == len(datas[0]) for s in sts)
    encoder.set_pipelines(4)


# ---------------------------------------------------------------------------------------------
# AC-strategy search (U4) with the proposals' hooks (H8 / H9 / H10) and the general transform +
# quantise kernel: strategy map, entropy estimates, coefficients and the whole codestream bit-exact.
ACS_STAGES = ("acs", "acs_entropy", "raw_qf", "dc_quant", "nzeros", "coeffs") + ENTROPY_STAGES


def compare_full(pkg, oracle, enc, img, distance, effort, proposal):
    data, st = enc.encode(img, distance, effort, proposal, 0)
    ora = oracle.encode(img, distance, effort, proposal, 0)
    assert ora.error == ""
    d = pkg.frame_dims(img.shape[1], img.shape[0])
    for stg in ACS_STAGES:
        a, b = enc.dump(stg), ora.dump(stg)
        assert a.shape == b.shape, stg
        if stg == "coeffs":
            a, b = _valid_coeffs(a, d), _valid_coeffs(b, d)
        if stg == "acs_entropy":
            same = (a == b) | (np.isnan(a) & np.isnan(b))
        else:
            same = a == b
        if not same.all():
            bad = np.flatnonzero(~same.reshape(-1))
            raise AssertionError(f"{stg}: {bad.size} mismatches, first at {bad[:8]}: {a.reshape(-1)[bad[:8]]} vs {b.reshape(-1)[bad[:8]]}")
    dec = oracle.decode(data)
    assert dec.error == "", dec.error
    assert np.array_equal(dec.dump("acs"), enc.dump("acs"))
    return st


@pytest.mark.parametrize("proposal", [0, 1, 2, 3])
def test_acs_parity_proposals(pkg, oracle, encoder, proposal):
    st = compare_full(pkg, oracle, encoder, pkg.synth_image(512, 384, 3), 1.0, 7, proposal)


@pytest.mark.parametrize("w,h", [(64, 64), (8, 8), (1, 1), (257, 9), (264, 300), (100, 60)])
def test_acs_parity_sizes(pkg, oracle, encoder, w, h):
    compare_full(pkg, oracle, encoder, pkg.synth_image(w, h, 2 * w + h), 1.0, 7, 3)


@pytest.mark.parametrize("distance", [0.5, 2.0, 3.0, 4.5, 8.0, 12.0])
def test_acs_parity_distances(pkg, oracle, encoder, distance):
    compare_full(pkg, oracle, encoder, pkg.synth_image(320, 256, 17), distance, 7, 3)


def test_acs_parity_extremes(pkg, oracle, encoder):
    for fill in (0, 255):   # all-black: 0/0 -> NaN ratios (H7), NaN entropy under H9
        compare_full(pkg, oracle, encoder, np.full((64, 96, 3), fill, dtype=np.uint8), 1.0, 7, 3)
    rng = np.random.default_rng(5)
    compare_full(pkg, oracle, encoder, rng.integers(0, 256, size=(96, 128, 3), dtype=np.uint8), 1.0, 7, 3)
    img = pkg.synth_image(256, 256, 9)
    compare_full(pkg, oracle, encoder, img, 1.0, 4, 3)   # effort < 5: no search
    compare_full(pkg, oracle, encoder, img, 1.0, 9, 1)


def test_codestream_pins_gpu(pkg, encoder):
    """The committed golden pins (tests/golden/codestream_pins.json) hold for the CUDA path as well."""
    import hashlib, json, os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "codestream_pins.json")) as f:
        pins = json.load(f)
    for p in pins:
        data, _ = encoder.encode(pkg.synth_image(p["w"], p["h"], p["index"]), p["distance"], p["effort"], p["proposal"], p["flags"])
        assert len(data) == p["bytes"] and hashlib.sha256(data).hexdigest() == p["sha256"], p
        assert hashlib.sha256(encoder.dump("acs").tobytes()).hexdigest() == p["acs_sha256"], p


def test_gpu_codestream_decodes_to_the_image(pkg, oracle, encoder):
    """The CUDA path's codestream, decoded to pixels by the independent self-decoder, is the input image at the
    quality the distance asks for (all strategies + the combined proposal)."""
    img = pkg.synth_image(512, 384, 21)
    prev = 99.0
    for d, floor in ((0.5, 42.0), (1.0, 38.0), (3.0, 33.0)):
        data, st = encoder.encode(img, d, 7, pkg.PROPOSAL_COMBINED, 0)
        rec = oracle.decode_pixels(data, 512, 384)
        assert rec is not None
        mse = np.mean((img.astype(np.float64) - rec.astype(np.float64)) ** 2)
        p = 10.0 * np.log10(255.0 ** 2 / mse)
        assert floor < p < prev, (d, p)
        prev = p


def test_full_size_search_codestream_equals_oracle(pkg, oracle, encoder):
    """BASELINE config 2/3 scale with the full search: the 3840x2160 codestream of the combined proposal is
    byte-identical to the oracle's (4 DC groups, 135 AC groups, every transform size)."""
    img = pkg.synth_image(3840, 2160, 2)
    data, st = encoder.encode(img, 1.0, 7, pkg.PROPOSAL_COMBINED, 0)
    ref = oracle.encode(img, 1.0, 7, pkg.PROPOSAL_COMBINED, 0)
    assert np.array_equal(encoder.dump("acs"), ref.dump("acs"))
    assert np.array_equal(np.frombuffer(data, dtype=np.uint8), ref.dump("codestream"))
    acs = encoder.dump("acs")
    first = acs[acs >= 128] & 0x7F
    assert [int((first == s).sum()) for s in range(27)] == list(st.acs_histogram)   # stats report the partition
    assert sum(1 for c in st.acs_histogram if c) >= 6


@pytest.mark.parametrize("w,h,proposal,flags", [(256, 200, 3, 0), (512, 384, 0, 1), (131, 77, 1, 0), (640, 480, 2, 0), (8, 8, 3, 0),
                                                (1, 1, 0, 1), (257, 9, 3, 0)])
def test_quality_stats_equal_oracle(pkg, oracle, encoder, w, h, proposal, flags):
    """FLAG_QUALITY: the device reconstructs the frame it coded (dequantise, DC -> LLF, chroma-from-luma, inverse
    transforms of every strategy, XYB -> 8-bit sRGB) and returns the per-channel squared error against the input:
    bit-exact integers against oracle/jxo_recon.cc; the PSNR follows calculate_psnr (image_reader.rs:604-606)."""
    img = pkg.synth_image(w, h, 5)
    for distance in (0.5, 1.0, 3.0):
        data, st = encoder.encode(img, distance, 7, proposal, flags | pkg.FLAG_QUALITY)
        f = oracle.encode(img, distance, 7, proposal, flags)
        assert f.error == ""
        want = [int(v) for v in f.sse(img)]
        f.close()
        assert st.sse == want, (distance, st.sse, want)
        mse = sum(want) / (3.0 * w * h)
        assert (np.isinf(st.psnr) and st.psnr > 0) if mse == 0 else abs(st.psnr - 10.0 * np.log10(255.0 ** 2 / mse)) < 1e-9
        # the codestream does not depend on the flag
        data0, st0 = encoder.encode(img, distance, 7, proposal, flags)
        assert data0 == data and st0.sse is None
        # and the statistics describe what the BITSTREAM carries: decoding the bytes with the independent self-decoder
        # (entropy decode, dequantise, inverse transforms, colour) gives the same squared error
        rec = oracle.decode_pixels(data, w, h)
        assert rec is not None
        d2 = (rec.astype(np.int64) - img.astype(np.int64)) ** 2
        assert [int(d2[:, :, c].sum()) for c in range(3)] == st.sse


def test_quality_stats_batch_and_full_size(pkg, oracle, encoder):
    """The quality stage inside the pipelined batch path, at 4K (size-independent property: the statistics of an
    image do not depend on how many others are in flight)."""
    w, h = 3840, 2160
    imgs = [pkg.synth_image(w, h, 40 + i) for i in range(3)]
    single = [encoder.encode(im, 1.0, 7, 0, pkg.FLAG_FIXED_DCT8 | pkg.FLAG_QUALITY)[1] for im in imgs]
    encoder.set_pipelines(3)
    datas, sts = encoder.encode_batch(imgs, 1.0, 7, 0, pkg.FLAG_FIXED_DCT8 | pkg.FLAG_QUALITY)
    for a, b in zip(single, sts):
        assert a.sse == b.sse and a.psnr == b.psnr and 30.0 < a.psnr < 60.0


def test_randomised_parity_sweep():
    """60 random (size, content, distance, effort, proposal, flags) cases: codestream and quality statistics equal the
    oracle's (tools/stress_parity.py runs the same sweep at any length; 5 300 cases were clean in round 1, 2 000 of them on
    the round's last commit)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "stress_parity.py"), "60", "11"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_two_devices_in_one_process(pkg, oracle):
    """Contexts on different GPUs of one process (the C ABI allows it: jxlb200_create(device)) give the same bytes;
    covers per-device state such as kernel attributes and constant tables.  Needs two visible GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    img = pkg.synth_image(520, 264, 77)
    outs = []
    for dev in (1, 0, 1):
        with pkg.Encoder(dev) as enc:
            for (proposal, flags) in ((3, pkg.FLAG_QUALITY), (0, pkg.FLAG_FIXED_DCT8)):
                data, st = enc.encode(img, 1.0, 7, proposal, flags)
                outs.append((dev, proposal, data, st.sse))
    ref = {}
    for dev, proposal, data, sse in outs:
        ref.setdefault(proposal, (data, sse))
        assert (data, sse) == ref[proposal], (dev, proposal)
    want = oracle.encode(img, 1.0, 7, 3, 0)
    assert ref[3][0] == want.dump("codestream").tobytes()
    want.close()
