"""GPU parity at BASELINE.json's exact sizes and settings (configs[2..4]) through the C ABI, plus the reference-pinned
homogeneity vectors (rows H1-H7) run through the CUDA kernel itself, and the per-group stream taps.

The oracle needs ~7 s per 1920x1080 search encode and ~2 min for 7680x4320 on one core, so the oracle encodes of a test
run side by side on host threads (ctypes releases the GIL)."""
import base64
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWEEP = (0.5, 1.0, 1.5, 2.0, 2.5, 3.0)     # BASELINE configs[4]: "distance sweep 0.5-3.0 (combined.diff)"


def test_1080p_combined_at_every_sweep_distance(pkg, oracle, encoder):
    """configs[4] (one image of the 4096): 1920x1080 (8 x 5 ragged AC groups, 30 x 17 tiles with a ragged last tile row),
    combined.diff, effort 7, each distance of the sweep: strategy map and codestream equal the oracle's."""
    img = pkg.synth_image(1920, 1080, 17)
    with ThreadPoolExecutor(max_workers=len(SWEEP)) as ex:
        oras = list(ex.map(lambda d: oracle.encode(img, d, 7, 3, 0), SWEEP))
    sizes = []
    for d, ora in zip(SWEEP, oras):
        assert ora.error == ""
        data, st = encoder.encode(img, d, 7, 3, 0)
        assert np.array_equal(encoder.dump("acs"), ora.dump("acs")), d
        assert np.array_equal(np.frombuffer(data, dtype=np.uint8), ora.dump("codestream")), d
        assert st.num_groups == 40 and st.num_dc_groups == 1
        sizes.append(len(data))
    assert all(a > b for a, b in zip(sizes, sizes[1:]))       # bpp falls along the sweep


def test_8k_partitioning_equals_oracle(pkg, oracle, encoder):
    """configs[2]: 7680x4320 with the homogeneity-partitioning proposal, full search: 12 DC groups, 510 AC groups,
    8160 tiles.  Strategy map, quantised coefficients' statistics and the codestream equal the oracle's."""
    img = pkg.synth_image(7680, 4320, 23)
    with ThreadPoolExecutor(max_workers=1) as ex:
        fut = ex.submit(oracle.encode, img, 1.0, 7, 1, 0)
        data, st = encoder.encode(img, 1.0, 7, 1, 0)
        acs, nz = encoder.dump("acs"), encoder.dump("nzeros")
        ora = fut.result()
    assert ora.error == ""
    assert st.num_dc_groups == 12 and st.num_groups == 510
    assert np.array_equal(acs, ora.dump("acs"))
    assert np.array_equal(nz, ora.dump("nzeros"))
    assert np.array_equal(np.frombuffer(data, dtype=np.uint8), ora.dump("codestream"))
    first = acs[acs >= 128] & 0x7F
    assert np.isin(first, (3, 12, 13)).any() and np.isin(first, (18, 19, 20)).any()   # the override and the 64-level both act


def test_1080p_factored_entropy_batch_equals_singles(pkg, encoder):
    """configs[3]: a batch of 64 1080p images with the factored-entropy proposal through jxlb200_encode_batch equals the
    64 one-at-a-time encodes (each of which the other tests tie to the oracle)."""
    imgs = [pkg.synth_image(1920, 1080, 100 + i) for i in range(8)]
    batch = [imgs[i % 8] for i in range(64)]
    dists = [SWEEP[i % len(SWEEP)] for i in range(64)]
    encoder.set_pipelines(16)
    datas, sts = encoder.encode_batch(batch, dists, 7, 2, 0)
    encoder.set_pipelines(4)
    singles = {}
    for i in range(64):
        key = (i % 8, dists[i])
        if key not in singles:
            singles[key] = encoder.encode(batch[i], dists[i], 7, 2, 0)[0]
        assert datas[i] == singles[key], i
        assert sts[i].codestream_bytes == len(datas[i])


def _cases():
    with open(os.path.join(ROOT, "tests", "golden", "homogeneity_cases.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_homogeneity_golden_vectors_on_the_gpu(pkg, encoder, case):
    """rows H1-H7: the hand-computed vectors of tests/golden/ (independent numpy restatement of
    proposals/homogeneity-partitioning.diff:17-211, tools/make_golden.py) through the CUDA kernel with exactly their
    planes, stride and ysize (jxlb200_debug_homogeneity).  Tolerance 2e-6 relative: the golden evaluates the final
    sqrt in float, the diff's `0.3 * sqrt(...)` promotes to double (H5)."""
    shape = tuple(case["shape"])
    dec = lambda s: np.frombuffer(base64.b64decode(s), dtype=np.float32).reshape(shape)
    r = encoder.homogeneity_map(dec(case["X"]), dec(case["Y"]), dec(case["B"]), case["d"])
    got = r[case["py"] // 8, case["px"] // 8]
    for g, w in zip(got, case["r"]):
        if w is None:
            assert np.isnan(g)
        elif isinstance(w, str):
            assert np.isinf(g) and (g > 0) == (w == "inf")
        else:
            assert abs(g - w) <= 2e-6 * max(1.0, abs(w)), (got, case["r"])


def test_group_stream_taps(pkg, oracle, encoder):
    """JXLB200_STAGE_GROUP_STREAMS / GROUP_OFFSETS: every AC group's rANS section as the oracle's per-group writer holds it
    (the north star's "bit-exact group bitstreams")."""
    for (w, h, flags) in ((520, 260, 0), (264, 300, 1), (1000, 700, 0)):
        img = pkg.synth_image(w, h, w)
        encoder.encode(img, 1.0, 7, 3, flags)
        ora = oracle.encode(img, 1.0, 7, 3, flags)
        assert np.array_equal(encoder.dump("group_offsets"), ora.dump("group_offsets"))
        assert np.array_equal(encoder.dump("group_streams"), ora.dump("group_streams"))


def test_forced_strategy_map_parity(pkg, oracle, encoder):
    """Row U5 through the C ABI: a caller-supplied strategy map (jxlb200_debug_set_strategy_map + JXLB200_FLAG_FORCED_ACS) that
    uses every transform of the subset, DCT32X8 / DCT8X32 included: coefficients, statistics, codestream and the quality
    statistics equal the oracle's."""
    from test_oracle_entropy import every_strategy_map
    from test_gpu_parity import _valid_coeffs
    for (w, h, seed, dist) in ((264, 200, 5, 1.0), (520, 260, 9, 0.5), (100, 60, 2, 3.0)):
        img = pkg.synth_image(w, h, 70 + seed)
        d = pkg.frame_dims(w, h)
        acs = every_strategy_map(d["bys"], d["bxs"], seed=seed)
        encoder.set_strategy_map(acs)
        data, st = encoder.encode(img, dist, 7, 0, pkg.FLAG_FORCED_ACS | pkg.FLAG_QUALITY)
        ora = oracle.encode_forced(img, acs, dist, 7, 0, 0)
        assert ora.error == ""
        for stage in ("acs", "raw_qf", "dc_quant", "nzeros"):
            assert np.array_equal(encoder.dump(stage), ora.dump(stage)), stage
        assert np.array_equal(_valid_coeffs(encoder.dump("coeffs"), d), _valid_coeffs(ora.dump("coeffs"), d))
        assert np.array_equal(np.frombuffer(data, dtype=np.uint8), ora.dump("codestream"))
        assert st.sse == [int(v) for v in ora.sse(img)]
    with pytest.raises(pkg.EncodeError):
        bad = acs.copy(); bad[0, 0] = 5 | 0x80          # a 32x32 transform declared on a map that does not cover it
        encoder.set_strategy_map(bad)
    with pytest.raises(pkg.EncodeError):
        encoder.encode(pkg.synth_image(64, 64, 1), 1.0, 7, 0, pkg.FLAG_FORCED_ACS)   # map of another frame size


def test_six_worker_threads_with_their_own_contexts(pkg, oracle):
    """The reference's threading contract (benchmark.rs:97-103, config.rs:22): up to six OS threads call the encoder at the
    same time, each with its own context.  Every thread's codestreams equal the oracle's."""
    import threading
    jobs = [(264 + 8 * k, 200 - 8 * k, 40 + k, (0.5, 1.0, 2.0)[k % 3], k % 4) for k in range(6)]
    want = {}
    for (w, h, seed, dist, prop) in jobs:
        want[(w, h, seed)] = oracle.encode(pkg.synth_image(w, h, seed), dist, 7, prop, 0).dump("codestream").tobytes()
    errors = []

    def worker(job):
        w, h, seed, dist, prop = job
        try:
            img = pkg.synth_image(w, h, seed)
            with pkg.Encoder(0) as enc:
                for _ in range(4):
                    data, st = enc.encode(img, dist, 7, prop, 0)
                    if data != want[(w, h, seed)]:
                        errors.append(("mismatch", job))
        except Exception as e:  # noqa: BLE001
            errors.append((repr(e), job))

    threads = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_forked_single_frame_equals_the_single_stream_one(pkg, oracle, monkeypatch):
    """A lone frame runs independent stages on auxiliary streams ($JXLB200_FORK, default on); with the fork switched off
    the same bytes and the same intermediates come out, and both equal the oracle."""
    img = pkg.synth_image(520, 392, 77)
    ora = oracle.encode(img, 1.0, 7, 3, 0)
    out = {}
    for fork in ("1", "0"):
        monkeypatch.setenv("JXLB200_FORK", fork)
        with pkg.Encoder(0) as enc:
            for _ in range(3):
                data, st = enc.encode(img, 1.0, 7, 3, pkg.FLAG_QUALITY)
            out[fork] = (data, st.sse, enc.dump("homog").tobytes(), enc.dump("dc_quant").tobytes(), enc.dump("tokens").tobytes())
    assert out["1"] == out["0"]
    assert np.array_equal(np.frombuffer(out["1"][0], dtype=np.uint8), ora.dump("codestream"))
    assert out["1"][1] == [int(v) for v in ora.sse(img)]


def test_forced_strategy_map_in_batch_mode(pkg, oracle):
    """jxlb200_debug_set_strategy_map reaches every pipeline of the context: a batch with JXLB200_FLAG_FORCED_ACS gives the
    single encode's bytes on every pipeline."""
    from test_oracle_entropy import every_strategy_map
    w, h = 264, 200
    d = pkg.frame_dims(w, h)
    acs = every_strategy_map(d["bys"], d["bxs"], seed=3)
    imgs = [pkg.synth_image(w, h, 90 + i) for i in range(6)]
    with pkg.Encoder(0) as enc:
        enc.set_strategy_map(acs)
        enc.set_pipelines(4)
        datas, _ = enc.encode_batch(imgs, 1.0, 7, 0, pkg.FLAG_FORCED_ACS)
        for img, data in zip(imgs, datas):
            ora = oracle.encode_forced(img, acs, 1.0, 7, 0, 0)
            assert np.array_equal(np.frombuffer(data, dtype=np.uint8), ora.dump("codestream"))


@pytest.mark.parametrize("w,h,distance,proposal", [(520, 392, 1.0, 3), (264, 200, 3.0, 0), (100, 60, 0.5, 1), (1000, 700, 2.0, 2)])
def test_gaborish_parity(pkg, oracle, encoder, w, h, distance, proposal):
    """Row U3 (opt-in JXLB200_FLAG_GABORISH): the sharpened planes, everything computed from them, the codestream (loop
    filter bit set) and the quality statistics (reconstruction through the decoder's blur) equal the oracle's; the
    oracle's own decoder reads the stream back to the same pixels."""
    from test_gpu_parity import compare_all
    img = pkg.synth_image(w, h, 11 + w)
    flags = pkg.FLAG_GABORISH
    stages = ("xyb", "mask1x1", "homog", "acs", "raw_qf", "dc_quant", "nzeros", "coeffs", "codestream")
    if proposal == 0:
        stages = tuple(s for s in stages if s != "homog")      # (only the proposals compute the map on the device)
    compare_all(pkg, oracle, encoder, img, distance, 7, proposal, flags, stages)
    data, st = encoder.encode(img, distance, 7, proposal, flags | pkg.FLAG_QUALITY)
    ora = oracle.encode(img, distance, 7, proposal, flags)
    assert st.sse == [int(v) for v in ora.sse(img)]
    plain, _ = encoder.encode(img, distance, 7, proposal, 0)
    assert plain != data
    dec = oracle.decode_pixels(data, w, h)
    sse = ((dec.astype(np.int64) - img.astype(np.int64)) ** 2).reshape(-1, 3).sum(0)
    assert [int(v) for v in sse] == st.sse


@pytest.mark.parametrize("w,h,distance,proposal,extra", [(520, 392, 1.0, 3, 0), (264, 200, 3.0, 0, 16), (100, 60, 0.5, 2, 0),
                                                         (1000, 700, 2.0, 1, 16), (264, 136, 1.0, 0, 1)])
def test_chroma_from_luma_parity(pkg, oracle, encoder, w, h, distance, proposal, extra):
    """Row U3 (opt-in JXLB200_FLAG_CFL): the fitted colour-correlation map, the search that applies it per tile (the X channel's
    chroma-from-luma path is live here), coefficients, the map's modular stream, the codestream and the quality statistics
    equal the oracle's — alone, with Gaborish, and on the fixed-DCT8 path; the oracle's decoder reads the stream back."""
    from test_gpu_parity import compare_all
    img = pkg.synth_image(w, h, 31 + w)
    flags = pkg.FLAG_CFL | extra
    stages = ("cmap", "acs", "raw_qf", "dc_quant", "nzeros", "coeffs", "codestream")
    compare_all(pkg, oracle, encoder, img, distance, 7, proposal, flags, stages)
    assert np.any(encoder.dump("cmap") != 0)
    data, st = encoder.encode(img, distance, 7, proposal, flags | pkg.FLAG_QUALITY)
    ora = oracle.encode(img, distance, 7, proposal, flags)
    assert st.sse == [int(v) for v in ora.sse(img)]
    dec = oracle.decode_pixels(data, w, h)
    sse = ((dec.astype(np.int64) - img.astype(np.int64)) ** 2).reshape(-1, 3).sum(0)
    assert [int(v) for v in sse] == st.sse


def test_optional_filters_in_batch_mode_and_forced_map_at_low_effort(pkg, oracle, encoder):
    """A batch with JXLB200_FLAG_GABORISH | JXLB200_FLAG_CFL on four pipelines gives the single encodes' bytes (per-pipeline
    sharpened planes and factor tables); a forced strategy map below effort 5 (no quant adjustment in the coefficient
    stage) equals the oracle."""
    from test_oracle_entropy import every_strategy_map
    flags = pkg.FLAG_GABORISH | pkg.FLAG_CFL
    imgs = [pkg.synth_image(264 + 8 * i, 200, 60 + i) for i in range(6)]
    with pkg.Encoder(0) as enc:
        enc.set_pipelines(4)
        datas, _ = enc.encode_batch(imgs, [0.5, 1.0, 2.0, 1.0, 3.0, 1.5], 7, 3, flags)
    for img, data, d in zip(imgs, datas, [0.5, 1.0, 2.0, 1.0, 3.0, 1.5]):
        assert np.array_equal(np.frombuffer(data, dtype=np.uint8), oracle.encode(img, d, 7, 3, flags).dump("codestream"))
    w, h = 264, 200
    dd = pkg.frame_dims(w, h)
    acs = every_strategy_map(dd["bys"], dd["bxs"], seed=11)
    img = pkg.synth_image(w, h, 5)
    encoder.set_strategy_map(acs)
    data, st = encoder.encode(img, 1.0, 3, 0, pkg.FLAG_FORCED_ACS | pkg.FLAG_QUALITY)
    ora = oracle.encode_forced(img, acs, 1.0, 3, 0, 0)
    assert np.array_equal(np.frombuffer(data, dtype=np.uint8), ora.dump("codestream"))
    assert st.sse == [int(v) for v in ora.sse(img)]
