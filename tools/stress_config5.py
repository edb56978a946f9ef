"""BASELINE config 5 in miniature on one GPU: N synthetic 1080p images, combined proposal, distance sweep
0.5..3.0 round-robin by image index, pushed through jxlb200_encode_batch; every K-th codestream is decoded by
the self-decoder and compared with a one-at-a-time encode.  Prints throughput and bpp per distance."""
import importlib
import sys
import time

sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import torch

pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
import oracle_lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w, h = 1920, 1080
base = [torch.from_numpy(pkg.synth_image(w, h, i)).pin_memory().numpy() for i in range(8)]
imgs = [base[i % 8] for i in range(N)]
dists = [pkg.distance_for_image(i) for i in range(N)]
ora = oracle_lib.load(rebuild=False)
with pkg.Encoder(0) as enc:
    enc.set_pipelines(32)
    enc.encode_batch(imgs[:32], dists[:32], 7, pkg.PROPOSAL_COMBINED, 0)
    t0 = time.perf_counter()
    datas, sts = enc.encode_batch(imgs, dists, 7, pkg.PROPOSAL_COMBINED, 0)
    dt = time.perf_counter() - t0
    print(f"{N} x 1080p combined: {dt:.2f} s, {N * w * h / 1e6 / dt:.0f} MP/s e2e, {N / dt:.0f} images/s")
    for d in sorted(set(dists)):
        b = [s.bpp for s, dd in zip(sts, dists) if dd == d]
        print(f"  distance {d}: bpp {np.mean(b):.3f}")
    for i in range(0, N, max(1, N // 6)):
        single, _ = enc.encode(imgs[i], dists[i], 7, pkg.PROPOSAL_COMBINED, 0)
        assert single == datas[i], i
        rec = ora.decode_pixels(datas[i], w, h)
        mse = np.mean((imgs[i].astype(np.float64) - rec.astype(np.float64)) ** 2)
        print(f"  image {i}: d={dists[i]} {len(datas[i])} bytes, batch == single, decodes, PSNR {10 * np.log10(255 ** 2 / mse):.2f} dB")
print("ok")
