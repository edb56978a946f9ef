"""Randomised parity sweep: random sizes (1..400 x 1..300 unless $STRESS_MAXW / $STRESS_MAXH say otherwise), seeds, distances (0.05..20), efforts, proposals and flags; the CUDA
path's codestream and quality statistics must equal the oracle's for every case.  Images mix the synthetic generator with
random rectangles, pure noise, black and saturated areas.  Usage: python tools/stress_parity.py [cases] [seed]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
import oracle_lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
MAXW, MAXH = int(os.environ.get("STRESS_MAXW", 400)), int(os.environ.get("STRESS_MAXH", 300))   # e.g. 2300 x 2200: frames that cross a DC group
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ora = oracle_lib.load(rebuild=False)
t0 = time.perf_counter()
bad = 0
with pkg.Encoder(0) as enc:
    for i in range(N):
        w, h = int(rng.integers(1, MAXW)), int(rng.integers(1, MAXH))
        img = pkg.synth_image(w, h, int(rng.integers(0, 1 << 30))).copy()
        kind = int(rng.integers(0, 5))
        if kind == 1:
            img[:] = rng.integers(0, 256, img.shape, dtype=np.uint8)
        elif kind == 2:
            for _ in range(6):
                x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
                img[y0:y0 + int(rng.integers(1, 80)), x0:x0 + int(rng.integers(1, 80))] = rng.integers(0, 256, 3, dtype=np.uint8)
        elif kind == 3:
            img[: h // 2] = 0
            img[:, : w // 3] = 255
        distance = float(np.round(np.exp(rng.uniform(np.log(0.05), np.log(20.0))), 3))
        effort = int(rng.choice([3, 5, 7, 9]))
        proposal = int(rng.integers(0, 4))
        flags = int(rng.choice([0, 0, 1, 2, 3])) | (pkg.FLAG_GABORISH if rng.random() < 0.25 else 0) | (pkg.FLAG_CFL if rng.random() < 0.3 else 0)
        data, st = enc.encode(img, distance, effort, proposal, flags | pkg.FLAG_QUALITY)
        f = ora.encode(img, distance, effort, proposal, flags)
        want = f.dump("codestream").tobytes()
        sse = [int(v) for v in f.sse(img)]
        f.close()
        ok = data == want and st.sse == sse
        if not ok:
            bad += 1
            print("MISMATCH", i, (w, h), kind, distance, effort, proposal, flags, len(data), len(want), st.sse, sse, flush=True)
print(f"{N} cases, {bad} mismatches, {time.perf_counter() - t0:.1f} s")
sys.exit(1 if bad else 0)
