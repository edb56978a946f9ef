#!/usr/bin/env python
"""Stage-by-stage GPU vs oracle comparison with mismatch details (debug helper for gpurun)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
o = oracle_lib.load(rebuild=False)
cases = [(64, 64, 1.0, 7, 0, 0), (256, 200, 1.0, 7, 3, 0), (200, 120, 2.0, 5, 1, 0), (264, 300, 0.5, 7, 2, 0), (520, 260, 3.0, 9, 3, 0),
         (320, 256, 2.0, 7, 0, 0)]
if len(sys.argv) > 1:
    cases = [tuple(float(v) if i == 2 else int(v) for i, v in enumerate(a.split(","))) for a in sys.argv[1:]]
with pkg.Encoder(0) as enc:
    for (w, h, d, e, prop, flags) in cases:
        if (w, h) == (320, 256):
            yy, xx = np.mgrid[0:256, 0:320]
            img = np.stack([80 + xx * 0.3 + yy * 0.1, 90 + yy * 0.25, 100 + (xx + yy) * 0.15], axis=-1).astype(np.uint8)
        else:
            img = pkg.synth_image(w, h, w + h)
        data, st = enc.encode(img, d, e, prop, flags | pkg.FLAG_QUALITY)
        ora = o.encode(img, d, e, prop, flags)
        dims = o.dims(w, h)
        print(f"--- {w}x{h} d={d} e={e} prop={prop}: {len(data)} bytes, total {st.total_ms:.3f} ms, acs {st.stage_ms[4]:.3f} coeff {st.stage_ms[5]:.3f}")
        for stage in ("xyb", "qf_float", "mask1x1", "homog", "acs", "acs_entropy", "raw_qf", "dc_quant", "nzeros", "coeffs", "tokens", "codestream"):
            a, b = enc.dump(stage), ora.dump(stage)
            if a.shape != b.shape:
                print(f"  {stage}: SHAPE {a.shape} vs {b.shape}"); continue
            av, bv = a.view(np.uint8), b.view(np.uint8)
            if np.array_equal(av, bv):
                print(f"  {stage}: ok"); continue
            if a.dtype == np.float32:
                bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))))
            else:
                bad = np.flatnonzero(a != b)
            print(f"  {stage}: {bad.size} mismatches of {a.size}; first {bad[:6]} gpu {a[bad[:6]]} oracle {b[bad[:6]]}")
            if stage in ("acs", "acs_entropy") and bad.size:
                bxs = dims["bxs"]
                print("    blocks (bx,by):", [(int(i % bxs), int(i // bxs)) for i in bad[:8]])
        sse = [int(v) for v in ora.sse(img)]
        print("  sse:", "ok" if list(st.sse) == sse else f"gpu {list(st.sse)} oracle {sse}")
        print("  acs histogram:", {i: int(v) for i, v in enumerate(st.acs_histogram) if v})
