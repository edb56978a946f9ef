#!/usr/bin/env python
"""One warm-up encode + one measured encode of a workload (target of ncu launch lists / --set full captures)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = int(sys.argv[1]), int(sys.argv[2])
d, effort, proposal, flags = float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
n = int(sys.argv[7]) if len(sys.argv) > 7 else 2
img = pkg.synth_image(w, h, 5)
with pkg.Encoder(0) as enc:
    for i in range(n):
        data, st = enc.encode(img, d, effort, proposal, flags)
    print(len(data), "bytes", st.total_ms, "ms", [round(v, 3) for v in st.stage_ms], "launches", st.kernel_launches)
