"""Small encodes that touch every kernel; run under compute-sanitizer on the GPU box."""
import importlib
import sys

sys.path.insert(0, "/root/repo")
import numpy as np

pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
with pkg.Encoder(0) as enc:
    for (w, h, prop, flags) in ((264, 300, 3, 0), (8, 8, 0, 1), (1, 1, 3, 0), (257, 9, 1, 0), (520, 260, 0, 1), (96, 64, 2, 0)):
        data, st = enc.encode(pkg.synth_image(w, h, w + h), 1.0, 7, prop, flags)
        print(w, h, prop, flags, len(data), st.num_clusters, flush=True)
    datas, _ = enc.encode_batch([pkg.synth_image(100, 60, i) for i in range(3)], [0.5, 1.0, 3.0], 7, 3, 0)
    print("batch", [len(d) for d in datas])
