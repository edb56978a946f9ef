"""Small encodes that touch every kernel; run under compute-sanitizer on the GPU box."""
import importlib
import sys

sys.path.insert(0, "/root/repo")
import numpy as np

pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
with pkg.Encoder(0) as enc:
    Q = pkg.FLAG_QUALITY
    for (w, h, prop, flags) in ((264, 300, 3, Q), (8, 8, 0, 1), (1, 1, 3, 0), (257, 9, 1, Q), (520, 260, 0, 1 | Q), (96, 64, 2, 0), (320, 256, 0, Q)):
        if (w, h) == (320, 256):     # smooth gradient: 64-sized transforms
            yy, xx = np.mgrid[0:256, 0:320]
            img = np.stack([80 + xx * 0.3 + yy * 0.1, 90 + yy * 0.25, 100 + (xx + yy) * 0.15], axis=-1).astype(np.uint8)
        else:
            img = pkg.synth_image(w, h, w + h)
        data, st = enc.encode(img, 1.0, 5 if w == 96 else 7, prop, flags)
        print(w, h, prop, flags, len(data), st.num_clusters, flush=True)
    datas, _ = enc.encode_batch([pkg.synth_image(100, 60, i) for i in range(3)], [0.5, 1.0, 3.0], 7, 3, 0)
    print("batch", [len(d) for d in datas])
