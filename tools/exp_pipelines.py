"""Batch throughput of HBM-resident 4K frames for the knobs in the environment (JXLB200_ANS_WARPS, JXLB200_ANS_GPW, ...)."""
import sys, importlib, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = 3840, 2160
imgs = [pkg.synth_image(w, h, i) for i in range(8)]
d = [torch.from_numpy(im).cuda() for im in imgs]
names = ["h2d","xyb","aq","homog","acs","coeff","tok","histo","ans","dc","asm"]
enc = pkg.Encoder(0)
for P, B in ((32, 64),):
    enc.set_pipelines(P)
    ptrs = [d[i % 8].data_ptr() for i in range(B)]
    for _ in range(3):
        enc.encode_batch_device(ptrs, w, h, 3 * w, 1.0, 7, 0, 1)
    best = 1e9
    for _ in range(5):
        sts, ms = enc.encode_batch_device(ptrs, w, h, 3 * w, 1.0, 7, 0, 1)
        best = min(best, ms)
    sm = np.mean([s.stage_ms[:11] for s in sts], axis=0)
    print(f"warps={os.environ.get('JXLB200_ANS_WARPS','-')} gpw={os.environ.get('JXLB200_ANS_GPW','-')} P={P} B={B} dev_ms={best:.2f} per_img={best/B:.3f} GP/s={B*w*h/1e6/best:.2f}")
    print("   ", {n: round(float(v), 2) for n, v in zip(names, sm)})
