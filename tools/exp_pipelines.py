import sys, importlib, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = 3840, 2160
imgs = [pkg.synth_image(w, h, i) for i in range(8)]
d = [torch.from_numpy(im).cuda() for im in imgs]
names = ["h2d","xyb","aq","homog","acs","coeff","tok","histo","ans","dc","asm"]
enc = pkg.Encoder(0)
for P, B in ((32, 64), (48, 96), (64, 128)):
    enc.set_pipelines(P)
    ptrs = [d[i % 8].data_ptr() for i in range(B)]
    for _ in range(2):
        enc.encode_batch_device(ptrs, w, h, 3 * w, 1.0, 7, 0, 1)
    t0 = time.perf_counter()
    sts, ms = enc.encode_batch_device(ptrs, w, h, 3 * w, 1.0, 7, 0, 1)
    wall = (time.perf_counter() - t0) * 1e3
    sm = np.mean([s.stage_ms[:11] for s in sts], axis=0)
    print(f"P={P} B={B} dev_ms={ms:.2f} wall={wall:.2f} per_img={ms/B:.3f} GP/s={B*w*h/1e6/ms:.2f} latency_per_img={np.mean([s.total_ms for s in sts]):.2f}")
    print("   ", {n: round(float(v), 2) for n, v in zip(names, sm)})
