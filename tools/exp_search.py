"""Full-search path (combined proposal): single-image stage times and batch throughput, 4K and 1080p."""
import sys, importlib, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
enc = pkg.Encoder(0)
for (w, h, B) in ((3840, 2160, 32), (1920, 1080, 128)):
    imgs = [torch.from_numpy(pkg.synth_image(w, h, i)).cuda() for i in range(8)]
    for i in range(3):
        st = enc.encode_device(imgs[0].data_ptr(), w, h, 3 * w, 1.0, 7, 3, 0)
    enc.set_pipelines(32)
    ptrs = [imgs[i % 8].data_ptr() for i in range(B)]
    best = 1e9
    for _ in range(4):
        sts, ms = enc.encode_batch_device(ptrs, w, h, 3 * w, 1.0, 7, 3, 0)
        best = min(best, ms)
    sm = np.mean([x.stage_ms[:11] for x in sts], axis=0)
    print("   batch stage ms:", " ".join("%s=%.2f" % (n, v) for n, v in zip(["h2d","xyb","aq","homog","acs","coeff","tok","histo","ans","dc","asm"], sm)))
    print("%s %dx%d single: acs %.3f coeff %.3f total %.3f ms | batch %d: %.2f ms = %.0f MP/s" % (
        os.path.basename(os.environ.get("JXLB200_LIB", "default")), w, h, st.stage_ms[4], st.stage_ms[5], st.total_ms, B, best,
        B * w * h / 1e6 / (best / 1e3)))
