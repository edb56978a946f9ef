"""One 4K DCT8 encode with the quality stage (used for the ncu captures of k_dct8_quant_v4 and k_recon_sse)."""
import sys, importlib
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = 3840, 2160
d = torch.from_numpy(pkg.synth_image(w, h, 0)).cuda()
enc = pkg.Encoder(0)
for i in range(4):
    st = enc.encode_device(d.data_ptr(), w, h, 3 * w, 1.0, 7, 0, pkg.FLAG_FIXED_DCT8 | pkg.FLAG_QUALITY)
print("coeff ms", st.stage_ms[5], "quality ms", st.stage_ms[12], "psnr", st.psnr, "sse", st.sse)
