"""One 1080p full-search encode (combined proposal) — used for ncu captures of k_acs / k_coeff_general."""
import sys, importlib
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = 1920, 1080
d = torch.from_numpy(pkg.synth_image(w, h, 0)).cuda()
enc = pkg.Encoder(0)
for i in range(3):
    st = enc.encode_device(d.data_ptr(), w, h, 3 * w, 1.0, 7, 3, 0)
print("acs", st.stage_ms[4], "coeff", st.stage_ms[5])
