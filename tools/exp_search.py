import sys, importlib
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = 3840, 2160
img = pkg.synth_image(w, h, 0)
d = torch.from_numpy(img).cuda()
enc = pkg.Encoder(0)
for i in range(3):
    st = enc.encode_device(d.data_ptr(), w, h, 3 * w, 1.0, 7, 3, 0)
print("acs %.3f coeff %.3f total %.3f bytes %d" % (st.stage_ms[4], st.stage_ms[5], st.total_ms, st.codestream_bytes))
