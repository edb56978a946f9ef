import sys, importlib, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = 3840, 2160
d = torch.from_numpy(pkg.synth_image(w, h, 0)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
enc = pkg.Encoder(0)
ts = []
for i in range(8):
    flush.fill_(i); torch.cuda.synchronize()
    st = enc.encode_device(d.data_ptr(), w, h, 3 * w, 1.0, 7, 0, 1)
    ts.append(st.stage_ms[5])
print("v2" if os.environ.get("JXLB200_DCT8_V2") else "v1", "coeff ms", np.round(ts[3:], 4), "bytes", st.codestream_bytes)
