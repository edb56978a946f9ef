"""Times the DCT8 transform+quantise variants on a 4K frame (stage_ms[COEFF], CUDA events, L2 flushed)."""
import sys, importlib, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = 3840, 2160
d = torch.from_numpy(pkg.synth_image(w, h, 0)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
enc = pkg.Encoder(0)
ts = []
for i in range(10):
    flush.fill_(i); torch.cuda.synchronize()
    st = enc.encode_device(d.data_ptr(), w, h, 3 * w, 1.0, 7, 0, 1)
    ts.append(st.stage_ms[5])
print("variant", os.environ.get("JXLB200_DCT8", "default"), "rows", os.environ.get("JXLB200_DCT8_ROWS", "-"), "tps", os.environ.get("JXLB200_DCT8_TPS", "-"), "coeff ms", np.round(ts[3:], 4),
      "min", round(min(ts[3:]), 4), "bytes", st.codestream_bytes)
