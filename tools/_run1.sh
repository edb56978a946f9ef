timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
python - <<'PY'
import sys, importlib
sys.path.insert(0, "/root/repo")
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = 3840, 2160
d = torch.from_numpy(pkg.synth_image(w, h, 0)).cuda()
enc = pkg.Encoder(0)
ts = []
for i in range(8):
    st = enc.encode_device(d.data_ptr(), w, h, 3 * w, 1.0, 7, 0, 1)
    ts.append(st.stage_ms[6])
print("tokenize ms", np.round(ts[3:], 4))
PY
python tools/exp_pipelines.py
