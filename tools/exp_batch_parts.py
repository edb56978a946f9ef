"""Batch throughput of HBM-resident 1080p frames with parts of the pipeline switched off (where the time of a frame goes
when 32 frames are in flight): full search | forced strategy map (no search, general coefficient path) | fixed DCT8 | effort 5."""
import sys, importlib, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("jpeg-xl-lossy-image-compression-thesis_b200")
w, h = 1920, 1080
img = pkg.synth_image(w, h, 5)
d = torch.from_numpy(img).cuda()
enc = pkg.Encoder(0)
data, st = enc.encode(img, 1.0, 7, 3, 0)
print("clusters", int(enc.dump("num_clusters")[0]), flush=True)
acs = enc.dump("acs").reshape((h + 7) // 8, (w + 7) // 8)
enc.set_strategy_map(acs)
P, B = 32, 128
enc.set_pipelines(P)
ptrs = [d.data_ptr()] * B
MODES = os.environ.get("EXP_MODES", "full,forced_map,fixed_dct8,effort5,full_noproposal,full_gab_cfl").split(",")
for name, effort, prop, flags in (("full", 7, 3, 0), ("forced_map", 7, 3, pkg.FLAG_FORCED_ACS), ("fixed_dct8", 7, 3, pkg.FLAG_FIXED_DCT8), ("effort5", 5, 3, 0), ("full_noproposal", 7, 0, 0),
                                   ("full_gab_cfl", 7, 3, pkg.FLAG_GABORISH | pkg.FLAG_CFL)):
    if name not in MODES:
        continue
    for _ in range(2):
        enc.encode_batch_device(ptrs, w, h, 3 * w, 1.0, effort, prop, flags)
    best = 1e9
    for _ in range(4):
        sts, ms = enc.encode_batch_device(ptrs, w, h, 3 * w, 1.0, effort, prop, flags)
        best = min(best, ms)
    print(f"{name:16s} {best / B * 1000:8.1f} us/frame  {B * w * h / 1e3 / best:8.1f} MP/s", flush=True)
