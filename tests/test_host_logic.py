"""Host-side logic that needs no GPU: input readers of the file-level execute_cjxl twin, the T3 harness' "not run" path."""
import importlib
import io
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_png_ppm_and_npy_inputs_read_alike(pkg, tmp_path):
    """The harness hands PNG paths to cjxl (benchmark.rs:654-660); the Python twin reads them (and PPM / .npy)."""
    from PIL import Image
    mod = importlib.import_module(pkg.__name__ + ".encoder")
    img = pkg.synth_image(40, 24, 3)
    png, ppm, npy = (str(tmp_path / n) for n in ("a.png", "a.ppm", "a.npy"))
    Image.fromarray(img).save(png)
    with open(ppm, "wb") as f:
        f.write(b"P6\n# comment\n40 24\n255\n" + img.tobytes())
    np.save(npy, img)
    for path in (png, ppm, npy):
        got = mod._read_image(path)
        assert got.dtype == np.uint8 and np.array_equal(got, img), path
    # RGBA and greyscale PNGs are converted to RGB8, the only colour type the path codes (image_reader.rs ColorType::Rgb8)
    rgba = str(tmp_path / "b.png")
    Image.fromarray(np.dstack([img, np.full(img.shape[:2], 255, np.uint8)]), "RGBA").save(rgba)
    assert np.array_equal(mod._read_image(rgba), img)


def test_t3_harness_reports_not_run_without_libjxl(tmp_path):
    env = dict(os.environ)
    for k in ("JXLB200_DJXL", "JXLB200_CJXL"):
        env.pop(k, None)
    env["PATH"] = "/usr/bin:/bin"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "t3_conformance.py")], capture_output=True, text=True, env=env,
                       timeout=120)
    if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "bin", "djxl")):
        return
    assert r.returncode == 0 and "T3: not run" in r.stdout
    assert json.load(open(os.path.join(ROOT, "gpurun_out", "t3_conformance.json")))["t3"] == "not run"
