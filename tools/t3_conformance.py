#!/usr/bin/env python
"""Tier T3 conformance harness: decode this encoder's codestreams with the REFERENCE decoder and
compare size / quality with the reference encoder, the way the thesis harness measures them
(benchmark-jpegxl/src/docker_manager.rs:136 cjxl, :155 ssimulacra2, :174 butteraugli_main; image_reader.rs:370 decode).

libjxl is not in the reference tree and cannot be installed offline, so this tool looks for binaries
  $JXLB200_DJXL / $JXLB200_CJXL / $JXLB200_SSIMULACRA2 / $JXLB200_BUTTERAUGLI, baseline/_ref/bin/<tool>, PATH
and prints "T3: not run (<tool> not found)" when the decoder is missing.  With the tools present it checks, per case:
  * djxl decodes the file (exit code 0) to the input's dimensions;
  * bpp within 0.5 % of `cjxl --distance=D --effort=E` on the same image      (north star);
  * SSIMULACRA2 and Butteraugli (3-norm) within 0.1 of the reference encoder's (north star);
  * PSNR of the djxl decode equals the PSNR jxlb200_stats reports for its own reconstruction within 0.05 dB.
Needs a GPU for the encodes (no CPU fallback).  Writes gpurun_out/t3_conformance.json."""
from __future__ import annotations

import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "jpeg-xl-lossy-image-compression-thesis_b200"
# (w, h, distance, effort, proposal, flags); flags 48 = JXLB200_FLAG_GABORISH | JXLB200_FLAG_CFL, the loop filter and the colour
# correlation map that cjxl's defaults use (with this repo's own sharpening kernel and fit: the bpp / metric deltas of those
# cases measure exactly that difference)
CASES = [(512, 512, 1.0, 7, 0, 0), (1920, 1080, 0.5, 7, 3, 0), (1920, 1080, 3.0, 7, 3, 0), (3840, 2160, 1.0, 7, 1, 0),
         (200, 120, 8.0, 5, 2, 0), (512, 512, 1.0, 7, 0, 16), (512, 512, 1.0, 7, 0, 32), (1920, 1080, 1.0, 7, 0, 48)]

# What this harness is the first chance to check (none of it can be pinned offline; DESIGN.md sections 2 and 3):
#  * every [UPSTREAM] table and constant of the U-rows: quant weights, coefficient orders, context tables, AQ constants,
#    the entropy-cost multipliers of the search, the default Gaborish weights the decoder applies;
#  * the numerics the diffs leave open: contraction of `acc += a * b` in the proposals' scalar loops (fused here),
#    `sqrt` / `0.3 * sqrt` evaluated in double, the planes' pitch and padded height used as `src_stride` / `src_ysize`;
#  * the defined-behaviour choices at row / column 0 of the modified Laplacian and in TryMergeAcs.


def find_tool(name):
    env = os.environ.get("JXLB200_" + name.upper().replace("_MAIN", ""))
    for cand in (env, os.path.join(ROOT, "baseline", "_ref", "bin", name), shutil.which(name)):
        if cand and os.path.isfile(cand) and os.access(cand, os.X_OK):
            return cand
    return None


def write_ppm(path, img):
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img).tobytes())


def read_ppm(path):
    data = open(path, "rb").read()
    parts = data.split(None, 4)
    w, h = int(parts[1]), int(parts[2])
    return np.frombuffer(parts[4], dtype=np.uint8, count=w * h * 3).reshape(h, w, 3)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def metric(tool, orig, dist):
    """first number the tool prints (ssimulacra2: stdout; butteraugli_main: 3-norm line), as metrics.rs parses them"""
    r = subprocess.run([tool, orig, dist], capture_output=True, text=True)
    for tok in (r.stdout + " " + r.stderr).replace(":", " ").split():
        try:
            return float(tok)
        except ValueError:
            continue
    return None


def main():
    djxl, cjxl = find_tool("djxl"), find_tool("cjxl")
    ssim, butter = find_tool("ssimulacra2"), find_tool("butteraugli_main")
    out_path = os.path.join(ROOT, "gpurun_out", "t3_conformance.json")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    if not djxl:
        msg = "T3: not run (djxl not found: set $JXLB200_DJXL or put libjxl's tools under baseline/_ref/bin)"
        print(msg)
        json.dump({"t3": "not run", "reason": "djxl not found"}, open(out_path, "w"))
        return 0
    pkg = importlib.import_module(PKG)
    results, ok = [], True
    with pkg.Encoder(0) as enc, tempfile.TemporaryDirectory() as tmp:
        for i, (w, h, d, e, prop, fl) in enumerate(CASES):
            img = pkg.synth_image(w, h, 500 + i)
            data, st = enc.encode(img, d, e, prop, pkg.FLAG_QUALITY | fl)
            src, jxl, dec = (os.path.join(tmp, f"{i}.{x}") for x in ("ppm", "jxl", "dec.ppm"))
            write_ppm(src, img)
            open(jxl, "wb").write(data)
            r = subprocess.run([djxl, jxl, dec], capture_output=True, text=True)
            rec = {"case": [w, h, d, e, prop, fl], "bytes": len(data), "bpp": st.bpp, "djxl_rc": r.returncode}
            if r.returncode != 0:
                rec["djxl_stderr"] = r.stderr[-400:]
                ok = False
            else:
                out = read_ppm(dec)
                rec["decoded_shape_ok"] = out.shape == img.shape
                rec["psnr_djxl"] = psnr(img, out)
                rec["psnr_stats"] = st.psnr
                rec["psnr_ok"] = abs(rec["psnr_djxl"] - st.psnr) < 0.05
                ok = ok and rec["decoded_shape_ok"] and rec["psnr_ok"]
                if cjxl and prop == 0:            # the proposals need a patched libjxl build: compare the unpatched case
                    ref_jxl, ref_dec = os.path.join(tmp, f"{i}.ref.jxl"), os.path.join(tmp, f"{i}.ref.ppm")
                    subprocess.run([cjxl, src, ref_jxl, f"--distance={d}", f"--effort={e}"], check=True, capture_output=True)
                    subprocess.run([djxl, ref_jxl, ref_dec], check=True, capture_output=True)
                    ref_bpp = 8.0 * os.path.getsize(ref_jxl) / (w * h)
                    rec["ref_bpp"] = ref_bpp
                    rec["bpp_within_0.5pct"] = abs(st.bpp - ref_bpp) <= 0.005 * ref_bpp
                    for name, tool in (("ssimulacra2", ssim), ("butteraugli", butter)):
                        if tool:
                            a, b = metric(tool, src, dec), metric(tool, src, ref_dec)
                            rec[name] = {"ours": a, "reference": b, "within_0.1": a is not None and b is not None and abs(a - b) <= 0.1}
            results.append(rec)
            print(json.dumps(rec))
    json.dump({"t3": "ran", "decodes": ok, "results": results}, open(out_path, "w"), indent=1)
    print("T3:", "all codestreams decoded with the reference djxl" if ok else "FAILED (see results)")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
