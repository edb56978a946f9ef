"""The C-ABI library loads and exports every symbol include/jxlb200.h declares (no GPU calls)."""
import ctypes
import os
import re

import pytest


def test_exports_match_header(pkg):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "jxlb200.h")).read()
    names = set(re.findall(r"\b(jxlb200_[a-z_]+)\s*\(", header))
    assert {"jxlb200_create", "jxlb200_destroy", "jxlb200_encode", "jxlb200_encode_device", "jxlb200_fetch",
            "jxlb200_encode_batch", "jxlb200_free", "jxlb200_dump", "jxlb200_last_error", "jxlb200_dims",
            "jxlb200_abi_version"} <= names
    lib = ctypes.CDLL(pkg.library_path())
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/jxlb200.h but not exported"
    assert pkg.load_library().jxlb200_abi_version() == 3


def test_dims_helper(pkg, oracle):
    for w, h in ((512, 512), (3840, 2160), (7680, 4320), (1920, 1080), (1, 1), (257, 9)):
        assert pkg.frame_dims(w, h) == oracle.dims(w, h)
    d = pkg.frame_dims(3840, 2160)
    assert (d["bxs"] * d["bys"], d["txs"] * d["tys"], d["num_groups"], d["num_dc_groups"]) == (129600, 2040, 135, 4)


def test_no_gpu_means_no_encoder(pkg):
    import torch
    if torch.cuda.is_available():
        return
    try:
        pkg.Encoder(0)
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Encoder() must fail loudly without a GPU")


def _build_c_client(tmp_path):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "jpeg-xl-lossy-image-compression-thesis_b200")
    exe = os.path.join(str(tmp_path), "cabi_client")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cabi_client.c"), "-o", exe, "-L", libdir, "-ljxlb200",
                           "-Wl,-rpath," + libdir])
    return exe


def test_header_is_plain_c_and_links(tmp_path):
    """include/jxlb200.h compiles as pedantic C99 and a C program links against libjxlb200.so; without a GPU the client
    sees jxlb200_create() == NULL (no CPU fallback) and exits cleanly."""
    import subprocess
    import torch
    exe = _build_c_client(tmp_path)
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "returned NULL" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_client_encodes(tmp_path):
    """The same C program on a B200: encode, stats, error contract, release."""
    import subprocess
    r = subprocess.run([_build_c_client(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok:"), r.stdout + r.stderr
