"""ctypes binding of include/jxlb200.h and the host-side mirror of the reference's
encoder interface.

Reference seam (paths relative to the reference root):
  * ``DockerManager::execute_cjxl(input_file, output_file, distance: f64, effort: u32)``
    benchmark-jpegxl/src/docker_manager.rs:100-137 — `cjxl in out --distance=D --effort=E`
  * ``compress_from_png`` old_test_jxl.py:450-473 — `cjxl -d D -e E in out`
  * failure => log and "skip" (benchmark-jpegxl/src/benchmark.rs:661-677)
  * proposal = which proposals/*.diff was applied before rebuilding libjxl
    (benchmark.rs:460-484); here a runtime enum.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

PROPOSAL_NONE, PROPOSAL_PARTITIONING, PROPOSAL_FACTORED_ENTROPY, PROPOSAL_COMBINED = 0, 1, 2, 3
FLAG_FIXED_DCT8, FLAG_UNIFORM_QF, FLAG_QUALITY, FLAG_FORCED_ACS, FLAG_GABORISH, FLAG_CFL = 1, 2, 4, 8, 16, 32

# stage id -> (name, numpy dtype)   (JXLB200_STAGE_* in include/jxlb200.h)
STAGES = {
    1: ("xyb", np.float32), 2: ("qf_float", np.float32), 3: ("mask1x1", np.float32), 4: ("homog", np.float32),
    5: ("acs", np.uint8), 6: ("raw_qf", np.int32), 7: ("quant_params", np.int32), 8: ("coeffs", np.int16),
    9: ("dc_quant", np.int16), 10: ("nzeros", np.uint8), 11: ("tokens", np.uint32), 12: ("histograms", np.uint32),
    13: ("context_map", np.uint8), 14: ("group_streams", np.uint8), 15: ("codestream", np.uint8),
    16: ("mask", np.float32), 17: ("cmap", np.int8), 18: ("token_offsets", np.uint32),
    19: ("group_offsets", np.uint32), 20: ("acs_entropy", np.float32), 21: ("num_clusters", np.int32),
}
STAGE_ID = {v[0]: k for k, v in STAGES.items()}


class EncodeError(RuntimeError):
    """The C ABI returned a negative code (the harness maps this to "skip")."""


class _Image(ctypes.Structure):
    _fields_ = [("pixels", ctypes.c_void_p), ("width", ctypes.c_uint32), ("height", ctypes.c_uint32),
                ("stride", ctypes.c_size_t)]


class _Params(ctypes.Structure):
    _fields_ = [("distance", ctypes.c_float), ("effort", ctypes.c_uint32), ("proposal", ctypes.c_uint32),
                ("flags", ctypes.c_uint32)]


class _Stats(ctypes.Structure):
    _fields_ = [("codestream_bytes", ctypes.c_uint64), ("bpp", ctypes.c_double), ("width", ctypes.c_uint32),
                ("height", ctypes.c_uint32), ("num_groups", ctypes.c_uint32), ("num_dc_groups", ctypes.c_uint32),
                ("global_scale", ctypes.c_uint32), ("quant_dc", ctypes.c_uint32), ("num_tokens", ctypes.c_uint64),
                ("num_clusters", ctypes.c_uint32), ("acs_histogram", ctypes.c_uint32 * 27),
                ("stage_ms", ctypes.c_float * 16), ("total_ms", ctypes.c_float), ("kernel_launches", ctypes.c_uint32),
                ("quality_valid", ctypes.c_uint32), ("sse", ctypes.c_uint64 * 3), ("psnr", ctypes.c_double)]


@dataclass
class Stats:
    codestream_bytes: int = 0
    bpp: float = 0.0
    width: int = 0
    height: int = 0
    num_groups: int = 0
    num_dc_groups: int = 0
    global_scale: int = 0
    quant_dc: int = 0
    num_tokens: int = 0
    num_clusters: int = 0
    acs_histogram: list = field(default_factory=list)
    stage_ms: list = field(default_factory=list)
    total_ms: float = 0.0
    kernel_launches: int = 0
    sse: Optional[list] = None      # FLAG_QUALITY: per-channel squared error of the decoded image against the input
    psnr: Optional[float] = None    # FLAG_QUALITY: PSNR over all samples, dB

    @classmethod
    def _from_c(cls, s: _Stats) -> "Stats":
        return cls(int(s.codestream_bytes), float(s.bpp), int(s.width), int(s.height), int(s.num_groups),
                   int(s.num_dc_groups), int(s.global_scale), int(s.quant_dc), int(s.num_tokens),
                   int(s.num_clusters), list(s.acs_histogram), list(s.stage_ms), float(s.total_ms), int(s.kernel_launches),
                   [int(v) for v in s.sse] if s.quality_valid else None, float(s.psnr) if s.quality_valid else None)


def library_path() -> str:
    """In-tree libjxlb200.so ($JXLB200_LIB overrides it, e.g. to compare kernel variants)."""
    return os.environ.get("JXLB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libjxlb200.so")


_LIB = None


def load_library() -> ctypes.CDLL:
    """Loads libjxlb200.so and declares every symbol of include/jxlb200.h.  Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    vp, u8pp, szp = ctypes.c_void_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8)), ctypes.POINTER(ctypes.c_size_t)
    lib.jxlb200_abi_version.restype = ctypes.c_int
    lib.jxlb200_create.restype = vp
    lib.jxlb200_create.argtypes = [ctypes.c_int]
    lib.jxlb200_destroy.argtypes = [vp]
    lib.jxlb200_last_error.restype = ctypes.c_char_p
    lib.jxlb200_last_error.argtypes = [vp]
    lib.jxlb200_encode.restype = ctypes.c_int
    lib.jxlb200_encode.argtypes = [vp, ctypes.POINTER(_Image), ctypes.POINTER(_Params), u8pp, szp,
                                   ctypes.POINTER(_Stats)]
    lib.jxlb200_encode_device.restype = ctypes.c_int
    lib.jxlb200_encode_device.argtypes = [vp, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_size_t,
                                          ctypes.POINTER(_Params), ctypes.POINTER(_Stats)]
    lib.jxlb200_fetch.restype = ctypes.c_int
    lib.jxlb200_fetch.argtypes = [vp, u8pp, szp]
    lib.jxlb200_encode_batch.restype = ctypes.c_int
    lib.jxlb200_encode_batch.argtypes = [vp, ctypes.POINTER(_Image), ctypes.POINTER(_Params), ctypes.c_size_t,
                                         u8pp, szp, ctypes.POINTER(_Stats)]
    lib.jxlb200_encode_batch_device.restype = ctypes.c_int
    lib.jxlb200_encode_batch_device.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint32),
                                                ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_size_t),
                                                ctypes.POINTER(_Params), ctypes.c_size_t, ctypes.POINTER(_Stats),
                                                ctypes.POINTER(ctypes.c_float)]
    lib.jxlb200_set_pipelines.restype = ctypes.c_int
    lib.jxlb200_set_pipelines.argtypes = [vp, ctypes.c_int]
    lib.jxlb200_free.argtypes = [vp]
    lib.jxlb200_dump.restype = ctypes.c_int64
    lib.jxlb200_dump.argtypes = [vp, ctypes.c_int, vp, ctypes.c_size_t]
    lib.jxlb200_dims.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int32)]
    lib.jxlb200_debug_set_strategy_map.restype = ctypes.c_int
    lib.jxlb200_debug_set_strategy_map.argtypes = [vp, vp, ctypes.c_uint32, ctypes.c_uint32]
    lib.jxlb200_debug_homogeneity.restype = ctypes.c_int
    lib.jxlb200_debug_homogeneity.argtypes = [vp, vp, vp, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float, vp]
    _LIB = lib
    return lib


_DIM_NAMES = ("xsize", "ysize", "xs_pad", "ys_pad", "pitch", "bxs", "bys", "gxs", "gys", "num_groups", "dgxs",
              "dgys", "num_dc_groups", "txs", "tys")


def frame_dims(width: int, height: int) -> dict:
    d = (ctypes.c_int32 * 16)()
    load_library().jxlb200_dims(width, height, d)
    return dict(zip(_DIM_NAMES, list(d)))


class Encoder:
    """One encoder context = one CUDA stream + device arenas; owned by one thread (the
    reference gives each worker thread its own DockerManager, benchmark.rs:97-103)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._ctx = self._lib.jxlb200_create(device)
        if not self._ctx:
            raise RuntimeError("jxlb200_create failed: no usable sm_100 CUDA device (there is no CPU fallback)")

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.jxlb200_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _err(self) -> str:
        return self._lib.jxlb200_last_error(self._ctx).decode("utf-8", "replace")

    def encode(self, image: np.ndarray, distance: float = 1.0, effort: int = 7, proposal: int = PROPOSAL_NONE,
               flags: int = 0) -> tuple[bytes, Stats]:
        """``execute_cjxl`` on an in-memory (h, w, 3) uint8 image -> (codestream bytes, Stats)."""
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise EncodeError("image must be (h, w, 3) uint8")
        if image.strides[2] != 1 or image.strides[1] != 3:
            image = np.ascontiguousarray(image)
        img = _Image(image.ctypes.data, image.shape[1], image.shape[0], image.strides[0])
        par = _Params(distance, effort, proposal, flags)
        out = ctypes.POINTER(ctypes.c_uint8)()
        n = ctypes.c_size_t(0)
        st = _Stats()
        rc = self._lib.jxlb200_encode(self._ctx, ctypes.byref(img), ctypes.byref(par), ctypes.byref(out),
                                      ctypes.byref(n), ctypes.byref(st))
        if rc != 0:
            raise EncodeError(self._err())
        data = ctypes.string_at(out, n.value)
        self._lib.jxlb200_free(out)
        return data, Stats._from_c(st)

    def encode_device(self, d_ptr: int, width: int, height: int, stride: int, distance: float = 1.0,
                      effort: int = 7, proposal: int = PROPOSAL_NONE, flags: int = 0) -> Stats:
        """Encode an RGB8 image already resident in device memory (``d_ptr`` = device address)."""
        par = _Params(distance, effort, proposal, flags)
        st = _Stats()
        rc = self._lib.jxlb200_encode_device(self._ctx, d_ptr, width, height, stride, ctypes.byref(par),
                                             ctypes.byref(st))
        if rc != 0:
            raise EncodeError(self._err())
        return Stats._from_c(st)

    def set_pipelines(self, n: int) -> None:
        """How many images the batch calls keep in flight (one CUDA stream + arenas each)."""
        if self._lib.jxlb200_set_pipelines(self._ctx, n) != 0:
            raise EncodeError(self._err())

    def encode_batch(self, images, distance=1.0, effort: int = 7, proposal: int = PROPOSAL_NONE, flags: int = 0):
        """The per-image loop of ``JXLCompressionBenchmark::run`` (benchmark.rs:637-660) as one call:
        ``images`` is a list of (h, w, 3) uint8 arrays (pinned memory is copied without staging);
        ``distance`` may be a scalar or one value per image.  Returns ([bytes], [Stats])."""
        n = len(images)
        dist = list(distance) if hasattr(distance, "__len__") else [distance] * n
        imgs = (_Image * n)()
        pars = (_Params * n)()
        keep = []
        for i, im in enumerate(images):
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise EncodeError("image must be (h, w, 3) uint8")
            if im.strides[2] != 1 or im.strides[1] != 3:
                im = np.ascontiguousarray(im)
            keep.append(im)
            imgs[i] = _Image(im.ctypes.data, im.shape[1], im.shape[0], im.strides[0])
            pars[i] = _Params(dist[i], effort, proposal, flags)
        outs = (ctypes.POINTER(ctypes.c_uint8) * n)()
        lens = (ctypes.c_size_t * n)()
        sts = (_Stats * n)()
        rc = self._lib.jxlb200_encode_batch(self._ctx, imgs, pars, n, outs, lens, sts)
        if rc != 0:
            raise EncodeError(self._err())
        data = []
        for i in range(n):
            data.append(ctypes.string_at(outs[i], lens[i]))
            self._lib.jxlb200_free(outs[i])
        return data, [Stats._from_c(s) for s in sts]

    def encode_batch_device(self, d_ptrs, width: int, height: int, stride: int, distance=1.0, effort: int = 7,
                            proposal: int = PROPOSAL_NONE, flags: int = 0):
        """Batch over RGB8 images resident in device memory (``d_ptrs`` = device addresses).  Returns
        ([Stats], device_ms) where device_ms is the CUDA-event time of the whole batch."""
        n = len(d_ptrs)
        dist = list(distance) if hasattr(distance, "__len__") else [distance] * n
        ptrs = (ctypes.c_void_p * n)(*d_ptrs)
        ws = (ctypes.c_uint32 * n)(*([width] * n))
        hs = (ctypes.c_uint32 * n)(*([height] * n))
        ss = (ctypes.c_size_t * n)(*([stride] * n))
        pars = (_Params * n)(*[_Params(dist[i], effort, proposal, flags) for i in range(n)])
        sts = (_Stats * n)()
        ms = ctypes.c_float(0.0)
        rc = self._lib.jxlb200_encode_batch_device(self._ctx, ptrs, ws, hs, ss, pars, n, sts, ctypes.byref(ms))
        if rc != 0:
            raise EncodeError(self._err())
        return [Stats._from_c(s) for s in sts], float(ms.value)

    def fetch(self) -> bytes:
        out = ctypes.POINTER(ctypes.c_uint8)()
        n = ctypes.c_size_t(0)
        if self._lib.jxlb200_fetch(self._ctx, ctypes.byref(out), ctypes.byref(n)) != 0:
            raise EncodeError(self._err())
        data = ctypes.string_at(out, n.value)
        self._lib.jxlb200_free(out)
        return data

    def set_strategy_map(self, acs) -> None:
        """AC-strategy map (bys x bxs uint8, raw strategy | 0x80 on first blocks) that encodes with FLAG_FORCED_ACS use
        instead of the search: jxlb200_debug_set_strategy_map."""
        acs = np.ascontiguousarray(acs, dtype=np.uint8)
        if self._lib.jxlb200_debug_set_strategy_map(self._ctx, acs.ctypes.data, acs.shape[1], acs.shape[0]) != 0:
            raise EncodeError(self._err())

    def homogeneity_map(self, x, y, b, distance: float) -> np.ndarray:
        """r_h, r_v, r_d of every 8x8 block of caller-supplied XYB planes (rows x stride float32 arrays; the stride is the
        reference's `src_stride` bound, the row count its `src_ysize`): jxlb200_debug_homogeneity."""
        x, y, b = (np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, b))
        rows, stride = y.shape
        out = np.zeros((rows // 8, stride // 8, 3), dtype=np.float32)
        rc = self._lib.jxlb200_debug_homogeneity(self._ctx, x.ctypes.data, y.ctypes.data, b.ctypes.data, stride, rows,
                                                 float(distance), out.ctypes.data)
        if rc != 0:
            raise EncodeError(self._err())
        return out

    def dump(self, stage) -> np.ndarray:
        """Intermediate of the last encode as a flat numpy array (parity taps)."""
        sid = STAGE_ID[stage] if isinstance(stage, str) else int(stage)
        n = self._lib.jxlb200_dump(self._ctx, sid, None, 0)
        if n < 0:
            raise EncodeError(self._err())
        buf = np.empty(n, dtype=np.uint8)
        if n:
            self._lib.jxlb200_dump(self._ctx, sid, buf.ctypes.data, n)
        return buf.view(STAGES[sid][1])

    # -- the reference's file-level interface ---------------------------------------------
    def execute_cjxl(self, input_file: str, output_file: str, distance: float, effort: int,
                     proposal: int = PROPOSAL_NONE) -> tuple[Optional[str], Optional[str]]:
        """File-level twin of ``DockerManager::execute_cjxl``: reads a binary PPM (P6) or .npy image,
        writes ``output_file``; returns ``(stdout, None)`` on success or ``(None, stderr)`` on failure,
        so the caller can keep the reference's skip-and-continue behaviour."""
        try:
            img = _read_image(input_file)
            data, st = self.encode(img, distance, effort, proposal)
            os.makedirs(os.path.dirname(os.path.abspath(output_file)) or ".", exist_ok=True)
            with open(output_file, "wb") as f:
                f.write(data)
            return (f"Compressed to {len(data)} bytes ({st.bpp:.3f} bpp).", None)
        except (EncodeError, OSError, ValueError) as e:
            return (None, str(e))


def _read_image(path: str) -> np.ndarray:
    """The harness hands PNG paths to cjxl (benchmark.rs:654-660); binary PPM and .npy are accepted as well."""
    if path.endswith(".npy"):
        return np.load(path)
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] == b"\x89PNG\r\n\x1a\n":
        try:
            from PIL import Image
        except ImportError as e:
            raise ValueError("PNG input needs Pillow") from e
        import io
        with Image.open(io.BytesIO(data)) as im:
            return np.ascontiguousarray(np.asarray(im.convert("RGB"), dtype=np.uint8))
    if data[:2] != b"P6":
        raise ValueError("only PNG, binary PPM (P6) and .npy inputs are supported")
    parts, pos = [], 2
    while len(parts) < 3:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        parts.append(int(data[pos:end]))
        pos = end
    w, h, maxv = parts
    if maxv != 255:
        raise ValueError("only 8-bit PPM is supported")
    return np.frombuffer(data, dtype=np.uint8, count=w * h * 3, offset=pos + 1).reshape(h, w, 3)
