"""ctypes wrapper of the CPU oracle (oracle/_build/libjxo.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke(); never by the product package."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libjxo.so")

STAGES = {
    1: ("xyb", np.float32), 2: ("qf_float", np.float32), 3: ("mask1x1", np.float32), 4: ("homog", np.float32),
    5: ("acs", np.uint8), 6: ("raw_qf", np.int32), 7: ("quant_params", np.int32), 8: ("coeffs", np.int16),
    9: ("dc_quant", np.int16), 10: ("nzeros", np.uint8), 11: ("tokens", np.uint32), 12: ("histograms", np.uint32),
    13: ("context_map", np.uint8), 14: ("group_streams", np.uint8), 15: ("codestream", np.uint8),
    16: ("mask", np.float32), 17: ("cmap", np.int8), 18: ("token_offsets", np.uint32),
    19: ("group_offsets", np.uint32), 20: ("acs_entropy", np.float32), 21: ("num_clusters", np.int32),
}
STAGE_ID = {v[0]: k for k, v in STAGES.items()}


class Params(ctypes.Structure):
    _fields_ = [("distance", ctypes.c_float), ("effort", ctypes.c_uint32), ("proposal", ctypes.c_uint32),
                ("flags", ctypes.c_uint32)]


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        vp = ctypes.c_void_p
        lib.jxo_encode.restype = vp
        lib.jxo_encode.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(Params)]
        lib.jxo_encode_forced.restype = vp
        lib.jxo_encode_forced.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(Params), vp, ctypes.c_size_t]
        lib.jxo_decode.restype = vp
        lib.jxo_decode.argtypes = [vp, ctypes.c_size_t]
        lib.jxo_reconstruct.restype = ctypes.c_int
        lib.jxo_reconstruct.argtypes = [vp, vp]
        lib.jxo_sse.restype = ctypes.c_int
        lib.jxo_sse.argtypes = [vp, vp, ctypes.c_size_t, vp]
        lib.jxo_error.restype = ctypes.c_char_p
        lib.jxo_error.argtypes = [vp]
        lib.jxo_free.argtypes = [vp]
        lib.jxo_dump.restype = ctypes.c_size_t
        lib.jxo_dump.argtypes = [vp, ctypes.c_int, vp, ctypes.c_size_t]
        lib.jxo_dims.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
        fp = ctypes.POINTER(ctypes.c_float)
        lib.jxo_homogeneity_indices.argtypes = [vp, vp, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                                ctypes.c_size_t, ctypes.c_float, fp]
        lib.jxo_homogeneity.restype = ctypes.c_float
        lib.jxo_homogeneity.argtypes = [vp, vp, vp] + [ctypes.c_size_t] * 8 + [ctypes.c_float]
        lib.jxo_homogeneity_partition.restype = ctypes.c_int
        lib.jxo_homogeneity_partition.argtypes = [ctypes.c_float] * 4
        lib.jxo_factored_entropy.restype = ctypes.c_float
        lib.jxo_factored_entropy.argtypes = [ctypes.c_float] * 4
        lib.jxo_dct2d.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        lib.jxo_idct2d.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int]
        lib.jxo_transform.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp]
        lib.jxo_inverse_transform.argtypes = [ctypes.c_int, vp, vp, ctypes.c_int]
        lib.jxo_quant_weights.restype = ctypes.c_int
        lib.jxo_quant_weights.argtypes = [ctypes.c_int, vp, ctypes.c_size_t]
        lib.jxo_natural_order.restype = ctypes.c_int
        lib.jxo_natural_order.argtypes = [ctypes.c_int, vp, ctypes.c_size_t]
        lib.jxo_cbrt.restype = ctypes.c_float
        lib.jxo_cbrt.argtypes = [ctypes.c_float]
        lib.jxo_srgb_lut.argtypes = [vp]

    def encode(self, image, distance=1.0, effort=7, proposal=0, flags=0):
        image = np.ascontiguousarray(image)
        p = Params(distance, effort, proposal, flags)
        h = self.lib.jxo_encode(image.ctypes.data, image.shape[1], image.shape[0], image.strides[0], ctypes.byref(p))
        return Frame(self, h)

    def encode_forced(self, image, acs, distance=1.0, effort=7, proposal=0, flags=0):
        """encode with the given AC-strategy map (uint8, raw strategy | 0x80 on first blocks) instead of the search"""
        image = np.ascontiguousarray(image)
        acs = np.ascontiguousarray(acs, dtype=np.uint8)
        p = Params(distance, effort, proposal, flags)
        h = self.lib.jxo_encode_forced(image.ctypes.data, image.shape[1], image.shape[0], image.strides[0], ctypes.byref(p),
                                       acs.ctypes.data, acs.size)
        return Frame(self, h)

    def decode(self, codestream):
        """self-decoder: codestream bytes -> Frame holding dc_quant / acs / raw_qf / coeffs / nzeros"""
        buf = np.frombuffer(bytes(codestream), dtype=np.uint8)
        return Frame(self, self.lib.jxo_decode(buf.ctypes.data, buf.size))

    def decode_pixels(self, codestream, w, h):
        """self-decoder to pixels: returns (h, w, 3) uint8 sRGB, or None when the stream is rejected"""
        fr = self.decode(codestream)
        if fr.error:
            return None
        out = np.zeros((h, w, 3), dtype=np.uint8)
        ok = self.lib.jxo_reconstruct(fr.h, out.ctypes.data)
        return out if ok else None

    def dims(self, w, h):
        d = (ctypes.c_int32 * 16)()
        self.lib.jxo_dims(w, h, d)
        names = ("xsize", "ysize", "xs_pad", "ys_pad", "pitch", "bxs", "bys", "gxs", "gys", "num_groups", "dgxs",
                 "dgys", "num_dc_groups", "txs", "tys")
        return dict(zip(names, list(d)))

    def homogeneity_indices(self, x, y, b, px, py, d):
        x, y, b = (np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, b))
        out = (ctypes.c_float * 3)()
        self.lib.jxo_homogeneity_indices(x.ctypes.data, y.ctypes.data, b.ctypes.data, x.shape[1], x.shape[0], px, py,
                                         d, out)
        return np.array(list(out), dtype=np.float32)

    def homogeneity(self, x, y, b, px, py, xs, ys, bx, by, d):
        x, y, b = (np.ascontiguousarray(a, dtype=np.float32) for a in (x, y, b))
        return self.lib.jxo_homogeneity(x.ctypes.data, y.ctypes.data, b.ctypes.data, x.shape[1], x.shape[0], px, py,
                                        xs, ys, bx, by, d)

    def transform(self, strategy, px):
        px = np.ascontiguousarray(px, dtype=np.float32)
        out = np.zeros(px.size, dtype=np.float32)
        self.lib.jxo_transform(strategy, px.ctypes.data, px.shape[1], out.ctypes.data)
        return out

    def inverse_transform(self, strategy, coef, rows, cols):
        coef = np.ascontiguousarray(coef, dtype=np.float32)
        out = np.zeros((rows, cols), dtype=np.float32)
        self.lib.jxo_inverse_transform(strategy, coef.ctypes.data, out.ctypes.data, cols)
        return out

    def quant_weights(self, kind):
        n = self.lib.jxo_quant_weights(kind, None, 0)
        out = np.zeros(3 * n, dtype=np.float32)
        self.lib.jxo_quant_weights(kind, out.ctypes.data, out.size)
        return out.reshape(3, n)

    def natural_order(self, strategy):
        n = self.lib.jxo_natural_order(strategy, None, 0)
        out = np.zeros(n, dtype=np.uint16)
        self.lib.jxo_natural_order(strategy, out.ctypes.data, n)
        return out


class Frame:
    def __init__(self, oracle, handle):
        self.o, self.h = oracle, handle
        err = oracle.lib.jxo_error(handle)
        self.error = err.decode() if err else ""

    def sse(self, image):
        """per-channel sum of squared errors of this frame's reconstruction against `image` (h, w, 3) uint8"""
        image = np.ascontiguousarray(image)
        out = np.zeros(3, dtype=np.uint64)
        ok = self.o.lib.jxo_sse(self.h, image.ctypes.data, image.strides[0], out.ctypes.data)
        return out if ok else None

    def dump(self, stage):
        sid = STAGE_ID[stage] if isinstance(stage, str) else int(stage)
        n = self.o.lib.jxo_dump(self.h, sid, None, 0)
        buf = np.empty(n, dtype=np.uint8)
        if n:
            self.o.lib.jxo_dump(self.h, sid, buf.ctypes.data, n)
        return buf.view(STAGES[sid][1])

    def close(self):
        if self.h:
            self.o.lib.jxo_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


_ORACLE = None


def load(rebuild=True):
    global _ORACLE
    if _ORACLE is None:
        if rebuild or not os.path.exists(SO):
            build()
        _ORACLE = Oracle(ctypes.CDLL(SO))
    return _ORACLE
