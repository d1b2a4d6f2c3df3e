"""ctypes binding of the C-ABI in include/jpgenc_b200.h (the drop-in boundary of the encode path).

This is the binding a maintainer of the reference would write for a Python harness; the tests and bench.py call
the CUDA path exclusively through it.  There is no fallback of any kind: if the shared library is missing or no
B200 is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libjpgenc_b200.so"

OK = 0
ERR_CUDA, ERR_ARG, ERR_NO_DEVICE, ERR_FORMAT, ERR_IO, ERR_CAPACITY, ERR_NOMEM = -1, -2, -3, -4, -5, -6, -7

u8p = C.POINTER(C.c_uint8)
i16p = C.POINTER(C.c_int16)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


class HuffTable(C.Structure):
    _fields_ = [
        ("code_msb", C.c_uint32 * 256),
        ("length", C.c_uint8 * 256),
        ("counts", C.c_uint8 * 16),
        ("symbols", C.c_uint8 * 256),
        ("nsymbols", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("real_w", C.c_uint32), ("real_h", C.c_uint32), ("mcu_w", C.c_uint32), ("mcu_h", C.c_uint32),
        ("n_blocks", C.c_uint64), ("refined_blocks", C.c_uint64), ("scan_bits", C.c_uint64),
        ("scan_bytes", C.c_uint64), ("stuffed_ff", C.c_uint64),
        ("ms_k1", C.c_float), ("ms_forward", C.c_float), ("ms_stats", C.c_float), ("ms_entropy", C.c_float),
        ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
        ("sum_ms_k1", C.c_double), ("sum_ms_forward", C.c_double), ("sum_ms_stats", C.c_double), ("sum_ms_entropy", C.c_double),
        ("timed_encodes", C.c_uint64),
    ]


# every symbol the header declares: name -> (restype, argtypes)
_SIGNATURES = {
    "jpgenc_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "jpgenc_destroy": (None, [C.c_void_p]),
    "jpgenc_last_error": (C.c_char_p, [C.c_void_p]),
    "jpgenc_stream": (C.c_void_p, [C.c_void_p]),
    "jpgenc_synchronize": (C.c_int, [C.c_void_p]),
    "jpgenc_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "jpgenc_set_qtables": (C.c_int, [C.c_void_p, u8p, u8p]),
    "jpgenc_set_dct_constants": (C.c_int, [C.c_void_p, f64p, f64p]),
    "jpgenc_upload_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "jpgenc_bind_device_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "jpgenc_color_dct_quant": (C.c_int, [C.c_void_p]),
    "jpgenc_get_coefficients": (C.c_int, [C.c_void_p, i16p]),
    "jpgenc_set_coefficients": (C.c_int, [C.c_void_p, i32p, i32p, i32p, C.c_uint32, C.c_uint32]),
    "jpgenc_set_coefficients_mcu": (C.c_int, [C.c_void_p, i16p, C.c_uint32, C.c_uint32]),
    "jpgenc_symbol_stats": (C.c_int, [C.c_void_p, u32p, u64p]),
    "jpgenc_set_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "jpgenc_build_huffman": (C.c_int, [u32p, u64p, C.POINTER(HuffTable)]),
    "jpgenc_build_huffman_containers": (C.c_int, [u32p, u64p, C.POINTER(HuffTable)]),
    "jpgenc_build_huffman_arrays": (C.c_int, [u32p, u64p, C.POINTER(HuffTable)]),
    "jpgenc_build_huffman_device": (C.c_int, [C.c_void_p, C.c_uint32, u32p, u64p, C.POINTER(HuffTable)]),
    "jpgenc_entropy_encode": (C.c_int, [C.c_void_p, C.POINTER(HuffTable), u64p]),
    "jpgenc_download_scan": (C.c_int, [C.c_void_p, u8p, C.c_uint64]),
    "jpgenc_ppm_info": (C.c_int, [C.c_char_p, C.c_size_t, u32p, u32p, u32p, C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "jpgenc_ppm_samples": (C.c_int, [C.c_char_p, C.c_size_t, u8p]),
    "jpgenc_write_headers": (C.c_size_t, [C.c_uint32, C.c_uint32, u8p, u8p, C.POINTER(HuffTable), u8p]),
    "jpgenc_encode_bound": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, u64p]),
    "jpgenc_assemble_last": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, u64p]),
    "jpgenc_encode_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, u64p]),
    "jpgenc_encode_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                       C.c_void_p, C.c_uint64, u64p]),
    "jpgenc_encode_ppm_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "jpgenc_debug_counter": (C.c_int, [C.c_void_p, C.c_int, u32p]),
    "jpgenc_stage_subsample_dims": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, u32p, u32p]),
    "jpgenc_stage_subsample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]),
    "jpgenc_stage_dct": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]),
    "jpgenc_dct_quant_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, u8p, u64p]),
    "jpgenc_bind_host_to_device_numa": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "jpgenc_dev_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "jpgenc_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jpgenc_host_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "jpgenc_host_free_pinned": (C.c_int, [C.c_void_p]),
    "jpgenc_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "jpgenc_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "jpgenc_synth_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "jpgenc_synth_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "jpgenc_flush_l2": (C.c_int, [C.c_void_p]),
    "jpgenc_timer_begin": (C.c_int, [C.c_void_p]),
    "jpgenc_timer_end": (C.c_int, [C.c_void_p, f32p]),
    "jpgenc_launch_count": (C.c_uint64, [C.c_void_p]),
    "jpgenc_encode_frames_device": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.POINTER(C.c_void_p), u64p, u64p]),
    "jpgenc_encode_frames": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.POINTER(C.c_void_p), u64p, u64p]),
    "jpgenc_encode_frames_packed": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_void_p, C.c_uint64, u64p, u64p, u64p]),
    "jpgenc_batch_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "jpgenc_batch_destroy": (None, [C.c_void_p]),
    "jpgenc_batch_last_error": (C.c_char_p, [C.c_void_p]),
    "jpgenc_batch_set_qtables": (C.c_int, [C.c_void_p, u8p, u8p]),
    "jpgenc_batch_encode": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.POINTER(C.c_void_p), u64p, u64p]),
    "jpgenc_batch_encode_device": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.c_uint32,
                                             C.POINTER(C.c_void_p), u64p, u64p]),
}

_lib = None


def load_library() -> C.CDLL:
    """Load libjpgenc_b200.so and bind every declared symbol; raises if the library was not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); there is no fallback path")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class JpgencError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"jpgenc error {code}: {text}")
        self.code = code


def _np_ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


class Encoder:
    """One GPU context (= one CUDA stream + its device buffers).  Methods mirror the C-ABI one to one."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.jpgenc_create(device, C.byref(h))
        if rc != OK:
            raise JpgencError(rc, (self.lib.jpgenc_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.jpgenc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise JpgencError(rc, (self.lib.jpgenc_last_error(self.h) or b"").decode())

    # ---- parameters -----------------------------------------------------------------------------
    def set_stage_timing(self, level: int):
        """0: no event records between the kernels (default); 1: around the K1 fast kernel; 2: around every stage -- what fills
        stats().ms_k1 / ms_forward / ms_stats / ms_entropy and their running sums"""
        self._check(self.lib.jpgenc_set_stage_timing(self.h, int(level)))

    def set_qtables(self, qy, qc):
        qy = np.ascontiguousarray(qy, np.uint8).reshape(64)
        qc = np.ascontiguousarray(qc, np.uint8).reshape(64)
        self._check(self.lib.jpgenc_set_qtables(self.h, _np_ptr(qy, u8p), _np_ptr(qc, u8p)))

    # ---- input ----------------------------------------------------------------------------------
    def upload_rgb(self, rgb: np.ndarray, maxval: int = 255):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w, _ = rgb.shape
        self._check(self.lib.jpgenc_upload_rgb(self.h, rgb.ctypes.data, w, h, maxval))

    def upload_rgb_ptr(self, host_ptr: int, w: int, h: int, maxval: int = 255):
        self._check(self.lib.jpgenc_upload_rgb(self.h, host_ptr, w, h, maxval))

    def bind_device_rgb(self, dev_ptr: int, w: int, h: int, maxval: int = 255):
        self._check(self.lib.jpgenc_bind_device_rgb(self.h, dev_ptr, w, h, maxval))

    # ---- stages ---------------------------------------------------------------------------------
    def color_dct_quant(self):
        self._check(self.lib.jpgenc_color_dct_quant(self.h))

    def get_coefficients(self) -> np.ndarray:
        st = self.stats()
        out = np.empty((st.n_blocks // 6, 6, 64), np.int16)
        self._check(self.lib.jpgenc_get_coefficients(self.h, _np_ptr(out, i16p)))
        return out

    def set_coefficients(self, q_y, q_cb, q_cr):
        q_y, q_cb, q_cr = (np.ascontiguousarray(a, np.int32) for a in (q_y, q_cb, q_cr))
        mh, mw = q_y.shape[0] // 16, q_y.shape[1] // 16
        self._check(self.lib.jpgenc_set_coefficients(self.h, _np_ptr(q_y, i32p), _np_ptr(q_cb, i32p), _np_ptr(q_cr, i32p), mw, mh))

    def set_coefficients_mcu(self, coef, mcu_w, mcu_h):
        coef = np.ascontiguousarray(coef, np.int16)
        assert coef.size == mcu_w * mcu_h * 6 * 64
        self._check(self.lib.jpgenc_set_coefficients_mcu(self.h, _np_ptr(coef, i16p), mcu_w, mcu_h))

    def symbol_stats(self):
        count = np.zeros((4, 256), np.uint32)
        first = np.zeros((4, 256), np.uint64)
        self._check(self.lib.jpgenc_symbol_stats(self.h, _np_ptr(count, u32p), _np_ptr(first, u64p)))
        return count, first

    def build_huffman(self, count, first):
        tabs = (HuffTable * 4)()
        for t in range(4):
            c = np.ascontiguousarray(count[t], np.uint32)
            f = np.ascontiguousarray(first[t], np.uint64)
            rc = self.lib.jpgenc_build_huffman(_np_ptr(c, u32p), _np_ptr(f, u64p), C.byref(tabs[t]))
            if rc != OK:
                raise JpgencError(rc, f"jpgenc_build_huffman(table {t})")
        return tabs

    def build_huffman_device(self, count, first):
        """n tables at once on the device: count (n, 256) u32, first (n, 256) u64 -> HuffTable array"""
        c = np.ascontiguousarray(count, np.uint32).reshape(-1, 256)
        f = np.ascontiguousarray(first, np.uint64).reshape(-1, 256)
        n = c.shape[0]
        tabs = (HuffTable * n)()
        self._check(self.lib.jpgenc_build_huffman_device(self.h, n, _np_ptr(c, u32p), _np_ptr(f, u64p), tabs))
        return tabs

    def entropy_encode(self, tabs) -> int:
        n = C.c_uint64()
        self._check(self.lib.jpgenc_entropy_encode(self.h, tabs, C.byref(n)))
        return n.value

    def download_scan(self) -> np.ndarray:
        n = self.stats().scan_bytes
        out = np.empty(n, np.uint8)
        self._check(self.lib.jpgenc_download_scan(self.h, _np_ptr(out, u8p), n))
        return out

    def headers(self, tabs, w, h, qy, qc) -> np.ndarray:
        qy = np.ascontiguousarray(qy, np.uint8).reshape(64)
        qc = np.ascontiguousarray(qc, np.uint8).reshape(64)
        n = self.lib.jpgenc_write_headers(w, h, _np_ptr(qy, u8p), _np_ptr(qc, u8p), tabs, None)
        out = np.empty(n, np.uint8)
        self.lib.jpgenc_write_headers(w, h, _np_ptr(qy, u8p), _np_ptr(qc, u8p), tabs, _np_ptr(out, u8p))
        return out

    # ---- whole image ----------------------------------------------------------------------------
    def encode_bound(self, out: np.ndarray | None = None) -> int:
        n = C.c_uint64()
        if out is None:
            self._check(self.lib.jpgenc_encode_bound(self.h, None, 0, C.byref(n)))
        else:
            self._check(self.lib.jpgenc_encode_bound(self.h, out.ctypes.data, out.size, C.byref(n)))
        return n.value

    def encode_rgb(self, rgb: np.ndarray, maxval: int = 255) -> bytes:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        h, w, _ = rgb.shape
        cap = max(4096, rgb.size // 2 + 4096)
        while True:
            out = np.empty(cap, np.uint8)
            n = C.c_uint64()
            rc = self.lib.jpgenc_encode_rgb(self.h, rgb.ctypes.data, w, h, maxval, out.ctypes.data, cap, C.byref(n))
            if rc == ERR_CAPACITY:
                cap = int(n.value) + 16
                continue
            self._check(rc)
            return out[: n.value].tobytes()

    def encode_planes(self, p0: np.ndarray, p1: np.ndarray, p2: np.ndarray, real_w: int, real_h: int, ycbcr: bool = False) -> bytes:
        """three (H16, W16) float64 planes as the reference's Image holds them -> JPEG bytes (exact FP64 path for every block)"""
        planes = [np.ascontiguousarray(p, np.float64) for p in (p0, p1, p2)]
        h16, w16 = planes[0].shape
        n = C.c_uint64()
        self._check(self.lib.jpgenc_encode_planes(self.h, planes[0].ctypes.data, planes[1].ctypes.data, planes[2].ctypes.data, w16, h16, real_w, real_h,
                                                  1 if ycbcr else 0, None, 0, C.byref(n)))
        out = np.empty(n.value, np.uint8)
        self._check(self.lib.jpgenc_assemble_last(self.h, out.ctypes.data, out.size, C.byref(n)))
        return out[: n.value].tobytes()

    def debug_counter(self, index: int) -> int:
        v = C.c_uint32()
        self._check(self.lib.jpgenc_debug_counter(self.h, index, C.byref(v)))
        return v.value

    def stage_subsample(self, plane: np.ndarray, mode: int) -> np.ndarray:
        """Image::applySubsampling(mode) on one plane of doubles (mode = order of Image::SubsamplingMode)"""
        plane = np.ascontiguousarray(plane, np.float64)
        h, w = plane.shape
        ow, oh = C.c_uint32(), C.c_uint32()
        self._check(self.lib.jpgenc_stage_subsample_dims(mode, w, h, C.byref(ow), C.byref(oh)))
        out = np.empty((oh.value, ow.value), np.float64)
        self._check(self.lib.jpgenc_stage_subsample(self.h, plane.ctypes.data, w, h, mode, out.ctypes.data))
        return out

    def stage_dct(self, plane: np.ndarray, mode: int) -> np.ndarray:
        """Image::applyDCT(mode) on one plane of doubles (mode 0 Simple, 1 Matrix, 2 Arai)"""
        plane = np.ascontiguousarray(plane, np.float64)
        h, w = plane.shape
        out = np.empty_like(plane)
        self._check(self.lib.jpgenc_stage_dct(self.h, plane.ctypes.data, w, h, mode, out.ctypes.data))
        return out

    def encode_rgb_into(self, host_ptr: int, w: int, h: int, out_ptr: int, cap: int, maxval: int = 255) -> int:
        n = C.c_uint64()
        self._check(self.lib.jpgenc_encode_rgb(self.h, host_ptr, w, h, maxval, out_ptr, cap, C.byref(n)))
        return n.value

    def encode_frames_device(self, frame_ptrs, w: int, h: int, out_ptrs=None, caps=None, maxval: int = 255, host_frames=False):
        """equally sized frames (device memory, or host memory with host_frames=True) through every kernel together;
        returns the JPEG sizes"""
        n = len(frame_ptrs)
        frames = (C.c_void_p * n)(*frame_ptrs)
        sizes = (C.c_uint64 * n)()
        outs = (C.c_void_p * n)(*out_ptrs) if out_ptrs is not None else None
        capv = (C.c_uint64 * n)(*caps) if caps is not None else None
        fn = self.lib.jpgenc_encode_frames if host_frames else self.lib.jpgenc_encode_frames_device
        self._check(fn(self.h, n, frames, w, h, maxval, outs, capv, sizes))
        return [int(x) for x in sizes]

    def encode_frames_packed(self, frame_ptrs, w: int, h: int, out_ptr: int | None, cap: int, maxval: int = 255, host_frames=False):
        """equally sized frames -> complete files back to back in ONE buffer (one device-to-host copy per pass);
        returns (offsets, sizes, total bytes)"""
        n = len(frame_ptrs)
        frames = (C.c_void_p * n)(*frame_ptrs)
        sizes = (C.c_uint64 * n)()
        offs = (C.c_uint64 * n)()
        total = C.c_uint64()
        self._check(self.lib.jpgenc_encode_frames_packed(self.h, n, frames, 0 if host_frames else 1, w, h, maxval, out_ptr, cap,
                                                         offs, sizes, C.byref(total)))
        return [int(x) for x in offs], [int(x) for x in sizes], int(total.value)

    def encode_ppm_file(self, src: str, dst: str):
        self._check(self.lib.jpgenc_encode_ppm_file(self.h, src.encode(), dst.encode()))

    # ---- microbench -----------------------------------------------------------------------------
    def dct_quant_blocks(self, dev_in: int, dev_out: int, nblocks: int, q, want_refined=True) -> int:
        q = np.ascontiguousarray(q, np.uint8).reshape(64)
        n = C.c_uint64()
        self._check(self.lib.jpgenc_dct_quant_blocks(self.h, dev_in, dev_out, nblocks, _np_ptr(q, u8p),
                                                     C.byref(n) if want_refined else None))
        return n.value

    # ---- helpers --------------------------------------------------------------------------------
    def stats(self) -> Stats:
        s = Stats()
        self._check(self.lib.jpgenc_get_stats(self.h, C.byref(s)))
        return s

    def synchronize(self):
        self._check(self.lib.jpgenc_synchronize(self.h))

    def stream(self) -> int:
        return self.lib.jpgenc_stream(self.h)

    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self.lib.jpgenc_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p: int):
        self._check(self.lib.jpgenc_dev_free(self.h, p))

    def h2d(self, dev: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self._check(self.lib.jpgenc_memcpy_h2d(self.h, dev, arr.ctypes.data, arr.nbytes))

    def d2h(self, arr: np.ndarray, dev: int):
        assert arr.flags.c_contiguous
        self._check(self.lib.jpgenc_memcpy_d2h(self.h, arr.ctypes.data, dev, arr.nbytes))

    def synth_rgb(self, dev: int, w: int, h: int, seed: int = 0):
        self._check(self.lib.jpgenc_synth_rgb(self.h, dev, w, h, seed))

    def synth_blocks(self, dev: int, nblocks: int):
        self._check(self.lib.jpgenc_synth_blocks(self.h, dev, nblocks))

    def flush_l2(self):
        self._check(self.lib.jpgenc_flush_l2(self.h))

    def timer_begin(self):
        self._check(self.lib.jpgenc_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        self._check(self.lib.jpgenc_timer_end(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return self.lib.jpgenc_launch_count(self.h)


class Batch:
    """Several contexts on one GPU driven by host threads inside the library: independent frames in, JPEG files out."""

    def __init__(self, device: int = 0, workers: int = 4):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.jpgenc_batch_create(device, workers, C.byref(h))
        if rc != OK:
            raise JpgencError(rc, (self.lib.jpgenc_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.jpgenc_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise JpgencError(rc, (self.lib.jpgenc_batch_last_error(self.h) or b"").decode())

    def encode_ptrs(self, frame_ptrs, w: int, h: int, out_ptrs, caps, maxval: int = 255, device_frames: bool = False):
        """raw-pointer form (pinned host or device frames); out_ptrs may be None for device frames -> sizes only"""
        n = len(frame_ptrs)
        frames = (C.c_void_p * n)(*frame_ptrs)
        sizes = (C.c_uint64 * n)()
        outs = (C.c_void_p * n)(*out_ptrs) if out_ptrs is not None else None
        capv = (C.c_uint64 * n)(*caps) if caps is not None else None
        fn = self.lib.jpgenc_batch_encode_device if device_frames else self.lib.jpgenc_batch_encode
        self._check(fn(self.h, n, frames, w, h, maxval, outs, capv, sizes))
        return [int(x) for x in sizes]

    def encode(self, frames, maxval: int = 255):
        """list of HxWx3 uint8 arrays of one size -> list of JPEG byte strings"""
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
        h, w, _ = frames[0].shape
        cap = max(4096, frames[0].size + 4096)
        outs = [np.empty(cap, np.uint8) for _ in frames]
        sizes = self.encode_ptrs([f.ctypes.data for f in frames], w, h, [o.ctypes.data for o in outs], [cap] * len(frames), maxval)
        return [o[:n].tobytes() for o, n in zip(outs, sizes)]


def bind_host_to_device_numa(device: int):
    """-> (numa node or -1, CPUs in the new affinity mask); see jpgenc_bind_host_to_device_numa"""
    node, cpus = C.c_int(-1), C.c_int(0)
    load_library().jpgenc_bind_host_to_device_numa(device, C.byref(node), C.byref(cpus))
    return node.value, cpus.value


def pinned_empty(nbytes: int):
    """(numpy view, raw pointer) over freshly allocated page-locked host memory."""
    lib = load_library()
    p = C.c_void_p()
    if lib.jpgenc_host_alloc_pinned(nbytes, C.byref(p)) != OK:
        raise MemoryError("cudaMallocHost failed")
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    return np.frombuffer(buf, np.uint8), p.value


def pinned_free(ptr: int):
    load_library().jpgenc_host_free_pinned(ptr)
