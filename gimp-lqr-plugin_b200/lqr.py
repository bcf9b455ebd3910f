"""ctypes mirror of the LqrCarver C API (include/lqr.h).

The same thin wrapper drives either library that exports the API:
  * the product: ``liblqr-1.so`` (plain-C shim that dlopen()s the CUDA engine), and
  * the CPU oracle ``oracle/liblqr_oracle.so`` (tests / bench baseline only).

Names, argument order and error behaviour follow the calls gimp-lqr-plugin makes
(reference src/render.c:222-248,318,366,376; src/io_functions.c:94,125,155-164,216-219), so the
parity tests read like the plug-in's own render path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

LQR_ERROR, LQR_OK, LQR_NOMEM, LQR_USRCANCEL = 0, 1, 2, 3
LQR_RES_ORDER_HOR, LQR_RES_ORDER_VERT = 0, 1
(LQR_EF_GRAD_NORM, LQR_EF_GRAD_SUMABS, LQR_EF_GRAD_XABS, LQR_EF_LUMA_GRAD_NORM,
 LQR_EF_LUMA_GRAD_SUMABS, LQR_EF_LUMA_GRAD_XABS, LQR_EF_NULL) = range(7)

PROGRESS_INIT = C.CFUNCTYPE(C.c_int, C.c_char_p)
PROGRESS_UPDATE = C.CFUNCTYPE(C.c_int, C.c_double)
PROGRESS_END = C.CFUNCTYPE(C.c_int, C.c_char_p)
VMAP_FUNC = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)

_P = C.c_void_p
_I = C.c_int

# name -> (restype, argtypes); this is also the list the symbol-export test checks
API = {
    "lqr_carver_new": (_P, [_P, _I, _I, _I]),
    "lqr_carver_destroy": (None, [_P]),
    "lqr_carver_init": (_I, [_P, _I, C.c_float]),
    "lqr_carver_attach": (_I, [_P, _P]),
    "lqr_carver_resize": (_I, [_P, _I, _I]),
    "lqr_carver_flatten": (_I, [_P]),
    "lqr_carver_bias_add_rgb_area": (_I, [_P, _P, _I, _I, _I, _I, _I, _I]),
    "lqr_carver_rigmask_add_rgb_area": (_I, [_P, _P, _I, _I, _I, _I, _I]),
    "lqr_carver_set_energy_function_builtin": (_I, [_P, _I]),
    "lqr_carver_set_resize_order": (None, [_P, _I]),
    "lqr_carver_set_progress": (None, [_P, _P]),
    "lqr_carver_set_side_switch_frequency": (None, [_P, C.c_uint]),
    "lqr_carver_set_enl_step": (_I, [_P, C.c_float]),
    "lqr_carver_set_dump_vmaps": (None, [_P]),
    "lqr_carver_set_no_dump_vmaps": (None, [_P]),
    "lqr_carver_get_width": (_I, [_P]),
    "lqr_carver_get_height": (_I, [_P]),
    "lqr_carver_get_ref_width": (_I, [_P]),
    "lqr_carver_get_ref_height": (_I, [_P]),
    "lqr_carver_get_channels": (_I, [_P]),
    "lqr_carver_get_orientation": (_I, [_P]),
    "lqr_carver_get_depth": (_I, [_P]),
    "lqr_carver_get_enl_step": (C.c_float, [_P]),
    "lqr_carver_scan_line": (_I, [_P, C.POINTER(_I), C.POINTER(_P)]),
    "lqr_carver_scan_by_row": (_I, [_P]),
    "lqr_carver_scan": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_P)]),
    "lqr_carver_scan_reset": (None, [_P]),
    "lqr_carver_get_true_energy": (_I, [_P, _P, _I]),
    "lqr_carver_list_start": (_P, [_P]),
    "lqr_carver_list_current": (_P, [_P]),
    "lqr_carver_list_next": (_P, [_P]),
    "lqr_vmap_dump": (_P, [_P]),
    "lqr_vmap_destroy": (None, [_P]),
    "lqr_vmap_get_data": (_P, [_P]),
    "lqr_vmap_get_width": (_I, [_P]),
    "lqr_vmap_get_height": (_I, [_P]),
    "lqr_vmap_get_depth": (_I, [_P]),
    "lqr_vmap_get_orientation": (_I, [_P]),
    "lqr_vmap_list_start": (_P, [_P]),
    "lqr_vmap_list_current": (_P, [_P]),
    "lqr_vmap_list_next": (_P, [_P]),
    "lqr_vmap_list_foreach": (_I, [_P, VMAP_FUNC, _P]),
    "lqr_progress_new": (_P, []),
    "lqr_progress_set_init": (_I, [_P, PROGRESS_INIT]),
    "lqr_progress_set_update": (_I, [_P, PROGRESS_UPDATE]),
    "lqr_progress_set_end": (_I, [_P, PROGRESS_END]),
    "lqr_progress_set_update_step": (_I, [_P, C.c_float]),
    "lqr_progress_set_init_width_message": (_I, [_P, C.c_char_p]),
    "lqr_progress_set_init_height_message": (_I, [_P, C.c_char_p]),
    "lqr_progress_set_end_width_message": (_I, [_P, C.c_char_p]),
    "lqr_progress_set_end_height_message": (_I, [_P, C.c_char_p]),
}

_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]


class LqrError(RuntimeError):
    pass


class LqrLib:
    """One loaded library exporting the lqr_* API."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.dll = C.CDLL(path, mode=os.RTLD_NOW | os.RTLD_LOCAL)
        for name, (res, args) in API.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def carver(self, image: np.ndarray) -> "Carver":
        return Carver(self, image)


def _malloc_copy(arr: np.ndarray) -> int:
    """The engine ADOPTS the pixel buffer and free()s it (render.c:220-223): hand it a malloc'ed copy."""
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    p = _libc.malloc(max(arr.nbytes, 1))
    if not p:
        raise MemoryError
    C.memmove(p, arr.ctypes.data, arr.nbytes)
    return p


def _check(ret: int, what: str):
    if ret != LQR_OK:
        raise LqrError(f"{what} -> LqrRetVal {ret}")


class VMap:
    def __init__(self, data: np.ndarray, depth: int, orientation: int):
        self.data, self.depth, self.orientation = data, depth, orientation

    @classmethod
    def from_handle(cls, lib: LqrLib, h: int) -> "VMap":
        w, hh = lib.lqr_vmap_get_width(h), lib.lqr_vmap_get_height(h)
        ptr = lib.lqr_vmap_get_data(h)
        data = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int)), shape=(hh, w)).copy()
        return cls(data, lib.lqr_vmap_get_depth(h), lib.lqr_vmap_get_orientation(h))


class Carver:
    """Mirror of an LqrCarver handle.  ``image`` is (H, W, C) or (H, W) uint8, C in 1..4."""

    def __init__(self, lib: LqrLib, image: np.ndarray, _attached_to: "Carver | None" = None):
        self.lib = lib
        image = np.ascontiguousarray(image, dtype=np.uint8)
        if image.ndim == 2:
            image = image[:, :, None]
        h, w, c = image.shape
        self.channels = c
        self.handle = lib.lqr_carver_new(_malloc_copy(image), w, h, c)
        if not self.handle:
            raise LqrError("lqr_carver_new -> NULL")
        self._owned = _attached_to is None
        self._keep = []  # progress callbacks must outlive the carver
        self.aux: list[Carver] = []

    # -- life cycle -------------------------------------------------------------------------
    def destroy(self):
        if self.handle and self._owned:
            self.lib.lqr_carver_destroy(self.handle)
        self.handle = None
        for a in self.aux:
            a.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def init(self, delta_x: int = 1, rigidity: float = 0.0):
        _check(self.lib.lqr_carver_init(self.handle, delta_x, rigidity), "lqr_carver_init")
        return self

    def attach(self, image: np.ndarray) -> "Carver":
        aux = Carver(self.lib, image, _attached_to=self)
        _check(self.lib.lqr_carver_attach(self.handle, aux.handle), "lqr_carver_attach")
        self.aux.append(aux)
        return aux

    # -- masks ------------------------------------------------------------------------------
    def bias_add_rgb_area(self, mask: np.ndarray, bias_factor: int, x_off: int = 0, y_off: int = 0):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        if mask.ndim == 2:
            mask = mask[:, :, None]
        h, w, c = mask.shape
        _check(self.lib.lqr_carver_bias_add_rgb_area(self.handle, mask.ctypes.data, bias_factor, c, w, h,
                                                     x_off, y_off), "lqr_carver_bias_add_rgb_area")

    def rigmask_add_rgb_area(self, mask: np.ndarray, x_off: int = 0, y_off: int = 0):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        if mask.ndim == 2:
            mask = mask[:, :, None]
        h, w, c = mask.shape
        _check(self.lib.lqr_carver_rigmask_add_rgb_area(self.handle, mask.ctypes.data, c, w, h, x_off, y_off),
               "lqr_carver_rigmask_add_rgb_area")

    # -- knobs ------------------------------------------------------------------------------
    def set_energy_function_builtin(self, ef: int):
        _check(self.lib.lqr_carver_set_energy_function_builtin(self.handle, ef),
               "lqr_carver_set_energy_function_builtin")

    def set_resize_order(self, order: int):
        self.lib.lqr_carver_set_resize_order(self.handle, order)

    def set_side_switch_frequency(self, f: int):
        self.lib.lqr_carver_set_side_switch_frequency(self.handle, f)

    def set_enl_step(self, s: float):
        _check(self.lib.lqr_carver_set_enl_step(self.handle, s), "lqr_carver_set_enl_step")

    def set_dump_vmaps(self):
        self.lib.lqr_carver_set_dump_vmaps(self.handle)

    def set_progress(self, on_init=None, on_update=None, on_end=None, update_step: float | None = None):
        """Install Python callbacks the way render.c:767-779 installs gimp_progress_*."""
        lib = self.lib
        p = lib.lqr_progress_new()
        cbs = []
        if on_init:
            cb = PROGRESS_INIT(lambda m: (on_init(m.decode()), LQR_OK)[1])
            lib.lqr_progress_set_init(p, cb)
            cbs.append(cb)
        if on_update:
            # a hook that returns an LqrRetVal other than LQR_OK asks to cancel (None counts as LQR_OK)
            cb = PROGRESS_UPDATE(lambda f: LQR_OK if (r := on_update(f)) is None else int(r))
            lib.lqr_progress_set_update(p, cb)
            cbs.append(cb)
        if on_end:
            cb = PROGRESS_END(lambda m: (on_end(m.decode()), LQR_OK)[1])
            lib.lqr_progress_set_end(p, cb)
            cbs.append(cb)
        if update_step is not None:
            lib.lqr_progress_set_update_step(p, update_step)
        lib.lqr_progress_set_init_width_message(p, b"Resizing width...")
        lib.lqr_progress_set_init_height_message(p, b"Resizing height...")
        self._keep.extend(cbs)
        lib.lqr_carver_set_progress(self.handle, p)

    # -- hot path ---------------------------------------------------------------------------
    def resize(self, w1: int, h1: int):
        _check(self.lib.lqr_carver_resize(self.handle, w1, h1), "lqr_carver_resize")

    def flatten(self):
        _check(self.lib.lqr_carver_flatten(self.handle), "lqr_carver_flatten")

    # -- getters ----------------------------------------------------------------------------
    @property
    def width(self):
        return self.lib.lqr_carver_get_width(self.handle)

    @property
    def height(self):
        return self.lib.lqr_carver_get_height(self.handle)

    @property
    def ref_width(self):
        return self.lib.lqr_carver_get_ref_width(self.handle)

    @property
    def ref_height(self):
        return self.lib.lqr_carver_get_ref_height(self.handle)

    @property
    def orientation(self):
        return self.lib.lqr_carver_get_orientation(self.handle)

    @property
    def depth(self):
        return self.lib.lqr_carver_get_depth(self.handle)

    @property
    def enl_step(self):
        return self.lib.lqr_carver_get_enl_step(self.handle)

    def info(self) -> dict:
        return dict(width=self.width, height=self.height, ref_width=self.ref_width, ref_height=self.ref_height,
                    orientation=self.orientation, depth=self.depth, channels=self.channels)

    # -- read-out: the loop of write_carver_to_layer (io_functions.c:155-164) --------------------
    def scan_image(self) -> np.ndarray:
        lib = self.lib
        w, h, c = self.width, self.height, self.channels
        out = np.zeros((h, w, c), dtype=np.uint8)
        n = C.c_int()
        line = C.c_void_p()
        by_row = bool(lib.lqr_carver_scan_by_row(self.handle))
        count = 0
        while lib.lqr_carver_scan_line(self.handle, C.byref(n), C.byref(line)):
            if by_row:
                buf = np.ctypeslib.as_array(C.cast(line, C.POINTER(C.c_ubyte)), shape=(w, c))
                out[n.value, :, :] = buf
            else:
                buf = np.ctypeslib.as_array(C.cast(line, C.POINTER(C.c_ubyte)), shape=(h, c))
                out[:, n.value, :] = buf
            count += 1
        expect = h if by_row else w
        if count != expect:
            raise LqrError(f"scan_line returned {count} lines, expected {expect}")
        return out

    def scan_pixels(self) -> np.ndarray:
        """Pixel-wise read-out through lqr_carver_scan (slow; small images only)."""
        lib = self.lib
        w, h, c = self.width, self.height, self.channels
        out = np.zeros((h, w, c), dtype=np.uint8)
        x, y, px = C.c_int(), C.c_int(), C.c_void_p()
        while lib.lqr_carver_scan(self.handle, C.byref(x), C.byref(y), C.byref(px)):
            out[y.value, x.value, :] = np.ctypeslib.as_array(C.cast(px, C.POINTER(C.c_ubyte)), shape=(c,))
        return out

    def true_energy(self, orientation: int = 0) -> np.ndarray:
        w, h = self.width, self.height
        buf = np.zeros((h, w), dtype=np.float32)
        _check(self.lib.lqr_carver_get_true_energy(self.handle, buf.ctypes.data, orientation),
               "lqr_carver_get_true_energy")
        return buf

    # -- seam maps --------------------------------------------------------------------------
    def vmap_dump(self) -> VMap:
        h = self.lib.lqr_vmap_dump(self.handle)
        if not h:
            raise LqrError("lqr_vmap_dump -> NULL")
        v = VMap.from_handle(self.lib, h)
        self.lib.lqr_vmap_destroy(h)
        return v

    def flushed_vmaps(self) -> list[VMap]:
        """write_all_vmaps (io_functions.c:292-314): lqr_vmap_list_foreach over the engine-owned list."""
        out: list[VMap] = []
        lib = self.lib

        def visit(vh, _data):
            out.append(VMap.from_handle(lib, vh))
            return LQR_OK

        cb = VMAP_FUNC(visit)
        _check(lib.lqr_vmap_list_foreach(lib.lqr_vmap_list_start(self.handle), cb, None), "lqr_vmap_list_foreach")
        return out

    def attached_handles(self) -> list[int]:
        lib = self.lib
        out = []
        it = lib.lqr_carver_list_start(self.handle)
        while it:
            out.append(lib.lqr_carver_list_current(it))
            it = lib.lqr_carver_list_next(it)
        return out


def batch_resize(lib: LqrLib, carvers, w1: int, h1: int):
    """lqr_b200_batch_resize (product only): the resize driver for a batch of independent carvers in lockstep -- carvers
    of equal geometry and knobs share every launch of the engine; others are resized one after the other.  Libraries
    without the entry point (the CPU oracle) get plain lqr_carver_resize calls."""
    fn = getattr(lib.dll, "lqr_b200_batch_resize", None)
    if fn is None:
        for c in carvers:
            c.resize(w1, h1)
        return
    fn.restype = _I
    fn.argtypes = [C.POINTER(_P), _I, _I, _I]
    arr = (_P * len(carvers))(*[c.handle for c in carvers])
    _check(fn(arr, len(carvers), w1, h1), "lqr_b200_batch_resize")
