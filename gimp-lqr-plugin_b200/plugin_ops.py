"""The plug-in's own host loops next to the hot path, bound to the CUDA engine's C ABI
(include/b200carve.h): the colouring of write_vmap_to_layer (reference src/io_functions.c:249-279) and the mask
reduction of guess_new_size (reference src/layers_combo.c:274-392).  Same argument meaning as the reference
functions; no CPU fallback -- the engine library must load and a CUDA device must be present."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ENGINE_PATH

GUESS_DIR_HOR, GUESS_DIR_VERT = 0, 1  # layers_combo.h GuessDir

_eng = None


def _engine():
    global _eng
    if _eng is None:
        eng = C.CDLL(ENGINE_PATH)
        eng.b200c_vmap_colour.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.c_void_p]
        eng.b200c_vmap_colour.restype = C.c_int
        eng.b200c_guess_new_size.argtypes = [C.c_void_p] + [C.c_int] * 9 + [C.POINTER(C.c_int)]
        eng.b200c_guess_new_size.restype = C.c_int
        eng.b200c_last_error.restype = C.c_char_p
        _eng = eng
    return _eng


def _check(rc, what):
    if rc != 1:  # B200C_OK
        raise RuntimeError(f"{what}: {_engine().b200c_last_error().decode()}")


def vmap_colour(vmap: np.ndarray, depth: int, colour_start, colour_end) -> np.ndarray:
    """write_vmap_to_layer's pixels: vmap (h, w) int32 seam orders -> (h, w, 4) uint8 RGBA."""
    vmap = np.ascontiguousarray(vmap, dtype=np.int32)
    h, w = vmap.shape
    out = np.empty((h, w, 4), dtype=np.uint8)
    cs, ce = (C.c_double * 3)(*colour_start), (C.c_double * 3)(*colour_end)
    _check(_engine().b200c_vmap_colour(vmap.ctypes.data, w, h, int(depth), cs, ce, out.ctypes.data), "vmap_colour")
    return out


def guess_new_size(mask: np.ndarray, has_alpha: bool, x_off: int, y_off: int, old_width: int, old_height: int,
                   direction: int) -> int:
    """guess_new_size for a discard mask (h, w, bpp) uint8 placed at (x_off, y_off) over the layer."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w, bpp = mask.shape
    res = C.c_int()
    _check(_engine().b200c_guess_new_size(mask.ctypes.data, w, h, bpp, int(bool(has_alpha)), x_off, y_off, old_width,
                                          old_height, direction, C.byref(res)), "guess_new_size")
    return res.value
