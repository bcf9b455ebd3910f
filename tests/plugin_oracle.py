"""ctypes binding of oracle/libplugin_oracle.so (the plug-in's own host loops restated in C; tests only)."""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(REPO, "oracle", "libplugin_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PATH):
            subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")])
        _lib = C.CDLL(PATH)
        _lib.plugin_oracle_vmap_colour.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                                   C.POINTER(C.c_double), C.c_void_p]
        _lib.plugin_oracle_vmap_colour.restype = None
        _lib.plugin_oracle_guess_new_size.argtypes = [C.c_void_p] + [C.c_int] * 9
        _lib.plugin_oracle_guess_new_size.restype = C.c_int
    return _lib


def vmap_colour(vmap, depth, colour_start, colour_end):
    vmap = np.ascontiguousarray(vmap, dtype=np.int32)
    h, w = vmap.shape
    out = np.empty((h, w, 4), dtype=np.uint8)
    lib().plugin_oracle_vmap_colour(vmap.ctypes.data, w, h, int(depth), (C.c_double * 3)(*colour_start),
                                    (C.c_double * 3)(*colour_end), out.ctypes.data)
    return out


def guess_new_size(mask, has_alpha, x_off, y_off, old_width, old_height, direction):
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    h, w, bpp = mask.shape
    return lib().plugin_oracle_guess_new_size(mask.ctypes.data, w, h, bpp, int(bool(has_alpha)), x_off, y_off, old_width,
                                              old_height, direction)
