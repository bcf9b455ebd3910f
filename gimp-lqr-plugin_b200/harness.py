"""ctypes binding of tests/harness/plugin_sequence.c: the plug-in's render path (render.c:220-248,318-376;
io_functions.c:155-164) replayed in C against a library exporting the LqrCarver API."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import REPO_DIR
from .render import PlugInVals

HARNESS_PATH = os.path.join(REPO_DIR, "tests", "harness", "libplugin_sequence.so")


class HarnessVals(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "height", "bpp", "new_width", "new_height", "pres_coeff", "disc_coeff")] + \
        [("rigidity", C.c_float), ("delta_x", C.c_int), ("enl_step", C.c_float)] + \
        [(n, C.c_int) for n in ("nrg_func", "res_order", "output_seams", "scaleback", "no_disc_on_enlarge", "mask_bpp",
                                "resize_aux_layers")]


class HarnessResult(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("out_width", "out_height", "n_vmaps", "vmap_width", "vmap_height", "vmap_depth",
                                       "n_progress_updates")] + \
        [(n, C.c_double) for n in ("ms_new", "ms_setup", "ms_resize", "ms_scan", "ms_total")]


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(HARNESS_PATH)
        _lib.harness_render.restype = C.c_int
        _lib.harness_render.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(HarnessVals), C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.POINTER(HarnessResult)]
    return _lib


def render(lib_path: str, layer: np.ndarray, vals: PlugInVals, pres=None, disc=None, rigmask=None, out=None):
    """Returns (image, first_vmap_or_None, HarnessResult).  Masks must have the layer's size.  `out`: an optional
    caller-owned uint8 buffer of at least max(h, new_h) * max(w, new_w) * bpp bytes that receives the layer (the
    plug-in writes into the drawable's existing pixel region); the returned image is then a view of it."""
    layer = np.ascontiguousarray(layer, dtype=np.uint8)
    h, w, bpp = layer.shape
    masks = [None if m is None else np.ascontiguousarray(m, dtype=np.uint8) for m in (pres, disc, rigmask)]
    mbpp = next((m.shape[2] for m in masks if m is not None), 0)
    hv = HarnessVals(w, h, bpp, vals.new_width, vals.new_height, vals.pres_coeff, vals.disc_coeff, vals.rigidity,
                     vals.delta_x, vals.enl_step, vals.nrg_func, vals.res_order, int(vals.output_seams),
                     int(vals.scaleback), int(vals.no_disc_on_enlarge), mbpp, int(vals.resize_aux_layers))
    need = max(h, vals.new_height) * max(w, vals.new_width) * bpp
    own_out = out is None
    if own_out:
        out = np.zeros(need, dtype=np.uint8)
    elif out.dtype != np.uint8 or not out.flags.c_contiguous or out.size < need:
        raise ValueError("out: contiguous uint8 buffer of at least %d bytes expected" % need)
    # the seam-map sink is only written with output_seams (a zeroed layer-sized array costs milliseconds per call)
    vmap = np.zeros((h, w), dtype=np.int32) if vals.output_seams else None
    res = HarnessResult()
    ptr = [None if m is None else m.ctypes.data for m in masks]
    ok = _load().harness_render(lib_path.encode(), layer.ctypes.data, C.byref(hv), ptr[0], ptr[1], ptr[2], out.ctypes.data,
                                None if vmap is None else vmap.ctypes.data, C.byref(res))
    if not ok:
        raise RuntimeError("harness_render failed")
    img = out.reshape(-1)[: res.out_width * res.out_height * bpp].reshape(res.out_height, res.out_width, bpp)
    if own_out:
        img = img.copy()
    vm = vmap.reshape(-1)[: res.vmap_width * res.vmap_height].reshape(res.vmap_height, res.vmap_width).copy() \
        if (res.n_vmaps and vmap is not None) else None
    return img, vm, res


def render_batch(lib_path: str, layers, vals: PlugInVals, in_flight: int = 16, keep_outputs: bool = False):
    """`layers` (equal-shaped uint8 images) through the plug-in call sequence, `in_flight` of them at a time on host
    threads created in C (harness_render_batch): no Python between the images.  Returns {"wall_ms", "ms_new", ...} with
    the per-phase times summed over the images, plus "outputs" (the resized images) when keep_outputs is set."""
    layers = [np.ascontiguousarray(a, dtype=np.uint8) for a in layers]
    h, w, bpp = layers[0].shape
    assert all(a.shape == (h, w, bpp) for a in layers)
    hv = HarnessVals(w, h, bpp, vals.new_width, vals.new_height, vals.pres_coeff, vals.disc_coeff, vals.rigidity,
                     vals.delta_x, vals.enl_step, vals.nrg_func, vals.res_order, 0, int(vals.scaleback),
                     int(vals.no_disc_on_enlarge), 0, 0)
    lib = _load()
    lib.harness_render_batch.restype = C.c_int
    lib.harness_render_batch.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                         C.POINTER(HarnessVals), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    ptrs = (C.c_void_p * len(layers))(*[a.ctypes.data for a in layers])
    outs, optrs = None, None
    if keep_outputs:
        outs = [np.zeros((max(h, vals.new_height), max(w, vals.new_width), bpp), dtype=np.uint8) for _ in layers]
        optrs = (C.c_void_p * len(layers))(*[a.ctypes.data for a in outs])
    sums = (C.c_double * 5)()
    wall = C.c_double()
    if not lib.harness_render_batch(lib_path.encode(), ptrs, optrs, len(layers), in_flight, C.byref(hv), sums, C.byref(wall)):
        raise RuntimeError("harness_render_batch failed")
    res = dict(zip(("ms_new", "ms_setup", "ms_resize", "ms_scan", "ms_total"), list(sums)), wall_ms=wall.value)
    if keep_outputs:
        nw, nh = vals.new_width, vals.new_height
        res["outputs"] = [a.reshape(-1)[: nw * nh * bpp].reshape(nh, nw, bpp).copy() for a in outs]
    return res


def render_lockstep(lib_path: str, layers, vals: PlugInVals, group: int = 32, in_flight: int = 2, keep_outputs: bool = False):
    """The same batch through harness_render_lockstep: groups of `group` images are set up one after the other, resized
    by ONE lqr_b200_batch_resize call (shared launches on the device, one host thread per group) and written back;
    `in_flight` groups at a time, so the copies of one group overlap the seams of another."""
    layers = [np.ascontiguousarray(a, dtype=np.uint8) for a in layers]
    h, w, bpp = layers[0].shape
    assert all(a.shape == (h, w, bpp) for a in layers)
    hv = HarnessVals(w, h, bpp, vals.new_width, vals.new_height, vals.pres_coeff, vals.disc_coeff, vals.rigidity,
                     vals.delta_x, vals.enl_step, vals.nrg_func, vals.res_order, 0, int(vals.scaleback),
                     int(vals.no_disc_on_enlarge), 0, 0)
    lib = _load()
    lib.harness_render_lockstep.restype = C.c_int
    lib.harness_render_lockstep.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int,
                                            C.POINTER(HarnessVals), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    ptrs = (C.c_void_p * len(layers))(*[a.ctypes.data for a in layers])
    outs, optrs = None, None
    if keep_outputs:
        outs = [np.zeros((max(h, vals.new_height), max(w, vals.new_width), bpp), dtype=np.uint8) for _ in layers]
        optrs = (C.c_void_p * len(layers))(*[a.ctypes.data for a in outs])
    sums = (C.c_double * 5)()
    wall = C.c_double()
    if not lib.harness_render_lockstep(lib_path.encode(), ptrs, optrs, len(layers), group, in_flight, C.byref(hv), sums,
                                       C.byref(wall)):
        raise RuntimeError("harness_render_lockstep failed")
    res = dict(zip(("ms_new", "ms_setup", "ms_resize", "ms_scan", "ms_total"), list(sums)), wall_ms=wall.value)
    if keep_outputs:
        nw, nh = vals.new_width, vals.new_height
        res["outputs"] = [a.reshape(-1)[: nw * nh * bpp].reshape(nh, nw, bpp).copy() for a in outs]
    return res

