"""Host-side mirror of the plug-in's render path, libgimp-free.

``render_noninteractive`` replays, against any library exporting the LqrCarver API, exactly the calls
that reference src/render.c makes: render_init_carver (render.c:220-248) followed by
render_noninteractive (render.c:318-376).  Field names and defaults are those of ``PlugInVals``
(reference src/main_common.h:34-60, defaults src/main.c:62-87).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import lqr

SCALEBACK_MODE_LQRBACK = 0


@dataclass
class PlugInVals:
    new_width: int = 100
    new_height: int = 100
    pres_coeff: int = 1000
    disc_coeff: int = 1000
    rigidity: float = 0.0
    delta_x: int = 1
    enl_step: float = 150.0
    resize_aux_layers: bool = True
    output_seams: bool = False
    nrg_func: int = lqr.LQR_EF_GRAD_XABS
    res_order: int = lqr.LQR_RES_ORDER_HOR
    scaleback: bool = False
    scaleback_mode: int = SCALEBACK_MODE_LQRBACK
    no_disc_on_enlarge: bool = True


@dataclass
class RenderResult:
    image: np.ndarray
    aux: list = field(default_factory=list)
    vmaps: list = field(default_factory=list)
    info: dict = field(default_factory=dict)
    progress: list = field(default_factory=list)


def _ignore_disc_mask(vals: PlugInVals, old_w, old_h, new_w, new_h) -> bool:
    """render.c:794-821."""
    if not vals.no_disc_on_enlarge:
        return False
    if vals.res_order == lqr.LQR_RES_ORDER_HOR:
        return new_w > old_w or (new_w == old_w and new_h > old_h)
    return new_h > old_h or (new_h == old_h and new_w > old_w)


def render_init_carver(lib: lqr.LqrLib, layer: np.ndarray, vals: PlugInVals, pres=None, disc=None, rigmask=None,
                       interactive: bool = False, progress_log: list | None = None) -> lqr.Carver:
    """render.c:104-273, the engine-facing part: masks are (array, x_off, y_off) or a bare array."""
    def unpack(m):
        if m is None:
            return None
        if isinstance(m, tuple):
            return m
        return (m, 0, 0)

    pres, disc, rigmask = unpack(pres), unpack(disc), unpack(rigmask)
    old_h, old_w = layer.shape[:2]
    rigidity = 3 * vals.rigidity if rigmask is not None else vals.rigidity  # render.c:781-792
    ignore_disc = (not interactive) and _ignore_disc_mask(vals, old_w, old_h, vals.new_width, vals.new_height)

    carver = lib.carver(layer)                                   # render.c:222
    carver.init(vals.delta_x, rigidity)                          # render.c:224
    if pres is not None and vals.pres_coeff != 0:                # update_bias, io_functions.c:78-81
        carver.bias_add_rgb_area(pres[0], vals.pres_coeff, pres[1], pres[2])
    if disc is not None and not ignore_disc and vals.disc_coeff != 0:
        carver.bias_add_rgb_area(disc[0], -vals.disc_coeff, disc[1], disc[2])
    if rigmask is not None:
        carver.rigmask_add_rgb_area(rigmask[0], rigmask[1], rigmask[2])
    carver.set_energy_function_builtin(vals.nrg_func)            # render.c:234
    carver.set_resize_order(vals.res_order)                      # render.c:235
    if progress_log is not None:                                 # render.c:236, 767-779
        carver.set_progress(on_init=lambda m: progress_log.append(("init", m)),
                            on_update=lambda f: progress_log.append(("update", f)),
                            on_end=lambda m: progress_log.append(("end", m)))
    else:
        carver.set_progress()
    carver.set_side_switch_frequency(2)                          # render.c:237
    carver.set_enl_step(vals.enl_step / 100)                     # render.c:238
    if not interactive and vals.output_seams:
        carver.set_dump_vmaps()                                  # render.c:239-242
    if vals.resize_aux_layers:                                   # render.c:243-248: aux layers are resized to the
        for m in (pres, disc, rigmask):                          # layer's size first (resize_unlock_aux_layer)
            if m is not None:
                carver.attach(_fit_aux(m, old_w, old_h))
    return carver


def _fit_aux(mask, w, h) -> np.ndarray:
    """gimp_layer_resize(layer, width, height, aux_x_off - x_off, aux_y_off - y_off) (render.c:866-879):
    crop/pad the aux layer to the main layer's rectangle, transparent fill."""
    arr, x_off, y_off = mask
    if arr.ndim == 2:
        arr = arr[:, :, None]
    mh, mw, c = arr.shape
    out = np.zeros((h, w, c), dtype=np.uint8)
    x0, y0 = max(0, x_off), max(0, y_off)
    x1, y1 = min(w, x_off + mw), min(h, y_off + mh)
    if x1 > x0 and y1 > y0:
        out[y0:y1, x0:x1] = arr[y0 - y_off:y1 - y_off, x0 - x_off:x1 - x_off]
    return out


def render_noninteractive(lib: lqr.LqrLib, layer: np.ndarray, vals: PlugInVals, pres=None, disc=None, rigmask=None,
                          log_progress: bool = False) -> RenderResult:
    """render.c:275-463 minus the libgimp calls."""
    old_h, old_w = layer.shape[:2]
    plog: list | None = [] if log_progress else None
    carver = render_init_carver(lib, layer, vals, pres, disc, rigmask, progress_log=plog)
    try:
        carver.resize(vals.new_width, vals.new_height)           # render.c:318
        if vals.scaleback and vals.scaleback_mode == SCALEBACK_MODE_LQRBACK:
            carver.flatten()                                     # render.c:325
            carver.resize(old_w, old_h)                          # render.c:328
        res = RenderResult(image=None, progress=plog or [])
        if vals.output_seams:
            res.vmaps = carver.flushed_vmaps()                   # render.c:340-346
        res.info = carver.info()
        res.image = carver.scan_image()                          # render.c:366
        for h in carver.attached_handles():                      # render.c:368-374
            aux = next(a for a in carver.aux if a.handle == h)
            res.aux.append(aux.scan_image())
        return res
    finally:
        carver.destroy()                                         # render.c:376
