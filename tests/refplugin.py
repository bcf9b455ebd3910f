"""ctypes driver of oracle/_ref/libplugin_ref_{b200,oracle}.so: the reference plug-in's OWN object code
(src/render.c + src/io_functions.c, compiled unmodified by oracle/Makefile against include/lqr.h and the in-memory
libgimp of oracle/gimpstub/) linked against the product shim or the CPU oracle.  Tests only."""
import ctypes as C
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(REPO, "oracle", "_ref")
COL_DEFAULT = (1.0, 1.0, 0.0, 0.2, 0.0, 0.0)  # PlugInColVals defaults, reference src/main.c:89-96


def path(flavour):
    return os.path.join(REF_DIR, f"libplugin_ref_{flavour}.so")


class RefPlugin:
    def __init__(self, flavour):
        self.dll = d = C.CDLL(path(flavour), mode=os.RTLD_NOW | os.RTLD_LOCAL)
        d.fg_layer_add.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_char_p]
        d.fg_layer_info.argtypes = [C.c_int, C.POINTER(C.c_int)]
        d.fg_layer_pixels.restype = C.c_void_p
        d.fg_layer_name.restype = C.c_char_p
        d.fg_image_layers.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int]
        d.fg_last_message.restype = C.c_char_p
        d.fg_progress_counts.argtypes = [C.POINTER(C.c_long)]
        d.ref_run_noninteractive.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                             C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]

    def layer(self, lid):
        info = (C.c_int * 6)()
        assert self.dll.fg_layer_info(lid, info), f"no layer {lid}"
        w, h, bpp, xo, yo, _ = list(info)
        px = np.ctypeslib.as_array(C.cast(self.dll.fg_layer_pixels(lid), C.POINTER(C.c_ubyte)), shape=(h, w, bpp)).copy()
        return dict(id=lid, pixels=px, x_off=xo, y_off=yo, name=self.dll.fg_layer_name(lid).decode())

    def run(self, image, vals, pres=None, disc=None, rigmask=None, layer_off=(0, 0), col=COL_DEFAULT, resize_canvas=False):
        """image: (H, W, C) uint8; masks: (array, x_off, y_off) in image coordinates.  Returns the layers of the image
        after the plug-in ran: {"main": ..., "pres"/"disc"/"rigmask": ..., "seams": [...]}."""
        d = self.dll
        d.fg_reset()
        h, w, c = image.shape
        img_id = d.fg_image_new(w + layer_off[0], h + layer_off[1], 0 if c >= 3 else 1)

        def add(arr, xo, yo, name):
            arr = np.ascontiguousarray(arr, dtype=np.uint8)
            if arr.ndim == 2:
                arr = arr[:, :, None]
            return d.fg_layer_add(img_id, arr.shape[1], arr.shape[0], arr.shape[2], xo, yo, arr.ctypes.data, name)

        main = add(image, layer_off[0], layer_off[1], b"layer")
        ids = {}
        for key, m in (("pres", pres), ("disc", disc), ("rigmask", rigmask)):
            ids[key] = add(m[0], m[1] + layer_off[0], m[2] + layer_off[1], key.encode()) if m is not None else 0
        iv = (C.c_int * 18)(vals.new_width, vals.new_height, ids["pres"], vals.pres_coeff, ids["disc"], vals.disc_coeff,
                            ids["rigmask"], vals.delta_x, int(vals.resize_aux_layers), int(resize_canvas), 0,
                            int(vals.output_seams), vals.nrg_func, vals.res_order, 0, int(vals.scaleback),
                            vals.scaleback_mode, int(vals.no_disc_on_enlarge))
        fv = (C.c_float * 2)(vals.rigidity, vals.enl_step)
        cv = (C.c_double * 6)(*col)
        oi, ol = C.c_int(), C.c_int()
        ok = d.ref_run_noninteractive(img_id, main, iv, fv, cv, C.byref(oi), C.byref(ol))
        assert ok, f"render failed: {d.fg_last_message().decode()}"
        buf = (C.c_int * 64)()
        n = d.fg_image_layers(oi.value, buf, 64)
        known = {ol.value: "main", **{v: k for k, v in ids.items() if v}}
        out = {"seams": []}
        for lid in list(buf)[:n]:
            lay = self.layer(lid)
            if lid in known:
                out[known[lid]] = lay
            else:
                out["seams"].append(lay)
        pc = (C.c_long * 4)()
        d.fg_progress_counts(pc)
        out["progress"] = list(pc)
        return out


def layers_equal(a, b):
    diffs = []
    for key in sorted(set(a) | set(b)):
        if key == "progress":
            if a[key] != b[key]:
                diffs.append(f"progress counts {a[key]} != {b[key]}")
            continue
        la, lb = a.get(key), b.get(key)
        if key == "seams":
            if len(la) != len(lb):
                diffs.append(f"#seam layers {len(la)} != {len(lb)}")
            pairs = zip(la, lb)
        else:
            if (la is None) != (lb is None):
                diffs.append(f"{key}: present in one run only")
                continue
            pairs = [(la, lb)]
        for x, y in pairs:
            if (x["x_off"], x["y_off"], x["name"]) != (y["x_off"], y["y_off"], y["name"]):
                diffs.append(f"{key}: placement/name differs")
            if x["pixels"].shape != y["pixels"].shape:
                diffs.append(f"{key}: shape {x['pixels'].shape} != {y['pixels'].shape}")
            elif not np.array_equal(x["pixels"], y["pixels"]):
                diffs.append(f"{key}: {np.count_nonzero((x['pixels'] != y['pixels']).any(axis=2))} pixels differ")
    return diffs
