"""Per-stage device time of one resize (B200C_TIMING=1 makes the engine bracket every launch with CUDA
events).  Usage: B200C_TIMING=1 python tools/stage_timing.py [W H SEAMS [DELTA_X]]"""
import ctypes as C
import importlib
import json
import os
import sys
import time

os.environ.setdefault("B200C_TIMING", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gimp-lqr-plugin_b200")

STAGES = ["init_raw", "energy_full", "mmap_full", "vpath", "carve", "energy_band", "mmap_update", "mmap_tail", "fix_parents", "finish_vsmap",
          "inflate", "flatten", "transpose", "readout", "vmap", "mask", "energy_export"]


def main():
    w, h, seams = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (3840, 2160, 200)))
    dx = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    rigidity = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
    lib = pkg.load_product()
    eng = C.CDLL(pkg.ENGINE_PATH)
    eng.b200c_stage_ms.restype = C.c_double
    eng.b200c_stage_ms.argtypes = [C.c_char_p, C.POINTER(C.c_long)]
    eng.b200c_launch_count.restype = C.c_long
    img = pkg.synth.smooth_noise(w, h, 4)
    out = {}
    for rep in range(2):
        eng.b200c_stage_reset()
        l0 = eng.b200c_launch_count()
        c = lib.carver(img)
        c.init(dx, rigidity)
        c.set_side_switch_frequency(2)
        t0 = time.perf_counter()
        c.resize(w - seams, h)
        lib.lqr_carver_scan_reset(c.handle)
        img_out = c.scan_image()
        t1 = time.perf_counter()
        c.destroy()
        out = {"w": w, "h": h, "seams": seams, "delta_x": dx, "wall_s_resize_plus_readout": t1 - t0,
               "seams_per_s_wall": seams / (t1 - t0), "launches": eng.b200c_launch_count() - l0, "stages": {}}
        for s in STAGES:
            n = C.c_long()
            ms = eng.b200c_stage_ms(s.encode(), C.byref(n))
            if n.value:
                out["stages"][s] = {"ms": round(ms, 4), "launches": n.value, "us_per_launch": round(1e3 * ms / n.value, 2)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
