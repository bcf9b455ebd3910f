"""Debug aid: compares the engine's internal maps with the oracle's after k seams and prints the first mismatch.
Usage: python tools/debug_maps.py W H K [DELTA_X [RIGIDITY]]"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gimp-lqr-plugin_b200")
w, h, kmax = (int(a) for a in sys.argv[1:4])
dx = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rig = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
product, oracle = pkg.load_product(), pkg.load_oracle()
eng = C.CDLL(pkg.ENGINE_PATH)
eng.b200c_debug_build.argtypes = [C.c_void_p, C.c_int]
eng.b200c_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long]
eng.b200c_debug_fetch.restype = C.c_long
eng.b200c_last_error.restype = C.c_char_p
oracle.dll.lqr_oracle_debug_build.argtypes = [C.c_void_p, C.c_int]
oracle.dll.lqr_oracle_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long]
oracle.dll.lqr_oracle_debug_fetch.restype = C.c_long
product.dll.lqr_b200_engine_handle.restype = C.c_void_p
product.dll.lqr_b200_engine_handle.argtypes = [C.c_void_p]
img = pkg.synth.smooth_noise(w, h, 4)


def fetch(fn, handle, what, n, dt):
    buf = np.zeros(n, dtype=dt)
    assert fn(handle, what, buf.ctypes.data, n) == n
    return buf


for k in [int(a) for a in os.environ.get("KS", ",".join(str(i) for i in range(1, kmax + 1))).split(",")]:
    co, cp = oracle.carver(img), product.carver(img)
    for c in (co, cp):
        c.init(dx, rig)
        c.set_side_switch_frequency(2 if os.environ.get("FREQ", "0") == "2" else 0)
    assert oracle.dll.lqr_oracle_debug_build(co.handle, k) == 1
    eh = product.dll.lqr_b200_engine_handle(cp.handle)
    assert eng.b200c_debug_build(eh, k) == 1, eng.b200c_last_error()
    n = w * h
    live = fetch(oracle.dll.lqr_oracle_debug_fetch, co.handle, 4, n, np.int32).reshape(h, w) == 0
    bad = False
    for what, name, dt in [(0, "en", np.float32), (1, "m", np.float32), (2, "least", np.int32)]:
        a = fetch(oracle.dll.lqr_oracle_debug_fetch, co.handle, what, n, dt).reshape(h, w)
        b = fetch(eng.b200c_debug_fetch, eh, what, n, dt).reshape(h, w)
        m = live.copy()
        if name == "least":
            m[0, :] = False
        d = (a != b) & m
        if d.any():
            ys, xs = np.nonzero(d)
            # physical x -> current x: count live pixels to the left
            y0, x0 = ys[0], xs[0]
            cur = int(live[y0, :x0].sum())
            print(f"k={k} {name}: {d.sum()} differ; first at row {y0} phys x {x0} (current x {cur}): oracle {a[y0, x0]!r} engine {b[y0, x0]!r}; rows {ys.min()}..{ys.max()}")
            bad = True
    a = fetch(oracle.dll.lqr_oracle_debug_fetch, co.handle, 5, h, np.int32)
    b = fetch(eng.b200c_debug_fetch, eh, 5, h, np.int32)
    if not np.array_equal(a, b):
        print(f"k={k} vpath_x differs at rows {np.nonzero(a != b)[0][:5]}")
        bad = True
    co.destroy()
    cp.destroy()
    print(f"k={k}: {'MISMATCH' if bad else 'ok'}", flush=True)
    if bad:
        break
