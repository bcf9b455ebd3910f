"""Cold-start cost of the engine in a fresh process (a GIMP plug-in is one process per invocation):
python tools/cold_start.py   -> wall time of each C-ABI call of the first and the second small resize."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gimp-lqr-plugin_b200")
img = pkg.synth.smooth_noise(640, 360, 4)
eng = C.CDLL(pkg.ENGINE_PATH)
eng.b200c_carver_new.restype = C.c_void_p
for rnd in range(2):
    out = []
    t = time.perf_counter()
    c = C.c_void_p(eng.b200c_carver_new(img.ctypes.data, 640, 360, 4))
    out.append(("new", time.perf_counter() - t)); t = time.perf_counter()
    eng.b200c_carver_init(c, 1, C.c_float(0.0))
    out.append(("init", time.perf_counter() - t)); t = time.perf_counter()
    eng.b200c_carver_build_maps(c, 41, 1, None, None)
    out.append(("build_maps(40 seams)", time.perf_counter() - t)); t = time.perf_counter()
    p = C.c_void_p()
    eng.b200c_carver_set_width(c, 600)
    eng.b200c_carver_readout(c, C.byref(p))
    out.append(("readout", time.perf_counter() - t)); t = time.perf_counter()
    eng.b200c_carver_destroy(c)
    out.append(("destroy", time.perf_counter() - t))
    print("round", rnd, {k: round(v * 1e3, 1) for k, v in out})
