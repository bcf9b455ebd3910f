import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gimp-lqr-plugin_b200")
synth = pkg.synth
w, h = 7680, 4320
img = synth.smooth_noise(w, h, 4, alpha="random")
pres, rig = synth.ellipse_mask(w, h), synth.band_mask(w, h)
print("mask shapes", pres.shape, rig.shape, pres.dtype)
lib = pkg.load_product()
for rnd in range(2):
    t = time.perf_counter(); c = lib.carver(img); t1 = time.perf_counter()
    c.init(2, 30.0); t2 = time.perf_counter()
    c.bias_add_rgb_area(pres, 1000, 0, 0) if hasattr(c, "bias_add_rgb_area") else None; t3 = time.perf_counter()
    c.rigmask_add_rgb_area(rig, 0, 0) if hasattr(c, "rigmask_add_rgb_area") else None; t4 = time.perf_counter()
    print("round", rnd, "new %.1f init %.1f bias %.1f rigmask %.1f ms" % ((t1-t)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3))
    c.destroy()
