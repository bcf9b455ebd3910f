"""Small workload for ncu captures: python tools/prof_run.py W H SEAMS [DELTA_X]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gimp-lqr-plugin_b200")
w, h, seams = (int(a) for a in sys.argv[1:4])
dx = int(sys.argv[4]) if len(sys.argv) > 4 else 1
lib = pkg.load_product()
img = pkg.synth.smooth_noise(w, h, 4)
c = lib.carver(img)
c.init(dx, 0.0)
c.set_side_switch_frequency(2)
c.resize(w - seams, h)
out = c.scan_image()
c.destroy()
print("done", out.shape)
