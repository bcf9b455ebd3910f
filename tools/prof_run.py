"""Small workload for ncu captures: python tools/prof_run.py W H SEAMS [DELTA_X [RIGIDITY [MASKS]]]
MASKS=1 adds config 3's preservation ellipse and rigidity band."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gimp-lqr-plugin_b200")
w, h, seams = (int(a) for a in sys.argv[1:4])
dx = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rigidity = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
masks = len(sys.argv) > 6 and sys.argv[6] == "1"
lib = pkg.load_product()
if masks:
    img = pkg.synth.smooth_noise(w, h, 4, alpha="random")
    vals = pkg.render.PlugInVals(new_width=w - seams, new_height=h, delta_x=dx, rigidity=rigidity)
    res = pkg.render.render_noninteractive(lib, img, vals, pres=(pkg.synth.ellipse_mask(w, h), 0, 0),
                                           rigmask=(pkg.synth.band_mask(w, h), 0, 0))
    print("done", res.image.shape)
else:
    img = pkg.synth.smooth_noise(w, h, 4)
    c = lib.carver(img)
    c.init(dx, rigidity)
    c.set_side_switch_frequency(2)
    c.resize(w - seams, h)
    out = c.scan_image()
    c.destroy()
    print("done", out.shape)
