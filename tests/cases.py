"""Shared case matrix: the same cases pin oracle-vs-first-principles on CPU and
CUDA-vs-oracle on the GPU."""
import importlib
import zlib

import numpy as np

pkg = importlib.import_module("gimp-lqr-plugin_b200")
synth, render, lqr = pkg.synth, pkg.render, pkg.lqr
V = render.PlugInVals


def _img(kind, w, h, c, **kw):
    return getattr(synth, kind)(w, h, c, **kw)


def case(name, kind, w, h, c, vals, pres=None, disc=None, rig=None, **kw):
    return dict(name=name, kind=kind, w=w, h=h, c=c, vals=vals, pres=pres, disc=disc, rig=rig, kw=kw)


def build_inputs(cs):
    w, h = cs["w"], cs["h"]
    img = _img(cs["kind"], w, h, cs["c"], **cs["kw"])

    def mk(spec):
        if spec is None:
            return None
        kind, mw, mh, mc, xo, yo = spec
        if kind == "ellipse":
            m = synth.ellipse_mask(mw, mh, channels=mc)
        elif kind == "band":
            m = synth.band_mask(mw, mh, channels=mc)
        else:
            m = synth.iid(mw, mh, mc, seed=zlib.crc32(kind.encode()) & 0xFFFF)  # stable across processes
        return (m, xo, yo)

    return img, mk(cs["pres"]), mk(cs["disc"]), mk(cs["rig"])


def run_case(lib, cs):
    img, pres, disc, rig = build_inputs(cs)
    return render.render_noninteractive(lib, img, cs["vals"], pres, disc, rig, log_progress=True)


# Small enough for the oracle to finish in well under a second each.
CASES = [
    case("rgba_shrink_w", "smooth_noise", 96, 64, 4, V(new_width=80, new_height=64, output_seams=True)),
    case("rgb_shrink_w", "smooth_noise", 97, 61, 3, V(new_width=77, new_height=61, output_seams=True)),
    case("gray_shrink_w", "smooth_noise", 64, 48, 1, V(new_width=50, new_height=48, output_seams=True)),
    case("graya_shrink_w", "smooth_noise", 64, 48, 2, V(new_width=50, new_height=48, output_seams=True),
         alpha="random"),
    case("rgba_alpha_holes", "smooth_noise", 80, 60, 4, V(new_width=64, new_height=60, output_seams=True),
         alpha="holes"),
    case("iid_shrink_w", "iid", 90, 70, 4, V(new_width=60, new_height=70, output_seams=True)),
    case("flat_ties", "flat", 40, 30, 4, V(new_width=30, new_height=30, output_seams=True)),
    case("ramp", "ramp", 70, 50, 3, V(new_width=55, new_height=50, output_seams=True)),
    case("shrink_h", "smooth_noise", 72, 96, 4, V(new_width=72, new_height=80, output_seams=True)),
    case("shrink_both_hor", "smooth_noise", 96, 80, 4, V(new_width=80, new_height=70, output_seams=True)),
    case("shrink_both_vert", "smooth_noise", 96, 80, 4,
         V(new_width=80, new_height=70, output_seams=True, res_order=lqr.LQR_RES_ORDER_VERT)),
    case("enlarge_w", "smooth_noise", 80, 60, 4, V(new_width=100, new_height=60, output_seams=True)),
    case("enlarge_h", "smooth_noise", 80, 60, 3, V(new_width=80, new_height=75, output_seams=True)),
    case("enlarge_multistep", "smooth_noise", 60, 40, 4, V(new_width=130, new_height=40, output_seams=True)),
    case("enlarge_step_small", "smooth_noise", 64, 40, 4,
         V(new_width=80, new_height=40, output_seams=True, enl_step=110.0)),
    case("bidirectional_cfg5", "smooth_noise", 96, 54, 4, V(new_width=86, new_height=59, output_seams=True)),
    case("minus_one", "smooth_noise", 50, 40, 4, V(new_width=49, new_height=40, output_seams=True)),
    case("plus_one", "smooth_noise", 50, 40, 4, V(new_width=51, new_height=40, output_seams=True)),
    case("delta_x0", "smooth_noise", 64, 48, 4, V(new_width=50, new_height=48, delta_x=0, output_seams=True)),
    case("delta_x2", "smooth_noise", 64, 48, 4, V(new_width=50, new_height=48, delta_x=2, output_seams=True)),
    case("delta_x10", "smooth_noise", 64, 48, 4, V(new_width=50, new_height=48, delta_x=10, output_seams=True)),
    case("rigidity", "smooth_noise", 64, 48, 4,
         V(new_width=50, new_height=48, delta_x=2, rigidity=0.2, output_seams=True)),
    case("rigidity_big", "smooth_noise", 64, 48, 4,
         V(new_width=50, new_height=48, delta_x=3, rigidity=1000.0, output_seams=True)),
    case("rigmask", "smooth_noise", 80, 60, 4,
         V(new_width=60, new_height=60, delta_x=2, rigidity=10.0, output_seams=True), rig=("band", 80, 60, 4, 0, 0)),
    case("pres_mask", "smooth_noise", 80, 60, 4, V(new_width=60, new_height=60, output_seams=True),
         pres=("ellipse", 80, 60, 4, 0, 0)),
    case("disc_mask", "smooth_noise", 80, 60, 4, V(new_width=60, new_height=60, output_seams=True),
         disc=("ellipse", 80, 60, 4, 0, 0)),
    case("pres_disc_offset", "smooth_noise", 80, 60, 4, V(new_width=64, new_height=60, output_seams=True),
         pres=("ellipse", 50, 40, 4, -10, 30), disc=("noiseA", 100, 30, 2, 20, -5)),
    case("disc_ignored_on_enlarge", "smooth_noise", 60, 40, 4, V(new_width=70, new_height=40, output_seams=True),
         disc=("ellipse", 60, 40, 4, 0, 0)),
    case("cfg3_masks", "smooth_noise", 120, 68, 4,
         V(new_width=104, new_height=68, delta_x=2, rigidity=10.0, output_seams=True),
         pres=("ellipse", 120, 68, 4, 0, 0), rig=("band", 120, 68, 4, 0, 0), alpha="random"),
    case("lqrback", "smooth_noise", 80, 60, 4, V(new_width=64, new_height=50, scaleback=True, output_seams=True)),
    case("to_width_1", "smooth_noise", 12, 9, 3, V(new_width=1, new_height=9, output_seams=True)),
    case("tiny_w2", "iid", 2, 7, 4, V(new_width=1, new_height=7, output_seams=True)),
    case("tiny_h1", "iid", 9, 1, 4, V(new_width=6, new_height=1, output_seams=True)),
    case("tiny_h2", "iid", 9, 2, 3, V(new_width=6, new_height=2, output_seams=True)),
    case("tiny_3x3", "iid", 3, 3, 1, V(new_width=2, new_height=2, output_seams=True)),
    case("wide_band", "smooth_noise", 700, 40, 4, V(new_width=690, new_height=40, output_seams=True)),
]
# medium sizes: exercise the staged/windowed kernels (window prediction, capacity fallback, chunked backtrack)
CASES += [
    case("mid_dx8_fallback", "smooth_noise", 1500, 160, 4, V(new_width=1494, new_height=160, delta_x=8, output_seams=True)),
    case("mid_dx3_rigmask", "smooth_noise", 900, 200, 4,
         V(new_width=888, new_height=200, delta_x=3, rigidity=5.0, output_seams=True), rig=("band", 900, 200, 4, 0, 0)),
    case("mid_null_dx0", "smooth_noise", 300, 100, 4,
         V(new_width=290, new_height=100, delta_x=0, nrg_func=6, output_seams=True), pres=("noiseB", 300, 100, 2, 0, 0)),
    case("mid_null_dx1", "smooth_noise", 300, 100, 4,
         V(new_width=290, new_height=100, delta_x=1, nrg_func=6, output_seams=True), pres=("noiseB", 300, 100, 2, 0, 0)),
    case("mid_iid_dx2", "iid", 1200, 150, 4, V(new_width=1190, new_height=150, delta_x=2, output_seams=True)),
    case("mid_dx32", "smooth_noise", 400, 120, 3, V(new_width=394, new_height=120, delta_x=32, output_seams=True)),
    case("mid_dx40_generic", "smooth_noise", 400, 120, 3, V(new_width=396, new_height=120, delta_x=40, output_seams=True)),
    case("tall", "smooth_noise", 60, 900, 4, V(new_width=52, new_height=900, output_seams=True)),
    case("tall_dx0", "smooth_noise", 60, 500, 4, V(new_width=54, new_height=500, delta_x=0, output_seams=True)),
    case("wide_shrink_h", "smooth_noise", 900, 60, 4, V(new_width=900, new_height=52, output_seams=True)),
    case("mid_flat_ties", "flat", 500, 300, 4, V(new_width=470, new_height=300, output_seams=True)),
    case("mid_enlarge", "smooth_noise", 640, 360, 4, V(new_width=700, new_height=380, output_seams=True)),
]
# geometry limits of the tiled band DP: taller than its row table (4608 rows -> generic update kernel), and bands wider
# than its 12 segments (flat image: every cell ties, the band spans the image -> in-kernel wide-window row loop)
CASES += [
    case("taller_than_row_table", "smooth_noise", 40, 4700, 4, V(new_width=36, new_height=4700, output_seams=True)),
    case("wide_flat_band", "flat", 2600, 48, 4, V(new_width=2592, new_height=48, output_seams=True)),
    case("wide_iid_dx4", "iid", 2000, 64, 4, V(new_width=1994, new_height=64, delta_x=4, output_seams=True)),
]
for _ef in range(7):
    CASES.append(case(f"energy_fn_{_ef}", "smooth_noise", 64, 48, 4,
                      V(new_width=52, new_height=44, nrg_func=_ef, output_seams=True), alpha="random"))
CASES.append(case("batch_scm_cfg1", "smooth_noise", 128, 128, 3,
                  V(new_width=118, new_height=128, nrg_func=3, output_seams=True)))

CASE_IDS = [c["name"] for c in CASES]


def results_equal(a, b):
    """Bit-exact comparison of two RenderResults; returns a list of human-readable differences."""
    diffs = []
    if a.info != b.info:
        diffs.append(f"info {a.info} != {b.info}")
    if len(a.vmaps) != len(b.vmaps):
        diffs.append(f"#vmaps {len(a.vmaps)} != {len(b.vmaps)}")
    for i, (va, vb) in enumerate(zip(a.vmaps, b.vmaps)):
        if (va.depth, va.orientation, va.data.shape) != (vb.depth, vb.orientation, vb.data.shape):
            diffs.append(f"vmap{i} header differs")
        elif not np.array_equal(va.data, vb.data):
            bad = va.data != vb.data
            seams = np.union1d(va.data[bad], vb.data[bad])
            seams = seams[seams > 0]
            diffs.append(f"vmap{i}: {bad.sum()} cells differ, first diverging seam #{seams.min() if len(seams) else '?'}")
    if a.image.shape != b.image.shape:
        diffs.append(f"image shape {a.image.shape} != {b.image.shape}")
    elif not np.array_equal(a.image, b.image):
        diffs.append(f"image: {(a.image != b.image).any(axis=2).sum()} pixels differ")
    if len(a.aux) != len(b.aux):
        diffs.append(f"#aux {len(a.aux)} != {len(b.aux)}")
    for i, (xa, xb) in enumerate(zip(a.aux, b.aux)):
        if xa.shape != xb.shape or not np.array_equal(xa, xb):
            diffs.append(f"aux{i} differs")
    if a.progress != b.progress:
        diffs.append(f"progress log differs ({len(a.progress)} vs {len(b.progress)} events)")
    return diffs
