"""GPU parity tests: the CUDA path, called through the C ABI (liblqr-1.so -> libb200carve.so), against the
CPU oracle on the same seeded inputs.  Bars (BASELINE.json north_star): seam index sequences bit-exact
(compared through the seam-order maps, which hold every seam's pixel in every line), carved 8-bit pixels
exact, float energy within 1e-6 relative (we check bit-exact first and report the relative error).
"""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from cases import CASES, CASE_IDS, V, lqr, render, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(pkg, product):
    eng = C.CDLL(pkg.ENGINE_PATH)
    eng.b200c_debug_build.argtypes = [C.c_void_p, C.c_int]
    eng.b200c_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long]
    eng.b200c_debug_fetch.restype = C.c_long
    eng.b200c_last_error.restype = C.c_char_p
    eng.b200c_launch_count.restype = C.c_long
    product.dll.lqr_b200_engine_handle.restype = C.c_void_p
    product.dll.lqr_b200_engine_handle.argtypes = [C.c_void_p]
    return eng


@pytest.fixture(params=["fast", "narrow", "narrow_incta", "generic"])
def kernel_mode(request):
    """Every kernel set must match the oracle: "fast" (default: trapezoid-tiled band DP fed by 2-D TMA tiles, staged
    backtrack, strip-tiled full DP), "narrow" (B200C_BD_MAXSEG=1: the band DP hands every window wider than one segment
    -- i.e. nearly every row -- to the multi-SM tail kernel, which large images only reach on very wide bands),
    "narrow_incta" (the same with B200C_TAIL=0: the band kernel's own wide-window row loop, the fallback without
    cooperative launch) and "generic" (B200C_GENERIC=1: the single-CTA kernels everything falls back to).  Read at
    carver creation."""
    keys = ("B200C_GENERIC", "B200C_BD_MAXSEG", "B200C_TAIL")
    old = {k: os.environ.get(k) for k in keys}
    os.environ["B200C_GENERIC"] = "1" if request.param == "generic" else "0"
    if request.param.startswith("narrow"):
        os.environ["B200C_BD_MAXSEG"] = "1"
    else:
        os.environ.pop("B200C_BD_MAXSEG", None)
    if request.param == "narrow_incta":
        os.environ["B200C_TAIL"] = "0"
    else:
        os.environ.pop("B200C_TAIL", None)
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def test_device_is_blackwell(engine):
    """The engine is built for sm_100a only: the device the tests run on must be compute capability 10.x."""
    engine.b200c_device_count.restype = C.c_int
    assert engine.b200c_device_count() >= 1
    engine.b200c_device_cc.restype = C.c_int
    engine.b200c_device_cc.argtypes = [C.c_int]
    cc = engine.b200c_device_cc(0)
    assert cc // 10 == 10, f"compute capability {cc / 10} is not Blackwell sm_100"


@pytest.mark.parametrize("cs", CASES, ids=CASE_IDS)
def test_case_parity(product, oracle, cs, kernel_mode):
    want = cases.run_case(oracle, cs)
    got = cases.run_case(product, cs)
    diffs = cases.results_equal(got, want)
    assert not diffs, "; ".join(diffs)


@pytest.mark.parametrize("ef", range(7))
@pytest.mark.parametrize("c", [1, 2, 3, 4])
def test_energy_parity(product, oracle, ef, c):
    img = synth.smooth_noise(131, 97, c, alpha="random")
    out = []
    for lib in (product, oracle):
        with lib.carver(img) as cv:
            cv.init(1, 0.0)
            cv.set_energy_function_builtin(ef)
            out.append((cv.true_energy(0), cv.true_energy(1)))
    for got, want in zip(out[0], out[1]):
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
        assert rel.max() <= 1e-6, f"energy relative error {rel.max()}"  # north_star tolerance
        assert np.array_equal(got, want), "energy not bit-exact"


def _fetch(fn, handle, what, n, dtype):
    buf = np.zeros(n, dtype=dtype)
    got = fn(handle, what, buf.ctypes.data, n)
    assert got == n, f"fetch {what}: {got} != {n}"
    return buf


@pytest.mark.parametrize("n_seams,delta_x,rigidity,freq", [(0, 1, 0.0, 0), (1, 1, 0.0, 0), (7, 1, 0.0, 2),
                                                           (9, 2, 0.3, 2), (12, 3, 0.0, 0)])
def test_internal_maps_after_k_seams(product, oracle, engine, n_seams, delta_x, rigidity, freq, kernel_mode):
    """Energy, cumulative m-map (floats, bit-exact), parent map, index table, visibility map and the last
    seam after k iterations of the per-seam loop, before any inflate."""
    w, h = 150, 90
    img = synth.smooth_noise(w, h, 4, alpha="random")
    oracle.dll.lqr_oracle_debug_build.argtypes = [C.c_void_p, C.c_int]
    oracle.dll.lqr_oracle_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long]
    oracle.dll.lqr_oracle_debug_fetch.restype = C.c_long
    co = oracle.carver(img)
    cp = product.carver(img)
    for c in (co, cp):
        c.init(delta_x, rigidity)
        c.set_side_switch_frequency(freq)
    assert oracle.dll.lqr_oracle_debug_build(co.handle, n_seams) == lqr.LQR_OK
    eh = product.dll.lqr_b200_engine_handle(cp.handle)
    assert engine.b200c_debug_build(eh, n_seams) == 1, engine.b200c_last_error()
    n = w * h
    for what, name, dt in [(0, "en", np.float32), (1, "m", np.float32), (2, "least", np.int32),
                           (3, "raw", np.int32), (4, "vs", np.int32)]:
        a = _fetch(oracle.dll.lqr_oracle_debug_fetch, co.handle, what, n, dt).reshape(h, w)
        b = _fetch(engine.b200c_debug_fetch, eh, what, n, dt).reshape(h, w)
        if name == "raw":  # only the first w - n_seams entries of a row are live
            a, b = a[:, :w - n_seams], b[:, :w - n_seams]
        if name in ("m", "least", "en"):  # stale entries of carved pixels are not part of the state
            live = _fetch(oracle.dll.lqr_oracle_debug_fetch, co.handle, 4, n, np.int32).reshape(h, w) == 0
            if name == "least":
                live[0, :] = False  # row 0 has no parent
            a, b = a[live], b[live]
        assert np.array_equal(a, b), f"{name}: {np.count_nonzero(a != b)} entries differ after {n_seams} seams"
    if n_seams:
        a = _fetch(oracle.dll.lqr_oracle_debug_fetch, co.handle, 5, h, np.int32)
        b = _fetch(engine.b200c_debug_fetch, eh, 5, h, np.int32)
        assert np.array_equal(a, b), "vpath_x differs"
    co.destroy()
    cp.destroy()


def test_interactive_sequence(product, oracle, kernel_mode):
    """interface_I.c:504-529,615-633: one carver, many resizes, flatten in between, seam map dumps."""
    img = synth.smooth_noise(120, 90, 4)
    outs = []
    for lib in (product, oracle):
        log = []
        with lib.carver(img) as c:
            c.init(1, 0.0)
            c.set_side_switch_frequency(2)
            c.set_enl_step(1.5)
            aux = c.attach(synth.iid(120, 90, 3))
            for (tw, th) in [(100, 90), (110, 90), (120, 90), (135, 90), (100, 80), (120, 90)]:
                c.resize(tw, th)
                log.append((c.info(), c.scan_image().copy(), aux.scan_image().copy(), c.vmap_dump().data.copy()))
            c.flatten()
            c.resize(90, 100)
            log.append((c.info(), c.scan_image().copy(), aux.scan_image().copy(), c.vmap_dump().data.copy()))
        outs.append(log)
    for i, (a, b) in enumerate(zip(*outs)):
        assert a[0] == b[0], f"step {i}: info {a[0]} != {b[0]}"
        for k in (1, 2, 3):
            assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), f"step {i}: output {k} differs"
    # resize back to the reference size reproduces the original (help/en/index.wiki:130)
    assert np.array_equal(outs[0][2][1], img)


def test_scan_pixelwise(product, oracle):
    img = synth.smooth_noise(33, 21, 4)
    res = []
    for lib in (product, oracle):
        with lib.carver(img) as c:
            c.init(1, 0.0)
            c.resize(30, 25)
            res.append(c.scan_pixels())
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("w,h,seams", [(1920, 1080, 100), (3840, 2160, 40)])
def test_large_image_parity(product, oracle, w, h, seams):
    """BASELINE.json configs 2 and 4 geometry (fewer seams so the CPU oracle finishes in seconds)."""
    img = synth.smooth_noise(w, h, 4)
    vals = V(new_width=w - seams, new_height=h, output_seams=True)
    want = render.render_noninteractive(oracle, img, vals)
    got = render.render_noninteractive(product, img, vals)
    diffs = cases.results_equal(got, want)
    assert not diffs, "; ".join(diffs)


@pytest.mark.parametrize("w,h,seams", [(17001, 24, 12), (9001, 40, 10)])
def test_very_wide_rows_fall_back(product, oracle, w, h, seams):
    """Rows wider than the one-pass carve can stage (~15 k columns) and than the jump kernel's 64 partial arg-mins
    (16384 columns) take the plain carve and the single-CTA backtrack; rows wider than 8192 columns the strip launches of
    the full DP.  Odd widths: the index table's rows are not 16-byte aligned."""
    img = synth.smooth_noise(w, h, 4)
    vals = V(new_width=w - seams, new_height=h, output_seams=True)
    _assert_same(render.render_noninteractive(product, img, vals), render.render_noninteractive(oracle, img, vals))


def test_very_tall_delta4_backtrack(product, oracle):
    """delta_x 4 cuts the backtrack into blocks of 28 rows: over 7168 rows there are more sub-blocks (4 per block) than
    the chase kernel has threads."""
    w, h, seams = 72, 7400, 3
    img = synth.smooth_noise(w, h, 4)
    vals = V(new_width=w - seams, new_height=h, delta_x=4, output_seams=True)
    _assert_same(render.render_noninteractive(product, img, vals), render.render_noninteractive(oracle, img, vals))


def _assert_same(got, want):
    diffs = cases.results_equal(got, want)
    assert not diffs, "; ".join(diffs)


def test_full_size_config1(product, oracle):
    """BASELINE.json configs[0]: 512x512 RGB, 10 vertical seams with the fixed arguments of the batch script
    (batch/batch-gimp-lqr.scm:33-61: coefficients 1000, rigidity 0, delta_x 1, enl_step 150, nrg_func 3, res_order 0)."""
    w, h = 512, 512
    img = synth.smooth_noise(w, h, 3)
    vals = V(new_width=w - 10, new_height=h, pres_coeff=1000, disc_coeff=1000, rigidity=0.0, delta_x=1, enl_step=150.0,
             nrg_func=3, res_order=lqr.LQR_RES_ORDER_HOR, output_seams=True)
    _assert_same(render.render_noninteractive(product, img, vals), render.render_noninteractive(oracle, img, vals))


def test_full_size_config2(product, oracle):
    """BASELINE.json configs[1] at full size: 3840x2160 RGBA, all 200 seams, plug-in defaults (render.c:318)."""
    w, h, n = 3840, 2160, 200
    img = synth.smooth_noise(w, h, 4)
    vals = V(new_width=w - n, new_height=h, output_seams=True)
    _assert_same(render.render_noninteractive(product, img, vals, log_progress=True),
                 render.render_noninteractive(oracle, img, vals, log_progress=True))


def test_full_size_config3(product, oracle):
    """BASELINE.json configs[2] at full size: 7680x4320 RGBA (random alpha), 1000 seams, preservation ellipse
    (pres_coeff 1000), rigidity band mask, rigidity 10 (x3 with a mask, render.c:781-792), delta_x 2.  The CPU oracle
    needs about a minute for this one."""
    w, h, n = 7680, 4320, 1000
    img = synth.smooth_noise(w, h, 4, alpha="random")
    vals = V(new_width=w - n, new_height=h, delta_x=2, rigidity=10.0, output_seams=True)
    pres = synth.ellipse_mask(w, h)
    rig = synth.band_mask(w, h)
    got = render.render_noninteractive(product, img, vals, pres=pres, rigmask=rig)
    want = render.render_noninteractive(oracle, img, vals, pres=pres, rigmask=rig)
    _assert_same(got, want)


def test_full_size_config5(product, oracle):
    """BASELINE.json configs[4] at full size: 3840x2160 -> 3440x2360 (400 seams out, then 200 in along the height),
    order HOR, both seam maps (lqr_carver_set_dump_vmaps, render.c:239-242,340-346)."""
    w, h = 3840, 2160
    img = synth.smooth_noise(w, h, 4)
    vals = V(new_width=w - 400, new_height=h + 200, output_seams=True)
    got = render.render_noninteractive(product, img, vals)
    want = render.render_noninteractive(oracle, img, vals)
    assert len(got.vmaps) == 2 and got.image.shape == (h + 200, w - 400, 4)
    _assert_same(got, want)


def test_config3_geometry_scaled(product, oracle):
    """Config 3 (7680x4320, masks, rigidity, delta_x 2) at quarter scale; full size is covered by the
    size-independent property test below."""
    w, h = 1920, 1080
    img = synth.smooth_noise(w, h, 4, alpha="random")
    vals = V(new_width=w - 120, new_height=h, delta_x=2, rigidity=10.0, output_seams=True)
    pres = synth.ellipse_mask(w, h)
    rig = synth.band_mask(w, h)
    want = render.render_noninteractive(oracle, img, vals, pres=pres, rigmask=rig)
    got = render.render_noninteractive(product, img, vals, pres=pres, rigmask=rig)
    diffs = cases.results_equal(got, want)
    assert not diffs, "; ".join(diffs)


def test_full_size_properties_config2(product):
    """3840x2160 RGBA, 200 seams: size-independent properties (no oracle): one pixel per row per seam,
    connected seams, resize back to the reference is the identity, output = input minus the seam pixels."""
    w, h, n = 3840, 2160, 200
    img = synth.smooth_noise(w, h, 4)
    with product.carver(img) as c:
        c.init(1, 0.0)
        c.set_side_switch_frequency(2)
        c.resize(w - n, h)
        out = c.scan_image()
        vm = c.vmap_dump().data
        c.resize(w, h)
        back = c.scan_image()
    assert np.array_equal(back, img)
    counts = np.bincount(vm.ravel(), minlength=n + 1)
    assert np.array_equal(counts[1:], np.full(n, h))
    keep = vm == 0
    assert np.array_equal(out.reshape(h, w - n, 4), img[keep].reshape(h, w - n, 4))
    # connectivity of every seam in the coordinates current at its removal
    order = np.argsort(vm, axis=1, kind="stable")  # not needed for the check below, kept cheap:
    for k in (1, 2, n // 2, n):
        xs = np.argmax(vm == k, axis=1)
        removed_before = np.array([np.count_nonzero((vm[y, :x] > 0) & (vm[y, :x] < k)) for y, x in enumerate(xs)])
        assert np.abs(np.diff(xs - removed_before)).max() <= 1


def test_batch_images_in_flight(product, oracle):
    """Config 4 shape: several images in flight on one GPU (host threads, one stream per carver) give the same
    per-image results as the oracle run one by one."""
    import importlib
    batch = importlib.import_module("gimp-lqr-plugin_b200.batch")
    want = batch.carve_shard(oracle, range(12), 160, 90, 150, 90)
    got = batch.carve_shard(product, range(12), 160, 90, 150, 90, in_flight=6)
    assert got == want


def test_lanes_reused_across_carvers(product, oracle):
    """Streams and per-seam graph executables are pooled per process (lanes) and re-targeted with
    cudaGraphExecKernelNodeSetParams: carvers of different sizes, delta_x and rigidity that inherit a lane one after
    the other -- and several at once from host threads -- still match the oracle."""
    from concurrent.futures import ThreadPoolExecutor
    kinds = [(160, 90, 1, 0.0), (200, 120, 1, 0.0), (96, 140, 2, 0.0), (160, 90, 1, 0.5), (130, 77, 1, 0.0), (200, 120, 0, 0.0)]

    def run(lib, i):
        w, h, dx, rig = kinds[i % len(kinds)]
        img = synth.smooth_noise(w, h, 4, seed=500 + i)
        res = render.render_noninteractive(lib, img, V(new_width=w - 12, new_height=h, delta_x=dx, rigidity=rig,
                                                       output_seams=True))
        return res.image.tobytes(), res.vmaps[0].data.tobytes()

    want = [run(oracle, i) for i in range(12)]
    assert [run(product, i) for i in range(12)] == want
    with ThreadPoolExecutor(4) as ex:
        assert list(ex.map(lambda i: run(product, i), range(12))) == want


def test_c_batch_driver_matches_oracle(pkg, oracle):
    """tests/harness harness_render_batch: the plug-in call sequence from C host threads, 8 images in flight."""
    import importlib
    harness = importlib.import_module("gimp-lqr-plugin_b200.harness")
    w, h, n = 192, 108, 20
    imgs = [synth.smooth_noise(w, h, 4, seed=700 + i) for i in range(24)]
    vals = V(new_width=w - n, new_height=h)
    outs = harness.render_batch(pkg.SHIM_PATH, imgs, vals, in_flight=8, keep_outputs=True)["outputs"]
    for img, got in zip(imgs, outs):
        want = render.render_noninteractive(oracle, img, vals).image
        assert np.array_equal(got, want)


def test_range_extension_without_flatten(product, oracle, kernel_mode):
    """SURVEY.md section 8(f) rank 3 (interface_I.c:504-529): the interactive dialog asks for sizes outside the range
    computed so far; the engine runs build_maps again on the already inflated carver (more seams from the current
    minimum width, Appendix A.9's `vs >= 2*max_level - 1` bookkeeping) -- no flatten, earlier seams unchanged."""
    img = synth.smooth_noise(120, 90, 4)
    outs = []
    for lib in (product, oracle):
        log = []
        with lib.carver(img) as c:
            c.init(1, 0.0)
            c.set_side_switch_frequency(2)
            aux = c.attach(synth.iid(120, 90, 3))
            for tw in (110, 95, 128, 80, 150, 120):
                c.resize(tw, 90)
                log.append((c.info(), c.scan_image().copy(), aux.scan_image().copy(), c.vmap_dump().data.copy()))
        outs.append(log)
    for i, (a, b) in enumerate(zip(*outs)):
        assert a[0] == b[0], f"step {i}: info {a[0]} != {b[0]}"
        for k in (1, 2, 3):
            assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), f"step {i}: output {k} differs"
    # the first 10 seams are the same pixels in every later, deeper map
    first, last = outs[0][0][3], outs[0][-1][3]
    assert np.array_equal((first > 0) & (first <= 10), (last > 0) & (last <= 10))
    assert np.array_equal(outs[0][-1][1], img)  # back at the reference size: the original


@pytest.mark.parametrize("w,h,new_w,new_h,n", [(160, 90, 150, 90, 12), (200, 120, 176, 120, 5), (96, 140, 110, 150, 4)])
def test_batch_resize_lockstep(product, oracle, w, h, new_w, new_h, n):
    """lqr_b200_batch_resize: n independent images advanced by shared launches (image = blockIdx.z, argument blocks in a
    table in HBM) give, image by image, what n separate lqr_carver_resize calls give on the oracle -- pixels and seam
    maps, shrinking, enlarging and along both directions."""
    imgs = [synth.smooth_noise(w, h, 4, seed=900 + i) for i in range(n)]

    def run(lib, batched):
        cs = []
        for img in imgs:
            c = lib.carver(img)
            c.init(1, 0.0)
            c.set_side_switch_frequency(2)
            c.set_dump_vmaps()
            cs.append(c)
        if batched:
            lqr.batch_resize(lib, cs, new_w, new_h)
        else:
            for c in cs:
                c.resize(new_w, new_h)
        out = [(c.scan_image().copy(), [v.data.copy() for v in c.flushed_vmaps()]) for c in cs]
        for c in cs:
            c.destroy()
        return out

    want = run(oracle, False)
    got = run(product, True)
    for i, ((gi, gv), (wi, wv)) in enumerate(zip(got, want)):
        assert gi.shape == wi.shape and np.array_equal(gi, wi), f"image {i} differs"
        assert len(gv) == len(wv) and all(np.array_equal(a, b) for a, b in zip(gv, wv)), f"seam maps of image {i} differ"


def test_batch_resize_mixed_falls_back(product, oracle):
    """Carvers that do not agree in geometry are resized one after the other by the same call."""
    imgs = [synth.smooth_noise(120 + 8 * i, 80, 4, seed=950 + i) for i in range(3)]
    outs = []
    for lib in (product, oracle):
        cs = [lib.carver(im).init(1, 0.0) for im in imgs]
        if lib is product:
            lqr.batch_resize(lib, cs, 100, 80)
        else:
            for c in cs:
                c.resize(100, 80)
        outs.append([c.scan_image().copy() for c in cs])
        for c in cs:
            c.destroy()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_c_lockstep_driver_matches_oracle(pkg, oracle):
    """tests/harness harness_render_lockstep: groups of images set up, resized by one lqr_b200_batch_resize call and written
    back from C host threads, two groups in flight (the last group is a partial one)."""
    import importlib
    harness = importlib.import_module("gimp-lqr-plugin_b200.harness")
    w, h, n = 192, 108, 20
    imgs = [synth.smooth_noise(w, h, 4, seed=700 + i) for i in range(22)]
    vals = V(new_width=w - n, new_height=h)
    outs = harness.render_lockstep(pkg.SHIM_PATH, imgs, vals, group=8, in_flight=2, keep_outputs=True)["outputs"]
    for img, got in zip(imgs, outs):
        want = render.render_noninteractive(oracle, img, vals).image
        assert np.array_equal(got, want)


def test_progress_reports_completed_seams(product, oracle, engine):
    """LqrProgress update hooks (render.c:767-779) fire for COMPLETED seams: the engine marks every progress point with
    an event in its queue and delivers the callback once the device is past it.  At delivery of fraction i/total the
    mapped completed-seams word must show at least i-1 (the event guarantees i; the word is written by the first kernel
    of the next seam).  Same fractions, same count as the oracle."""
    w, h, n = 640, 480, 100
    img = synth.smooth_noise(w, h, 4)
    engine.b200c_carver_seams_done.restype = C.c_int
    engine.b200c_carver_seams_done.argtypes = [C.c_void_p]
    logs = {}
    for name, lib in (("oracle", oracle), ("product", product)):
        log = []
        with lib.carver(img) as c:
            c.init(1, 0.0)
            eh = product.dll.lqr_b200_engine_handle(c.handle) if lib is product else None
            c.set_progress(on_update=lambda f, eh=eh, log=log: log.append((f, engine.b200c_carver_seams_done(eh) if eh else -1)))
            c.resize(w - n, h)
        logs[name] = log
    assert [f for f, _ in logs["product"]] == [f for f, _ in logs["oracle"]]
    assert len(logs["product"]) >= 40
    for f, done in logs["product"]:
        assert done >= round(f * n) - 1, f"callback for {f:.2f} delivered with only {done} seams complete"
    # the callbacks are spread over the run, not fired while enqueueing: completed-seam counts grow from call to call
    dones = [d for _, d in logs["product"]]
    assert dones[-1] >= n - 4 and sum(b > a for a, b in zip(dones, dones[1:])) >= len(dones) // 2


def test_progress_cancel(product):
    """A hook that does not return LQR_OK cancels the resize: LQR_USRCANCEL, queued seams completed, handle destroyable."""
    img = synth.smooth_noise(320, 240, 4)
    calls = []
    with product.carver(img) as c:
        c.init(1, 0.0)

        def on_update(f):
            calls.append(f)
            return lqr.LQR_OK if len(calls) < 5 else lqr.LQR_ERROR

        c.set_progress(on_update=on_update)
        with pytest.raises(lqr.LqrError, match="LqrRetVal 3"):
            c.resize(220, 240)
    assert len(calls) == 5


@pytest.mark.parametrize("knob", ["B200C_SPLIT", "B200C_CLUSTER", "B200C_TRACE", "B200C_GRAPH"])
@pytest.mark.parametrize("name", ["mid_enlarge", "mid_dx3_rigmask", "wide_flat_band", "tall", "rgba_shrink_w"])
def test_fallback_paths(product, oracle, knob, name):
    """The round-2 paths toggled one at a time (read at carver creation): B200C_SPLIT=1 the carve as NEAR + FAR launches
    with FAR beside the band DP (off by default: measured slower), B200C_CLUSTER=0 the strip launches of the full DP instead of the
    cluster kernel, B200C_TRACE=0 the single-CTA backtrack, B200C_GRAPH=0 kernel-by-kernel launches."""
    cs = next(c for c in CASES if c["name"] == name)
    old = os.environ.get(knob)
    os.environ[knob] = "1" if knob == "B200C_SPLIT" else "0"  # the split carve is off by default, the others on
    try:
        got = cases.run_case(product, cs)
    finally:
        if old is None:
            os.environ.pop(knob, None)
        else:
            os.environ[knob] = old
    diffs = cases.results_equal(got, cases.run_case(oracle, cs))
    assert not diffs, "; ".join(diffs)

