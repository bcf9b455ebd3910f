"""BASELINE.json configs 2-5 through the product (liblqr-1.so -> libb200carve.so) at FULL size, with the
size-independent property checks (one pixel per line per seam, output = input minus seam pixels, expected number of
seam maps).  The oracle is not involved (it takes minutes at these sizes); small-size parity is in tests/.

Usage: python tools/run_configs.py [2 3 4 5] [--batch N]      (config 4: N images on this GPU, default 8)
"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gimp-lqr-plugin_b200")
synth, render = pkg.synth, pkg.render
V = render.PlugInVals


def check_vmap(vm, n, lines):
    counts = np.bincount(vm.ravel(), minlength=n + 1)
    assert counts.shape[0] == n + 1 and np.array_equal(counts[1:], np.full(n, lines)), "seam map: wrong pixel counts"


def run(lib, img, vals, **kw):
    t0 = time.perf_counter()
    res = render.render_noninteractive(lib, img, vals, **kw)
    return res, time.perf_counter() - t0


def main():
    which = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2, 3, 4, 5]
    batch = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 8
    lib = pkg.load_product()
    out = {}
    if 2 in which:
        w, h, n = 3840, 2160, 200
        img = synth.smooth_noise(w, h, 4)
        run(lib, img[:256, :512].copy(), V(new_width=500, new_height=256))  # warm-up (context, pools)
        res, dt = run(lib, img, V(new_width=w - n, new_height=h, output_seams=True))
        assert res.image.shape == (h, w - n, 4) and len(res.vmaps) == 1
        vm = res.vmaps[0].data
        check_vmap(vm, n, h)
        assert np.array_equal(res.image, img[vm == 0].reshape(h, w - n, 4))
        out["config2"] = {"what": "3840x2160 RGBA, 200 vertical seams", "wall_s": dt, "seams_per_s_e2e": n / dt}
    if 3 in which:
        w, h, n = 7680, 4320, 1000
        img = synth.smooth_noise(w, h, 4, alpha="random")
        pres, rig = synth.ellipse_mask(w, h), synth.band_mask(w, h)
        res, dt = run(lib, img, V(new_width=w - n, new_height=h, delta_x=2, rigidity=10.0, output_seams=True),
                      pres=(pres, 0, 0), rigmask=(rig, 0, 0))
        assert res.image.shape == (h, w - n, 4) and len(res.vmaps) == 1
        vm = res.vmaps[0].data
        check_vmap(vm, n, h)
        assert np.array_equal(res.image, img[vm == 0].reshape(h, w - n, 4))
        out["config3"] = {"what": "7680x4320 RGBA, 1000 seams, preservation + rigidity masks, delta_x 2, rigidity 10",
                          "wall_s": dt, "seams_per_s_e2e": n / dt}
        # the same through the plug-in's call sequence in C (no Python between the calls)
        harness = importlib.import_module("gimp-lqr-plugin_b200.harness")
        img_c, vm_c, hres = harness.render(pkg.SHIM_PATH, img, V(new_width=w - n, new_height=h, delta_x=2, rigidity=10.0,
                                                                 output_seams=True), pres=pres, rigmask=rig)
        assert np.array_equal(img_c, res.image) and np.array_equal(vm_c, vm)
        out["config3_c_harness"] = {k: round(getattr(hres, k), 2) for k in ("ms_new", "ms_setup", "ms_resize", "ms_scan", "ms_total")}
        out["config3_c_harness"]["seams_per_s_e2e"] = n / (hres.ms_total * 1e-3)
    if 4 in which:
        w, h, n = 1920, 1080, 100
        t0 = time.perf_counter()
        for i in range(batch):
            img = synth.smooth_noise(w, h, 4, seed=synth.SEED + i)
            res = render.render_noninteractive(lib, img, V(new_width=w - n, new_height=h))
            assert res.image.shape == (h, w - n, 4)
        dt = time.perf_counter() - t0
        out["config4"] = {"what": f"{batch} of the 256 x 1920x1080 RGBA images, 100 seams each, sequentially on one GPU "
                                  "(incl. synthetic image generation on the host)", "wall_s": dt,
                          "seams_per_s_e2e": batch * n / dt}
    if 44 in which:
        # config 4 with several images in flight on this GPU: one host thread + one CUDA stream per image (the engine
        # gives every carver its own stream); the row-serial chains of different images overlap on different SMs
        harness = importlib.import_module("gimp-lqr-plugin_b200.harness")
        w, h, n = 1920, 1080, int(os.environ.get("B200C_SEAMS", "100"))
        nthreads = int(os.environ.get("B200C_THREADS", "16"))
        distinct = [synth.smooth_noise(w, h, 4, seed=synth.SEED + i) for i in range(min(batch, 64))]
        imgs = [distinct[i % len(distinct)] for i in range(batch)]  # 64 distinct images, repeated (host time to make them)
        # warm-up with the same number of images in flight: staging buffers, streams and graph executables are pooled
        harness.render_batch(pkg.SHIM_PATH, imgs[:min(batch, 2 * nthreads)], V(new_width=w - n, new_height=h), in_flight=nthreads)
        import ctypes as C
        eng = C.CDLL(pkg.ENGINE_PATH)
        eng.b200c_hostprof_ms.restype = C.c_double
        names = ["new_common", "new_alloc", "up_pinned", "up_lockwait", "up_memcpy", "up_enqueue", "up_sync", "graph_build",
                 "graph_launch", "loop_sync", "readout", "destroy", "g_begin", "g_launches", "g_inst_destroy", "g_end"]
        prof0 = [eng.b200c_hostprof_ms(i) for i in range(len(names))]
        # the plug-in's call sequence AND the host threads in C (tests/harness): no Python between the images
        r = harness.render_batch(pkg.SHIM_PATH, imgs, V(new_width=w - n, new_height=h), in_flight=nthreads)
        dt = r["wall_ms"] * 1e-3
        out["config4_phase_ms_per_image"] = {k: round(v / batch, 3) for k, v in r.items() if k != "wall_ms"}
        out["config4_engine_host_ms_per_image"] = {nm: round((eng.b200c_hostprof_ms(i) - prof0[i]) / batch, 3)
                                                   for i, nm in enumerate(names)}
        out["config4_concurrent"] = {"what": f"{batch} x 1920x1080 RGBA, 100 seams each, {nthreads} images in flight on one GPU "
                                             "(host threads, one stream per carver)", "wall_s": dt,
                                     "seams_per_s_e2e": batch * n / dt}
    if 45 in which:
        # config 4's CPU side: the oracle port of liblqr through the same C batch driver, 1 thread and all cores
        harness = importlib.import_module("gimp-lqr-plugin_b200.harness")
        w, h, n = 1920, 1080, 100
        ncores = os.cpu_count() or 1
        imgs = [synth.smooth_noise(w, h, 4, seed=synth.SEED + i) for i in range(ncores)]
        r1 = harness.render_batch(pkg.ORACLE_PATH, imgs[:2], V(new_width=w - n, new_height=h), in_flight=1)
        rn = harness.render_batch(pkg.ORACLE_PATH, imgs, V(new_width=w - n, new_height=h), in_flight=ncores)
        out["config4_cpu_oracle"] = {"what": "oracle port of liblqr, 1920x1080 RGBA, 100 seams per image",
                                     "seams_per_s_1_thread": 2 * n / (r1["wall_ms"] * 1e-3),
                                     "threads": ncores, "seams_per_s_all_threads": ncores * n / (rn["wall_ms"] * 1e-3)}
    if 5 in which:
        w, h = 3840, 2160
        img = synth.smooth_noise(w, h, 4)
        res, dt = run(lib, img, V(new_width=w - 400, new_height=h + 200, output_seams=True))
        assert res.image.shape == (h + 200, w - 400, 4) and len(res.vmaps) == 2
        check_vmap(res.vmaps[0].data, 400, h)
        check_vmap(res.vmaps[1].data, 200, w - 400)
        out["config5"] = {"what": "3840x2160 -> 3440x2360 (W-400, H+200) with seam maps", "wall_s": dt,
                          "seams_per_s_e2e": 600 / dt}
        # the output path of the seam maps (write_vmap_to_layer's colouring, io_functions.c:249-279): engine vs CPU loop
        ops = importlib.import_module("gimp-lqr-plugin_b200.plugin_ops")
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        import plugin_oracle
        cs, ce = (1.0, 1.0, 0.0), (1.0, 0.0, 0.0)
        vm = res.vmaps[0]
        ops.vmap_colour(vm.data, vm.depth, cs, ce)
        t0 = time.perf_counter()
        got = [ops.vmap_colour(v.data, v.depth, cs, ce) for v in res.vmaps]
        t_gpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        want = [plugin_oracle.vmap_colour(v.data, v.depth, cs, ce) for v in res.vmaps]
        t_cpu = time.perf_counter() - t0
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
        out["config5_vmap_colour"] = {"what": "both seam maps coloured as RGBA, host buffers in and out",
                                      "engine_ms": t_gpu * 1e3, "cpu_loop_ms": t_cpu * 1e3,
                                      "mpixel_per_s_engine": sum(v.data.size for v in res.vmaps) / t_gpu / 1e6}
    if os.environ.get("B200C_TIMING"):
        import ctypes as C
        eng = C.CDLL(pkg.ENGINE_PATH)
        eng.b200c_stage_ms.restype = C.c_double
        eng.b200c_stage_ms.argtypes = [C.c_char_p, C.POINTER(C.c_long)]
        st = {}
        for name in ["energy_full", "mmap_full", "vpath", "carve", "energy_band", "mmap_update", "mmap_tail", "fix_parents", "inflate",
                     "readout", "mask", "gather_rig", "transpose", "flatten", "vmap"]:
            n = C.c_long()
            ms = eng.b200c_stage_ms(name.encode(), C.byref(n))
            if n.value:
                st[name] = {"ms": round(ms, 2), "launches": n.value, "us_per_launch": round(1e3 * ms / n.value, 1)}
        out["stages_all_configs"] = st
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
