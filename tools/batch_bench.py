"""Config 4 (BASELINE.json: 256 x 1920x1080 RGBA, 100 seams each) on ONE GPU through the lockstep batch engine.
python tools/batch_bench.py [N_IMAGES [GROUP ...]]   -> device-resident seams/s per group size, then the C-ABI path."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("gimp-lqr-plugin_b200")
harness = importlib.import_module("gimp-lqr-plugin_b200.harness")
W, H, SEAMS, CH = 1920, 1080, 100, 4


def bind():
    eng = C.CDLL(pkg.ENGINE_PATH)
    P, I = C.c_void_p, C.c_int
    for name, (res, args) in {
        "b200c_carver_new_device": (P, [P, I, I, I]), "b200c_carver_destroy": (None, [P]),
        "b200c_carver_init": (I, [P, I, C.c_float]), "b200c_carver_set_energy_function": (I, [P, I]),
        "b200c_carver_set_side_switch_frequency": (I, [P, C.c_uint]), "b200c_batch_build_maps": (I, [C.POINTER(P), I, I]),
        "b200c_carver_set_width": (I, [P, I]), "b200c_carver_readout_device": (I, [P, P]), "b200c_set_stream": (I, [P]),
        "b200c_last_error": (C.c_char_p, []), "b200c_set_timing": (None, [I]), "b200c_stage_reset": (None, []),
        "b200c_stage_ms": (C.c_double, [C.c_char_p, C.POINTER(C.c_long)]),
    }.items():
        fn = getattr(eng, name)
        fn.restype, fn.argtypes = res, args
    return eng


def device_batch(eng, d_in, d_out, group):
    """all images of d_in (n, H, W, CH) through lockstep sessions of `group` carvers"""
    n = d_in.shape[0]
    for g0 in range(0, n, group):
        g = min(group, n - g0)
        cs = []
        for i in range(g):
            c = eng.b200c_carver_new_device(d_in[g0 + i].data_ptr(), W, H, CH)
            assert c, eng.b200c_last_error()
            assert eng.b200c_carver_init(c, 1, 0.0) == 1 and eng.b200c_carver_set_energy_function(c, 2) == 1
            assert eng.b200c_carver_set_side_switch_frequency(c, 2) == 1
            cs.append(c)
        arr = (C.c_void_p * g)(*cs)
        assert eng.b200c_batch_build_maps(arr, g, SEAMS + 1) == 1, eng.b200c_last_error()
        for i, c in enumerate(cs):
            assert eng.b200c_carver_set_width(c, W - SEAMS) == 1
            assert eng.b200c_carver_readout_device(c, d_out[g0 + i].data_ptr()) == 1
        for c in cs:
            eng.b200c_carver_destroy(c)


def device_batch_threads(eng, torch, dev, d_in, d_out, group, threads):
    """the same, `threads` lockstep sessions in flight: every host thread owns a stream and takes every threads-th
    group; all streams start after one event and the returned time spans until the last of them is done"""
    import threading
    n = d_in.shape[0]
    main = torch.cuda.current_stream()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    streams = [torch.cuda.Stream(device=dev) for _ in range(threads)]
    dones = [torch.cuda.Event() for _ in range(threads)]
    start.record(main)
    errs = []

    def work(t):
        try:
            with torch.cuda.stream(streams[t]):
                streams[t].wait_event(start)
                eng.b200c_set_stream(C.c_void_p(streams[t].cuda_stream))
                for g0 in range(t * group, n, threads * group):
                    device_batch(eng, d_in[g0:g0 + group], d_out[g0:g0 + group], group)
                dones[t].record(streams[t])
                eng.b200c_set_stream(None)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]
    for d in dones:
        main.wait_event(d)
    end.record(main)
    end.synchronize()
    return start.elapsed_time(end)


def main():
    if os.environ.get("BATCH_THREADS"):
        eng = bind()
        pkg.load_product()
        dev = torch.device("cuda", 0)
        n = int(sys.argv[1])
        imgs = [pkg.synth.smooth_noise(W, H, CH, seed=pkg.synth.SEED + i) for i in range(16)]
        d_in = torch.from_numpy(np.stack([imgs[i % 16] for i in range(n)])).to(dev)
        d_out = torch.empty((n, H, W - SEAMS, CH), dtype=torch.uint8, device=dev)
        for spec in sys.argv[2:]:
            group, threads = (int(v) for v in spec.split("x"))
            device_batch_threads(eng, torch, dev, d_in[: group * threads], d_out[: group * threads], group, threads)
            torch.cuda.synchronize()
            ms = device_batch_threads(eng, torch, dev, d_in, d_out, group, threads)
            print(f"{n} images, groups of {group}, {threads} sessions in flight: {n * SEAMS / (ms * 1e-3):10.0f} seams/s ({ms:.1f} ms)", flush=True)
        return
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    groups = [int(a) for a in sys.argv[2:]] or [32, 64, 128]
    eng = bind()
    pkg.load_product()
    dev = torch.device("cuda", 0)
    imgs = [pkg.synth.smooth_noise(W, H, CH, seed=pkg.synth.SEED + i) for i in range(min(n, 16))]
    host = np.stack([imgs[i % len(imgs)] for i in range(n)])
    d_in = torch.from_numpy(host).to(dev)
    d_out = torch.empty((n, H, W - SEAMS, CH), dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    ref = None
    with torch.cuda.stream(stream):
        eng.b200c_set_stream(C.c_void_p(stream.cuda_stream))
        for group in groups:
            device_batch(eng, d_in[: min(n, group)], d_out[: min(n, group)], group)  # warm-up
            stream.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            device_batch(eng, d_in, d_out, group)
            b.record(stream)
            stream.synchronize()
            ms = a.elapsed_time(b)
            print(f"device-resident, group {group:4d}: {n * SEAMS / (ms * 1e-3):10.0f} seams/s  ({ms:.1f} ms for {n} images)", flush=True)
            out = d_out[:len(imgs)].cpu().numpy()
            if ref is None:
                ref = out
            else:
                assert np.array_equal(ref, out), "group sizes disagree"
        if os.environ.get("BATCH_STAGES"):
            group = groups[-1]
            eng.b200c_set_timing(1)
            eng.b200c_stage_reset()
            device_batch(eng, d_in[:group], d_out[:group], group)
            stream.synchronize()
            for s_ in ["energy_full", "mmap_full", "seam_jumps", "vpath", "carve", "energy_band", "mmap_update", "mmap_tail",
                       "fix_parents", "inflate", "readout"]:
                nl = C.c_long()
                ms = eng.b200c_stage_ms(s_.encode(), C.byref(nl))
                if nl.value:
                    print(f"  stage {s_:12s}: {ms:8.2f} ms in {nl.value:5d} launches = {1e3 * ms / nl.value:8.1f} us/launch (group {group})", flush=True)
            eng.b200c_set_timing(0)
        eng.b200c_set_stream(None)
    if os.environ.get("BATCH_STAGES"):
        return
    vals = pkg.render.PlugInVals(new_width=W - SEAMS, new_height=H)
    layers = [host[i] for i in range(n)]
    for group, fl in [(16, 4), (16, 8), (8, 12), (32, 4), (8, 16)]:
        harness.render_lockstep(pkg.SHIM_PATH, layers[: group * fl], vals, group=group, in_flight=fl)
        r = harness.render_lockstep(pkg.SHIM_PATH, layers, vals, group=group, in_flight=fl)
        print(f"C ABI, host buffers, group {group} x {fl} in flight: {n * SEAMS / (r['wall_ms'] * 1e-3):10.0f} seams/s  ({r['wall_ms']:.1f} ms)", flush=True)
    r = harness.render_batch(pkg.SHIM_PATH, layers, vals, in_flight=16)
    print(f"C ABI, host buffers, thread per image, 16 in flight: {n * SEAMS / (r['wall_ms'] * 1e-3):10.0f} seams/s", flush=True)


if __name__ == "__main__":
    main()
