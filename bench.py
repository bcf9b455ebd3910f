#!/usr/bin/env python
"""bench.py -- headline benchmark of the seam-carving hot path (BASELINE.json: seams/sec, 4K RGBA).

Workload (config.workload): BASELINE.json configs[1] -- one 3840x2160 RGBA image per GPU, 3840 -> 3640
(200 vertical seams), plug-in defaults (GRAD_XABS, delta_x 1, rigidity 0, side-switch frequency 2,
src/main.c:62-87 + src/render.c:237).  One "step" = one full pass of the hot path over that image:
carver set-up, energy map, m-map DP, 200 x (backtrack, carve, band energy, band DP), inflate, read-out.

  value : seams/s with the input image already resident in HBM (device pointer in, device pointer out),
          timed with CUDA events on the stream the engine's kernels run on, max over ranks.
  e2e   : the same metric through the reference-facing C ABI (liblqr-1.so: lqr_carver_new .. scan_line)
          with HOST buffers: the H2D upload of the image and the D2H read-out are inside the timed region.
  N > 1 : one process per GPU (torchrun), every rank carves its own image (weak scaling, no collective on
          the data path); value = all ranks' seams / max-over-ranks time.
  --impl reference : the CPU oracle port of liblqr (liblqr itself is not vendored in the reference tree)
          on one host core -- liblqr is single-threaded and one image's seams are strictly sequential.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
pkg = importlib.import_module("gimp-lqr-plugin_b200")

W, H, SEAMS, CH = 3840, 2160, 200, 4
WORKLOAD = f"{W}x{H} RGBA, {SEAMS} vertical seams ({W}->{W - SEAMS}), plug-in defaults, one image per GPU"
METRIC, UNIT = "seams_per_sec_4k_rgba", "seams/s"


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_engine():
    eng = C.CDLL(pkg.ENGINE_PATH)
    P, I = C.c_void_p, C.c_int
    sig = {
        "b200c_carver_new_device": (P, [P, I, I, I]), "b200c_carver_destroy": (None, [P]),
        "b200c_carver_init": (I, [P, I, C.c_float]), "b200c_carver_set_energy_function": (I, [P, I]),
        "b200c_carver_set_side_switch_frequency": (I, [P, C.c_uint]),
        "b200c_carver_build_maps": (I, [P, I, I, P, P]), "b200c_carver_set_width": (I, [P, I]),
        "b200c_carver_readout_device": (I, [P, P]), "b200c_carver_sync": (I, [P]), "b200c_set_stream": (I, [P]),
        "b200c_set_device": (I, [I]), "b200c_set_timing": (None, [I]), "b200c_launch_count": (C.c_long, []),
        "b200c_stage_ms": (C.c_double, [C.c_char_p, C.POINTER(C.c_long)]), "b200c_stage_reset": (None, []),
        "b200c_update_cells": (C.c_ulonglong, []), "b200c_last_error": (C.c_char_p, []),
        "b200c_device_count": (I, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(eng, name)
        fn.restype, fn.argtypes = res, args
    return eng


def device_step(eng, d_in_ptr, d_out_ptr):
    """The resize driver of lqr_carver_resize for a pure width shrink, on device buffers."""
    c = eng.b200c_carver_new_device(d_in_ptr, W, H, CH)
    if not c:
        raise RuntimeError(eng.b200c_last_error().decode())
    try:
        ok = eng.b200c_carver_init(c, 1, 0.0) == 1
        ok = ok and eng.b200c_carver_set_energy_function(c, 2) == 1
        ok = ok and eng.b200c_carver_set_side_switch_frequency(c, 2) == 1
        ok = ok and eng.b200c_carver_build_maps(c, SEAMS + 1, 1, None, None) == 1
        ok = ok and eng.b200c_carver_set_width(c, W - SEAMS) == 1
        ok = ok and eng.b200c_carver_readout_device(c, d_out_ptr) == 1
        if not ok:
            raise RuntimeError(eng.b200c_last_error().decode())
    finally:
        eng.b200c_carver_destroy(c)


_ABI_OUT = None  # the layer's pixel region the plug-in writes the result into: allocated once, like the drawable


def abi_step(lib_path, img, seams=SEAMS):
    """The plug-in's own call sequence in C (tests/harness/plugin_sequence.c: render_init_carver + render_noninteractive +
    write_carver_to_layer, render.c:220-248,318,366,376; io_functions.c:155-164) on host buffers."""
    harness = importlib.import_module("gimp-lqr-plugin_b200.harness")
    vals = pkg.render.PlugInVals(new_width=W - seams, new_height=H)
    global _ABI_OUT
    if _ABI_OUT is None:
        _ABI_OUT = np.zeros(W * H * CH, dtype=np.uint8)
    out, _, res = harness.render(lib_path, img, vals, out=_ABI_OUT)
    return out, res


def run_reference(args, rank):
    if rank != 0:
        return
    img = pkg.synth.smooth_noise(W, H, CH)
    est_step = 2.5  # s, one full config-2 pass on one core
    seams = SEAMS
    if (args.steps + args.warmup) * est_step > 170:
        seams = max(20, int(SEAMS * 170 / ((args.steps + args.warmup) * est_step)))
    def step():
        return abi_step(pkg.ORACLE_PATH, img, seams)[0]

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = seams * args.steps / dt
    sample = (f"{args.steps} x ({W}x{H} RGBA, {seams} seams of the {SEAMS}-seam workload) through the Lqr API incl. "
              "scan_line read-out; C oracle restating liblqr (liblqr itself is not in the reference tree), gcc -O2, "
              "1 thread: liblqr is single-threaded and one image's seams are sequential")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "seams_per_step": seams},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline():
    img = pkg.synth.smooth_noise(W, H, CH)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        abi_step(pkg.ORACLE_PATH, img)
        times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    return {"value": SEAMS / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"3 x the full workload ({WORKLOAD}) through the Lqr API, median; C oracle restating liblqr "
                      "(unverified against real liblqr), gcc -O2, 1 host core (liblqr is single-threaded)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-in-flight", type=int, default=8,
                    help="N=1 only: also time 2x this many images of the workload, this many in flight (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)

    eng = bind_engine()
    pkg.load_product()  # fails loudly when the CUDA engine is not built
    eng.b200c_set_device(local_rank)

    img = pkg.synth.smooth_noise(W, H, CH, seed=pkg.synth.SEED + rank)
    d_in = torch.from_numpy(img).to(dev)
    d_out = torch.empty((H, W - SEAMS, CH), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: device-resident, CUDA events on the engine's stream ----------------------
    with torch.cuda.stream(stream):
        eng.b200c_set_stream(C.c_void_p(stream.cuda_stream))
        for _ in range(warmup):
            device_step(eng, d_in.data_ptr(), d_out.data_ptr())
        stream.synchronize()
        barrier()
        launches0 = eng.b200c_launch_count()
        evs = []
        wall0 = time.perf_counter()
        with ClockSampler(local_rank) as clk:
            for _ in range(args.steps):
                flush.zero_()  # L2 flush between timed iterations (outside the event bracket)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                device_step(eng, d_in.data_ptr(), d_out.data_ptr())
                b.record(stream)
                evs.append((a, b))
            stream.synchronize()
            barrier()
        wall = time.perf_counter() - wall0
        launches = eng.b200c_launch_count() - launches0
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        out_dev = d_out.cpu().numpy()

        # ---------------- per-kernel durations (separate instrumented pass, same stream) ------------
        eng.b200c_set_timing(1)
        eng.b200c_stage_reset()
        cells0 = eng.b200c_update_cells()
        prof_steps = 2
        for _ in range(prof_steps):
            flush.zero_()
            device_step(eng, d_in.data_ptr(), d_out.data_ptr())
        stream.synchronize()
        stages = {}
        for s in ["energy_full", "mmap_full", "seam_jumps", "vpath", "carve", "energy_band", "mmap_update", "mmap_tail", "fix_parents", "inflate", "readout"]:
            n = C.c_long()
            ms = eng.b200c_stage_ms(s.encode(), C.byref(n))
            if n.value:
                stages[s] = {"ms_per_step": ms / prof_steps, "launches_per_step": n.value / prof_steps,
                             "us_per_launch": 1e3 * ms / n.value}
        cells = (eng.b200c_update_cells() - cells0) / prof_steps
        eng.b200c_set_timing(0)
        eng.b200c_set_stream(None)

    # ---------------- e2e: host buffers through the C ABI -------------------------------------------
    for _ in range(2):
        out_abi, _ = abi_step(pkg.SHIM_PATH, img)
    barrier()
    t0 = time.perf_counter()
    phases = {"ms_new": 0.0, "ms_setup": 0.0, "ms_resize": 0.0, "ms_scan": 0.0}
    for _ in range(args.steps):
        out_abi, hres = abi_step(pkg.SHIM_PATH, img)
        for k in phases:
            phases[k] += getattr(hres, k) / args.steps
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(out_abi, out_dev), "device-resident and C-ABI paths disagree"

    # ---------------- several images in flight on this GPU (a batch host; SURVEY.md config 4's shape) ---------------
    # Same workload and C-ABI path as e2e, host threads in C (harness_render_batch), one stream per image: the row-serial
    # chains of different images overlap on different SMs.  Reported beside the headline, not as it.
    batch_line = None
    if world == 1 and args.batch_in_flight > 0:
        harness = importlib.import_module("gimp-lqr-plugin_b200.harness")
        vals = pkg.render.PlugInVals(new_width=W - SEAMS, new_height=H)
        k = args.batch_in_flight
        harness.render_batch(pkg.SHIM_PATH, [img] * k, vals, in_flight=k)  # warm-up: staging buffers, lanes, graphs
        runs = [harness.render_batch(pkg.SHIM_PATH, [img] * (2 * k), vals, in_flight=k) for _ in range(3)]
        r = sorted(runs, key=lambda x: x["wall_ms"])[1]  # median of three batches
        batch_line = {"value": 2 * k * SEAMS / (r["wall_ms"] * 1e-3), "unit": UNIT, "images": 2 * k, "in_flight": k,
                      "wall_ms": r["wall_ms"], "batches": "median of 3",
                      "path": "tests/harness harness_render_batch -> liblqr-1.so, one host thread + one stream per image"}

    # ---------------- reduce over ranks ------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = t.tolist()
    value = world * SEAMS * args.steps / (dev_ms_max / 1e3)
    e2e_value = world * SEAMS * args.steps / (e2e_ms_max / 1e3)

    if rank == 0:
        peak, peak_src = peaks()
        total_stage = sum(v["ms_per_step"] for v in stages.values()) or 1.0
        dom = max(stages, key=lambda k: stages[k]["ms_per_step"])
        # algorithmic bytes per launch of the dominant kernel
        if dom == "mmap_update":
            n_l = stages[dom]["launches_per_step"]
            alg_bytes = 13.0 * cells / max(n_l, 1)  # per evaluated band cell: en 4 + m 4 + parent 1 read, m 4 written
            note = ("k_band_dp: 13 B per evaluated band cell (compact maps); row-serial chain of h dependent rows on one "
                    "SM -- latency-bound, not bandwidth-bound, see DESIGN.md section 4")
        elif dom == "mmap_full":
            alg_bytes = 8.0 * (W - SEAMS / 2) * H
            note = "full DP: 8 B/px (en read + m written)"
        else:
            alg_bytes = 8.0 * W * H / 2
            note = "index-table shift: 8 B x (W - x_seam) per row"
        dur_s = stages[dom]["us_per_launch"] * 1e-6
        achieved = alg_bytes / dur_s / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of one k_band_dp launch, from the committed ncu --set full capture
        # (profiles/r01_band_dp_ncu_summary.txt: 7.73 MB read, 0 written back before the kernel ends -- the maps are L2-resident)
        traffic = 7.73e6 if dom == "mmap_update" else None
        mf = stages.get("mmap_full")
        roof_full = None
        if mf:
            # launches_per_step counts the row-block launches of all full passes; one pass = all its row blocks
            passes = 3.0
            t_pass = mf["ms_per_step"] / passes * 1e-3
            a_full = 8.0 * (W - SEAMS / 2) * H / t_pass / 1e9
            roof_full = {"bound": "hbm", "achieved": a_full, "peak": peak, "unit": "GB/s", "frac": a_full / peak,
                         "traffic": None, "kernel": "k_mmap_full_strips (one full m-map DP pass = h/32 launches, 8 B/px)",
                         "ms_per_pass": t_pass * 1e3}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "flushed between timed steps (256 MiB write)",
                       "timing": "CUDA events on the engine's stream per step, max over ranks",
                       "mpixel_per_s_carved": value * (W - SEAMS / 2) * H / 1e6},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": W * H * CH,
                    "d2h_bytes_per_step": (W - SEAMS) * H * CH, "ms_per_step": e2e_ms_max / args.steps,
                    "phases_ms": {k: round(v, 3) for k, v in phases.items()},
                    "path": "tests/harness/plugin_sequence.c -> liblqr-1.so C ABI: lqr_carver_new .. resize .. scan_line "
                            "loop, pageable host buffers as the plug-in passes them (rank 0 phases); the harness keeps freed layer-sized "
                            "blocks in the malloc heap (mallopt) and writes into a preallocated pixel region, as a "
                            "long-running host does"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": dom, "share_of_step": stages[dom]["ms_per_step"] / total_stage,
                         "peak_source": peak_src, "note": note},
            "roofline_mmap_full": roof_full,
            "kernels": {k: {kk: round(vv, 4) for kk, vv in v.items()} for k, v in stages.items()},
            "wall_s_timed_region": wall,
        }
        if batch_line:
            line["batch_in_flight"] = batch_line
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
