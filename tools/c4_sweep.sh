#!/bin/bash
# SURVEY.md config 4 on one GPU: N images of 1920x1080 RGBA, 100 seams each, T images in flight (C host threads through
# liblqr-1.so, tests/harness harness_render_batch).  Prints per-image phase times (ms, summed over threads / images) and e2e seams/s.
run() { echo "== $*"; env "$@" timeout 200 python tools/run_configs.py 44 --batch ${BATCH:-96} 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['config4_phase_ms_per_image'], round(d['config4_concurrent']['seams_per_s_e2e'])); print({k: v for k, v in d['config4_engine_host_ms_per_image'].items() if v >= 0.2})"; }
for t in ${THREADS:-1 2 4 8 12 16 24}; do run B200C_THREADS=$t; done
