run() { echo "== $*"; env "$@" timeout 200 python tools/run_configs.py 44 --batch 96 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['config4_phase_ms_per_image'], round(d['config4_concurrent']['seams_per_s_e2e']))"; }
for t in 1 4 8 12 16 24 32; do run B200C_THREADS=$t; done
