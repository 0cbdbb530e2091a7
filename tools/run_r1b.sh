set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "int8 or spatial or refit or overlap or diffeo_apply" 2>&1 | tail -3
timeout 120 python tools/spatial_time.py 4096 65536 2>&1 | tail -1
timeout 120 python tools/spatial_time.py 16384 65536 2>&1 | tail -1
timeout 120 python tools/ozaki_time.py 4096 262144 2>&1 | tail -1
timeout 120 python tools/ozaki_time.py 16384 65536 2>&1 | tail -1
