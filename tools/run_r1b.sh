set -x
python tools/ozaki_time.py 4096 262144 2>&1 | tail -3
python tools/power_probe.py 4096 1048576 2>&1 | grep -E "serialised|A0"
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
