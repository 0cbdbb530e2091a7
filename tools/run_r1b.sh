set -x
python tools/pipe_ab.py 4096 1048576 2>&1 | tail -4
python tools/pipe_ab.py 16384 131072 2>&1 | tail -4
python tools/pipe_ab.py 4096 1048576 int8x6 2>&1 | tail -2
python -m pytest tests -m gpu -x -q -k "int8 or refit" 2>&1 | tail -3
