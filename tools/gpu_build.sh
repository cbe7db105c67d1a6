timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/time_build.py 2>&1 | tail -6
