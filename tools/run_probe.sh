python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/perf_probe.py 2>&1 | grep "potr\|lml"
