python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/perf_probe.py 500,10 1000,20 2>&1 | grep -v dgemm
python tools/one_eval.py 500 10 1 5 | tail -2
