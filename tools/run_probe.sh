python -m pytest tests/test_gpu_gp.py -m gpu -x -q 2>&1 | tail -2
python tools/perf_probe.py 2>&1 | grep "build"
python tools/c35_probe.py c5 2>&1 | head -1
