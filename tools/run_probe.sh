python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/one_eval.py 500 10 1 4 | tail -2
python tools/one_eval.py 1000 20 1 3 | tail -1
python tools/one_eval.py 200 5 1 3 64 | tail -1
