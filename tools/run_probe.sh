python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/dbg_clocks.py | tail -16
python tools/one_eval.py 500 10 1 4
python tools/one_eval.py 1000 20 1 3
