python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/one_eval.py 500 10 1 4 | tail -2
GEGP_SPLIT_TILES=400 python tools/one_eval.py 500 10 1 4 | tail -1
GEGP_SPLIT_TILES=100000 python tools/one_eval.py 500 10 1 4 | tail -1
python tools/one_eval.py 500 10 0 4 | tail -1
python tools/one_eval.py 1000 20 1 3 | tail -1
GEGP_SPLIT_TILES=100000 python tools/one_eval.py 1000 20 1 3 | tail -1
python tools/one_eval.py 1000 20 0 3 | tail -1
GEGP_SPLIT_TILES=100000 python tools/one_eval.py 1000 20 0 3 | tail -1
python tools/one_eval.py 200 5 1 3 64 | tail -1
python bench.py --steps 10 --warmup 3 --no-c3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['phases']['cholesky_ms'])"
