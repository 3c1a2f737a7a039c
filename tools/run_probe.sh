python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for c in 0 5; do
GEGP_GEMM_CFG=$c python tools/one_eval.py 500 10 1 5 | tail -2
GEGP_GEMM_CFG=$c python tools/one_eval.py 500 10 0 5 | tail -1
done
python tools/one_eval.py 1000 20 1 3 | tail -1
python bench.py --steps 20 --warmup 3 --no-c3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['phases']['cholesky_ms'], d['roofline']['frac'])"
