python tools/one_eval.py 1000 50 1 2 2>&1 | tee gpurun_out/c5_lml_grad.log
python bench.py --steps 10 --warmup 3 --no-c3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['batched'])"
