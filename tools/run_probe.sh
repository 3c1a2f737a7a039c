python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/e2e_probe.py 2>&1 | head -8
python bench.py --steps 20 --warmup 3 --no-c3 > gpurun_out/bench_b200_v3.json 2> gpurun_out/bench_b200_v3.err; tail -2 gpurun_out/bench_b200_v3.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_b200_v3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['phases'])
"
