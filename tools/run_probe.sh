python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for c in 0 4; do
echo "== GEGP_GEMM_CFG=$c"
GEGP_GEMM_CFG=$c python tools/one_eval.py 500 10 1 5 | tail -2
GEGP_GEMM_CFG=$c python tools/one_eval.py 1000 20 1 3 | tail -1
GEGP_GEMM_CFG=$c python tools/c4_scan.py 1024 1 | tail -1
GEGP_GEMM_CFG=$c python tools/c4_scan.py 1024 0 | tail -1
done
python tools/perf_probe.py 500,10 2>&1 | head -6
