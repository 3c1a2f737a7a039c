set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t3_all.log
python tools/perf_probe.py > gpurun_out/perf3.log 2>&1
python tools/gemm_one.py 8192 && ncu --set full --clock-control none --import-source on -k regex:gemm_f64 -c 1 -o gpurun_out/gemm_8192 -f python tools/gemm_one.py 8192 > gpurun_out/ncu_gemm.log 2>&1
tail -30 gpurun_out/t3_all.log
cat gpurun_out/perf3.log
tail -5 gpurun_out/ncu_gemm.log
