python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/one_eval.py 1000 20 1 2 | tail -1
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 30 -c 1 -o gpurun_out/gemm64_c2 -f python tools/one_eval.py 500 10 1 1 > gpurun_out/ncu_g64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:potf2_inv -s 20 -c 1 -o gpurun_out/potf2_v2 -f python tools/one_eval.py 500 10 1 1 > gpurun_out/ncu_potf2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:leaf_trsm -s 5 -c 1 -o gpurun_out/ltrsm_v2 -f python tools/one_eval.py 500 10 1 1 > gpurun_out/ncu_ltrsm.log 2>&1
ls gpurun_out/*.ncu-rep
