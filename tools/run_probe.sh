ncu --set full --clock-control none --import-source on -k regex:build_cov -c 1 -o gpurun_out/build_c3 -f python tools/one_eval.py 1000 20 0 1 > gpurun_out/ncu_build.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_nt -s 8 -c 1 -o gpurun_out/gemm_tma_c3 -f python tools/one_eval.py 1000 20 1 1 > gpurun_out/ncu_tma3.log 2>&1
ls -la gpurun_out/
