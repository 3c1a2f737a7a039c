# Round-end measurement set (one GPU): tests, both bench arms, launch lists, ncu --set full captures, probes.
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err
tail -3 gpurun_out/bench_b200.err
python tools/perf_probe.py > gpurun_out/perf_probe.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_v3_c2.csv python bench.py --steps 2 --warmup 3 --no-phases > gpurun_out/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_v3_c3.csv python tools/one_eval.py 1000 20 1 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 60 -c 1 -o gpurun_out/gemm64_c2 -f python tools/one_eval.py 500 10 1 1 > gpurun_out/ncu_g64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_nt -s 8 -c 1 -o gpurun_out/gemm_tma_c3 -f python tools/one_eval.py 1000 20 1 1 > gpurun_out/ncu_tma3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:build_cov_fast -s 0 -c 1 -o gpurun_out/buildfast_c3 -f python tools/one_eval.py 1000 20 1 1 > gpurun_out/ncu_build.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:potf2_inv -s 20 -c 1 -o gpurun_out/potf2_c2 -f python tools/one_eval.py 500 10 1 1 > gpurun_out/ncu_potf2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:leaf_trsm -s 10 -c 1 -o gpurun_out/ltrsm_c2 -f python tools/one_eval.py 500 10 1 1 > gpurun_out/ncu_ltrsm.log 2>&1
python tools/c35_probe.py c3 > gpurun_out/c3_predict.log 2>&1
python tools/c4_scan.py 1024 1 > gpurun_out/c4_scan.log 2>&1
python tools/c4_scan.py 1024 0 >> gpurun_out/c4_scan.log 2>&1
cat gpurun_out/c3_predict.log gpurun_out/c4_scan.log
