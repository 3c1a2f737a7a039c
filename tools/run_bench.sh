set -x
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err
cat gpurun_out/bench_ref.json gpurun_out/bench_b200.json
tail -3 gpurun_out/bench_b200.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_v2_c2.csv python bench.py --steps 2 --warmup 3 --no-phases > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_nt -s 5 -c 1 -o gpurun_out/gemm_tma_c2 -f python tools/one_eval.py 500 10 1 1 > gpurun_out/ncu_tma.log 2>&1
tail -2 gpurun_out/ncu_tma.log
