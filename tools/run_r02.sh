# Round-2 measurement set (one GPU).  Part A: tests + bench lines; part B (arg "ncu"): launch lists and --set full captures.
O=gpurun_out
if [ "$1" != "ncu" ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $O/r2v2_gputests.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2v2_bench_ref.json 2> $O/r2v2_bench_ref.err
python bench.py --steps 20 --warmup 3 > $O/r2v2_bench_c2.json 2> $O/r2v2_bench_c2.err
python bench.py --workload c3 --impl reference --steps 1 --warmup 0 > $O/r2v2_bench_c3_ref.json 2> $O/r2v2_bench_c3_ref.err
python bench.py --workload c3 --steps 5 --warmup 3 > $O/r2v2_bench_c3.json 2> $O/r2v2_bench_c3.err
tail -2 $O/r2v2_bench_c2.err $O/r2v2_bench_c3.err
python tools/c35_probe.py c3 > $O/r2v2_c3_predict.log 2>&1
python tools/c35_probe.py c5 > $O/r2v2_c5_modes.log 2>&1
python tools/one_eval.py 1000 50 1 2 > $O/r2v2_c5_lml_grad.log 2>&1
python tools/c4_scan.py 1024 1 > $O/r2v2_c4_scan.log 2>&1
python tools/c4_scan.py 1024 0 >> $O/r2v2_c4_scan.log 2>&1
python tools/c4_scan.py 128 0 >> $O/r2v2_c4_scan.log 2>&1
python tools/chol_probe.py > $O/r2v2_chol.log 2>&1
cat $O/r2v2_c3_predict.log $O/r2v2_c5_modes.log $O/r2v2_c5_lml_grad.log $O/r2v2_c4_scan.log $O/r2v2_chol.log
else
python bench.py --steps 2 --warmup 3 --no-phases > $O/r2v2_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_r02v2_c2.csv python bench.py --steps 2 --warmup 3 --no-phases > $O/r2v2_ncu_launch.log 2>&1
python tools/one_eval.py 1000 20 1 2 > $O/r2v2_plain_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_r02v2_c3.csv python tools/one_eval.py 1000 20 1 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_nt -s 5 -c 1 -o $O/r02v2_tma_uut_c3 -f python tools/one_eval.py 1000 20 1 2 > $O/r2v2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 150 -c 1 -o $O/r02v2_gemm64_c2 -f python tools/one_eval.py 500 10 1 2 > $O/r2v2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lml_grad_kernel -s 1 -c 1 -o $O/r02v2_lmlgrad_c2 -f python tools/one_eval.py 500 10 1 2 > $O/r2v2_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lml_grad_kernel -s 1 -c 1 -o $O/r02v2_lmlgrad_c3 -f python tools/one_eval.py 1000 20 1 2 > $O/r2v2_ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:build_cov_fast -s 1 -c 1 -o $O/r02v2_build_c3 -f python tools/one_eval.py 1000 20 1 2 > $O/r2v2_ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_k128 -s 60 -c 1 -o $O/r02v2_k128_c2 -f python tools/one_eval.py 500 10 1 2 > $O/r2v2_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:leaf_trsm -s 60 -c 1 -o $O/r02v2_ltrsm_c2 -f python tools/one_eval.py 500 10 1 2 > $O/r2v2_ncu_g.log 2>&1
tail -1 $O/r2v2_ncu_?.log
fi
