# Round-2 measurement set (one GPU), tag v3.  Part A: tests + bench lines; part B (arg "ncu"): launch lists and --set full captures.
O=gpurun_out
T=r2v3
if [ "$1" != "ncu" ]; then
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $O/${T}_gputests.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
python bench.py --steps 20 --warmup 3 > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
python bench.py --workload c3 --impl reference --steps 1 --warmup 0 > $O/${T}_bench_c3_ref.json 2> $O/${T}_bench_c3_ref.err
python bench.py --workload c3 --steps 5 --warmup 3 > $O/${T}_bench_c3.json 2> $O/${T}_bench_c3.err
tail -2 $O/${T}_bench_c2.err $O/${T}_bench_c3.err
python tools/c35_probe.py c3 > $O/${T}_c3_predict.log 2>&1
python tools/c35_probe.py c5 > $O/${T}_c5_modes.log 2>&1
python tools/one_eval.py 1000 50 1 2 > $O/${T}_c5_lml_grad.log 2>&1
python tools/c4_scan.py 1024 1 > $O/${T}_c4_scan.log 2>&1
python tools/c4_scan.py 1024 0 >> $O/${T}_c4_scan.log 2>&1
python tools/c4_scan.py 128 0 >> $O/${T}_c4_scan.log 2>&1
python tools/chol_probe.py > $O/${T}_chol.log 2>&1
cat $O/${T}_c3_predict.log $O/${T}_c5_modes.log $O/${T}_c5_lml_grad.log $O/${T}_c4_scan.log $O/${T}_chol.log
else
python bench.py --steps 2 --warmup 3 --no-phases > $O/${T}_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_r02v3_c2.csv python bench.py --steps 2 --warmup 3 --no-phases > $O/${T}_ncu_launch.log 2>&1
python tools/one_eval.py 1000 20 1 2 > $O/${T}_plain_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_r02v3_c3.csv python tools/one_eval.py 1000 20 1 2 > /dev/null 2>&1
# c2: TMA launches of one evaluation are the root pair's first product and Kinv_ab = U_ab U_bb^T (the second of each evaluation)
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_nt -s 3 -c 1 -o $O/r02v3_tma_kinvab_c2 -f python tools/one_eval.py 500 10 1 2 > $O/${T}_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tma_nt -s 5 -c 1 -o $O/r02v3_tma_uut_c3 -f python tools/one_eval.py 1000 20 1 2 > $O/${T}_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 150 -c 1 -o $O/r02v3_gemm64_c2 -f python tools/one_eval.py 500 10 1 2 > $O/${T}_ncu_c.log 2>&1
tail -1 $O/${T}_ncu_?.log
fi
