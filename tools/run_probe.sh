python tools/dbg_leaf.py 40 | head -4
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t5_all.log
python tools/perf_probe.py 500,10 > gpurun_out/perf5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches5.csv python tools/perf_probe.py 500,10 > gpurun_out/ncu5.log 2>&1
tail -30 gpurun_out/t5_all.log
cat gpurun_out/perf5.log
