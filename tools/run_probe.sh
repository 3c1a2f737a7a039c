python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t8_all.log
tail -5 gpurun_out/t8_all.log
python tools/one_eval.py 500 10 0 4
python tools/one_eval.py 500 10 1 4
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches8_nograd.csv python tools/one_eval.py 500 10 0 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches8_grad.csv python tools/one_eval.py 500 10 1 2 > /dev/null 2>&1
