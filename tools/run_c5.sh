timeout 600 python tools/c35_probe.py c5 2>&1 | tee gpurun_out/c5_modes.log
timeout 300 python tools/c35_probe.py c3 2>&1 | tee gpurun_out/c3_predict.log
