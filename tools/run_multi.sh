G=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 20 --warmup 3 > gpurun_out/bench_g$G.json 2> gpurun_out/bench_g$G.err
tail -c 300 gpurun_out/bench_g$G.err; cut -c1-700 gpurun_out/bench_g$G.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 tools/c4_scan.py 1024 1 2>&1 | grep -v Warn | tee gpurun_out/c4_scan_g$G.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 tools/c4_scan.py 1024 0 2>&1 | grep -v Warn | tee -a gpurun_out/c4_scan_g$G.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $G --steps 1 --warmup 1 | cut -c1-200
