#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpus.txt
free -g >> gpurun_out/gpus.txt
NG=${NG:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/dist_check.log 2>&1
echo "dist_check exit $?" | tee -a gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $NG --steps 2 --warmup 1 --qubits ${SMALLQ:-31} > gpurun_out/bench_dist_small.json 2> gpurun_out/bench_dist_small.err
echo "bench small exit $?" | tee -a gpurun_out/summary.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $NG --steps 2 --warmup 1 > gpurun_out/bench_dist.json 2> gpurun_out/bench_dist.err
echo "bench full exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/dist_check.log; tail -n 3 gpurun_out/bench_dist_small.err gpurun_out/bench_dist.err
cat gpurun_out/bench_dist_small.json gpurun_out/bench_dist.json | cut -c1-1500
