#!/bin/bash
# eight B200s: the default bench as the driver launches it, and the C5 end-to-end workload
O=gpurun_out
mkdir -p $O
nvidia-smi -L | wc -l
echo "== bench, N=8 (default workload C4, strong scaling; the driver's flags)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2_bench_c4_8gpu.json 2> $O/r2_bench_c4_8gpu.err; echo "rc=$?"; tail -c 200 $O/r2_bench_c4_8gpu.json; tail -2 $O/r2_bench_c4_8gpu.err
echo "== bench --workload C5, N=8"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload C5 --steps 5 --warmup 3 --no-secondary > $O/r2_bench_c5_8gpu.json 2> $O/r2_bench_c5_8gpu.err; echo "rc=$?"; tail -c 200 $O/r2_bench_c5_8gpu.json; tail -2 $O/r2_bench_c5_8gpu.err
