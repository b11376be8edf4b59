#!/bin/bash
# two B200s: the NCCL-gathered result against one GPU (tests/test_gpu_multi.py) and the default bench at N=2
O=gpurun_out
mkdir -p $O
nvidia-smi -L
echo "== 2-GPU parity test"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/r2_pytest_2gpu.log 2>&1; echo "rc=$?"; tail -4 $O/r2_pytest_2gpu.log
echo "== bench, N=2 (default workload, strong scaling)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2_bench_c4_2gpu.json 2> $O/r2_bench_c4_2gpu.err; echo "rc=$?"; tail -c 300 $O/r2_bench_c4_2gpu.json; tail -3 $O/r2_bench_c4_2gpu.err
echo "== reference arm under torchrun, N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/r2_bench_reference_2gpu.json 2> $O/r2_bench_reference_2gpu.err; echo "rc=$?"; tail -c 300 $O/r2_bench_reference_2gpu.json
