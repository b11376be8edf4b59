#!/bin/bash
# two B200s: the NCCL-gathered results against one GPU (tests/test_gpu_multi.py, incl. the 2-D sharded nmfp), the
# single-GPU block test, and the benches at N=2
O=gpurun_out
mkdir -p $O
nvidia-smi -L
echo "== 2-GPU parity test + nmfp tests"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_nmfp.py -m gpu -q > $O/r2_pytest_2gpu.log 2>&1; echo "rc=$?"; tail -4 $O/r2_pytest_2gpu.log
echo "== bench C3, N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --workload C3 --steps 5 --warmup 3 --no-secondary > $O/r2_bench_c3_2gpu.json 2> $O/r2_bench_c3_2gpu.err; echo "rc=$?"; tail -c 300 $O/r2_bench_c3_2gpu.json; tail -3 $O/r2_bench_c3_2gpu.err
echo "== bench C5, N=2"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --workload C5 --steps 2 --warmup 1 --no-secondary > $O/r2_bench_c5_2gpu.json 2> $O/r2_bench_c5_2gpu.err; echo "rc=$?"; tail -c 300 $O/r2_bench_c5_2gpu.json; tail -3 $O/r2_bench_c5_2gpu.err
