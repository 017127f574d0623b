#!/bin/bash
# Multi-GPU evidence on an N-GPU box:  gpurun --gpus N --timeout 900 -- 'bash profiles/capture_multi.sh N r2'
#   tests/multi_gpu_check.py (sharded E-step + in-kernel all-reduce against a one-GPU run), pytest's multi-GPU tests, the bench line
#   at N GPUs (strong scaling of the 3 Gbp job; weak figure, parity check and exchange cost inside the line).
n=${1:-2}
tag=${2:-r2}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29541 tests/multi_gpu_check.py --workload cfg2 > $out/${tag}_multi_gpu_check_n$n.txt 2>&1
tail -3 $out/${tag}_multi_gpu_check_n$n.txt
python -m pytest tests/test_gpu_multi.py -x -q > $out/${tag}_pytest_multi_n$n.txt 2>&1
tail -2 $out/${tag}_pytest_multi_n$n.txt
run 29542 bench.py --gpus $n > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
cat $out/${tag}_bench_n$n.json | cut -c1-1500
run 29543 bench.py --gpus $n --allreduce nccl --no-cpu-baseline --no-binary > $out/${tag}_bench_n${n}_nccl.json 2> $out/${tag}_bench_n${n}_nccl.err
cut -c1-400 $out/${tag}_bench_n${n}_nccl.json
