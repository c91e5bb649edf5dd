#!/bin/bash
# Round-2 GPU call 15 (2 GPUs): bench.py under torchrun at N=2 -- DP parity on the shared 64-clip global batch, other configs, NCCL all-reduce
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c15
mkdir -p $O
timeout 300 python -m pytest tests/test_path_gpu.py -m gpu -q -k "device_prompt" > $O/pytest_fix.log 2>&1; tail -3 $O/pytest_fix.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3) > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 2500 $O/bench_n2.json; echo; tail -5 $O/bench_n2.err
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0) > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
head -c 300 $O/bench_ref_n2.json; echo
