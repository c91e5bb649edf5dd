#!/bin/bash
# Round-2 GPU call 12: convergent MMA-issue warps (uniform-register descriptors) in the encoder attention kernels
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c12
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_attn_fwd" > $O/pytest_attn.log 2>&1
tail -4 $O/pytest_attn.log
timeout 300 python tools/time_attn.py 2 6 7 8 12 14 15 > $O/time_attn.log 2>&1
cat $O/time_attn.log
timeout 300 python tools/attn_trace.py 8 > $O/attn_trace.log 2>&1
cat $O/attn_trace.log
B="--no-cpu-baseline --no-e2e --no-other-configs --no-dp-parity --steps 8 --warmup 3"
TA_ATTN_TC=14 timeout 300 python bench.py $B > $O/bench_tc14.json 2> $O/bench_tc14.err
head -c 330 $O/bench_tc14.json | tail -c 200; echo
