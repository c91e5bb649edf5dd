#!/bin/bash
# Round-2 GPU call 9: 64-key tiles, three CTAs per SM (modes 14-16)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c09
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_attn_fwd" > $O/pytest_attn.log 2>&1
tail -4 $O/pytest_attn.log
timeout 300 python tools/time_attn.py 2 6 8 14 15 16 > $O/time_attn.log 2>&1
cat $O/time_attn.log
TA_ATTN_TC=14 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd5 -s 1 -c 1 -o $O/ncu_attn14 -f python tools/prof_kernels.py attn_enc > $O/ncu_attn14.log 2>&1
B="--no-cpu-baseline --no-e2e --no-other-configs --no-dp-parity --steps 8 --warmup 3"
TA_ATTN_TC=14 timeout 300 python bench.py $B > $O/bench_tc14.json 2> $O/bench_tc14.err
head -c 330 $O/bench_tc14.json | tail -c 200; echo
