#!/bin/bash
# Round-2 GPU call 5: P packed on the integer pipes (F2FP off the XU pipe) -- all attention forward kernels
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c05
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_attn" > $O/pytest_attn.log 2>&1
tail -5 $O/pytest_attn.log
timeout 300 python tools/time_attn.py 2 3 11 7 8 10 12 13 > $O/time_attn.log 2>&1
cat $O/time_attn.log
B="--no-cpu-baseline --no-e2e --no-other-configs --no-dp-parity --steps 8 --warmup 3"
TA_ATTN_TC=2 timeout 300 python bench.py $B --trace-kernels $O/trace_tc2.txt > $O/bench_tc2.json 2> $O/bench_tc2.err
TA_ATTN_TC=8 timeout 300 python bench.py $B --trace-kernels $O/trace_tc8.txt > $O/bench_tc8.json 2> $O/bench_tc8.err
for f in $O/bench_tc*.json; do echo $f; head -c 330 $f | tail -c 200; echo; done
head -8 $O/trace_tc2.txt; head -5 $O/trace_tc8.txt
timeout 600 python -m pytest tests/test_path_gpu.py -m gpu -q -k "full_size or small_configs or greedy" > $O/pytest_path.log 2>&1
tail -4 $O/pytest_path.log
