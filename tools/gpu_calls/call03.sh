#!/bin/bash
# Round-2 GPU call 3: ping-pong persistent encoder attention (ta_attn_set_tc 6 / 7): parity, timing, whole-step effect
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c03
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "test_attn_fwd" > $O/pytest_attn.log 2>&1
tail -5 $O/pytest_attn.log
timeout 300 python -m pytest tests/test_ddp_gpu.py tests/test_path_gpu.py -m gpu -q -k "ddp or generic_projector" > $O/pytest_fixed.log 2>&1
tail -3 $O/pytest_fixed.log
timeout 300 python tools/time_attn.py 2 6 7 > $O/time_attn.log 2>&1
cat $O/time_attn.log
B="--no-cpu-baseline --no-e2e --no-other-configs --no-dp-parity --steps 8 --warmup 3"
TA_ATTN_TC=2 timeout 300 python bench.py $B > $O/bench_tc2.json 2> $O/bench_tc2.err
TA_ATTN_TC=6 timeout 300 python bench.py $B --trace-kernels $O/trace_tc6.txt > $O/bench_tc6.json 2> $O/bench_tc6.err
TA_ATTN_TC=7 timeout 300 python bench.py $B > $O/bench_tc7.json 2> $O/bench_tc7.err
for f in $O/bench_tc*.json; do echo $f; head -c 330 $f | tail -c 200; echo; done
TA_ATTN_TC=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd3 -s 1 -c 1 -o $O/ncu_attn3 -f \
    python tools/prof_kernels.py attn_enc > $O/ncu_attn3.log 2>&1
ls -la $O
