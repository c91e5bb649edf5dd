#!/bin/bash
# Round-2 GPU call 13: whole-warp MMA issue in every tcgen05 kernel (GEMM 1-CTA / 2-CTA, attention fwd x4, attention bwd): tests + step
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c13
mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1
tail -6 $O/pytest_gpu.log
timeout 300 python tools/time_attn.py 2 14 15 > $O/time_attn.log 2>&1; cat $O/time_attn.log
timeout 300 python tools/time_ffn.py > $O/time_ffn.log 2>&1; cat $O/time_ffn.log
B="--no-cpu-baseline --no-e2e --no-other-configs --no-dp-parity --steps 8 --warmup 3"
TA_ATTN_TC=2 timeout 300 python bench.py $B --trace-kernels $O/trace_tc2.txt > $O/bench_tc2.json 2> $O/bench_tc2.err
TA_ATTN_TC=15 timeout 300 python bench.py $B --trace-kernels $O/trace_tc15.txt > $O/bench_tc15.json 2> $O/bench_tc15.err
for f in $O/bench_tc*.json; do echo $f; head -c 330 $f | tail -c 200; echo; done
head -22 $O/trace_tc15.txt
