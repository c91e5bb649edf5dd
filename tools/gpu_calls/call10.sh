#!/bin/bash
# Round-2 GPU call 10: mbarrier wait modes (0 suspend hint / 1 try_wait / 2 test_wait spin): attention variants, timeline, whole step
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c10
mkdir -p $O
for v in "" w1 w2; do
  echo "=== variant '$v'" | tee -a $O/time_attn.log
  TA_LIB_VARIANT=$v timeout 300 python tools/time_attn.py 2 6 8 14 >> $O/time_attn.log 2>&1
done
cat $O/time_attn.log
for v in w1 w2; do
  echo "=== variant '$v'" >> $O/attn_trace.log
  TA_LIB_VARIANT=$v timeout 300 python tools/attn_trace.py 8 6 >> $O/attn_trace.log 2>&1
done
cat $O/attn_trace.log
B="--no-cpu-baseline --no-e2e --no-other-configs --no-dp-parity --steps 8 --warmup 3"
for v in "" w1 w2; do
  TA_LIB_VARIANT=$v timeout 300 python bench.py $B > $O/bench_$v.json 2> $O/bench_$v.err
  echo "variant '$v'"; head -c 330 $O/bench_$v.json | tail -c 200; echo
done
TA_LIB_VARIANT=w2 TA_ATTN_TC=14 timeout 300 python bench.py $B --trace-kernels $O/trace_w2_tc14.txt > $O/bench_w2_tc14.json 2> $O/bench_w2_tc14.err
head -c 330 $O/bench_w2_tc14.json | tail -c 200; echo; head -14 $O/trace_w2_tc14.txt
TA_LIB_VARIANT=w2 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn or gemm" > $O/pytest_w2.log 2>&1; tail -3 $O/pytest_w2.log
