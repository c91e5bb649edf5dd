#!/bin/bash
# Round-2 GPU call 25: pipelined 64-query attention backward (variant 2), decoder forward ring variants, new 260-token full-size fixture
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c25
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn_bwd" > $O/pytest_bwd.log 2>&1; tail -4 $O/pytest_bwd.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn_fwd_decoder" > $O/pytest_fwd.log 2>&1; tail -2 $O/pytest_fwd.log
timeout 300 python tools/time_lm_attn.py > $O/time_lm_attn.log 2>&1; cat $O/time_lm_attn.log
timeout 1200 python -m pytest tests/test_path_gpu.py -m gpu -q > $O/pytest_path.log 2>&1; tail -5 $O/pytest_path.log; grep -n "full_b" $O/pytest_path.log | head
for v in 2 1; do
  TA_ATTN_BWD_VARIANT=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_bwd$v.txt > $O/bench_bwd$v.json 2> $O/bench_bwd$v.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_bwd$v.json") if l.startswith("{")][-1]
print("attn_bwd_variant=$v", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"))
P
  grep -n "attn_tc_bwd\|attn_tc_fwd6" $O/trace_bwd$v.txt
done
