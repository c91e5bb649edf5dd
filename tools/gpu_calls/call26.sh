#!/bin/bash
# Round-2 GPU call 26: occupancy diagnostic of the decoder forward kernel; bench A/B of ring 3 vs 4
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c26
mkdir -p $O
timeout 300 python tools/time_lm_attn.py > $O/time_lm_attn.log 2>&1; cat $O/time_lm_attn.log
for v in 4 3 4 3; do
  TA_ATTN_TC_LM=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_ring$v.txt > $O/bench_ring$v.json 2> $O/bench_ring$v.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_ring$v.json") if l.startswith("{")][-1]
print("ring=$v", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"))
P
  grep -n "attn_tc_fwd6" $O/trace_ring$v.txt
done
