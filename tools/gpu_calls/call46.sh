#!/bin/bash
# Round-2 GPU call 46: encoder attention with one exponential in eight on the FMA pipe (mode 15) vs all on the SFUs (mode 14), in the step
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c46
mkdir -p $O
for v in 14 15 14 15; do
  TA_ATTN_TC=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --no-e2e > $O/bench_tc$v.json 2> $O/bench_tc$v.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_tc$v.json") if l.startswith("{")][-1]
print("attn_tc=$v", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"), round(d["roofline"]["encoder_attention"]["us"],1))
P
done
