#!/bin/bash
# Round-2 GPU call 30: where are the inter-kernel gaps?  (bench --trace-kernels now lists idle time by the kernel that follows it)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c30
mkdir -p $O
timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_gaps.txt > $O/bench.json 2> $O/bench.err
sed -n '/^gaps/,$p' $O/trace_gaps.txt
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench.json") if l.startswith("{")][-1]
print("step", d["ms_per_step"], d["gpu_launches"])
P
