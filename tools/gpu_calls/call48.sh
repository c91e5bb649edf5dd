#!/bin/bash
# Round-2 GPU call 48: final binary (norm kernels at 4 rows per block) -- full GPU suite + bench
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c48
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_mlp.txt > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench.json") if l.startswith("{")][-1]
print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"], d["gpu_launches"])
P
grep -n "layernorm\|rmsnorm" $O/trace_mlp.txt | head -4
