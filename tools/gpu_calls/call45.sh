#!/bin/bash
# Round-2 GPU call 45: SwiGLU-backward with THREE in-place box sets and a 3-stage operand ring
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c45
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "swiglu" > $O/pytest_swiglu.log 2>&1; tail -2 $O/pytest_swiglu.log
timeout 300 python tools/time_ffn.py > $O/time_ffn.log 2>&1; sed -n 3,8p $O/time_ffn.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_mlp.txt > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench.json") if l.startswith("{")][-1]
print("step", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"))
P
grep -n "256, 6\|256, 5," $O/trace_mlp.txt
