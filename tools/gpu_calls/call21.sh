#!/bin/bash
# Round-2 GPU call 21: encoder LayerNorm walking rows last-to-first (L2 reuse of the producer's tail) -- same-box A/B with in-situ traces
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c21
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "layernorm or encoder" > $O/pytest_ln.log 2>&1; tail -2 $O/pytest_ln.log
for r in 1 0 1 0; do
  TA_LN_REVERSE=$r timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_rev$r.txt > $O/bench_rev$r.json 2> $O/bench_rev$r.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_rev$r.json") if l.startswith("{")][-1]
print("ln_reverse=$r", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"))
P
  grep -n "layernorm_bf16\|256, 7\|256, 1, false" $O/trace_rev$r.txt
done
