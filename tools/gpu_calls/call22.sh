#!/bin/bash
# Round-2 GPU call 22: persistent prefetching LayerNorm; full GPU suite on the current tree; default bench line
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c22
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_mlp.txt > $O/bench_short.json 2> $O/bench_short.err
grep -n "layernorm_bf16" $O/trace_mlp.txt; head -3 $O/trace_mlp.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<P
import json
for f in ("bench_short","bench"):
    d=[json.loads(l) for l in open("$O/%s.json"%f) if l.startswith("{")][-1]
    print(f, d["ms_per_step"], d["clocks"], d.get("loss"), {k:(v.get("ms_per_step")) for k,v in (d.get("other_configs") or {}).items()})
P
