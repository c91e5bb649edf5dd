#!/bin/bash
# Round-2 GPU call 47: final tree -- full GPU suite, smoke(), default bench line, reference arm
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c47
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench.json") if l.startswith("{")][-1]
print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"], d["gpu_launches"])
print({k:(v.get("ms_per_step")) for k,v in (d.get("other_configs") or {}).items()})
P
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; head -c 400 $O/bench_ref.json; echo
