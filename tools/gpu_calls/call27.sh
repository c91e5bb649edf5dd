#!/bin/bash
# Round-2 GPU call 27: full GPU suite + timings + default bench line on the current tree (longest-first grids, 4-slot ring default)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c27
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 300 python tools/time_lm_attn.py > $O/time_lm_attn.log 2>&1; cat $O/time_lm_attn.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench.json") if l.startswith("{")][-1]
print("bench", d["value"], d["ms_per_step"], d["e2e"], d["clocks"], d.get("loss"), d.get("parity"))
print({k:(v.get("ms_per_step")) for k,v in (d.get("other_configs") or {}).items()})
print({k:(v.get("tflops"), v.get("frac")) for k,v in d["roofline"]["qwen3_ffn"].items() if isinstance(v, dict)}, d["roofline"]["achieved"], d["roofline"]["frac"])
P
